"""Oracle for the Double64 path WITH a functional: the SCF step restated in mpmath (50 digits).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure-Python loops over mpmath numbers: small discretizations
only (order 6, 8 cells, 64 Gauss points: a step takes seconds).

Why it exists.  The reference runs the same generic Julia code with T = Double64
(src/discretization/discretization.jl:289, assemble.jl:277-319); it commits a Double64 trajectory only for a model
WITHOUT a functional (examples/basic_example_doublefloat), so the double-double kernels that evaluate the functional
(ddk_xc), assemble Vxc (ddk_H), solve the eigenproblem (ddk_eig) and run Brent on Exc (ddk_decide) had no independent
check beyond Float64 accuracy.  This module is that check: the formulas of oracle/scf.py (each of which cites the
reference function it follows) evaluated in 50-digit arithmetic ON THE SAME double-double tables the device
receives, with the number-type mixture of a Double64 run (`1/4pi` and `4pi` are Float64
constants, the functionals as in oracle/xc_dd.py).  It is pinned itself: rounded to Float64 precision its steps must
reproduce oracle/scf.py on the same discretization (tests/test_dd_host.py).

Scope: nspin = 1, OptimizedAufbau with full / singly partial last shells (the two-fold degenerate branch raises),
ODA with Brent (oracle/brent.py's state machine in mpmath numbers) or frozen_t.
"""
from __future__ import annotations

import mpmath as mp

from . import xc_dd
from .brent import brent_minimize

F64_INV4PI = mp.mpf(0.07957747154594767)   # Julia `1/4pi` (assemble.jl:123): a Float64 constant
F64_FOURPI = mp.mpf(12.566370614359172)    # Julia `4pi`: Int * Irrational -> Float64 (density.jl:121, energies.jl:221)


def _mp_array(dd_arr):
    """DD array -> nested lists of mpf (exact: hi + lo)."""
    import numpy as np
    hi, lo = np.asarray(dd_arr.hi), np.asarray(dd_arr.lo)

    def rec(h, l):
        if h.ndim == 0:
            return mp.mpf(float(h)) + mp.mpf(float(l))
        return [rec(h[i], l[i]) for i in range(h.shape[0])]
    return rec(hi, lo)


class OracleMP:
    def __init__(self, basis, ops_dd, quad_dd, *, Z, N, lh, nh, ex=True, ec=True, hartree=1, scftol="1e-20", tinit="0.6",
                 frozen_t=False, degen_tol="1e-3", abstol_ls=None, reltol_ls=None, maxiter_ls=100, dps=50):
        mp.mp.dps = dps
        self.b = basis
        self.Z, self.N, self.hartree = mp.mpf(Z), mp.mpf(N), mp.mpf(hartree)
        self.lh, self.nh, self.L = lh, nh, lh + 1
        self.ex, self.ec = bool(ex), bool(ec)
        self.scftol, self.tinit, self.frozen_t = mp.mpf(scftol), mp.mpf(tinit), frozen_t
        self.degen_tol = mp.mpf(degen_tol)
        eps_dd = mp.mpf("4.930380657631324e-32")          # eps(Double64); ODA tolerances default to 1e4 eps (oda/types.jl:39-40)
        self.abstol_ls = mp.mpf(abstol_ls) if abstol_ls else 10000 * eps_dd
        self.reltol_ls = mp.mpf(reltol_ls) if reltol_ls else 10000 * eps_dd
        self.maxiter_ls = maxiter_ls
        self.C, self.n, self.Nh = basis.ncells, basis.ordermax + 1, basis.size
        self.Rmax = mp.mpf(float(basis.mesh.points[-1]))
        self.Ib = [list(map(int, v)) for v in basis.cells_to_indices]
        self.Ig = [list(map(int, v)) for v in basis.cells_to_generators]
        loc = {k: _mp_array(getattr(ops_dd, k + "_loc")) for k in ("M0", "A", "Mm1", "Mm2")}
        self.F = _mp_array(ops_dd.F_loc)                  # [C][n][n][n], generator order
        self.M0, self.A, self.Mm1, self.Mm2 = (self._assemble(loc[k]) for k in ("M0", "A", "Mm1", "Mm2"))
        q = quad_dd
        self.Q = q.npoints
        self.x, self.w, self.y, self.wy = (_mp_array(v) for v in (q.x, q.w, q.y, q.wy))
        self.Pgx, self.Pgy = _mp_array(q.Pgenx), _mp_array(q.Pgeny)   # [n][Q]
        self.inv = _mp_array(q.invshifts)                              # [C][2]
        self.ycell = [int(v) for v in q.ycell]
        self.Kin = [(self.A + l * (l + 1) * self.Mm2) / 2 for l in range(self.L)]
        self.Coulomb = -self.Z * self.Mm1
        # M0 = Lc Lc^T once (generalized eigenproblems of every step)
        self.Lc = mp.cholesky(self.M0)
        self.Lci = mp.inverse(self.Lc)

    def _assemble(self, loc):
        G = mp.zeros(self.Nh, self.Nh)
        for c in range(self.C):
            for a, ga in zip(self.Ib[c], self.Ig[c]):
                for b_, gb in zip(self.Ib[c], self.Ig[c]):
                    G[a, b_] += loc[c][ga][gb]
        return G

    # ---- assemble_hartree!, tensor contractions (assemble.jl:63-83, free allocation.jl:17-45)
    def tensor_matrix(self, D):
        B = [mp.mpf(0)] * self.Nh
        for c in range(self.C):
            Ib, Ig = self.Ib[c], self.Ig[c]
            for m, gm in zip(Ib, Ig):
                acc = mp.mpf(0)
                for i, gi in zip(Ib, Ig):
                    for j, gj in zip(Ib, Ig):
                        acc += D[i, j] * self.F[c][gi][gj][gm]
                B[m] += acc
        return mp.matrix(B)

    def poisson(self, B):
        return mp.lu_solve(self.A, B)

    def hartree_matrix(self, W):
        H = mp.zeros(self.Nh, self.Nh)
        for c in range(self.C):
            Ib, Ig = self.Ib[c], self.Ig[c]
            for i, gi in zip(Ib, Ig):
                for j, gj in zip(Ib, Ig):
                    acc = mp.mpf(0)
                    for m, gm in zip(Ib, Ig):
                        acc += W[m] * self.F[c][gi][gj][gm]
                    H[i, j] += acc
        H = (H + self.N / self.Rmax * self.M0) * self.hartree
        return (H + H.T) / 2

    # ---- optimized_eval_density! + evaluate_vrho! + fill_local_matrix! (assemble.jl:91-125,277-286; local matrix.jl:129-155)
    def cell_density(self, D, c):
        Ib, Ig = self.Ib[c], self.Ig[c]
        inv1, inv2 = self.inv[c]
        out = []
        for q in range(self.Q):
            acc = mp.mpf(0)
            for i, gi in zip(Ib, Ig):
                for j, gj in zip(Ib, Ig):
                    acc += D[i, j] * (self.Pgx[gi][q] * self.Pgx[gj][q])
            X = inv1 * self.x[q] + inv2
            out.append(acc / X ** 2 * F64_INV4PI)
        return out

    def vxc_matrix(self, D):
        V = mp.zeros(self.Nh, self.Nh)
        if not (self.ex or self.ec):
            return V
        for c in range(self.C):
            rho = self.cell_density(D, c)
            f = [xc_dd.vrho(rho[q], self.ex, self.ec) * self.w[q] for q in range(self.Q)]
            inv1 = self.inv[c][0]
            Ib, Ig = self.Ib[c], self.Ig[c]
            for i, gi in zip(Ib, Ig):
                for j, gj in zip(Ib, Ig):
                    acc = mp.mpf(0)
                    for q in range(self.Q):
                        acc += (self.Pgx[gi][q] * self.Pgx[gj][q]) * f[q]
                    V[i, j] += acc * inv1
        return (V + V.T) / 2

    # ---- find_orbital! (oda/loop.jl:65-77): lowest nh pairs of (H_l, M0), U^T M0 U = I
    def find_orbital(self, Har, Vxc):
        eps, U = [], []
        for l in range(self.L):
            H = self.Kin[l] + self.Coulomb + Vxc + Har
            Cm = self.Lci * H * self.Lci.T
            Cm = (Cm + Cm.T) / 2
            w, V = mp.eigsy(Cm)
            order = sorted(range(self.Nh), key=lambda i: w[i])[: self.nh]
            eps.append([w[i] for i in order])
            U.append([self.Lci.T * V[:, i] for i in order])
        return eps, U

    # ---- aufbau! (optimized.jl:99-222), nspin = 1, capacities 4l + 2
    def aufbau(self, eps):
        flat = [(eps[l][k], l + self.L * k, l, k) for k in range(self.nh) for l in range(self.L)]
        flat.sort(key=lambda v: (v[0], v[1]))            # stable sortperm of vec(eps) (column-major index l + L k)
        n = [[mp.mpf(0)] * self.nh for _ in range(self.L)]
        left = self.N
        i = 0
        while left > 0 and i < len(flat):
            e0, _, l, k = flat[i]
            if i + 1 < len(flat) and abs(flat[i + 1][0] - e0) < self.degen_tol and left < (4 * l + 2) + (4 * flat[i + 1][2] + 2):
                raise NotImplementedError("two-fold degenerate last shell: outside the scope of the mpmath oracle")
            cap = mp.mpf(4 * l + 2)
            if left >= cap:
                n[l][k] = cap
                left -= cap
            else:
                n[l][k] = left
                break
            i += 1
        return n

    # ---- density! and the energies (density.jl:15-41, energies.jl:7-141,217-233)
    def density(self, U, n):
        D = mp.zeros(self.Nh, self.Nh)
        for k in range(self.nh):
            for l in range(self.L):
                if n[l][k] != 0:
                    u = U[l][k]
                    D += n[l][k] * (u * u.T)
        return D

    def grid_density(self, D):
        out = []
        for q in range(self.Q):
            c = self.ycell[q]
            Ib, Ig = self.Ib[c], self.Ig[c]
            s = mp.mpf(0)
            for a, ga in zip(Ib, Ig):
                for b_, gb in zip(Ib, Ig):
                    s += self.Pgy[ga][q] * D[a, b_] * self.Pgy[gb][q]
            out.append(s / (F64_FOURPI * self.y[q] ** 2))
        return out

    def exc_of_grid(self, rho):
        if not (self.ex or self.ec):
            return mp.mpf(0)
        acc = mp.mpf(0)
        for q in range(self.Q):
            acc += self.wy[q] * (rho[q] * xc_dd.zk(rho[q], self.ex, self.ec) * self.y[q] ** 2)
        return F64_FOURPI * acc

    def orbital_sum(self, U, n, mats):
        e = mp.mpf(0)
        for l in range(self.L):
            for k in range(self.nh):
                if n[l][k] != 0:
                    u = U[l][k]
                    e += n[l][k] * (u.T * (mats[l] * u))[0]
        return e

    def initial_state(self):
        return dict(D=mp.zeros(self.Nh, self.Nh), rho_y=[mp.mpf(0)] * self.Q, B=mp.matrix([0] * self.Nh),
                    W=mp.matrix([0] * self.Nh), E=dict(Etot=mp.mpf(0), Ekin=mp.mpf(0), Ecou=mp.mpf(0), Ehar=mp.mpf(0),
                                                       Eexc=mp.mpf(0)), t=self.tinit, niter=0, stop=mp.mpf(0))

    def scf_step(self, st):
        """scf_step!(alg::ODA, solver) (oda/loop.jl:1-46) in mpmath; mutates and returns st (adds eps, n)."""
        Dp, Ep = st["D"], st["E"]
        har_on = self.hartree != 0
        Har = self.hartree_matrix(st["W"]) if har_on else mp.zeros(self.Nh, self.Nh)
        Vxc = self.vxc_matrix(Dp)
        eps, U = self.find_orbital(Har, Vxc)
        n = self.aufbau(eps)
        D = self.density(U, n)
        B = self.tensor_matrix(D) if har_on else mp.matrix([0] * self.Nh)
        W = self.poisson(B) if har_on else mp.matrix([0] * self.Nh)
        n2r = self.N ** 2 / self.Rmax
        Ekin = self.orbital_sum(U, n, self.Kin)
        Ecou = self.orbital_sum(U, n, [self.Coulomb] * self.L)
        Ehar = ((B.T * W)[0] + n2r) / 2 if har_on else mp.mpf(0)
        rho_new = self.grid_density(D)
        Eexc = self.exc_of_grid(rho_new)
        sne = sum(n[l][k] * eps[l][k] for l in range(self.L) for k in range(self.nh))
        corr = sum(Vxc[i, j] * D[i, j] for i in range(self.Nh) for j in range(self.Nh)) if (self.ex or self.ec) else mp.mpf(0)
        Etot = sne - Ehar + Eexc - corr
        t = st["t"]
        if st["niter"] > 0:
            h01 = ((st["B"].T * W)[0] + n2r) / 2 if har_on else mp.mpf(0)
            rho_p = st["rho_y"]
            F = lambda tt: self.exc_of_grid([tt * a + (1 - tt) * b_ for a, b_ in zip(rho_new, rho_p)])
            if self.frozen_t:
                etot = None
            else:
                A0, A1 = Ekin + Ecou, Ep["Ekin"] + Ep["Ecou"]
                qa = Ehar + Ep["Ehar"] - 2 * h01
                qb = (A0 - A1) + 2 * (h01 - Ep["Ehar"])
                qc = A1 + Ep["Ehar"]
                if self.ex or self.ec:
                    t, etot, _ = brent_minimize(lambda tt: qa * tt ** 2 + qb * tt + qc + F(tt), mp.mpf(0), mp.mpf(1),
                                                abs_tol=self.abstol_ls, rel_tol=self.reltol_ls, iterations=self.maxiter_ls,
                                                sqrt=mp.sqrt)
                else:
                    t = min(max(-qb / (2 * qa), mp.mpf(0)), mp.mpf(1)) if qa > 0 else (mp.mpf(0) if qc <= qc + qb + qa else mp.mpf(1))
                    etot = qc + qb * t + qa * t ** 2
            ekin = t * Ekin + (1 - t) * Ep["Ekin"]
            ecou = t * Ecou + (1 - t) * Ep["Ecou"]
            ehar = t ** 2 * Ehar + (1 - t) ** 2 * Ep["Ehar"] + 2 * t * (1 - t) * h01
            if self.frozen_t:
                eexc = F(t)
                etot = ekin + ecou + ehar + eexc
            else:
                eexc = etot - ekin - ecou - ehar
            Dn = t * D + (1 - t) * Dp
            rho_y = [t * a + (1 - t) * b_ for a, b_ in zip(rho_new, rho_p)]
            Bn, Wn = t * B + (1 - t) * st["B"], t * W + (1 - t) * st["W"]
            E = dict(Etot=etot, Ekin=ekin, Ecou=ecou, Ehar=ehar, Eexc=eexc)
        else:
            Dn, rho_y, Bn, Wn = D, rho_new, B, W
            E = dict(Etot=Etot, Ekin=Ekin, Ecou=Ecou, Ehar=Ehar, Eexc=Eexc)
        dn = Dn - Dp
        frob = mp.sqrt(sum(dn[i, j] ** 2 for i in range(self.Nh) for j in range(self.Nh)))
        stop = max(frob, abs(E["Etot"] - Ep["Etot"]))
        st.update(D=Dn, rho_y=rho_y, B=Bn, W=Wn, E=E, t=t, niter=st["niter"] + 1, stop=stop, eps=eps, n=n)
        return st
