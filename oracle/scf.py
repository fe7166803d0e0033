"""Oracle: the SCF hot path (ODA + OptimizedAufbau) restated on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

One `Oracle` object = one (model, discretization, algorithm) triple of the
reference.  Every method cites the reference function it follows:

    assemble_hartree            src/discretization/assemble.jl:63-83
    assemble_vxc                src/discretization/assemble.jl:91-125,277-319
                                src/fem/local matrix.jl:129-155, src/fem/matrices.jl:138-148
    find_orbital                src/algorithms/oda/loop.jl:65-77
    aufbau                      src/algorithms/aufbau/optimized.jl:99-306
    density                     src/discretization/density.jl:15-41
    eval_density / exc_energy   src/discretization/density.jl:97-126, energies.jl:217-233
    hartree_energy / _mix       src/discretization/energies.jl:96-141
    kinetic / coulomb energy    src/discretization/energies.jl:47-87
    total_energy                src/discretization/energies.jl:7-39
    line_search_energy          src/algorithms/optimization/line_search.jl:45-80
    scf_step / groundstate      src/algorithms/oda/loop.jl:1-116, src/solver/groundstate.jl:53-62

Inputs are the density-independent tables (duck-typed `tables`, produced either
by oracle.fem_reference.ReferenceTables -- the reference's own arithmetic -- or
by the package's stable builder through tests/helpers.py) so that "identical
tables" parity can be checked.

Array conventions follow Julia: eps[l, k, sigma], n[l, k, sigma], U[i, k, l, sigma],
D[i, j, sigma]; the spin axis is always present here (size 1 when nspin == 1).
Dense linear algebra is SciPy/LAPACK; `eig="sqrtinv"` reproduces the reference's
S = sqrt(inv(M0)) route literally, `eig="generalized"` solves the same pencil by
Cholesky reduction (default; see DESIGN.md "parity protocol").
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

from .brent import brent_minimize
from .xc import XC

FOURPI = 4.0 * np.pi
INV4PI = 1.0 / (4.0 * np.pi)   # Julia `1/4π` == 1/(4π) (assemble.jl:123)


@dataclass
class ODAOptions:
    """ODA(; scftol, tinit, ...) + OptimizedAufbau(; ...)  (oda/types.jl:29-45, optimized.jl:18-27)."""
    scftol: float = 1e-10
    tinit: float = 0.6
    frozen_t: bool = False
    maxiter_ls: int = 100
    abstol_ls: float = 1e4 * np.finfo(np.float64).eps
    reltol_ls: float = 1e4 * np.finfo(np.float64).eps
    max_degen: int = 2
    degen_tol: float = 1e-3
    handle_degen_spin_1iter: bool = True
    # occupation scheme: "optimized" (aufbau/optimized.jl), "smeared" (aufbau/smeared.jl, needs temp),
    # "frozen" (aufbau/frozen.jl, needs nfix in the layout of n)
    aufbau_kind: str = "optimized"
    temp: float = 0.0
    nfix: object = None


@dataclass
class Energies:
    Etot: float = 0.0
    Ekin: float = 0.0
    Ecou: float = 0.0
    Ehar: float = 0.0
    Eexc: float = 0.0

    def copy(self):
        return Energies(self.Etot, self.Ekin, self.Ecou, self.Ehar, self.Eexc)


class DegeneracyError(RuntimeError):
    """More than two-fold degeneracy in the last shell (optimized.jl:139)."""


@dataclass
class State:
    D: np.ndarray
    U: np.ndarray
    eps: np.ndarray
    n: np.ndarray
    energies: Energies = field(default_factory=Energies)
    t: float = 0.6
    niter: int = 0
    stop: float = 0.0


class Oracle:
    def __init__(self, tables, *, Z, N, hartree=1.0, ex=None, ec=None, nspin=1,
                 lh=0, nh=10, options: ODAOptions | None = None, eig="generalized"):
        self.tb = tables
        self.Z, self.N, self.hartree = float(Z), float(N), float(hartree)
        self.xc = XC(ex, ec, nspin)
        self.nspin, self.lh, self.nh = nspin, lh, nh
        self.L = lh + 1
        self.opt = options or ODAOptions()
        self.eig = eig
        tb = tables
        self.Nh = tb.size
        self.C = len(tb.mesh) - 1
        self.Rmax = float(tb.mesh[-1])
        self.n_loc = tb.p + 1
        self.Ib = [np.asarray(v) for v in tb.c2i]
        self.Ig = [np.asarray(v) for v in tb.c2g]
        # fixed operators (assemble.jl:7-27, discretization.jl:335-339)
        self.Kin = [(tb.A + l * (l + 1) * tb.Mm2) / 2 for l in range(self.L)]
        self.Coulomb = -self.Z * tb.Mm1
        self.Hfix = [self.Kin[l] + self.Coulomb for l in range(self.L)]
        self._Acho = sla.cho_factor(tb.A) if self.hartree != 0 else None
        if eig == "sqrtinv":
            lam, V = sla.eigh(tb.M0)
            self.S = (V / np.sqrt(lam)) @ V.T     # sqrt(inv(Symmetric(M0))), discretization.jl:331
        # cell of each global grid point, grouped
        self._ysel = [np.where(tb.ycell == c)[0] for c in range(self.C)]

    # ------------------------------------------------------------------ Hartree
    def tensor_matrix(self, Dsum):
        """B_m = sum_ij D_ij F_ijm  (free allocation.jl:17-36)."""
        B = np.zeros(self.Nh)
        for c in range(self.C):
            Ib = self.Ib[c]
            B[Ib] += np.einsum("ijk,ij->k", self.tb.F_cells[c], Dsum[np.ix_(Ib, Ib)])
        return B

    def tensor_vector(self, W):
        """FW_ij = sum_m W_m F_ijm  (free allocation.jl:38-45)."""
        FW = np.zeros((self.Nh, self.Nh))
        for c in range(self.C):
            Ib = self.Ib[c]
            FW[np.ix_(Ib, Ib)] += np.einsum("ijk,k->ij", self.tb.F_cells[c], W[Ib])
        return FW

    def poisson(self, B):
        return sla.cho_solve(self._Acho, B)

    def assemble_hartree(self, D):
        if self.hartree == 0:
            return np.zeros((self.Nh, self.Nh)), np.zeros(self.Nh)
        B = self.tensor_matrix(D.sum(axis=2))
        W = self.poisson(B)
        Har = self.tensor_vector(W) + self.N / self.Rmax * self.tb.M0
        Har = Har * self.hartree
        return (Har + Har.T) / 2, W

    # ------------------------------------------------------------------ XC potential
    def _dloc_gen(self, Dm, c):
        n = self.n_loc
        Dg = np.zeros((n, n))
        Dg[np.ix_(self.Ig[c], self.Ig[c])] = Dm[np.ix_(self.Ib[c], self.Ib[c])]
        return Dg

    def cell_density(self, Dm, c):
        """optimized_eval_density! on the Gauss nodes of cell c (assemble.jl:91-125)."""
        tb = self.tb
        inv1, inv2 = tb.invshifts[c]
        X = inv1 * tb.x + inv2
        rho = self._dloc_gen(Dm, c).reshape(-1) @ tb.Qgenx
        rho = rho / X ** 2
        rho = rho * INV4PI
        return rho, X

    def assemble_vxc(self, D):
        """VxcUP / VxcDOWN (assemble.jl:277-319)."""
        tb = self.tb
        n = self.n_loc
        V = np.zeros((self.Nh, self.Nh, self.nspin))
        if not self.xc.active:
            return V
        for c in range(self.C):
            rhos = [self.cell_density(D[:, :, s], c)[0] for s in range(self.nspin)]
            v = self.xc.vrho(*rhos)
            v = (v,) if self.nspin == 1 else v
            inv1 = tb.invshifts[c][0]
            for s in range(self.nspin):
                fx = v[s] * tb.w
                K = (tb.Qgenx @ fx) * inv1
                K = K.reshape(n, n).T
                V[np.ix_(self.Ib[c], self.Ib[c], [s])] += K[np.ix_(self.Ig[c], self.Ig[c])][:, :, None]
        for s in range(self.nspin):
            V[:, :, s] = (V[:, :, s] + V[:, :, s].T) / 2
        return V

    # ------------------------------------------------------------------ eigenproblem
    def find_orbital(self, Har, Vxc):
        U = np.zeros((self.Nh, self.nh, self.L, self.nspin))
        eps = np.zeros((self.L, self.nh, self.nspin))
        for s in range(self.nspin):
            for l in range(self.L):
                H = self.Hfix[l] + Vxc[:, :, s] + Har
                if self.eig == "sqrtinv":
                    SHS = self.S @ H @ self.S
                    lam, V = sla.eigh((SHS + SHS.T) / 2, subset_by_index=[0, self.nh - 1])
                    V = self.S @ V
                else:
                    lam, V = sla.eigh(H, self.tb.M0, subset_by_index=[0, self.nh - 1])
                eps[l, :, s] = lam
                U[:, :, l, s] = V
        return U, eps

    # ------------------------------------------------------------------ densities / energies
    def density(self, U, n):
        D = np.zeros((self.Nh, self.Nh, self.nspin))
        for s in range(self.nspin):
            for k in range(self.nh):
                for l in range(self.L):
                    if n[l, k, s] != 0:
                        u = U[:, k, l, s]
                        D[:, :, s] += n[l, k, s] * np.outer(u, u)
        return D

    def eval_density_grid(self, Dm):
        """eval_density!(rho, disc, D, y) on the global Gauss grid (density.jl:97-126)."""
        tb = self.tb
        rho = np.zeros(tb.y.shape[0])
        for c in range(self.C):
            sel = self._ysel[c]
            if sel.size == 0:
                continue
            g = tb.Pgeny[np.ix_(self.Ig[c], sel)]
            Dl = Dm[np.ix_(self.Ib[c], self.Ib[c])]
            rho[sel] = np.einsum("aq,ab,bq->q", g, Dl, g) / (FOURPI * tb.y[sel] ** 2)
        return rho

    def exc_energy(self, D):
        if not self.xc.active:
            return 0.0
        tb = self.tb
        rhos = [self.eval_density_grid(D[:, :, s]) for s in range(self.nspin)]
        zk = self.xc.zk(*rhos)
        fx = sum(rhos) * zk * tb.y ** 2
        return FOURPI * float(np.dot(tb.wy, fx))

    def hartree_energy(self, D):
        if self.hartree == 0:
            return 0.0
        B = self.tensor_matrix(D.sum(axis=2))
        W = self.poisson(B)
        return (float(B @ W) + self.N ** 2 / self.Rmax) / 2

    def hartree_mix_energy(self, D0, D1):
        if self.hartree == 0:
            return 0.0
        W = self.poisson(self.tensor_matrix(D0.sum(axis=2)))
        B = self.tensor_matrix(D1.sum(axis=2))
        return (float(B @ W) + self.N ** 2 / self.Rmax) / 2

    def _orbital_sum(self, U, n, mats):
        e = 0.0
        for s in range(self.nspin):
            for l in range(self.L):
                for k in range(self.nh):
                    if n[l, k, s] != 0:
                        u = U[:, k, l, s]
                        e += n[l, k, s] * float(u @ (mats[l] @ u))
        return e

    def kinetic_energy(self, U, n):
        return self._orbital_sum(U, n, self.Kin)

    def coulomb_energy(self, U, n):
        return self._orbital_sum(U, n, [self.Coulomb] * self.L)

    def total_energy(self, D, n, eps, Vxc):
        e = float(np.sum(n * eps))
        ehar = self.hartree_energy(D)
        if self.xc.active:
            exc = self.exc_energy(D)
            corr = sum(float(np.sum(Vxc[:, :, s] * D[:, :, s])) for s in range(self.nspin))
            return e - ehar + exc - corr
        return e - ehar

    # ------------------------------------------------------------------ line search
    def line_search_energy(self, k0, k1, c0, c1, h0, h1, h01, F, *, maxiter=100,
                           abstol=None, reltol=None):
        eps64 = np.finfo(np.float64).eps
        abstol = 1e4 * eps64 if abstol is None else abstol
        reltol = 1e4 * eps64 if reltol is None else reltol
        A0, A1 = k0 + c0, k1 + c1
        a = h0 + h1 - 2 * h01
        b = (A0 - A1) + 2 * (h01 - h1)
        c = A1 + h1
        if F is not None:
            t, fmin, _ = brent_minimize(lambda t: a * t ** 2 + b * t + c + F(t), 0.0, 1.0,
                                        abs_tol=abstol, rel_tol=reltol, iterations=maxiter)
            return t, fmin
        if a > 0:
            t = min(max(-b / (2 * a), 0.0), 1.0)
        else:
            t = 0.0 if c <= c + b + a else 1.0
        return t, c + b * t + a * t ** 2

    # ------------------------------------------------------------------ aufbau
    def flat_index(self, idx):
        """convert_index (discretization.jl:379-392), 0-based flat -> (l, k, sigma)."""
        s, r = divmod(idx, self.L * self.nh)
        k, l = divmod(r, self.L)
        return l, k, s

    def degeneracy(self, idx):
        l, _, _ = self.flat_index(idx)
        return 4 * l + 2 if self.nspin == 1 else 2 * l + 1

    def aufbau(self, eps, U, niter):
        """aufbau!(...::OptimizedAufbauCache, niter).  Returns (n, post) where post is
        None or (D1, Energies) when the two-fold branch ran (optimized.jl:228-306)."""
        opt = self.opt
        if opt.aufbau_kind == "frozen":      # aufbau!(…, ::FrozenAufbauCache, …), frozen.jl:84-88
            return np.array(opt.nfix, dtype=np.float64).reshape(eps.shape), None
        if opt.aufbau_kind == "smeared":     # aufbau!(…, ::SmearedAufbauCache, …), smeared.jl:68-101
            ef = eps.reshape(-1, order="F")
            deg = np.array([self.degeneracy(i) for i in range(ef.size)], dtype=np.float64)
            T = opt.temp
            lo, hi = ef.min() - 50 * T, ef.max() + 50 * T
            for _ in range(100):
                mid = (lo + hi) / 2
                occ = 0.0
                for i in range(ef.size):
                    occ += deg[i] / (1 + np.exp((ef[i] - mid) / T))
                if occ < self.N:
                    lo = mid
                else:
                    hi = mid
            mu = (lo + hi) / 2
            with np.errstate(over="ignore"):
                nf = deg / (1 + np.exp((ef - mu) / T))
            return nf.reshape(eps.shape, order="F"), None
        ef = eps.reshape(-1, order="F")
        order = np.argsort(ef, kind="stable")
        nf = np.zeros(ef.size)
        left = self.N
        i = 0
        shape = eps.shape
        R = lambda v: v.reshape(shape, order="F")
        while left > 0 and i < len(order):
            blk = [order[i]]
            j = i + 1
            while j < len(order) and abs(ef[order[j]] - ef[blk[0]]) < opt.degen_tol \
                    and len(blk) < opt.max_degen:
                blk.append(order[j])
                j += 1
            i += len(blk)
            degs = [self.degeneracy(o) for o in blk]
            tot = sum(degs)
            if left >= tot:
                for o, d in zip(blk, degs):
                    nf[o] = d
                left -= tot
            elif len(blk) == 1:
                nf[blk[0]] = left
                break
            elif len(blk) == 2:
                l1, k1, s1 = self.flat_index(blk[0])
                l2, k2, s2 = self.flat_index(blk[1])
                if opt.handle_degen_spin_1iter and niter == 0 and l1 == l2 and k1 == k2 and s1 != s2:
                    nf[blk[0]] = left / 2
                    nf[blk[1]] = left / 2
                    break
                n10 = left if degs[0] >= left else degs[0]
                n20 = left - n10
                nf[blk[0]], nf[blk[1]] = n10, n20
                D1 = self.density(U, R(nf))
                kin0 = self.kinetic_energy(U, R(nf))
                cou0 = self.coulomb_energy(U, R(nf))
                har0 = self.hartree_energy(D1)
                n21 = left if degs[1] >= left else degs[1]
                n11 = left - n21
                nf[blk[0]], nf[blk[1]] = n11, n21
                D2 = self.density(U, R(nf))
                kin1 = self.kinetic_energy(U, R(nf))
                cou1 = self.coulomb_energy(U, R(nf))
                har1 = self.hartree_energy(D2)
                har01 = self.hartree_mix_energy(D1, D2)
                F = (lambda t: self.exc_energy(t * D1 + (1 - t) * D2)) if self.xc.active else None
                td, emin = self.line_search_energy(kin0, kin1, cou0, cou1, har0, har1, har01, F)
                nf[blk[0]] = td * n10 + (1 - td) * n11
                nf[blk[1]] = td * n20 + (1 - td) * n21
                D1 = td * D1 + (1 - td) * D2
                e = Energies()
                e.Etot = emin
                e.Ekin = td * kin0 + (1 - td) * kin1
                e.Ecou = td * cou0 + (1 - td) * cou1
                e.Ehar = td ** 2 * har0 + (1 - td) ** 2 * har1 + 2 * td * (1 - td) * har01
                e.Eexc = e.Etot - e.Ekin - e.Ecou - e.Ehar
                return R(nf).copy(), (D1, e)
            else:
                raise DegeneracyError(f"degeneracy {len(blk)} > 2 not implemented")
        return R(nf).copy(), None

    # ------------------------------------------------------------------ SCF
    def initial_state(self):
        return State(D=np.zeros((self.Nh, self.Nh, self.nspin)),
                     U=np.zeros((self.Nh, self.nh, self.L, self.nspin)),
                     eps=np.zeros((self.L, self.nh, self.nspin)),
                     n=np.zeros((self.L, self.nh, self.nspin)), t=self.opt.tinit)

    def hamiltonian_parts(self, D):
        Har, W = self.assemble_hartree(D)
        Vxc = self.assemble_vxc(D)
        return Har, Vxc, W

    def scf_step(self, st: State, *, detail: dict | None = None):
        """scf_step!(alg::ODA, solver) (oda/loop.jl:1-46); mutates and returns st."""
        opt = self.opt
        Dprev = st.D.copy()
        eprev = st.energies.copy()
        Har, Vxc, _ = self.hamiltonian_parts(Dprev)
        U, eps = self.find_orbital(Har, Vxc)
        n, post = self.aufbau(eps, U, st.niter)
        e = Energies()
        if post is not None:
            D, e = post
        else:
            D = self.density(U, n)
            e.Etot = self.total_energy(D, n, eps, Vxc)
            e.Ekin = self.kinetic_energy(U, n)
            e.Ecou = self.coulomb_energy(U, n)
            e.Ehar = self.hartree_energy(D)
            e.Eexc = self.exc_energy(D)
        if detail is not None:
            detail.update(Har=Har, Vxc=Vxc, U=U, eps=eps, n=n, Dtrial=D.copy(), etrial=e.copy())
        t = st.t
        if st.niter > 0:
            har01 = self.hartree_mix_energy(D, Dprev)
            Dn = D
            F = (lambda tt: self.exc_energy(tt * Dn + (1 - tt) * Dprev)) if self.xc.active else None
            if opt.frozen_t:
                t = st.t
                etot = None
            else:
                t, etot = self.line_search_energy(e.Ekin, eprev.Ekin, e.Ecou, eprev.Ecou,
                                                  e.Ehar, eprev.Ehar, har01, F,
                                                  maxiter=opt.maxiter_ls, abstol=opt.abstol_ls,
                                                  reltol=opt.reltol_ls)
            D = t * D + (1 - t) * Dprev
            ekin = t * e.Ekin + (1 - t) * eprev.Ekin
            ecou = t * e.Ecou + (1 - t) * eprev.Ecou
            ehar = t ** 2 * e.Ehar + (1 - t) ** 2 * eprev.Ehar + 2 * t * (1 - t) * har01
            if opt.frozen_t:
                eexc = F(t) if F is not None else 0.0
                etot = ekin + ecou + ehar + eexc
            else:
                eexc = etot - ekin - ecou - ehar
            e = Energies(etot, ekin, ecou, ehar, eexc)
        stop = max(float(np.linalg.norm(D - Dprev)), abs(e.Etot - eprev.Etot))
        st.D, st.U, st.eps, st.n, st.energies, st.t, st.stop = D, U, eps, n, e, t, stop
        st.niter += 1
        return st

    def groundstate(self, maxiter=100, log=None):
        """solve!(solver) (groundstate.jl:53-62).  Returns (state, history)."""
        st = self.initial_state()
        hist = []
        while st.niter < maxiter:
            self.scf_step(st)
            e = st.energies
            rec = dict(iter=st.niter - 1, stop=st.stop, Etot=e.Etot, Ekin=e.Ekin, Ecou=e.Ecou,
                       Ehar=e.Ehar, Eexc=e.Eexc, t=st.t)
            hist.append(rec)
            if log:
                log(rec)
            if st.stop < self.opt.scftol:
                break
        return st, hist
