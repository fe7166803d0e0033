"""Double64 host tables: the density-independent FEM operators in double-double.

The reference's `T = Double64` path (`P1IntLegendreBasis(mesh, Double64; ...)`,
`examples/basic_example_doublefloat/run.jl:55-61`) builds every table of `init_cache!`
(`src/discretization/discretization.jl:320-341`) in DoubleFloats arithmetic.  This module is the host
mirror for `KSEDiscretization(..., dtype="double64")` (the part the Double64
SCF path needs): Gauss nodes, weights and generator values are produced with mpmath at 50 digits (three-term
recurrences, as `femops.py` does in extended precision), the O(C n^3 nq) contractions run in vectorised
double-double NumPy arithmetic (TwoSum / Dekker TwoProd), and every number reaches the C ABI as an
interleaved (hi, lo) pair -- the memory layout of Julia's `Vector{Double64}`.

Mesh points stay Float64-valued, as in the reference (`exprange` evaluates in Float64 and stores into
`zeros(T, n)`, `src/fem/mesh.jl:134-143`).
"""
from __future__ import annotations

from dataclasses import dataclass

import mpmath as mp
import numpy as np

from .basis import FEMBasis

_SPLIT = 134217729.0  # 2^27 + 1


def two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def _split(a):
    t = _SPLIT * a
    hi = t - (t - a)
    return hi, a - hi


def two_prod(a, b):
    p = a * b
    ah, al = _split(a)
    bh, bl = _split(b)
    return p, ((ah * bh - p) + ah * bl + al * bh) + al * bl


class DD:
    """Array of double-double numbers (hi + lo)."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi = np.asarray(hi, dtype=np.float64)
        self.lo = np.zeros_like(self.hi) if lo is None else np.asarray(lo, dtype=np.float64)

    @property
    def shape(self):
        return self.hi.shape

    def __getitem__(self, idx):
        return DD(self.hi[idx], self.lo[idx])

    def __setitem__(self, idx, v):
        self.hi[idx] = v.hi
        self.lo[idx] = v.lo

    def copy(self):
        return DD(self.hi.copy(), self.lo.copy())

    def __neg__(self):
        return DD(-self.hi, -self.lo)

    def __add__(self, o):
        o = o if isinstance(o, DD) else DD(np.asarray(o, dtype=np.float64))
        s, e = two_sum(self.hi, o.hi)
        t, f = two_sum(self.lo, o.lo)
        e = e + t
        s, e = two_sum(s, e)
        e = e + f
        s, e = two_sum(s, e)
        return DD(s, e)

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, o):
        if not isinstance(o, DD):
            o = DD(np.asarray(o, dtype=np.float64))
        p, e = two_prod(self.hi, o.hi)
        e = e + (self.hi * o.lo + self.lo * o.hi)
        s, e = two_sum(p, e)
        return DD(s, e)

    __rmul__ = __mul__

    def __truediv__(self, o):
        """Long division with three quotient terms."""
        if not isinstance(o, DD):
            o = DD(np.asarray(o, dtype=np.float64))
        q1 = self.hi / o.hi
        r = self - o * DD(q1)
        q2 = r.hi / o.hi
        r = r - o * DD(q2)
        q3 = r.hi / o.hi
        s, e = two_sum(q1, q2)
        return DD(s, e) + DD(q3)

    def sum(self, axis=-1):
        """Sequential double-double sum along one axis."""
        hi = np.moveaxis(self.hi, axis, 0)
        lo = np.moveaxis(self.lo, axis, 0)
        acc = DD(np.zeros(hi.shape[1:]), np.zeros(hi.shape[1:]))
        for q in range(hi.shape[0]):
            acc = acc + DD(hi[q], lo[q])
        return acc

    def pairs(self):
        """Interleaved (hi, lo): trailing axis of length 2, C-contiguous."""
        return np.ascontiguousarray(np.stack([self.hi, self.lo], axis=-1))

    def to_mp(self):
        flat_h, flat_l = self.hi.reshape(-1), self.lo.reshape(-1)
        return [mp.mpf(float(h)) + mp.mpf(float(l)) for h, l in zip(flat_h, flat_l)]


def from_mp(values, shape=None) -> DD:
    hi = np.array([float(v) for v in values], dtype=np.float64)
    lo = np.array([float(v - mp.mpf(float(h))) for v, h in zip(values, hi)], dtype=np.float64)
    if shape is not None:
        hi, lo = hi.reshape(shape), lo.reshape(shape)
    return DD(hi, lo)


def pair_to_mp(hi, lo):
    return mp.mpf(float(hi)) + mp.mpf(float(lo))


# ---------------------------------------------------------------------------- mpmath tables
def gauss_legendre_mp(n: int):
    """n-point Gauss-Legendre rule on [-1, 1] at the working precision of mpmath (ascending nodes)."""
    xs, ws = [], []
    for k in range(1, n + 1):
        x = mp.cos(mp.pi * (mp.mpf(k) - mp.mpf(1) / 4) / (mp.mpf(n) + mp.mpf(1) / 2))
        for _ in range(100):
            p0, p1 = mp.mpf(1), x
            for m in range(2, n + 1):
                p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
            dp = n * (x * p1 - p0) / (x * x - 1)
            dx = p1 / dp
            x = x - dx
            if abs(dx) < mp.mpf(10) ** (-(mp.mp.dps - 5)):
                break
        p0, p1 = mp.mpf(1), x
        for m in range(2, n + 1):
            p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
        dp = n * (x * p1 - p0) / (x * x - 1)
        xs.append(x)
        ws.append(2 / ((1 - x * x) * dp * dp))
    order = sorted(range(n), key=lambda i: xs[i])
    return [xs[i] for i in order], [ws[i] for i in order]


def generator_values_mp(p: int, xs):
    """Values and derivatives of the p+1 generators (src/fem/generators.jl:34-57) at the mp points xs."""
    G = [[None] * len(xs) for _ in range(p + 1)]
    dG = [[None] * len(xs) for _ in range(p + 1)]
    for q, x in enumerate(xs):
        P = [mp.mpf(1), x]
        for m in range(2, max(p, 1) + 1):
            P.append(((2 * m - 1) * x * P[m - 1] - (m - 1) * P[m - 2]) / m)
        G[0][q], dG[0][q] = 1 + x, mp.mpf(1)
        G[1][q], dG[1][q] = 1 - x, mp.mpf(-1)
        for k in range(1, p):
            G[k + 1][q] = (P[k + 1] - P[k - 1]) / (2 * k + 1)
            dG[k + 1][q] = P[k]
    return G, dG


@dataclass
class FEMOperatorsDD:
    """Cell-local operator blocks in double-double, generator order (as `FEMOperators`)."""

    basis: FEMBasis
    M0_loc: DD   # (C, n, n)
    A_loc: DD
    Mm1_loc: DD
    Mm2_loc: DD
    F_loc: DD    # (C, n, n, n)

    def assemble_pairs(self, loc: DD) -> np.ndarray:
        """Dense global matrix (Nh, Nh, 2).  Only hat-hat diagonal entries receive two contributions (from
        neighbouring cells); even and odd cells are scattered separately, then added in double-double."""
        b = self.basis
        parts = []
        for parity in (0, 1):
            hi = np.zeros((b.size, b.size))
            lo = np.zeros((b.size, b.size))
            for c in range(parity, b.ncells, 2):
                Ib = np.asarray(b.cells_to_indices[c])
                Ig = np.asarray(b.cells_to_generators[c])
                hi[np.ix_(Ib, Ib)] = loc.hi[c][np.ix_(Ig, Ig)]
                lo[np.ix_(Ib, Ib)] = loc.lo[c][np.ix_(Ig, Ig)]
            parts.append(DD(hi, lo))
        return (parts[0] + parts[1]).pairs()

    def F_coo_pairs(self):
        """Mass tensor as COO (i, j, m, (hi, lo)), 0-based -- the content of the reference's
        `Dict{NTuple{3,Int},T}` (`src/fem/matrices.jl:297-343`).  The only key two cells contribute to is
        (h, h, h) of a hat; it is merged in double-double."""
        b = self.basis
        I, J, M, H, Lo = [], [], [], [], []
        carry = None   # (hat index, hi, lo) of the right-hat triple of the previous cell
        for c in range(b.ncells):
            Ib = np.asarray(b.cells_to_indices[c])
            Ig = np.asarray(b.cells_to_generators[c])
            sub = np.ix_(Ig, Ig, Ig)
            th, tl = self.F_loc.hi[c][sub].copy(), self.F_loc.lo[c][sub].copy()
            ii, jj, mm = np.meshgrid(Ib, Ib, Ib, indexing="ij")
            keep = np.ones(th.shape, dtype=bool)
            if carry is not None:
                j1 = int(np.where(Ig == 1)[0][0])       # left hat of this cell == right hat of the previous one
                merged = DD(np.array(carry[1]), np.array(carry[2])) + DD(np.array(th[j1, j1, j1]), np.array(tl[j1, j1, j1]))
                th[j1, j1, j1], tl[j1, j1, j1] = float(merged.hi), float(merged.lo)
            if 0 in Ig and c + 1 < b.ncells:
                j0 = int(np.where(Ig == 0)[0][0])
                carry = (int(Ib[j0]), float(th[j0, j0, j0]), float(tl[j0, j0, j0]))
                keep[j0, j0, j0] = False                # emitted (merged) by the next cell
            else:
                carry = None
            I.append(ii[keep]); J.append(jj[keep]); M.append(mm[keep]); H.append(th[keep]); Lo.append(tl[keep])
        cat = lambda v, dt: np.ascontiguousarray(np.concatenate(v), dtype=dt)
        vals = np.ascontiguousarray(np.stack([cat(H, np.float64), cat(Lo, np.float64)], axis=-1))
        return cat(I, np.int64), cat(J, np.int64), cat(M, np.int64), vals


def build_femops_dd(basis: FEMBasis, *, nquad: int | None = None, dps: int = 50, ctx=None) -> FEMOperatorsDD:
    """All density-independent blocks in double-double (`init_cache!`,
    `src/discretization/discretization.jl:320-341`): the polynomial parts are integrated exactly by the
    Gauss rule, the 1/r and 1/r^2 weights converge geometrically (3p + 80 nodes reach 1e-34 on
    exponential meshes).  With `ctx` (a `_capi.Context`) the blocks are built on the device
    (`aks_dd_build_operators`, csrc/k_dd_tables.cuh: the same quadratures in double-double kernels, milliseconds
    instead of 4.5 s at order 20); without, by the mpmath / NumPy mirror below (what the CPU tests check it against)."""
    p = basis.ordermax
    n = p + 1
    C = basis.ncells
    nq = nquad or (3 * p + 80)
    if ctx is not None:
        from . import _capi
        present = np.zeros((C, n), dtype=np.int32)
        for c in range(C):
            present[c, np.asarray(basis.cells_to_generators[c])] = 1
        pts = np.ascontiguousarray(basis.mesh.points, dtype=np.float64)
        out = [np.zeros((C, n, n, 2)) for _ in range(4)] + [np.zeros((C, n, n, n, 2))]
        ctx.check(_capi.lib.aks_dd_build_operators(ctx.h, p, C, _capi.dptr(pts), _capi.i32ptr(present), nq,
                                                   *[_capi.dptr(o) for o in out]))
        M0, A, Mm1, Mm2, F = (DD(np.ascontiguousarray(o[..., 0]), np.ascontiguousarray(o[..., 1])) for o in out)
        return FEMOperatorsDD(basis, M0, A, Mm1, Mm2, F)
    with mp.workdps(dps):
        xs, ws = gauss_legendre_mp(nq)
        Gm, dGm = generator_values_mp(p, xs)
        G = from_mp([v for row in Gm for v in row], (n, nq))
        dG = from_mp([v for row in dGm for v in row], (n, nq))
        w = from_mp(ws)
        pts = [mp.mpf(float(v)) for v in basis.mesh.points]
        GG = G[:, None, :] * G[None, :, :]                       # (n, n, nq)
        Khat = (GG * w[None, None, :]).sum(axis=-1)
        Kdhat = ((dG[:, None, :] * dG[None, :, :]) * w[None, None, :]).sum(axis=-1)
        z = lambda *s: DD(np.zeros(s), np.zeros(s))
        M0, A, Mm1, Mm2, F = z(C, n, n), z(C, n, n), z(C, n, n), z(C, n, n), z(C, n, n, n)
        for c in range(C):
            a, b = pts[c], pts[c + 1]
            h2 = (b - a) / 2
            mid = a + h2
            present = np.zeros(n)
            present[np.asarray(basis.cells_to_generators[c])] = 1.0
            m2 = np.outer(present, present)
            m3 = m2[:, :, None] * present[None, None, :]
            h2d, ih2d = from_mp([h2], ()), from_mp([1 / h2], ())
            M0[c] = (Khat * h2d) * m2
            A[c] = (Kdhat * ih2d) * m2
            if a == 0:
                # first cell: phi / r is a polynomial (`fill_local_matrix_withsingularity!`,
                # src/fem/local matrix.jl:36-75); the hat 1+x divides exactly
                Gr_m = [[Gm[g][q] / (h2 * (1 + xs[q])) for q in range(nq)] for g in range(n)]
                Gr_m[0] = [1 / h2] * nq
                Gr = from_mp([v for row in Gr_m for v in row], (n, nq))
                GrG = Gr[:, None, :] * G[None, :, :]
                m1 = (GrG * w[None, None, :]).sum(axis=-1)
                m1 = (m1 + DD(m1.hi.T, m1.lo.T)) * 0.5
                Mm1[c] = (m1 * h2d) * m2
                Mm2[c] = (((Gr[:, None, :] * Gr[None, :, :]) * w[None, None, :]).sum(axis=-1) * h2d) * m2
                GGw = GG * w[None, None, :]
                F[c] = ((Gr[:, None, None, :] * GGw[None, :, :, :]).sum(axis=-1) * h2d) * m3
            else:
                r = [h2 * x + mid for x in xs]
                wr = from_mp([ws[q] / r[q] for q in range(nq)])
                wr2 = from_mp([ws[q] / (r[q] * r[q]) for q in range(nq)])
                Mm1[c] = ((GG * wr[None, None, :]).sum(axis=-1) * h2d) * m2
                Mm2[c] = ((GG * wr2[None, None, :]).sum(axis=-1) * h2d) * m2
                GGw = GG * wr[None, None, :]
                F[c] = ((G[:, None, None, :] * GGw[None, :, :, :]).sum(axis=-1) * h2d) * m3
    return FEMOperatorsDD(basis, M0, A, Mm1, Mm2, F)


# ---------------------------------------------------------------------------- Gauss tables in double-double
def legendre_dd(nmax: int, x: DD):
    """P_0..P_nmax at the points x (Bonnet recurrence in double-double, vectorised over the points)."""
    P = [DD(np.ones(x.shape)), x.copy()]
    for m in range(2, nmax + 1):
        P.append((x * P[m - 1] * float(2 * m - 1) - P[m - 2] * float(m - 1)) / float(m))
    return P[:nmax + 1]


def generator_values_dd(p: int, x: DD) -> DD:
    """Generator values (p+1, len(x)) at reference points x (src/fem/generators.jl:34-57)."""
    P = legendre_dd(max(p, 1), x)
    one = DD(np.ones(x.shape))
    rows = [one + x, one - x]
    for k in range(1, p):
        rows.append((P[k + 1] - P[k - 1]) / float(2 * k + 1))
    return DD(np.stack([r.hi for r in rows]), np.stack([r.lo for r in rows]))


def gauss_legendre_dd(n: int, ctx=None):
    """n-point Gauss-Legendre rule in double-double: Newton on P_n from the extended-precision nodes of
    `quadrature.gauss_legendre` (the reference calls `QuadGK.gauss(Double64, n)`,
    src/fem/integration methods.jl:50).  With `ctx` the rule is computed on the device (`aks_dd_gauss_legendre`)."""
    if ctx is not None:
        from . import _capi
        x, w = np.zeros((n, 2)), np.zeros((n, 2))
        ctx.check(_capi.lib.aks_dd_gauss_legendre(ctx.h, int(n), _capi.dptr(x), _capi.dptr(w)))
        return DD(x[:, 0].copy(), x[:, 1].copy()), DD(w[:, 0].copy(), w[:, 1].copy())
    from .quadrature import gauss_legendre
    x0, _ = gauss_legendre(n)
    hi = np.asarray(x0, dtype=np.float64)
    lo = np.asarray(x0 - hi.astype(np.longdouble), dtype=np.float64)
    x = DD(hi, lo)
    one = DD(np.ones(n))
    for _ in range(3):
        P = legendre_dd(n, x)
        dp = (x * P[n] - P[n - 1]) * float(n) / (x * x - one)
        x = x - P[n] / dp
    # exact antisymmetry of the nodes
    x = (x - DD(x.hi[::-1].copy(), x.lo[::-1].copy())) * 0.5
    P = legendre_dd(n, x)
    dp = (x * P[n] - P[n - 1]) * float(n) / (x * x - one)
    w = DD(2.0 * np.ones(n)) / ((one - x * x) * dp * dp)
    w = (w + DD(w.hi[::-1].copy(), w.lo[::-1].copy())) * 0.5
    return x, w


class GaussLegendreDD:
    """`GaussLegendre(basis, npoints)` with T = Double64 (src/fem/integration methods.jl:30-70): nodes and
    weights on [-1, 1], the global energy grid y, wy, the generator values at the nodes (`Pgenx`) and -- what
    the reference recomputes in every `eval_density!` -- the cell and the generator values of every grid point."""

    def __init__(self, basis: FEMBasis, npoints: int = 1000, ctx=None):
        self.npoints = int(npoints)
        p, C = basis.ordermax, basis.ncells
        self.x, self.w = gauss_legendre_dd(self.npoints, ctx=ctx)   # ctx: on the device
        self.Pgenx = generator_values_dd(p, self.x)                     # (p+1, Q)
        pts = np.asarray(basis.mesh.points, dtype=np.float64)
        half = DD(np.array(pts[-1])) * 0.5                               # Rmax / 2 (the mesh starts at 0)
        a = DD(np.array(pts[0]))
        half = (DD(np.array(pts[-1])) - a) * 0.5
        self.y = self.x * half + (a + half)
        self.wy = self.w * half
        self.ycell = np.clip(np.searchsorted(pts, self.y.hi, side="right") - 1, 0, C - 1).astype(np.int64)
        # affine maps in double-double (src/fem/basis.jl:185-190)
        left, right = DD(pts[:-1].copy()), DD(pts[1:].copy())
        h = right - left
        self.invshifts = DD(np.stack([(h * 0.5).hi, (left + h * 0.5).hi], axis=1),
                            np.stack([(h * 0.5).lo, (left + h * 0.5).lo], axis=1))   # (C, 2)
        c1 = DD(2.0 * np.ones(C)) / h
        c0 = -(left * c1) - DD(np.ones(C))
        xi = self.y * c1[self.ycell] + c0[self.ycell]
        self.Pgeny = generator_values_dd(p, xi)                          # (p+1, G)
