"""Oracle: FEM tables built the way the reference builds them (monomial
coefficients + Horner + exact monomial integration), in float64.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with the same arithmetic structure:
  * Legendre / IntLegendre coefficient tables  -- src/fem/legendre polynomial.jl:34-96
  * PolySet Horner evaluation, products, definite/indefinite integrals, factorisation
                                               -- src/fem/polynomial.jl:96-106,123-164,203-271
  * P1IntLegendreGenerator / P1IntLegendreBasis -- src/fem/generators.jl:25-57,79-138
  * BasisCache products, shifts                 -- src/fem/basis.jl:16-46,124-137,185-190
  * local matrices (exact, 1/x, 1/x^2, singular first cell)
                                               -- src/fem/local matrix.jl:5-123
  * global assembly of M0, A, M-1, M-2, F      -- src/fem/matrices.jl:120-150,208-225,297-343
  * GaussLegendre tables x,w,Pgenx,Qgenx,y,wy  -- src/fem/integration methods.jl:46-70

Third-party pieces restated by value, not by algorithm: Gauss-Legendre nodes
(FastGaussQuadrature; any accurate rule agrees to a few ulp) and the monomial
moments int X^k/(AX+B)^n dX (HypergeometricFunctions.pFq closed forms in
src/fem/integration_formula.jl:17-73; computed here by a converged
extended-precision quadrature).

Pinned by tests/test_oracle_golden.py::test_fem_matrices_golden against
test/reference_data/fem_matrices.jld2 (atol 1e-12, the reference's own tolerance).
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

_LD = np.longdouble


# ----------------------------------------------------------------------------- PolySet
def horner(coeffs: np.ndarray, x):
    """evaluate!(y, ps, x): rows of `coeffs` are polynomials, increasing powers
    (src/fem/polynomial.jl:71-106).  x scalar or vector -> (npoly,) or (npoly, nx)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 0:
        y = coeffs[:, -1].copy()
        for i in range(coeffs.shape[1] - 2, -1, -1):
            y = y * x + coeffs[:, i]
        return y
    y = np.zeros((coeffs.shape[0], x.shape[0])) + coeffs[:, -1][:, None]
    for i in range(coeffs.shape[1] - 2, -1, -1):
        y = y * x[None, :] + coeffs[:, i][:, None]
    return y


def polymul(ps: np.ndarray, qs: np.ndarray) -> np.ndarray:
    """mul(ps, qs): row i*np + j (0-based) = ps[j] * qs[i]  (src/fem/polynomial.jl:256-269)."""
    npo, degp = ps.shape
    nq, degq = qs.shape
    out = np.zeros((npo * nq, degp + degq - 1))
    M = np.zeros((degp, degp + degq - 1))
    for i in range(nq):
        for d in range(degq):
            for r in range(min(degp, degp + degq - 1 - d)):
                M[r, r + d] = qs[i, d]
        out[i * npo:(i + 1) * npo, :] = ps @ M
    return out


def integrate_sym(coeffs: np.ndarray) -> np.ndarray:
    """integrate!(y, ps, -1, 1), the a == -b branch (src/fem/polynomial.jl:138-149)."""
    m = coeffs.shape[1]
    s = (m + 1) // 2
    vec = np.array([2.0 / (2 * j - 1) for j in range(1, s + 1)])
    return coeffs[:, 0:m:2] @ vec


def legendre_coeffs(n: int, rational: bool = False):
    """Legendre(n, T): Bonnet recurrence on coefficient rows
    (src/fem/legendre polynomial.jl:34-58)."""
    if rational:
        P = [[Fraction(0)] * (n + 1) for _ in range(n + 1)]
        base = [[Fraction(1), 0, 0], [0, Fraction(1), 0], [Fraction(-1, 2), 0, Fraction(3, 2)]]
        for i in range(min(n, 2) + 1):
            for j in range(min(n, 2) + 1):
                P[i][j] = Fraction(base[i][j])
        for m in range(3, n + 1):
            for j in range(m):
                P[m][j + 1] = Fraction(2 * m - 1, m) * P[m - 1][j]
            for j in range(m - 1):
                P[m][j] -= Fraction(m - 1, m) * P[m - 2][j]
        return P
    P = np.zeros((n + 1, n + 1))
    base = np.array([[1.0, 0, 0], [0, 1.0, 0], [-0.5, 0, 1.5]])
    k = min(n, 2) + 1
    P[:k, :k] = base[:k, :k]
    for m in range(3, n + 1):
        P[m, 1:m + 1] = ((2 * m - 1) / m) * P[m - 1, 0:m]
        P[m, 0:m - 1] -= ((m - 1) / m) * P[m - 2, 0:m - 1]
    return P


def intlegendre_coeffs(n: int) -> np.ndarray:
    """IntLegendre(n, Float64): antiderivatives vanishing at -1; rows of degree >= 2
    from the *rational* Legendre table times 1 ./ (1:deg+1) (Float64), constant
    fixed by a Horner evaluation at -1 (src/fem/legendre polynomial.jl:78-96,
    src/fem/polynomial.jl:203-213)."""
    out = np.zeros((n + 1, n + 2))
    stored = np.array([[1.0, 1.0, 0.0], [-0.5, 0.0, 0.5]])
    if n <= 1:
        return stored[:n + 1, :n + 2].copy()
    out[:2, :3] = stored
    leg = legendre_coeffs(n, rational=True)
    xinv = 1.0 / np.arange(1, n + 2)
    for r in range(2, n + 1):
        for j in range(n + 1):
            out[r, j + 1] = float(leg[r][j]) * xinv[j]
    y = horner(out[2:], -1.0)
    out[2:, 0] -= y
    return out


def generator_coeffs(p: int):
    """P1IntLegendreGenerator(Float64; ordermax=p): (polynomials (p+1,p+1),
    derivpolynomials (p+1,p))  (src/fem/generators.jl:34-57)."""
    pol = np.zeros((p + 1, p + 1))
    der = np.zeros((p + 1, p))
    pol[0, 0] = 1.0
    pol[0, 1] = 1.0
    pol[1, 0] = 1.0
    pol[1, 1] = -1.0
    der[0, 0] = 1.0
    der[1, 0] = -1.0
    if p >= 2:
        leg = legendre_coeffs(p - 1)
        il = intlegendre_coeffs(p - 1)
        pol[2:, :] = il[1:, :]
        der[2:, :] = leg[1:, :]
    return pol, der


def factorize(ps: np.ndarray, root: float) -> np.ndarray:
    """factorize(ps, racine): q with (x - root) q = p for rows that vanish exactly
    at root, zero rows otherwise (src/fem/polynomial.jl:225-237)."""
    y = horner(ps, root)
    qs = np.zeros((ps.shape[0], ps.shape[1] - 1))
    for r in np.where(y == 0)[0]:
        q = np.zeros(ps.shape[1])
        q[0] = ps[r, 0] / (-root)
        for c in range(1, ps.shape[1]):
            q[c] = (ps[r, c] - q[c - 1]) / (-root)
        qs[r] = q[:-1]
    return qs


# ----------------------------------------------------------------------------- moments
def _gl_ld(n: int):
    k = np.arange(1, n + 1, dtype=_LD)
    x = np.cos(_LD(np.pi) * (k - _LD(0.25)) / (_LD(n) + _LD(0.5)))
    for _ in range(10):
        p0 = np.ones_like(x)
        p1 = x.copy()
        for m in range(2, n + 1):
            p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / _LD(m)
        dp = n * (x * p1 - p0) / (x * x - 1)
        x = x - p1 / dp
    p0 = np.ones_like(x)
    p1 = x.copy()
    for m in range(2, n + 1):
        p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / _LD(m)
    dp = n * (x * p1 - p0) / (x * x - 1)
    return x, 2 / ((1 - x * x) * dp * dp)


_XM, _WM = None, None


def monomial_moments(kmax: int, A: float, B: float, power: int) -> np.ndarray:
    """[ int_{-1}^{1} X^k / (A X + B)^power dX  for k = 0..kmax ]
    (values of _integration_monome_over_deg1/2, src/fem/integration_formula.jl:17-73)."""
    global _XM, _WM
    if _XM is None:
        _XM, _WM = _gl_ld(256)
    den = (_LD(A) * _XM + _LD(B)) ** power
    out = np.zeros(kmax + 1)
    xp = np.ones_like(_XM)
    for k in range(kmax + 1):
        out[k] = float(np.sum(_WM * xp / den))
        xp = xp * _XM
    return out


def gauss_nodes(n: int):
    """Q-point Gauss-Legendre rule on [-1,1] in float64, ascending
    (stands in for FastGaussQuadrature.gausslegendre)."""
    x, w = _gl_ld(n)
    o = np.argsort(x)
    return np.asarray(x[o], dtype=np.float64), np.asarray(w[o], dtype=np.float64)


# ----------------------------------------------------------------------------- basis maps
def basis_maps(npts: int, p: int):
    """cells_to_indices / cells_to_generators, 0-based (src/fem/generators.jl:79-138)."""
    C = npts - 1
    size = (p - 1) * C + (C - 1)
    c2i = [list(range(0, p))]
    c2g = [[0] + list(range(2, p + 1))]
    numbas = p  # 0-based index of the next basis function
    for _ in range(1, C - 1):
        numbas += p
        c2i.append([numbas - 2 * p] + list(range(numbas - p, numbas)))
        c2g.append([1, 0] + list(range(2, p + 1)))
    numbas += p - 1
    c2i.append([numbas - 2 * p + 1] + list(range(numbas - p + 1, numbas)))
    c2g.append([1] + list(range(2, p + 1)))
    assert numbas == size
    return size, c2i, c2g


def shift(a, b, m0, m1):
    c1 = (m1 - m0) / (b - a)
    c0 = -a * c1 + m0
    return c1, c0


class ReferenceTables:
    """Everything `init_cache!` + `GaussLegendre(basis, npoints)` produce, built
    with the reference's monomial arithmetic."""

    def __init__(self, mesh_points, p: int, npoints: int = 1000, *, need_Mm2=True,
                 need_F=True, tensors=("F",)):
        pts = np.asarray(mesh_points, dtype=np.float64)
        self.mesh = pts
        self.p = p
        n = p + 1
        C = len(pts) - 1
        self.size, self.c2i, self.c2g = basis_maps(len(pts), p)
        self.shifts = [shift(pts[c], pts[c + 1], -1.0, 1.0) for c in range(C)]
        self.invshifts = [shift(-1.0, 1.0, pts[c], pts[c + 1]) for c in range(C)]
        pol, der = generator_coeffs(p)
        self.pol, self.der = pol, der
        prodMG = polymul(pol, pol)
        prodMdG = polymul(der, der)
        self.prodMG = prodMG
        Nh = self.size
        # --- M0 / A : one local matrix rescaled per cell (matrices.jl:128-137, 215-223)
        inv1 = self.invshifts[0][0]
        K0 = (integrate_sym(prodMG) * inv1).reshape(n, n).T  # column-major reshape
        Kd0 = (integrate_sym(prodMdG) * inv1).reshape(n, n).T
        self.M0 = np.zeros((Nh, Nh))
        self.A = np.zeros((Nh, Nh))
        for c in range(C):
            Ib, Ig = np.array(self.c2i[c]), np.array(self.c2g[c])
            self.M0[np.ix_(Ib, Ib)] += K0[np.ix_(Ig, Ig)] * self.invshifts[c][0] / inv1
            self.A[np.ix_(Ib, Ib)] += (Kd0[np.ix_(Ig, Ig)] * self.shifts[c][0] ** 2
                                       * self.invshifts[c][0] / inv1)
        # --- weighted mass matrices (matrices.jl:138-148, local matrix.jl:36-123)
        self.Mm1 = self._weighted_mass(1)
        self.Mm2 = self._weighted_mass(2) if need_Mm2 else np.zeros((Nh, Nh))
        # --- mass tensor with 1/x weight, kept per cell in Ib order (matrices.jl:297-343)
        self.F_cells = self._weighted_tensor(1) if need_F else None
        # --- Gauss tables (integration methods.jl:46-70)
        self.npoints = npoints
        self.x, self.w = gauss_nodes(npoints)
        Rmax = pts[-1]
        self.y = Rmax / 2 * self.x + Rmax / 2
        self.wy = Rmax / 2 * self.w
        self.Pgenx = horner(pol, self.x)
        self.Qgenx = horner(prodMG, self.x)
        cell = np.clip(np.searchsorted(pts, self.y, side="right") - 1, 0, C - 1)
        self.ycell = cell
        xi = np.array([self.shifts[c][0] * yy + self.shifts[c][1] for c, yy in zip(cell, self.y)])
        self.Pgeny = horner(pol, xi)

    # local (p+1)^2 matrix of cell c for weight 1/r^power, generator order
    def _local_weighted(self, c: int, power: int, ps_products=None):
        n = self.p + 1
        a = self.mesh[c]
        phi = self.shifts[c]
        inv = self.invshifts[c]
        if a <= 0 <= self.mesh[c + 1]:
            gf = factorize(self.pol, phi[1])
            newps = polymul(gf, self.pol) if power == 1 else polymul(gf, gf)
            K = integrate_sym(newps) * inv[0] * phi[0] ** power
        else:
            mom = monomial_moments(self.prodMG.shape[1] - 1, inv[0], inv[1], power)
            K = (self.prodMG @ mom) * inv[0]
        return K.reshape(n, n).T

    def _weighted_mass(self, power: int):
        Nh = self.size
        out = np.zeros((Nh, Nh))
        for c in range(len(self.mesh) - 1):
            K = self._local_weighted(c, power)
            Ib, Ig = np.array(self.c2i[c]), np.array(self.c2g[c])
            out[np.ix_(Ib, Ib)] += K[np.ix_(Ig, Ig)]
        return out

    def _weighted_tensor(self, power: int):
        n = self.p + 1
        prodTG = polymul(self.prodMG, self.pol)
        cells = []
        for c in range(len(self.mesh) - 1):
            phi = self.shifts[c]
            inv = self.invshifts[c]
            if self.mesh[c] <= 0 <= self.mesh[c + 1]:
                gf = factorize(self.pol, phi[1])
                if power == 1:
                    newps = polymul(gf, self.prodMG)
                else:
                    newps = polymul(polymul(gf, gf), self.pol)
                T = integrate_sym(newps) * inv[0] * phi[0] ** power
            else:
                mom = monomial_moments(prodTG.shape[1] - 1, inv[0], inv[1], power)
                T = (prodTG @ mom) * inv[0]
            T = T.reshape(n, n, n).transpose(2, 1, 0)  # column-major reshape
            Ig = np.array(self.c2g[c])
            cells.append(np.ascontiguousarray(T[np.ix_(Ig, Ig, Ig)]))
        return cells

    def plain_tensor_dense(self):
        """mass_tensor(basis) with NoWeight, dense (tests only; tiny bases)."""
        n = self.p + 1
        prodTG = polymul(self.prodMG, self.pol)
        inv1 = self.invshifts[0][0]
        T0 = (integrate_sym(prodTG) * inv1).reshape(n, n, n).transpose(2, 1, 0)
        Nh = self.size
        out = np.zeros((Nh, Nh, Nh))
        for c in range(len(self.mesh) - 1):
            Ib, Ig = np.array(self.c2i[c]), np.array(self.c2g[c])
            out[np.ix_(Ib, Ib, Ib)] += T0[np.ix_(Ig, Ig, Ig)] * self.invshifts[c][0] / inv1
        return out

    def weighted_tensor_dense(self, power: int):
        cells = self._weighted_tensor(power)
        Nh = self.size
        out = np.zeros((Nh, Nh, Nh))
        for c, T in enumerate(cells):
            Ib = np.array(self.c2i[c])
            out[np.ix_(Ib, Ib, Ib)] += T
        return out
