"""Gauss-Legendre tables consumed every SCF iteration (host-built, device-resident).

API mirror of `GaussLegendre(basis, npoints=1000)`
(`src/fem/integration methods.jl:30-70`): nodes/weights `x, w` on [-1,1], the
global energy grid `y = Rmax/2 x + Rmax/2`, `wy = Rmax/2 w`, and the generator
values at the nodes `Pgenx[g, q]`.  Two things differ from the reference by
design (DESIGN.md "tables"):

* nodes come from Newton iterations on the Legendre recurrence in extended
  precision (the reference calls FastGaussQuadrature / QuadGK) and
* `Pgenx` is evaluated with the stable recurrence; the pair-product table
  `Qgenx` is *not* materialised for the device -- the kernels use the factored
  form B^T diag(w v) B (north star).  `qgenx()` builds it on request for the
  oracle, as exactly rounded products of `Pgenx` rows.

In addition the table holds what the reference recomputes on every energy
evaluation (`eval_density!`, `src/discretization/density.jl:97-126`): the cell
of every global grid point and the generator values there (`ycell`, `Pgeny`).
"""
from __future__ import annotations

import numpy as np

from .basis import FEMBasis, generator_values

_LD = np.longdouble


def gauss_legendre(n: int):
    """n-point Gauss-Legendre rule on [-1,1], ascending nodes; computed in
    extended precision and rounded to float64."""
    k = np.arange(1, n + 1, dtype=_LD)
    # Tricomi initial guess, then Newton on P_n
    theta = _LD(np.pi) * (k - _LD(0.25)) / (_LD(n) + _LD(0.5))
    x = np.cos(theta) * (1 - (_LD(n) - 1) / (8 * _LD(n) ** 3))
    for _ in range(8):
        p0 = np.ones_like(x)
        p1 = x.copy()
        for m in range(2, n + 1):
            p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / _LD(m)
        dp = n * (x * p1 - p0) / (x * x - 1)
        dx = p1 / dp
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    p0 = np.ones_like(x)
    p1 = x.copy()
    for m in range(2, n + 1):
        p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / _LD(m)
    dp = n * (x * p1 - p0) / (x * x - 1)
    w = 2 / ((1 - x * x) * dp * dp)
    order = np.argsort(x)
    x, w = x[order], w[order]
    # enforce exact antisymmetry of the nodes
    x = (x - x[::-1]) / 2
    w = (w + w[::-1]) / 2
    return x, w


class GaussLegendre:
    """Per-cell quadrature rule + global energy grid (see module docstring)."""

    def __init__(self, basis: FEMBasis, npoints: int = 1000):
        self.npoints = int(npoints)
        xl, wl = gauss_legendre(self.npoints)
        self.x = np.asarray(xl, dtype=np.float64)
        self.w = np.asarray(wl, dtype=np.float64)
        Rmax = basis.mesh.last
        # src/fem/integration methods.jl:52-54
        self.y = Rmax / 2 * self.x + Rmax / 2
        self.wy = Rmax / 2 * self.w
        G, _ = generator_values(basis.ordermax, self.x.astype(_LD))
        self.Pgenx = np.ascontiguousarray(np.asarray(G, dtype=np.float64))  # (p+1, Q)
        # global grid: owning cell (0-based) and generator values at phi(y)
        pts = basis.mesh.points
        cell = np.searchsorted(pts, self.y, side="right") - 1
        cell = np.clip(cell, 0, basis.ncells - 1)
        self.ycell = cell.astype(np.int32)
        c1 = basis.shifts[cell, 0]
        c0 = basis.shifts[cell, 1]
        xi = c1 * self.y + c0  # same float ops as `evaluate!`, src/fem/basis.jl:243-244
        Gy, _ = generator_values(basis.ordermax, xi.astype(_LD))
        self.Pgeny = np.ascontiguousarray(np.asarray(Gy, dtype=np.float64))  # (p+1, G)

    def qgenx(self) -> np.ndarray:
        """Pair-product table in the reference's layout: row i*(p+1)+j (0-based)
        holds g_j*g_i at the nodes (`src/fem/integration methods.jl:65`,
        `src/fem/polynomial.jl:256-271`)."""
        P = self.Pgenx
        n = P.shape[0]
        return (P[:, None, :] * P[None, :, :]).reshape(n * n, -1)
