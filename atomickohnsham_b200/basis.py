"""P1 + integrated-Legendre hierarchical FEM basis (host side).

API mirror of `P1IntLegendreBasis` / `FEMBasis` (`src/fem/generators.jl:79-138`,
`src/fem/basis.py`-equivalent `src/fem/basis.jl:107-190`).  What is shared
with the reference is the *definition*: generators on [-1,1]

    g_1 = 1+x,  g_2 = 1-x,  g_{k+2}(x) = int_{-1}^x P_k = (P_{k+1}-P_{k-1})/(2k+1),  k=1..p-1

(`src/fem/generators.jl:34-57`), the global numbering and the per-cell index /
generator maps, and the affine maps `shifts` /
`invshifts` (`src/fem/basis.jl:185-190`).  How values are produced is ours:
the reference stores monomial coefficients and evaluates by Horner, which is
noisy at order 20 (finding 1); here generator values come
from the three-term Legendre recurrence evaluated in extended precision.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mesh import Mesh, findindex

_LD = np.longdouble


def legendre_values(nmax: int, x: np.ndarray, dtype=_LD) -> np.ndarray:
    """P_0..P_nmax at points x, Bonnet recurrence in `dtype`. Shape (nmax+1, len(x))."""
    x = np.asarray(x, dtype=dtype)
    P = np.zeros((nmax + 1,) + x.shape, dtype=dtype)
    P[0] = 1
    if nmax >= 1:
        P[1] = x
    for m in range(2, nmax + 1):
        P[m] = ((2 * m - 1) * x * P[m - 1] - (m - 1) * P[m - 2]) / dtype(m)
    return P


def generator_values(ordermax: int, x: np.ndarray, dtype=_LD):
    """Values and reference-coordinate derivatives of the p+1 generators at x.

    Row g (0-based) is reference generator g+1.  Returns (G, dG), each of
    shape (ordermax+1, len(x)) in `dtype`.
    """
    p = ordermax
    x = np.asarray(x, dtype=dtype)
    P = legendre_values(max(p, 1), x, dtype)
    G = np.zeros((p + 1,) + x.shape, dtype=dtype)
    dG = np.zeros_like(G)
    G[0] = 1 + x
    dG[0] = 1
    G[1] = 1 - x
    dG[1] = -1
    for k in range(1, p):
        G[k + 1] = (P[k + 1] - P[k - 1]) / dtype(2 * k + 1)
        dG[k + 1] = P[k]
    return G, dG


def _shift(a: float, b: float, m0: float, m1: float):
    """Affine map [a,b] -> [m0,m1] as (c1, c0); same float operations as
    `shift` in `src/fem/basis.jl:185-190`."""
    c1 = (m1 - m0) / (b - a)
    c0 = -a * c1 + m0
    return (c1, c0)


@dataclass
class FEMBasis:
    """Index maps of the hierarchical basis (0-based here, 1-based in the C ABI).

    cells_to_indices[c][j]   : global basis index of the j-th local function of cell c
    cells_to_generators[c][j]: generator id (0-based row of `generator_values`)
    loc2glob[c, g]           : global index of generator g on cell c, or -1 (Dirichlet end)
    shifts[c]   = (c1, c0)   : physical -> reference map of cell c
    invshifts[c]= (c1, c0)   : reference -> physical map (c1 = h/2, c0 = midpoint)
    """

    mesh: Mesh
    ordermax: int
    size: int
    cells_to_indices: list
    cells_to_generators: list
    loc2glob: np.ndarray
    shifts: np.ndarray
    invshifts: np.ndarray
    binf: float = -1.0
    bsup: float = 1.0

    def __len__(self):
        return self.size

    @property
    def ncells(self) -> int:
        return self.mesh.ncells

    @property
    def nloc(self) -> int:
        return self.ordermax + 1

    def evaluate_all(self, x: float) -> np.ndarray:
        """Values of all basis functions at physical point x
        (`(pb::FEMBasis)(x)`, `src/fem/basis.jl:221-226`)."""
        out = np.zeros(self.size)
        c = findindex(self.mesh, x)
        if c < 1 or c > self.ncells:
            return out
        c -= 1
        c1, c0 = self.shifts[c]
        G, _ = generator_values(self.ordermax, np.array([c1 * x + c0]))
        for j, I in enumerate(self.cells_to_indices[c]):
            out[I] = float(G[self.cells_to_generators[c][j], 0])
        return out

    def evaluate(self, coeffs: np.ndarray, X) -> np.ndarray:
        """sum_i coeffs[i] phi_i(x) at each x of X; zero outside the mesh
        (`evaluate!`, `src/fem/basis.jl:301-319`)."""
        X = np.atleast_1d(np.asarray(X, dtype=np.float64))
        out = np.zeros(X.shape[0])
        for k, xk in enumerate(X):
            c = findindex(self.mesh, xk)
            if not (1 <= c <= self.ncells):
                continue
            c -= 1
            c1, c0 = self.shifts[c]
            G, _ = generator_values(self.ordermax, np.array([c1 * xk + c0]))
            Ib = self.cells_to_indices[c]
            Ig = self.cells_to_generators[c]
            out[k] = float(np.dot(np.asarray(G[Ig, 0], dtype=np.float64), coeffs[Ib]))
        return out


def P1IntLegendreBasis(mesh: Mesh, T=np.float64, *, ordermax: int) -> FEMBasis:
    """Build the hierarchical basis on `mesh` (`src/fem/generators.jl:79-138`).

    Basis size is (p-1)*C + (C-1).  Global numbering: hat at node 2, bubbles of
    cell 1, hat at node 3, bubbles of cell 2, ..., bubbles of cell C.
    """
    assert ordermax >= 1
    p = ordermax
    C = mesh.ncells
    assert C >= 2, "need at least two cells"
    size = (p - 1) * C + (C - 1)
    c2i, c2g = [], []
    # first cell: [hat(node 2), bubbles]; generators [1, 3..p+1] -> 0-based [0, 2..p]
    c2i.append(list(range(0, p)))
    c2g.append([0] + list(range(2, p + 1)))
    for c in range(1, C - 1):
        h = c * p  # 0-based global index of the hat at the right node of cell c
        c2i.append([h - p, h] + list(range(h + 1, h + p)))
        c2g.append([1, 0] + list(range(2, p + 1)))
    hC = (C - 1) * p
    c2i.append([hC - p] + list(range(hC, hC + p - 1)))
    c2g.append([1] + list(range(2, p + 1)))
    assert c2i[-1][-1] == size - 1
    loc2glob = -np.ones((C, p + 1), dtype=np.int64)
    for c in range(C):
        for I, g in zip(c2i[c], c2g[c]):
            loc2glob[c, g] = I
    pts = mesh.points
    shifts = np.array([_shift(pts[c], pts[c + 1], -1.0, 1.0) for c in range(C)])
    invshifts = np.array([_shift(-1.0, 1.0, pts[c], pts[c + 1]) for c in range(C)])
    return FEMBasis(mesh, p, size, c2i, c2g, loc2glob, shifts, invshifts)
