"""Radial meshes (host side of the drop-in API).

Mirrors the mesh constructors of the reference (`src/fem/mesh.jl:76-177`):
`linmesh`, `geometricmesh`, `polynomialmesh`, `expmesh`, `explinmesh` and the
cell lookup `findindex` (`src/fem/mesh.jl:64-75`).  Only the formulas are
shared; points are plain float64 numpy arrays (the Double64 path stores the
same float64-valued points, as the reference does: SURVEY.md Appendix A.1).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    """1-D mesh: `points` ascending, `name`, generator `params`."""

    points: np.ndarray
    name: str = ""
    params: dict = field(default_factory=dict)

    def __post_init__(self):
        self.points = np.ascontiguousarray(self.points, dtype=np.float64)

    def __len__(self):
        return self.points.shape[0]

    def __getitem__(self, i):
        return self.points[i]

    def __iter__(self):
        return iter(self.points)

    @property
    def first(self) -> float:
        return float(self.points[0])

    @property
    def last(self) -> float:
        return float(self.points[-1])

    @property
    def ncells(self) -> int:
        return len(self) - 1


def findindex(mesh: Mesh, x: float) -> int:
    """1-based cell index of `x` (reference semantics, `src/fem/mesh.jl:64-75`):
    right-owning at interior nodes, last node -> last cell, below the first
    node -> 0, above the last node -> len(mesh)+1."""
    pts = mesh.points
    if x <= pts[-1]:
        idx = int(np.searchsorted(pts, x, side="right"))  # searchsortedlast, 1-based
        if idx == len(pts):
            return idx - 1
        return idx
    return len(pts) + 1


def linmesh(a: float, b: float, n: int) -> Mesh:
    return Mesh(np.linspace(a, b, n), "Linear Mesh")


def geometricrange(a: float, b: float, n: int, *, s: float) -> np.ndarray:
    # src/fem/mesh.jl:89-106
    R = np.zeros(n)
    R[0] = a
    hn = (1.0 - s) / (1.0 - s ** (n - 1)) * (b - a)
    H = np.zeros(n - 1)
    H[-1] = hn
    for i in range(n - 2, 0, -1):
        H[i - 1] = s * H[i]
    for i in range(1, n):
        R[i] = R[i - 1] + H[i - 1]
    R[-1] = b
    return R


def geometricmesh(a: float, b: float, n: int, *, s: float) -> Mesh:
    return Mesh(geometricrange(a, b, n, s=s), "Geometric Mesh", {"s": s})


def polynomialrange(a: float, b: float, n: int, *, s: float) -> np.ndarray:
    # src/fem/mesh.jl:116-124
    R = np.zeros(n)
    R[0] = a
    R[-1] = b
    for i in range(2, n):
        R[i - 1] = ((i - 1) / (n - 1)) ** s * (float(b) - float(a)) + float(a)
    return R


def polynomialmesh(a: float, b: float, n: int, *, s: float) -> Mesh:
    return Mesh(polynomialrange(a, b, n, s=s), "Polynomial Mesh", {"s": s})


def exprange(a: float, b: float, n: int, *, s: float) -> np.ndarray:
    # src/fem/mesh.jl:134-143 : R[i] = (1+b-a)^(((i-1)/(n-1))^s) - 1 + a
    R = np.zeros(n)
    R[0] = a
    R[-1] = b
    for i in range(2, n):
        pw = ((i - 1) / (n - 1)) ** s
        R[i - 1] = (1 + b - a) ** pw - 1 + a
    return R


def expmesh(a: float, b: float, n: int, *, s: float) -> Mesh:
    """Exponentially graded mesh of `n` points on [a, b] (`src/fem/mesh.jl:153-155`)."""
    return Mesh(exprange(a, b, n, s=s), "Exponential Mesh", {"s": s})


def explinrange(a, b, n, *, rswitch, s=1.0, nlin=0):
    rexp = exprange(a, rswitch, n, s=s)
    if nlin > 0:
        rlin = np.linspace(rswitch, b, nlin + 1)
        return np.concatenate([rexp[:-1], rlin])
    return rexp


def explinmesh(a, b, n, *, rswitch, s=1.0, nlin):
    return Mesh(explinrange(a, b, n, rswitch=rswitch, s=s, nlin=nlin), "Exp-Lin Mesh",
                {"rswitch": rswitch, "nlin": nlin, "s": s})
