"""SCF algorithm options: `ODA` and `OptimizedAufbau`.

Mirrors `ODA(; scftol, tinit, frozen_t=false, aufbau=OptimizedAufbau(), maxiter_ls=100,
abstol_ls=1e4*eps, reltol_ls=abstol_ls)` (`src/algorithms/oda/types.jl:29-45`) and
`OptimizedAufbau(; max_degen=2, tol=1e-3, handle_degen_spin_1iter=true)`
(`src/algorithms/aufbau/optimized.jl:18-27`).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_EPS = float(np.finfo(np.float64).eps)


@dataclass
class OptimizedAufbau:
    max_degen: int = 2
    tol: float = 1e-3
    handle_degen_spin_1iter: bool = True


@dataclass
class ODA:
    scftol: float
    tinit: float
    frozen_t: bool = False
    aufbau: OptimizedAufbau = field(default_factory=OptimizedAufbau)
    maxiter_ls: int = 100
    abstol_ls: float = 1e4 * _EPS
    reltol_ls: float | None = None

    def __post_init__(self):
        if self.reltol_ls is None:
            self.reltol_ls = self.abstol_ls
