"""SCF algorithm options: `ODA` and `OptimizedAufbau`.

Mirrors `ODA(; scftol, tinit, frozen_t=false, aufbau=OptimizedAufbau(), maxiter_ls=100,
abstol_ls=1e4*eps, reltol_ls=abstol_ls)` (`src/algorithms/oda/types.jl:29-45`) and
`OptimizedAufbau(; max_degen=2, tol=1e-3, handle_degen_spin_1iter=true)`
(`src/algorithms/aufbau/optimized.jl:18-27`).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

_EPS = float(np.finfo(np.float64).eps)


@dataclass
class OptimizedAufbau:
    max_degen: int = 2
    tol: float = 1e-3
    handle_degen_spin_1iter: bool = True


class SmearedAufbau:
    """SmearedAufbau(; Temp): Fermi-Dirac occupations g_i f(eps_i; mu, Temp), mu by bisection
    (src/algorithms/aufbau/smeared.jl:24-31, 68-101).  Temp in Hartree, strictly positive."""

    def __init__(self, Temp):
        if not (Temp > 0):
            raise ValueError("Temp must be strictly positive.")
        self.Temp = float(Temp)


_SHELL_RE = re.compile(r"^(\d+)\s*([spdfghi])\s*(UP|DOWN|↑|↓)?$", re.IGNORECASE)


def parse_shell(label):
    """"2p" -> (2, 1); "3dUP" -> (3, 2, 1); "1s↓" -> (1, 0, 2)  (src/utils/orbital_labels.jl:68-84)."""
    m = _SHELL_RE.match(label.strip())
    if m is None:
        raise ValueError(f"Cannot parse '{label}'")
    n, l = int(m.group(1)), "spdfghi".index(m.group(2).lower())
    if not n > l:
        raise ValueError(f"Convention violated: need n>l, got n={n} l={l}")
    if m.group(3) is None:
        return (n, l)
    return (n, l, 1 if m.group(3).upper() in ("UP", "↑") else 2)


class FrozenAufbau:
    """FrozenAufbau(n::Dict): user-prescribed occupations by shell label, copied verbatim in every
    iteration (src/algorithms/aufbau/frozen.jl:18-88)."""

    def __init__(self, n):
        self.n = {str(k): float(v) for k, v in dict(n).items()}

    def nfix(self, lh, nh, nspin):
        """FrozenAufbauCache: the occupation array in the discretization's layout (L, nh[, 2])."""
        import numpy as np
        shells = [parse_shell(k) for k in self.n]
        longest = max(len(q) for q in shells)
        if longest < 3 and nspin > 1:
            raise ValueError('You have a spinned-model. Please provide occupation numbers with the spin index. '
                             'For example "1sUP".')
        if longest == 3 and nspin == 1:
            raise ValueError('You have a no spinned-model. Please provide occupation numbers without the spin '
                             'index. For example "1s".')
        out = np.zeros((lh + 1, nh) if nspin == 1 else (lh + 1, nh, 2))
        for key, q in zip(self.n, shells):
            idx = (q[1], q[0] - q[1] - 1) if len(q) == 2 else (q[1], q[0] - q[1] - 1, q[2] - 1)
            out[idx] = self.n[key]
        return out


@dataclass
class ODA:
    scftol: float
    tinit: float
    frozen_t: bool = False
    aufbau: object = field(default_factory=OptimizedAufbau)   # OptimizedAufbau | SmearedAufbau | FrozenAufbau
    maxiter_ls: int = 100
    abstol_ls: float = 1e4 * _EPS
    reltol_ls: float | None = None

    def __post_init__(self):
        if self.reltol_ls is None:
            self.reltol_ls = self.abstol_ls
