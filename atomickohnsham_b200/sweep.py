"""Independent-problem sweeps: the only axis along which this path shards.

A single atom (Nh <= ~800) is far too small to split; the reference itself only ever
parallelises over atoms (`Threads.@threads for z`, `examples/periodic table/process.jl:92`).
`basis_size` is the Z-dependent (lh, nh) rule of `examples/atomic density comparison/run.jl:58`;
`partition` is a static longest-processing-time split of the problem list over ranks, with
no data-path collective -- ranks only gather O(100 B) records per problem at the end.
"""
from __future__ import annotations

import math

from .model import BuiltinFunctional, KSEModel, NoFunctional


def basis_size(Z: int):
    """(lh, nh) = (min(ceil(Z/20), 3), clamp(Z, 2, 10))."""
    return min(math.ceil(Z / 20), 3), min(max(int(Z), 2), 10)


def periodic_table_models(Zs=range(1, 87), *, ex="lda_x", ec="lda_c_pw", charge=0, nspin=1):
    """Neutral atoms (or ions with N = Z - charge) for a sweep; returns (models, lh[], nh[])."""
    mk = lambda name: NoFunctional(nspin) if name is None else BuiltinFunctional(name, nspin)
    models, lhs, nhs = [], [], []
    for Z in Zs:
        N = Z - charge
        if N < 1:
            continue
        models.append(KSEModel(Z=Z, N=N, ex=mk(ex), ec=mk(ec)))
        lh, nh = basis_size(Z)
        lhs.append(lh)
        nhs.append(nh)
    return models, lhs, nhs


def problem_cost(Z: int) -> float:
    """Relative cost of one SCF iteration of atom Z: eigenpairs solved + cells assembled."""
    lh, nh = basis_size(Z)
    return (lh + 1) * nh + 4.0


def partition(costs, world: int):
    """Static LPT partition: returns `world` index lists with balanced total cost."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: loads[j])
        parts[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in parts]
