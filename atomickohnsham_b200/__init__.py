"""atomickohnsham_b200: B200-native SCF inner loop for radial Kohn-Sham atoms.

Host-side mirror of the AtomicKohnSham.jl public API for the hot path
(`KSEModel`, `KSEDiscretization`, `P1IntLegendreBasis` on exp/geometric meshes,
`ODA`, `groundstate`) over a C-ABI CUDA library (`include/aks_b200.h`).
"""
from .mesh import (Mesh, expmesh, geometricmesh, linmesh, polynomialmesh, explinmesh,  # noqa: F401
                   findindex)
from .basis import P1IntLegendreBasis, FEMBasis  # noqa: F401
from .quadrature import GaussLegendre  # noqa: F401
from .femops import FEMOperators, build_femops  # noqa: F401
from .dd import FEMOperatorsDD, GaussLegendreDD, build_femops_dd  # noqa: F401  (Double64 host tables)


def __getattr__(name):
    """Device-facing API (needs the built CUDA library; no CPU fallback)."""
    _dev = {"KSEModel": "model", "NoFunctional": "model", "BuiltinFunctional": "model", "Functional": "model",
            "RHF": "model", "Slater": "model", "LDA": "model", "ODA": "algorithms",
            "OptimizedAufbau": "algorithms", "SmearedAufbau": "algorithms", "FrozenAufbau": "algorithms",
            "parse_shell": "algorithms", "KSEDiscretization": "discretization",
            "groundstate": "solver", "KSESolver": "solver", "KSESolution": "solver", "Batch": "solver",
            "CallbackSet": "solver", "LogFileCallback": "solver", "LogBook": "solver",
            "eval_orbital": "postprocess", "eval_density": "postprocess", "eval_hartree": "postprocess",
            "eval_nuclear": "postprocess", "eval_kinetic_potential": "postprocess", "eval_vxc": "postprocess",
            "eval_effective_potential": "postprocess", "SanityChecks": "postprocess",
            "sanity_checks": "postprocess", "maximum_ionization": "sweep"}
    if name in _dev:
        import importlib
        return getattr(importlib.import_module(f".{_dev[name]}", __name__), name)
    raise AttributeError(name)
