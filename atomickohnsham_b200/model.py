"""KSEModel and the LDA/LSDA functional handles (host side of the drop-in API).

Mirrors `KSEModel(; Z, N, hartree=1, ex, ec)` (`src/physics/models.jl:19-58`),
`NoFunctional`, `BuiltinFunctional(:lda_x | :lda_c_pw; nspin)`
(`src/physics/exchange correlation/BuiltinFunctional.jl:13-39`) and the
shorthands `RHF`, `Slater` (`src/physics/models.jl:159-183`).  `Functional(...)`
is accepted as an alias of `BuiltinFunctional` for the two LDA identifiers the
reference uses through Libxc; GGA identifiers raise (out of scope).
"""
from __future__ import annotations

from dataclasses import dataclass

from ._capi import AKS_LDA_C_PW, AKS_LDA_X, AKS_XC_NONE


@dataclass(frozen=True)
class NoFunctional:
    n_spin: int = 1
    identifier: str = "none"

    @property
    def xc_id(self):
        return AKS_XC_NONE


@dataclass(frozen=True)
class BuiltinFunctional:
    identifier: str
    n_spin: int = 1

    def __post_init__(self):
        if self.n_spin not in (1, 2):
            raise ValueError("n_spin needs to be 1 or 2")
        if self.identifier not in ("lda_x", "lda_c_pw"):
            raise ValueError(f"{self.identifier} is not known for BuiltinFunctional.")

    @property
    def xc_id(self):
        return AKS_LDA_X if self.identifier == "lda_x" else AKS_LDA_C_PW


def Functional(identifier: str, n_spin: int = 1):
    ident = identifier.lstrip(":")
    if ident.startswith("gga"):
        raise NotImplementedError("GGA functionals are outside the B200 hot path (LDA/LSDA only)")
    return BuiltinFunctional(ident, n_spin)


class KSEModel:
    """Extended Kohn-Sham model of one atom or ion."""

    def __init__(self, *, Z, N, hartree=1, ex=None, ec=None):
        ex = ex if ex is not None else NoFunctional(1)
        ec = ec if ec is not None else NoFunctional(1)
        self.Z, self.N, self.hartree = Z, N, hartree
        self.exchange, self.correlation = ex, ec
        self.nspin = max(ex.n_spin, ec.n_spin)

    @property
    def has_exchcorr(self):
        return not (isinstance(self.exchange, NoFunctional) and isinstance(self.correlation, NoFunctional))


def RHF(*, Z, N):
    return KSEModel(Z=Z, N=N)


def Slater(*, Z, N, nspin: int = 1):
    return KSEModel(Z=Z, N=N, ex=BuiltinFunctional("lda_x", nspin), ec=NoFunctional(nspin))


def LDA(*, Z, N, nspin: int = 1):
    """lda_x + lda_c_pw (the functional pair of the reference's LDA examples)."""
    return KSEModel(Z=Z, N=N, ex=BuiltinFunctional("lda_x", nspin), ec=BuiltinFunctional("lda_c_pw", nspin))
