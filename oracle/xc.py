"""Oracle: LDA / LSDA exchange-correlation, the reference's builtin formulas.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows src/physics/exchange correlation/SlaterXa.jl:11-30 (Slater exchange),
PerdewWang.jl:15-101 (PW92 correlation, unpolarised and polarised) and the
clamp-at-zero / zero-density guards of BuiltinFunctional.jl:74-142.  Constants
are the float64 values Julia produces.

Libxc (`Functional(:lda_x)`, `Functional(:lda_c_pw)`) is an un-vendored
dependency (Libxc_jll 7.0.0); the reference's test/physics.jl:47-72 pins the
builtin formulas to Libxc at 1e-10 (unpolarised) at run time only, so
Libxc-specific thresholds are "parity unpinned" here.

All functions are vectorised over numpy arrays and return 0 where rho == 0.
"""
from __future__ import annotations

import numpy as np

CX = (3.0 / np.pi) ** (1.0 / 3.0)          # (3/pi)^(1/3) = 0.9847450218426965
CX6 = (6.0 / np.pi) ** (1.0 / 3.0)         # (6/pi)^(1/3)
PW92_P0 = (0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294)
PW92_P1 = (0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517)
PW92_PA = (0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671)
PW92_F2P0 = 4.0 / (9.0 * (2.0 ** (1.0 / 3.0) - 1.0))
TWO43M2 = 2.0 ** (4.0 / 3.0) - 2.0


def _G(rs, A, a1, b1, b2, b3, b4):
    Q0 = -2 * A * (1 + a1 * rs)
    Q1 = 2 * A * (b1 * np.sqrt(rs) + b2 * rs + b3 * rs ** 1.5 + b4 * rs ** 2)
    return Q0 * np.log(1 + 1 / Q1)


def _dG(rs, A, a1, b1, b2, b3, b4):
    Q0 = -2 * A * (1 + a1 * rs)
    dQ0 = -2 * A * a1
    sq = np.sqrt(rs)
    Q1 = 2 * A * (b1 * sq + b2 * rs + b3 * rs ** 1.5 + b4 * rs ** 2)
    dQ1 = 2 * A * (b1 / (2 * sq) + b2 + 1.5 * b3 * sq + 2 * b4 * rs)
    return dQ0 * np.log(1 + 1 / Q1) - Q0 * dQ1 / (Q1 * (Q1 + 1))


def _safe(rho):
    """(clamped rho, mask rho>0, rho with zeros replaced by 1 for safe evaluation)."""
    r = np.maximum(np.asarray(rho, dtype=np.float64), 0.0)
    m = r > 0
    return r, m, np.where(m, r, 1.0)


# ---------------------------------------------------------------- unpolarised
def slater_zk(rho):
    r, m, rs_ = _safe(rho)
    return np.where(m, -0.75 * CX * rs_ ** (1.0 / 3.0), 0.0)


def slater_vrho(rho):
    r, m, rs_ = _safe(rho)
    return np.where(m, -CX * rs_ ** (1.0 / 3.0), 0.0)


def pw92_zk(rho):
    r, m, rr = _safe(rho)
    rs = (3.0 / (4.0 * np.pi * rr)) ** (1.0 / 3.0)
    return np.where(m, _G(rs, *PW92_P0), 0.0)


def pw92_vrho(rho):
    r, m, rr = _safe(rho)
    rs = (3.0 / (4.0 * np.pi * rr)) ** (1.0 / 3.0)
    ec = _G(rs, *PW92_P0)
    return np.where(m, ec - rs / 3.0 * _dG(rs, *PW92_P0), 0.0)


# ---------------------------------------------------------------- polarised
def slater_zk_pol(ru, rd):
    ru, _, _ = _safe(ru)
    rd, _, _ = _safe(rd)
    rho = ru + rd
    m = rho > 0
    rr = np.where(m, rho, 1.0)
    return np.where(m, -0.75 * CX6 * (ru ** (4.0 / 3.0) + rd ** (4.0 / 3.0)) / rr, 0.0)


def slater_vrho_pol(ru, rd):
    return slater_vrho(2.0 * np.maximum(ru, 0.0)), slater_vrho(2.0 * np.maximum(rd, 0.0))


def _fz(z):
    return ((1 + z) ** (4.0 / 3.0) + (1 - z) ** (4.0 / 3.0) - 2) / TWO43M2


def _dfz(z):
    return (4.0 / 3.0) * ((1 + z) ** (1.0 / 3.0) - (1 - z) ** (1.0 / 3.0)) / TWO43M2


def _pw92_pol_parts(ru, rd):
    ru, _, _ = _safe(ru)
    rd, _, _ = _safe(rd)
    rho = ru + rd
    m = rho > 0
    rr = np.where(m, rho, 1.0)
    z = np.where(m, (ru - rd) / rr, 0.0)
    rs = (3.0 / (4.0 * np.pi * rr)) ** (1.0 / 3.0)
    return m, z, rs


def pw92_zk_pol(ru, rd):
    m, z, rs = _pw92_pol_parts(ru, rd)
    ec0 = _G(rs, *PW92_P0)
    ec1 = _G(rs, *PW92_P1)
    ac = -_G(rs, *PW92_PA)
    f = _fz(z)
    ec = ec0 + ac * (f / PW92_F2P0) * (1 - z ** 4) + (ec1 - ec0) * f * z ** 4
    return np.where(m, ec, 0.0)


def pw92_vrho_pol(ru, rd):
    m, z, rs = _pw92_pol_parts(ru, rd)
    ec0 = _G(rs, *PW92_P0)
    ec1 = _G(rs, *PW92_P1)
    ac = -_G(rs, *PW92_PA)
    d0 = _dG(rs, *PW92_P0)
    d1 = _dG(rs, *PW92_P1)
    dac = -_dG(rs, *PW92_PA)
    f = _fz(z)
    df = _dfz(z)
    z3, z4 = z ** 3, z ** 4
    ec = ec0 + ac * (f / PW92_F2P0) * (1 - z4) + (ec1 - ec0) * f * z4
    decdrs = d0 + dac * (f / PW92_F2P0) * (1 - z4) + (d1 - d0) * f * z4
    decdz = ac / PW92_F2P0 * (df * (1 - z4) - 4 * z3 * f) + (ec1 - ec0) * (df * z4 + 4 * z3 * f)
    common = ec - rs / 3.0 * decdrs
    return (np.where(m, common + (1 - z) * decdz, 0.0),
            np.where(m, common - (1 + z) * decdz, 0.0))


# ---------------------------------------------------------------- model-level
class XC:
    """`evaluate_vrho!` / `evaluate_zk!` of a model (src/physics/models.jl:97-116):
    exchange + correlation, either may be absent.  ex in {None,'lda_x'},
    ec in {None,'lda_c_pw'}."""

    def __init__(self, ex=None, ec=None, nspin=1):
        assert ex in (None, "lda_x") and ec in (None, "lda_c_pw")
        self.ex, self.ec, self.nspin = ex, ec, nspin

    @property
    def active(self):
        return self.ex is not None or self.ec is not None

    def zk(self, *rho):
        if self.nspin == 1:
            out = np.zeros_like(np.asarray(rho[0], dtype=np.float64))
            if self.ec:
                out = out + pw92_zk(rho[0])
            if self.ex:
                out = out + slater_zk(rho[0])
            return out
        out = np.zeros_like(np.asarray(rho[0], dtype=np.float64))
        if self.ec:
            out = out + pw92_zk_pol(*rho)
        if self.ex:
            out = out + slater_zk_pol(*rho)
        return out

    def vrho(self, *rho):
        if self.nspin == 1:
            out = np.zeros_like(np.asarray(rho[0], dtype=np.float64))
            if self.ec:
                out = out + pw92_vrho(rho[0])
            if self.ex:
                out = out + slater_vrho(rho[0])
            return out
        vu = np.zeros_like(np.asarray(rho[0], dtype=np.float64))
        vd = np.zeros_like(vu)
        if self.ec:
            a, b = pw92_vrho_pol(*rho)
            vu, vd = vu + a, vd + b
        if self.ex:
            a, b = slater_vrho_pol(*rho)
            vu, vd = vu + a, vd + b
        return vu, vd
