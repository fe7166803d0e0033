"""Pointwise LDA potentials on the host (numpy), for post-processing only (`eval_vxc`).

Slater exchange (`src/physics/exchange correlation/SlaterXa.jl:11-30`) and PW92 correlation
(`.../PerdewWang.jl:15-101`), with the builtin functionals' conventions: density clamped at 0 by the caller,
zero potential at rho = 0 (`BuiltinFunctional.jl:74-142`).  The SCF path evaluates the same formulas on the
device (`csrc/aks_device.cuh`).
"""
from __future__ import annotations

import numpy as np

_P0 = (0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294)
_P1 = (0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517)
_PA = (0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671)
_CX = (3.0 / np.pi) ** (1.0 / 3.0)
_F2P0 = 4.0 / (9.0 * (2.0 ** (1.0 / 3.0) - 1.0))
_D43 = 2.0 ** (4.0 / 3.0) - 2.0


def _G(rs, P):
    A, a1, b1, b2, b3, b4 = P
    sq = np.sqrt(rs)
    Q0 = -2.0 * A * (1.0 + a1 * rs)
    Q1 = 2.0 * A * (b1 * sq + b2 * rs + b3 * rs * sq + b4 * rs * rs)
    dQ1 = 2.0 * A * (b1 / (2.0 * sq) + b2 + 1.5 * b3 * sq + 2.0 * b4 * rs)
    lg = np.log(1.0 + 1.0 / Q1)
    return Q0 * lg, -2.0 * A * a1 * lg - Q0 * dQ1 / (Q1 * (Q1 + 1.0))


def vxc_unpolarised(rho, ex, ec):
    rho = np.asarray(rho, dtype=np.float64)
    v = np.zeros_like(rho)
    pos = rho > 0
    r = rho[pos]
    if ec == "lda_c_pw":
        rs = np.cbrt(3.0 / (4.0 * np.pi * r))
        G, dG = _G(rs, _P0)
        v[pos] += G - rs / 3.0 * dG
    if ex == "lda_x":
        v[pos] += -_CX * np.cbrt(r)
    return v


def vxc_polarised(ru, rd, ex, ec):
    ru = np.asarray(ru, dtype=np.float64)
    rd = np.asarray(rd, dtype=np.float64)
    vu, vd = np.zeros_like(ru), np.zeros_like(rd)
    rho = ru + rd
    pos = rho > 0
    if ec == "lda_c_pw" and np.any(pos):
        r, a, b = rho[pos], ru[pos], rd[pos]
        z = (a - b) / r
        rs = np.cbrt(3.0 / (4.0 * np.pi * r))
        e0, d0 = _G(rs, _P0)
        e1, d1 = _G(rs, _P1)
        ea, da = _G(rs, _PA)
        ac, dac = -ea, -da
        f = ((1.0 + z) ** (4.0 / 3.0) + (1.0 - z) ** (4.0 / 3.0) - 2.0) / _D43
        df = (4.0 / 3.0) * (np.cbrt(1.0 + z) - np.cbrt(1.0 - z)) / _D43
        z3, z4 = z ** 3, z ** 4
        ecz = e0 + ac * (f / _F2P0) * (1.0 - z4) + (e1 - e0) * f * z4
        decdrs = d0 + dac * (f / _F2P0) * (1.0 - z4) + (d1 - d0) * f * z4
        decdz = ac / _F2P0 * (df * (1.0 - z4) - 4.0 * z3 * f) + (e1 - e0) * (df * z4 + 4.0 * z3 * f)
        common = ecz - rs / 3.0 * decdrs
        vu[pos] += common + (1.0 - z) * decdz
        vd[pos] += common - (1.0 + z) * decdz
    if ex == "lda_x":
        vu[ru > 0] += -_CX * np.cbrt(2.0 * ru[ru > 0])
        vd[rd > 0] += -_CX * np.cbrt(2.0 * rd[rd > 0])
    return vu, vd
