"""Oracle: the builtin LDA functionals as the reference evaluates them with T = Double64, in mpmath.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows src/physics/exchange correlation/SlaterXa.jl:11-12 and PerdewWang.jl:15-43 (spin-unpolarised) with the
number-type mixture Julia produces in a Double64 run: `(3/pi)^(1/3)`, `-3/4 * ...`,
`4pi` and the PW92 parameters are Float64 values, the exponent of `rho^(1/3)` in Slater exchange is the
Float64 number 0.333... promoted to Double64 (NOT 1/3), the rational powers `^(1//3)`, `^(3//2)` of PW92
are converted in T and therefore exact; densities are clamped at 0 and the functionals vanish at rho = 0
(BuiltinFunctional.jl:112, PerdewWang.jl:35-38).

Parity status: **unpinned in Double64** -- the reference commits no Double64 output with a functional (its
own Double64 example runs without one, examples/basic_example_doublefloat/run.jl:9-20) and Julia is not in
this image.  The restatement is pinned where it can be: evaluated at Float64 inputs it must agree with
oracle/xc.py (itself pinned to the reference's Float64 goldens) to Float64 accuracy
(tests/test_dd_host.py::test_xc_dd_oracle_matches_float64_oracle).
"""
from __future__ import annotations

import mpmath as mp

CX = mp.mpf((3.0 / 3.141592653589793) ** (1.0 / 3.0))      # Float64 (3/pi)^(1/3) = 0.9847450218426965
CX34 = mp.mpf(-0.75 * (3.0 / 3.141592653589793) ** (1.0 / 3.0))
THIRD_F64 = mp.mpf(1.0 / 3.0)                               # the Float64 number, not 1/3
FOURPI = mp.mpf(4.0 * 3.141592653589793)                    # Float64 4 pi
PW92_P0 = tuple(mp.mpf(v) for v in (0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294))


def rs_of_rho(rho):
    return mp.cbrt(3 / (FOURPI * rho))


def _G(rs):
    A, a1, b1, b2, b3, b4 = PW92_P0
    Q0 = -2 * A * (1 + a1 * rs)
    Q1 = 2 * A * (b1 * mp.sqrt(rs) + b2 * rs + b3 * rs ** mp.mpf(1.5) + b4 * rs ** 2)
    return Q0 * mp.log(1 + 1 / Q1)


def _dG(rs):
    A, a1, b1, b2, b3, b4 = PW92_P0
    Q0 = -2 * A * (1 + a1 * rs)
    dQ0 = mp.mpf(-2 * 0.031091 * 0.21370)   # -2A * alpha1 is a Float64 product (both factors are Float64)
    sq = mp.sqrt(rs)
    Q1 = 2 * A * (b1 * sq + b2 * rs + b3 * rs ** mp.mpf(1.5) + b4 * rs ** 2)
    dQ1 = 2 * A * (b1 / (2 * sq) + b2 + mp.mpf(1.5 * 1.6382) * sq + 2 * b4 * rs)   # 3//2 * beta3 is a Float64 product
    return dQ0 * mp.log(1 + 1 / Q1) - Q0 * dQ1 / (Q1 * (Q1 + 1))


def vrho(rho, ex: bool, ec: bool):
    rho = max(mp.mpf(rho), mp.mpf(0))
    if rho == 0:
        return mp.mpf(0)
    v = mp.mpf(0)
    if ec:
        rs = rs_of_rho(rho)
        v += _G(rs) - rs / 3 * _dG(rs)
    if ex:
        v += -CX * rho ** THIRD_F64
    return v


def zk(rho, ex: bool, ec: bool):
    rho = max(mp.mpf(rho), mp.mpf(0))
    if rho == 0:
        return mp.mpf(0)
    e = mp.mpf(0)
    if ec:
        e += _G(rs_of_rho(rho))
    if ex:
        e += CX34 * rho ** THIRD_F64
    return e
