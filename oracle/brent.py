"""Oracle: Optim.jl `Brent()` bounded minimiser as the reference calls it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Call site: `optimize(f, 0, 1, Brent(); abs_tol, rel_tol, iterations)` in
src/algorithms/optimization/line_search.jl:59-67.  Optim.jl (1.13.3, pinned in
examples/Manifest.toml) is not vendored under /root/reference; this restates
its published algorithm (Brent 1973, "localmin", in Optim's
src/univariate/solvers/brent.jl form).  Pinned by the `t` column of the
committed sodium trajectory examples/basic_example/log.txt:35-57: first
evaluation at (3-sqrt 5)/2, boundary result 0.999999999991612, golden-section
tail values (tests/test_oracle_golden.py::test_brent_*).
"""
from __future__ import annotations

import math


def brent_minimize(f, lo: float, hi: float, *, abs_tol: float, rel_tol: float,
                   iterations: int = 100, sqrt=math.sqrt):
    """Returns (minimizer, minimum, n_f_calls).  Works on any number type with the usual operators (the Double64
    oracle passes mpmath numbers and `sqrt = mpmath.sqrt`: Optim forms the golden ratio in the type of the bounds)."""
    golden = (3 - sqrt(lo * 0 + 5)) / 2
    x = lo + golden * (hi - lo)
    fx = f(x)
    calls = 1
    step = 0.0
    old_step = 0.0
    ox = oox = x
    ofx = oofx = fx
    it = 0
    while it < iterations:
        p = 0.0
        q = 0.0
        x_tol = rel_tol * abs(x) + abs_tol
        mid = (hi + lo) / 2
        if abs(x - mid) <= 2 * x_tol - (hi - lo) / 2:
            break
        it += 1
        if abs(old_step) > x_tol:
            r = (x - ox) * (fx - oofx)
            q = (x - oox) * (fx - ofx)
            p = (x - oox) * q - (x - ox) * r
            q = 2 * (q - r)
            if q > 0:
                p = -p
            else:
                q = -q
        if abs(p) < abs(q * old_step / 2) and p < q * (hi - x) and p < q * (x - lo):
            old_step = step
            step = p / q
            xt = x + step
            if (xt - lo) < 2 * x_tol or (hi - xt) < 2 * x_tol:
                step = x_tol if x < mid else -x_tol
        else:
            old_step = (hi - x) if x < mid else (lo - x)
            step = golden * old_step
        if abs(step) >= x_tol:
            nx = x + step
        else:
            nx = x + (x_tol if step > 0 else -x_tol)
        nf = f(nx)
        calls += 1
        if nf < fx:
            if nx < x:
                hi = x
            else:
                lo = x
            oox, oofx = ox, ofx
            ox, ofx = x, fx
            x, fx = nx, nf
        else:
            if nx < x:
                lo = nx
            else:
                hi = nx
            if nf <= ofx or ox == x:
                oox, oofx = ox, ofx
                ox, ofx = nx, nf
            elif nf <= oofx or oox == x or oox == ox:
                oox, oofx = nx, nf
    return x, fx, calls
