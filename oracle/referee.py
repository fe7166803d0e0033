"""Referee for orbital energies: generalized eigenpairs of (H, M0) refined in double-double on the host.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Why it exists.  The reference computes `eigen(Symmetric(S*H*S), 1:nh)` with LAPACK
(src/algorithms/oda/loop.jl:65-77, S = sqrt(inv(M0)), src/discretization/discretization.jl:331) and the oracle
calls the same LAPACK through SciPy.  At order 20 (Nh = 799) the pencil has ||M0^-1 H|| ~ 1e7 and LAPACK's
eigenvalues carry an ABSOLUTE error of 1e-10 ... 5e-9 -- far above the 1e-10 RELATIVE tolerance the north star
asks for on levels like Ar 3p (-0.38 Ha).  Comparing a GPU eigenvalue with LAPACK at 1e-10 would therefore test
LAPACK, not the GPU.  The referee removes LAPACK from the comparison: starting from any approximate pair it
runs a Newton iteration on the eigenpair equations

        (H - lam M) x = 0,   x^T M x = 1

whose residuals are evaluated in double-double arithmetic (atomickohnsham_b200/dd.py: TwoSum / TwoProd on NumPy
arrays, 106 bits) and whose corrections come from a Float64 LU of the bordered Jacobian
[[H - lam M, -M x], [(M x)^T, 0]] (iterative refinement: the correction only needs a few digits, the residual
needs them all).  The iterate is kept as a (hi, lo) pair of vectors.  It converges quadratically to the eigenpair
of the GIVEN Float64 matrices; the returned residual norm is the certificate:
|lam - lam_exact| <= ||r||_{M^-1}^2 / gap.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from atomickohnsham_b200.dd import DD


def _dd_tree_sum(v: DD) -> DD:
    """Sum over the LAST axis by pairwise (tree) addition in double-double: log2(n) vectorised passes."""
    hi, lo = v.hi, v.lo
    while hi.shape[-1] > 1:
        m = hi.shape[-1]
        if m % 2:
            hi = np.concatenate([hi, np.zeros(hi.shape[:-1] + (1,))], axis=-1)
            lo = np.concatenate([lo, np.zeros(lo.shape[:-1] + (1,))], axis=-1)
            m += 1
        acc = DD(hi[..., : m // 2], lo[..., : m // 2]) + DD(hi[..., m // 2:], lo[..., m // 2:])
        hi, lo = acc.hi, acc.lo
    return DD(hi[..., 0], lo[..., 0])


def _dd_matvec(A: np.ndarray, x: DD) -> DD:
    """A x with A Float64 (exact in double-double) and x double-double; row sums accumulated in double-double."""
    return _dd_tree_sum(DD(A) * DD(x.hi[None, :], x.lo[None, :]))


def _dd_dot(a: DD, b: DD) -> DD:
    return _dd_tree_sum(a * b)


def refine_eigenpair(H: np.ndarray, M: np.ndarray, lam0: float, x0: np.ndarray, *, steps: int = 4):
    """Newton refinement of one eigenpair of the pencil (H, M) in double-double.

    Returns (lam_hi, lam_lo, x_hi, residual_norm) with lam = lam_hi + lam_lo the refined eigenvalue and
    residual_norm = ||(H - lam M) x||_2 / ||x||_2 evaluated in double-double (a converged pair shows ~1e-25)."""
    n = H.shape[0]
    x = DD(np.array(x0, dtype=np.float64), np.zeros(n))
    # M-normalise
    Mx = _dd_matvec(M, x)
    nrm = _dd_dot(x, Mx)
    sc = 1.0 / np.sqrt(float(nrm.hi))
    x = x * DD(np.float64(sc))
    lam = DD(np.float64(lam0))
    res = np.inf
    J = np.zeros((n + 1, n + 1))
    for _ in range(steps):
        Hx = _dd_matvec(H, x)
        Mx = _dd_matvec(M, x)
        xMx = _dd_dot(x, Mx)
        lam = _dd_dot(x, Hx) / xMx                       # Rayleigh quotient in double-double
        r = Hx - Mx * lam                                # residual in double-double
        g = (xMx - DD(np.float64(1.0))) * DD(np.float64(0.5))
        res = float(np.sqrt(np.sum(r.hi ** 2)) / np.sqrt(np.sum(x.hi ** 2)))
        if res < 1e-26 * max(1.0, abs(float(lam.hi))):
            break
        # Float64 Jacobian at the current point; its LU solves for the correction of (x, lam)
        J[:n, :n] = H - float(lam.hi) * M
        J[:n, n] = -Mx.hi
        J[n, :n] = Mx.hi
        J[n, n] = 0.0
        rhs = np.concatenate([-(r.hi + r.lo), [-(float(g.hi) + float(g.lo))]])
        d = sla.solve(J, rhs)
        x = x + DD(d[:n])
        lam = lam + DD(np.float64(d[n]))
    Hx = _dd_matvec(H, x)
    Mx = _dd_matvec(M, x)
    lam = _dd_dot(x, Hx) / _dd_dot(x, Mx)
    r = Hx - Mx * lam
    res = float(np.sqrt(np.sum((r.hi + r.lo) ** 2)) / np.sqrt(np.sum(x.hi ** 2)))
    return float(lam.hi), float(lam.lo), x.hi.copy(), res


def referee_levels(H: np.ndarray, M: np.ndarray, nlev: int):
    """Lowest `nlev` eigenvalues of (H, M): LAPACK pairs (which fix the index of every level; its errors are far
    below the level spacing) refined by `refine_eigenpair`.  Returns (lam_ref [nlev], lam_lapack [nlev],
    residual norms [nlev])."""
    w, V = sla.eigh(H, M, subset_by_index=[0, nlev - 1])
    ref, res = np.zeros(nlev), np.zeros(nlev)
    for k in range(nlev):
        hi, lo, _, rn = refine_eigenpair(H, M, w[k], V[:, k])
        ref[k], res[k] = hi + lo, rn
    return ref, w, res
