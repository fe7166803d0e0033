"""Post-processing of a `KSESolution` on arbitrary radial points, and the reference's sanity checks.

Host-side mirror of `src/solution/postprocess.jl:34-387` (`eval_orbital`, `eval_density`, `eval_hartree`,
`eval_nuclear`, `eval_kinetic_potential`, `eval_vxc`, `eval_effective_potential`) and of
`src/solution/sanity.jl:40-96` (`SanityChecks`).  These are the consumers of the hot path's outputs
(D, U, eps, n, W); they run on the host from the arrays `aks_get_state` returns.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .algorithms import parse_shell
from .basis import findindex, generator_values
from .model import NoFunctional


def _context(sol):
    if sol.context is None:
        raise ValueError("this KSESolution carries no (model, discretization) context")
    return sol.context


def _local(basis, x):
    """(global indices, generator values) of the basis functions alive at physical point x."""
    c = findindex(basis.mesh, x)
    if not (1 <= c <= basis.ncells):
        return None, None
    c -= 1
    c1, c0 = basis.shifts[c]
    G, _ = generator_values(basis.ordermax, np.array([c1 * x + c0]))
    Ib = np.asarray(basis.cells_to_indices[c])
    Ig = np.asarray(basis.cells_to_generators[c])
    return Ib, np.asarray(G[Ig, 0], dtype=np.float64)


def eval_orbital(sol, n, l=None, X=None, sigma=None, *, h10=False):
    """u_{nl}(r)/r (or the H^1_0 function u itself with h10=True) on X.  `n` may be a shell label such as
    "2p" or "3dUP" (postprocess.jl:34-62); sigma is 1 or 2 for spin-polarised solutions."""
    if isinstance(n, str):
        q = parse_shell(n)
        X = l if X is None else X
        n, l = q[0], q[1]
        sigma = q[2] if len(q) == 3 else sigma
    model, disc = _context(sol)
    if not 0 <= l <= n - 1:
        raise ValueError("Wrong number quantum. You should have 0 <= l <= n-1.")
    if disc.nspin == 2 and sigma is None:
        raise ValueError("The discretization is spin-polarized. Please give a spin sigma.")
    coeffs = sol.U[:, n - l - 1, l, (sigma or 1) - 1]
    X = np.atleast_1d(np.asarray(X, dtype=np.float64))
    v = disc.basis.evaluate(coeffs, X)
    return v if h10 else v / X


def eval_density(sol, X, sigma=None):
    """rho(X) = phi(X)^T D phi(X) / (4 pi X^2); spin-summed unless sigma is given (postprocess.jl:81-182)."""
    model, disc = _context(sol)
    X = np.atleast_1d(np.asarray(X, dtype=np.float64))
    D = sol.D
    spins = range(D.shape[2]) if sigma is None else [sigma - 1]
    out = np.zeros(X.shape[0])
    for k, x in enumerate(X):
        Ib, g = _local(disc.basis, x)
        if Ib is None:
            continue
        acc = 0.0
        for s in spins:
            acc += g @ D[np.ix_(Ib, Ib, [s])][:, :, 0] @ g
        out[k] = acc / (4.0 * np.pi * x * x)
    return out


def eval_hartree(sol, X):
    """V_H(X) = W(X)/X + N/Rmax (postprocess.jl:207-222); the boundary term alone when hartree == 0."""
    model, disc = _context(sol)
    X = np.atleast_1d(np.asarray(X, dtype=np.float64))
    boundary = model.N / disc.basis.mesh.last
    if model.hartree == 0 or sol.W is None or len(sol.W) == 0:
        return np.full(X.shape[0], boundary)
    return disc.basis.evaluate(np.asarray(sol.W), X) / X + boundary


def eval_nuclear(sol, X):
    """-Z/r, -inf at r = 0 (postprocess.jl:243-254)."""
    model, _ = _context(sol)
    X = np.atleast_1d(np.asarray(X, dtype=np.float64))
    with np.errstate(divide="ignore"):
        return np.where(X == 0.0, -np.inf, -model.Z / X)


def eval_kinetic_potential(l, X):
    """Centrifugal term l(l+1)/(2 r^2) (postprocess.jl:279-301)."""
    X = np.atleast_1d(np.asarray(X, dtype=np.float64))
    with np.errstate(divide="ignore"):
        return l * (l + 1) / (2.0 * X * X)


def eval_vxc(sol, X, sigma=None):
    """v_xc[rho](X) of the model's builtin LDA functionals (postprocess.jl:328-352)."""
    from ._xc_host import vxc_unpolarised, vxc_polarised
    model, disc = _context(sol)
    X = np.atleast_1d(np.asarray(X, dtype=np.float64))
    ex = None if isinstance(model.exchange, NoFunctional) else model.exchange.identifier
    ec = None if isinstance(model.correlation, NoFunctional) else model.correlation.identifier
    if disc.nspin == 1:
        return vxc_unpolarised(np.maximum(eval_density(sol, X), 0.0), ex, ec)
    vu, vd = vxc_polarised(np.maximum(eval_density(sol, X, 1), 0.0), np.maximum(eval_density(sol, X, 2), 0.0), ex, ec)
    if sigma is None:
        raise ValueError("The discretization is spin-polarized. Please give a spin sigma.")
    return vu if sigma == 1 else vd


def eval_effective_potential(sol, l, X, sigma=None):
    """-Z/r + l(l+1)/(2r^2) + hartree * V_H + v_xc (postprocess.jl:378-387)."""
    model, _ = _context(sol)
    v = eval_nuclear(sol, X) + eval_kinetic_potential(l, X) + model.hartree * eval_hartree(sol, X)
    if model.has_exchcorr:
        v = v + eval_vxc(sol, X, sigma)
    return v


@dataclass
class SanityChecks:
    """`SanityChecks` of the reference (sanity.jl:40-96); NaN where the quantity is unavailable."""
    energy_balance: float
    electron_count_error: float
    density_norm_error: float
    orthonormality_error: float
    virial_ratio: float


def sanity_checks(sol) -> SanityChecks:
    model, disc = _context(sol)
    e = sol.energies
    balance = e.Etot - (e.Ekin + e.Ecou + e.Ehar + e.Eexc)
    count = float(np.sum(sol.n) - model.N) if sol.n is not None else float("nan")
    if sol.D is not None:
        q = disc.fem_integration_method
        rho = eval_density(sol, q.y)
        dens = float(4.0 * np.pi * np.dot(q.wy, rho * q.y * q.y) - model.N)
    else:
        dens = float("nan")
    if sol.U is not None:
        M0 = disc.femops.M0
        orth = 0.0
        for s in range(sol.U.shape[3]):
            for l in range(sol.U.shape[2]):
                U = sol.U[:, :, l, s]
                orth = max(orth, float(np.max(np.abs(U.T @ M0 @ U - np.eye(U.shape[1])))))
    else:
        orth = float("nan")
    virial = float("nan") if e.Etot == 0 else e.Ekin / (-e.Etot)
    return SanityChecks(balance, count, dens, orth, virial)
