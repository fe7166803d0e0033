"""Density-independent FEM operators (host setup; inputs of the device path).

API mirror of `FEMOperators` + `init_cache!`
(`src/discretization/discretization.jl:27-46,320-341`) and of the assembly
entry points `mass_matrix`, `stiffness_matrix`, `mass_tensor`
(`src/fem/matrices.jl:61-150,192-225,231-343`):

    M0[i,j]  = int phi_i phi_j dr          A[i,j]   = int phi_i' phi_j' dr
    Mm1[i,j] = int phi_i phi_j / r dr      Mm2[i,j] = int phi_i phi_j / r^2 dr
    F[i,j,m] = int phi_i phi_j phi_m / r dr

The reference integrates monomial-coefficient products exactly
(`src/fem/local matrix.jl:81-123`, pFq moments in
`src/fem/integration_formula.jl`), which loses up to 4 digits at order 20
(SURVEY.md section 0).  Here every local block is integrated with a
Gauss-Legendre rule that is exact for the polynomial part and converged to
rounding for the 1/r, 1/r^2 weights, in extended precision, then rounded once.
The operators are kept *unassembled*: one (p+1)x(p+1) block (or (p+1)^3
tensor) per cell in generator order -- the layout the device consumes -- and
assembled to dense global matrices on request.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .basis import FEMBasis, generator_values
from .quadrature import gauss_legendre

_LD = np.longdouble


@dataclass
class FEMOperators:
    """Cell-local operator blocks in generator order; entries of generators that
    do not exist on a cell (Dirichlet ends) are zero."""

    basis: FEMBasis
    M0_loc: np.ndarray   # (C, n, n)
    A_loc: np.ndarray    # (C, n, n)
    Mm1_loc: np.ndarray  # (C, n, n)
    Mm2_loc: np.ndarray  # (C, n, n)
    F_loc: np.ndarray    # (C, n, n, n)

    # ---- dense global views (reference layout) ---------------------------------
    def _assemble(self, loc: np.ndarray) -> np.ndarray:
        b = self.basis
        out = np.zeros((b.size, b.size))
        for c in range(b.ncells):
            Ib = np.asarray(b.cells_to_indices[c])
            Ig = np.asarray(b.cells_to_generators[c])
            out[np.ix_(Ib, Ib)] += loc[c][np.ix_(Ig, Ig)]
        return out

    @property
    def M0(self):
        return self._assemble(self.M0_loc)

    @property
    def A(self):
        return self._assemble(self.A_loc)

    @property
    def Mm1(self):
        return self._assemble(self.Mm1_loc)

    @property
    def Mm2(self):
        return self._assemble(self.Mm2_loc)

    def F_coo(self):
        """Mass tensor as COO (i, j, m, val), 0-based, duplicates merged -- the
        content of the reference's `Dict{NTuple{3,Int},T}` (`src/fem/matrices.jl:297-343`)."""
        b = self.basis
        acc = {}
        for c in range(b.ncells):
            Ib = b.cells_to_indices[c]
            Ig = b.cells_to_generators[c]
            T = self.F_loc[c]
            for ii, i in enumerate(Ib):
                for jj, j in enumerate(Ib):
                    for kk, m in enumerate(Ib):
                        key = (i, j, m)
                        acc[key] = acc.get(key, 0.0) + T[Ig[ii], Ig[jj], Ig[kk]]
        keys = np.array(list(acc.keys()), dtype=np.int64)
        vals = np.array(list(acc.values()), dtype=np.float64)
        return keys[:, 0].copy(), keys[:, 1].copy(), keys[:, 2].copy(), vals


def build_femops(basis: FEMBasis, *, need_Mm2: bool = True, need_F: bool = True,
                 nquad: int | None = None) -> FEMOperators:
    """Assemble all density-independent blocks (`init_cache!`,
    `src/discretization/discretization.jl:320-341`)."""
    p = basis.ordermax
    n = p + 1
    C = basis.ncells
    nq = nquad or (3 * p + 64)
    x, w = gauss_legendre(nq)
    G, dG = generator_values(p, x)
    pts = basis.mesh.points.astype(_LD)
    Khat = np.einsum("iq,jq,q->ij", G, G, w)
    Kdhat = np.einsum("iq,jq,q->ij", dG, dG, w)
    M0 = np.zeros((C, n, n))
    A = np.zeros((C, n, n))
    Mm1 = np.zeros((C, n, n))
    Mm2 = np.zeros((C, n, n))
    F = np.zeros((C, n, n, n)) if need_F else np.zeros((C, 0, 0, 0))
    for c in range(C):
        a, b = pts[c], pts[c + 1]
        h2 = (b - a) / 2
        mid = a + h2
        present = np.zeros(n, dtype=bool)
        present[np.asarray(basis.cells_to_generators[c])] = True
        mask2 = np.outer(present, present)
        M0[c] = np.asarray(h2 * Khat, dtype=np.float64) * mask2
        A[c] = np.asarray(Kdhat / h2, dtype=np.float64) * mask2
        Gc = G * present[:, None].astype(_LD)
        if a == 0:
            # first cell touches r = 0: every present generator carries a factor (1+x),
            # so phi/r is a polynomial (`fill_local_matrix_withsingularity!`,
            # src/fem/local matrix.jl:36-75); divide once, analytically for the hat.
            Gr = Gc / (h2 * (1 + x))[None, :]      # phi_i / r
            Gr[0] = (1 / h2) * present[0]
            Mm1[c] = np.asarray(h2 * np.einsum("iq,jq,q->ij", Gr, Gc, w), dtype=np.float64)
            Mm2[c] = np.asarray(h2 * np.einsum("iq,jq,q->ij", Gr, Gr, w), dtype=np.float64)
            if need_F:
                F[c] = np.asarray(h2 * np.einsum("iq,jq,mq,q->ijm", Gr, Gc, Gc, w,
                                                 optimize=True), dtype=np.float64)
        else:
            r = h2 * x + mid
            Mm1[c] = np.asarray(h2 * np.einsum("iq,jq,q->ij", Gc, Gc, w / r), dtype=np.float64)
            Mm2[c] = np.asarray(h2 * np.einsum("iq,jq,q->ij", Gc, Gc, w / (r * r)),
                                dtype=np.float64)
            if need_F:
                F[c] = np.asarray(h2 * np.einsum("iq,jq,mq,q->ijm", Gc, Gc, Gc, w / r,
                                                 optimize=True), dtype=np.float64)
        # exact symmetry (the integrals are symmetric; rounding of einsum order is not)
        Mm1[c] = (Mm1[c] + Mm1[c].T) / 2
        Mm2[c] = (Mm2[c] + Mm2[c].T) / 2
    if not need_Mm2:
        Mm2[:] = 0
    return FEMOperators(basis, M0, A, Mm1, Mm2, F)
