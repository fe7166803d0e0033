"""Shared test helpers: build discretizations and adapt the package's tables to
the oracle's duck-typed `tables` input (identical-tables parity)."""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def make_mesh(spec):
    from atomickohnsham_b200 import mesh as m
    kind, a, b, n, s = spec
    return {"exp": m.expmesh, "geom": m.geometricmesh, "poly": m.polynomialmesh}[kind](a, b, n, s=s)


_TABLE_CACHE = {}


def package_tables(mesh_spec, p, npoints=1000):
    """(basis, femops, quad) from the package's stable builder, cached."""
    key = (tuple(mesh_spec), p, npoints)
    if key not in _TABLE_CACHE:
        from atomickohnsham_b200.basis import P1IntLegendreBasis
        from atomickohnsham_b200.femops import build_femops
        from atomickohnsham_b200.quadrature import GaussLegendre
        mesh = make_mesh(mesh_spec)
        basis = P1IntLegendreBasis(mesh, ordermax=p)
        ops = build_femops(basis)
        quad = GaussLegendre(basis, npoints)
        _TABLE_CACHE[key] = (basis, ops, quad)
    return _TABLE_CACHE[key]


def oracle_tables_from_package(basis, ops, quad):
    """Adapt package tables to the attribute set oracle.scf.Oracle reads."""
    C = basis.ncells
    F_cells = []
    for c in range(C):
        Ig = np.asarray(basis.cells_to_generators[c])
        F_cells.append(np.ascontiguousarray(ops.F_loc[c][np.ix_(Ig, Ig, Ig)]))
    return SimpleNamespace(
        mesh=basis.mesh.points, p=basis.ordermax, size=basis.size,
        c2i=basis.cells_to_indices, c2g=basis.cells_to_generators,
        shifts=[tuple(v) for v in basis.shifts], invshifts=[tuple(v) for v in basis.invshifts],
        M0=ops.M0, A=ops.A, Mm1=ops.Mm1, Mm2=ops.Mm2, F_cells=F_cells,
        x=quad.x, w=quad.w, Pgenx=quad.Pgenx, Qgenx=quad.qgenx(), y=quad.y, wy=quad.wy,
        ycell=quad.ycell, Pgeny=quad.Pgeny, npoints=quad.npoints)


def make_oracle(setup, *, tables="package", eig="generalized", **over):
    from oracle.scf import ODAOptions, Oracle
    s = dict(setup)
    s.update(over)
    npoints = s.get("npoints", 1000)
    if tables == "package":
        tb = oracle_tables_from_package(*package_tables(s["mesh"], s["p"], npoints))
    else:
        from oracle.fem_reference import ReferenceTables
        mesh = make_mesh(s["mesh"])
        tb = ReferenceTables(mesh.points, s["p"], npoints)
    opt = ODAOptions(scftol=s.get("scftol", 1e-10), tinit=s.get("tinit", 0.6),
                     frozen_t=s.get("frozen_t", False), max_degen=s.get("max_degen", 2),
                     degen_tol=s.get("degen_tol", 1e-3), aufbau_kind=s.get("aufbau_kind", "optimized"),
                     temp=s.get("temp", 0.0), nfix=s.get("nfix"))
    nspin = s.get("nspin", 1)
    return Oracle(tb, Z=s["Z"], N=s["N"], hartree=s.get("hartree", 1), ex=s.get("ex"), ec=s.get("ec"),
                  nspin=nspin, lh=s["lh"], nh=s["nh"], options=opt, eig=eig)
