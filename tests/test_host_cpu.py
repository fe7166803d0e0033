"""CPU-side checks of the product's host logic: the C-ABI library loads and exports every
symbol the header declares, fails loudly without a GPU, and the sweep partitioning / host
gather of the multi-GPU path works under a 2-rank gloo group."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "aks_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aks_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from atomickohnsham_b200.build import build_library
    build_library()
    from atomickohnsham_b200 import _capi
    names = _header_functions()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(_capi.lib, nm), nm
    assert set(_capi.EXPORTS) == set(names)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from atomickohnsham_b200 import _capi
    with pytest.raises(_capi.AKSError) as ei:
        _capi.Context(0)
    assert ei.value.code == _capi.AKS_ECUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "atomickohnsham_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "import_module(\"oracle" not in txt and "__import__(\"oracle" not in txt, f


def test_basis_maps_and_size_formula():
    """test/fem.jl:85-145: basis size (p-1)C+(C-1), continuity at nodes, zero outside."""
    from atomickohnsham_b200 import P1IntLegendreBasis, expmesh
    for npts in (5, 20):
        for p in (2, 4, 10):
            m = expmesh(0.0, 3.0, npts, s=1.3)
            b = P1IntLegendreBasis(m, ordermax=p)
            C = npts - 1
            assert b.size == (p - 1) * C + (C - 1)
            cover = np.zeros(b.size, dtype=int)
            for c in range(C):
                for I in b.cells_to_indices[c]:
                    cover[I] += 1
            assert cover.min() >= 1 and cover.max() == 2
    c = np.arange(1, b.size + 1, dtype=float)
    node = float(m.points[3])
    left = b.evaluate(c, [node - 1e-13])[0]
    right = b.evaluate(c, [node + 1e-13])[0]
    assert abs(left - right) < 1e-8
    assert b.evaluate(c, [m.first - 1e-12])[0] == 0.0


def test_mesh_generators():
    """test/fem.jl:1-83: endpoints, length, strict monotonicity; findindex sentinels."""
    from atomickohnsham_b200 import expmesh, findindex, geometricmesh, linmesh, polynomialmesh
    for n in (5, 20, 100):
        for m in (linmesh(0, 10, n), geometricmesh(0, 10, n, s=0.9), polynomialmesh(0, 10, n, s=2.0),
                  expmesh(0, 10, n, s=1.5)):
            assert len(m) == n and m.first == 0 and m.last == 10
            assert np.all(np.diff(m.points) > 0)
    m = expmesh(0, 10, 5, s=1.5)
    assert findindex(m, -1.0) == 0 and findindex(m, 11.0) == 6
    assert findindex(m, 10.0) == 4 and findindex(m, 0.0) == 1
    assert findindex(m, float(m.points[2])) == 3


def test_gauss_rule_and_quadrature_mass_matrix():
    """test/fem.jl:222-227: 60-point Gauss assembly with weight 1 reproduces the exact mass matrix."""
    from atomickohnsham_b200 import GaussLegendre, P1IntLegendreBasis, build_femops, expmesh
    b = P1IntLegendreBasis(expmesh(0.0, 3.0, 6, s=1.3), ordermax=4)
    ops = build_femops(b)
    q = GaussLegendre(b, 60)
    assert abs(q.w.sum() - 2.0) < 1e-14
    assert abs(np.dot(q.w, q.x ** 6) - 2.0 / 7.0) < 1e-14
    M = np.zeros((b.size, b.size))
    for c in range(b.ncells):
        Ib, Ig = np.array(b.cells_to_indices[c]), np.array(b.cells_to_generators[c])
        K = (q.Pgenx * q.w) @ q.Pgenx.T * b.invshifts[c][0]
        M[np.ix_(Ib, Ib)] += K[np.ix_(Ig, Ig)]
    assert np.allclose(M, ops.M0, atol=1e-12)
    for A in (ops.M0, ops.A, ops.Mm1, ops.Mm2):
        assert np.allclose(A, A.T, atol=1e-12)
        assert np.linalg.eigvalsh(A).min() > 0


def test_sweep_partition_balanced():
    from atomickohnsham_b200.sweep import basis_size, partition, problem_cost
    assert basis_size(1) == (1, 2) and basis_size(18) == (1, 10) and basis_size(21) == (2, 10) and basis_size(86) == (3, 10)
    costs = [problem_cost(Z) for Z in range(1, 87)]
    for world in (1, 2, 4, 8):
        parts = partition(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(86))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)


_GLOO_SCRIPT = r"""
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
os.environ["OMP_NUM_THREADS"] = "1"
import torch.distributed as dist
from atomickohnsham_b200.sweep import partition, problem_cost, basis_size
from helpers import make_oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
Zs = [1, 2, 3, 4, 5, 6]
mine = partition([problem_cost(Z) for Z in Zs], 2)[rank]
records = []
for i in mine:                      # this rank's share, solved for real (CPU stand-in of the rank's device batch)
    Z = Zs[i]
    lh, nh = basis_size(Z)
    o = make_oracle({{"Z": Z, "N": Z, "mesh": ["exp", 0, 60, 12, 1.5], "p": 4, "npoints": 60, "lh": lh, "nh": nh,
                     "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-9}})
    st, _ = o.groundstate(maxiter=60)
    records.append((Z, st.niter, st.energies.Etot, st.n.tolist()))
gathered = [None, None]
dist.all_gather_object(gathered, records)          # the only communication of the path: a host gather of records
if rank == 0:
    json.dump(sorted(r for part in gathered for r in part), open({out!r}, "w"))
dist.destroy_process_group()
"""


def test_two_rank_gloo_partitioned_sweep_matches_serial(tmp_path):
    """Multi-GPU path = LPT partition of independent problems + per-rank solves + a final host gather, no
    data-path collective.  Two gloo ranks solve their shares of a six-atom sweep FOR REAL (the CPU oracle stands in
    for the rank's device batch here; tests/test_gpu_round2.py runs the same with real device batches) and the
    gathered records equal the serial sweep, record for record."""
    import json
    from atomickohnsham_b200.sweep import basis_size
    from helpers import make_oracle
    out = str(tmp_path / "gather.json")
    script = tmp_path / "gloo_sweep.py"
    script.write_text(_GLOO_SCRIPT.format(root=ROOT, port=29500 + os.getpid() % 1000, out=out))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    got = json.load(open(out))
    assert [g[0] for g in got] == [1, 2, 3, 4, 5, 6], "host gather lost a problem"
    for Z, niter, etot, n in got:
        lh, nh = basis_size(Z)
        o = make_oracle({"Z": Z, "N": Z, "mesh": ["exp", 0, 60, 12, 1.5], "p": 4, "npoints": 60, "lh": lh, "nh": nh,
                         "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-9})
        st, _ = o.groundstate(maxiter=60)
        assert abs(st.energies.Etot - etot) <= 1e-12 * abs(etot) and st.n.tolist() == n, Z


def test_frozen_and_smeared_aufbau_host_and_oracle():
    """Behaviour pinned by the reference's own tests (test/algorithms.jl:43-175): shell parsing, the
    FrozenAufbauCache layout and its spin-suffix errors, SmearedAufbau's constructor check, electron-count
    conservation, bounds and monotonicity of the Fermi-Dirac occupations."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.algorithms import parse_shell
    assert parse_shell("1s") == (1, 0) and parse_shell("3dUP") == (3, 2, 1) and parse_shell("2p↓") == (2, 1, 2)
    with pytest.raises(ValueError):
        parse_shell("1p")
    af = aks.FrozenAufbau({"1s": 2.0, "2s": 1.5})
    n = af.nfix(lh=1, nh=2, nspin=1)
    assert n.shape == (2, 2) and n[0, 0] == 2.0 and n[0, 1] == 1.5 and n.sum() == 3.5
    with pytest.raises(ValueError):
        aks.FrozenAufbau({"1s": 1.0}).nfix(lh=0, nh=1, nspin=2)
    with pytest.raises(ValueError):
        aks.FrozenAufbau({"1sUP": 1.0}).nfix(lh=0, nh=1, nspin=1)
    for bad in (0.0, -0.1):
        with pytest.raises(ValueError):
            aks.SmearedAufbau(Temp=bad)
    # oracle: smeared occupations on the reference's test spectrum (lh = 1, nh = 2, N = 3)
    from oracle.scf import ODAOptions, Oracle

    class _O(Oracle):
        def __init__(self, opt):
            self.opt, self.N, self.nspin, self.L, self.nh = opt, 3.0, 1, 2, 2

        def flat_index(self, idx):
            return idx % 2, idx // 2, 0
    eps = np.array([-2.0, -1.0, 0.0, 1.0]).reshape((2, 2), order="F")
    cold, _ = _O(ODAOptions(aufbau_kind="smeared", temp=0.05)).aufbau(eps, None, 0)
    hot, _ = _O(ODAOptions(aufbau_kind="smeared", temp=1.0)).aufbau(eps, None, 0)
    deg = np.array([[2.0, 2.0], [6.0, 6.0]])
    for n in (cold, hot):
        assert abs(n.sum() - 3.0) < 1e-8 and np.all(n >= 0) and np.all(n <= deg)
        frac = (n / deg).reshape(-1, order="F")[np.argsort(eps.reshape(-1, order="F"))]
        assert np.all(np.diff(frac) <= 0)
    assert hot[1, 1] > cold[1, 1]
    fz, _ = _O(ODAOptions(aufbau_kind="frozen", nfix=n)).aufbau(eps + 100.0, None, 5)
    assert np.array_equal(fz, n)


def test_host_xc_matches_oracle():
    """The numpy LDA potentials used by post-processing (eval_vxc) against the oracle's restatement."""
    from atomickohnsham_b200._xc_host import vxc_polarised, vxc_unpolarised
    from oracle import xc as oxc
    rng = np.random.default_rng(3)
    rho = np.concatenate([[0.0, 1e-30], 10.0 ** rng.uniform(-12, 3, 200)])
    ref = oxc.slater_vrho(rho) + oxc.pw92_vrho(rho)
    assert np.allclose(vxc_unpolarised(rho, "lda_x", "lda_c_pw"), ref, rtol=1e-12, atol=1e-300)
    ru, rd = rho, rho[::-1].copy()
    vu, vd = vxc_polarised(ru, rd, "lda_x", "lda_c_pw")
    xu, xd = oxc.slater_vrho_pol(ru, rd)
    cu, cd = oxc.pw92_vrho_pol(ru, rd)
    assert np.allclose(vu, xu + cu, rtol=1e-11, atol=1e-300) and np.allclose(vd, xd + cd, rtol=1e-11, atol=1e-300)


def test_lane_balanced_order_is_a_balanced_permutation():
    """`sweep.lane_balanced_order` (host-side scheduling of a batch over the library's contiguous stream lanes): a
    permutation of the problems whose eight contiguous ranges carry nearly equal cost."""
    from atomickohnsham_b200.sweep import lane_balanced_order, periodic_table_models
    _, lhs, nhs = periodic_table_models(range(1, 87))
    costs = [(lhs[i] + 1.0) * nhs[i] + 4.0 for i in range(86)]
    order = lane_balanced_order(costs)
    assert sorted(order) == list(range(86))
    per = [sum(costs[i] for i in order[k * 86 // 8:(k + 1) * 86 // 8]) for k in range(8)]
    natural = [sum(costs[k * 86 // 8:(k + 1) * 86 // 8]) for k in range(8)]
    assert max(per) / min(per) < 1.35 < max(natural) / min(natural)
    assert lane_balanced_order([]) == [] and lane_balanced_order([3.0]) == [0]
