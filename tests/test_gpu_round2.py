"""GPU parity tests added in round 2 (all through the C ABI):

  * orbital energies at the north-star tolerance, 1e-10 RELATIVE, against the double-double referee
    (oracle/referee.py) instead of LAPACK -- C1, C2, C3 and six atoms of C5;
  * per-problem ODA options and iteration caps (aks_batch_set_oda_problem: the reference picks the algorithm per
    atom, examples/atomic density comparison/run.jl:74-76);
  * the device-driven run loop (device-side history, status polling every k steps) and the CUDA-graph replay of
    the step are bit-transparent;
  * no eigenpair of the windows is left uncertified on the C1-C5 workloads (AKS_EEIG never raised);
  * the partitioned sweep: two ranks (processes), real batches, host gather == one batch.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
pytestmark = pytest.mark.gpu

from helpers import make_oracle  # noqa: E402
from test_gpu_parity import CASES, gpu_batch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C2MESH = ["exp", 0, 300, 41, 1.5]


def _c5_setup(Z):
    from atomickohnsham_b200.sweep import basis_size
    lh, nh = basis_size(Z)
    return {"Z": Z, "N": Z, "mesh": C2MESH, "p": 20, "lh": lh, "nh": nh, "ex": "lda_x", "ec": "lda_c_pw",
            "scftol": 1e-10}


REFEREE_CASES = {"c1_hydrogen": CASES["h_slater_c1"], "c2_argon": CASES["ar_lda_c2"], "c3_scandium_lsda": CASES["sc_lsda_c3"],
                 **{f"c5_Z{Z}": _c5_setup(Z) for Z in (3, 10, 26, 36, 54, 86)}}


@pytest.mark.parametrize("name", list(REFEREE_CASES))
def test_orbital_energies_against_dd_referee(name):
    """North star: orbital energies within 1e-10 relative.  On identical input D (the oracle's 5th iterate) the
    GPU eigenvalues of every bound level are compared with the eigenvalues of the ORACLE's H and M0 refined in
    double-double (Newton iteration with double-double residuals, certificate: residual norm < 1e-20).  LAPACK --
    what the reference and the oracle call -- is measured against the same referee and is the inaccurate side at
    order 20 (its error is asserted in tests/test_oracle_golden.py::test_lapack_error_against_referee)."""
    from oracle.referee import refine_eigenpair
    import scipy.linalg as sla
    setup = REFEREE_CASES[name]
    o = make_oracle(setup)
    st = o.initial_state()
    for _ in range(5):
        o.scf_step(st)
    bt, *_ = gpu_batch(setup)
    e = st.energies
    bt.set_density(0, st.D, [e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc], niter=st.niter)
    eps_g, U_g, res = bt.eig_lowest(0)
    assert not bt.eig_info()[2].any(), "an eigenpair was not converged + certified"
    Har, Vxc, _ = o.hamiltonian_parts(st.D)
    worst_gpu, worst_lapack, nlev = 0.0, 0.0, 0
    for s in range(o.nspin):
        for l in range(o.L):
            H = o.Hfix[l] + Vxc[:, :, s] + Har
            w, V = sla.eigh(H, o.tb.M0, subset_by_index=[0, o.nh - 1])
            for k in range(o.nh):
                if w[k] >= -1e-3:      # bound levels: every level an aufbau of this atom can occupy
                    break
                hi, lo, _, rn = refine_eigenpair(H, o.tb.M0, w[k], V[:, k])
                ref = hi + lo
                assert rn < 1e-20 * max(1.0, abs(ref)), (name, l, k, rn)          # the referee's own certificate
                err = abs(eps_g[l, k, s] - ref) / abs(ref)
                assert err <= 1e-10, (name, s, l, k, eps_g[l, k, s], ref, err)     # NORTH-STAR TOLERANCE
                worst_gpu = max(worst_gpu, err)
                worst_lapack = max(worst_lapack, abs(w[k] - ref) / abs(ref))
                nlev += 1
    assert nlev >= 1
    print(f"[referee] {name}: {nlev} bound levels, worst relative error GPU {worst_gpu:.1e}, LAPACK {worst_lapack:.1e}")
    bt.close()


def test_c3_scandium_lsda_to_convergence():
    """Config C3 (Sc, LSDA, order 20, scftol 1e-8) run to the end on both sides.  The oracle is the checker:
    same status, Etot to 1e-10 relative, integer occupations bit-exact, the fractional 3d/4s pair to the accuracy
    the flat line search allows, either spin orientation (the two are exactly degenerate, see
    test_c3_scandium_lsda_steps)."""
    setup = CASES["sc_lsda_c3"]
    o = make_oracle(setup)
    st, hist = o.groundstate(maxiter=100)
    bt, *_ = gpu_batch(setup)
    final, _ = bt.run(100)
    r = final[0]
    conv_o = st.stop < setup["scftol"]
    print(f"[C3] oracle: niter {st.niter} stop {st.stop:.2e} Etot {st.energies.Etot!r}; gpu: niter {r['niter']} "
          f"stop {r['stop']:.2e} Etot {r['Etot']!r} status {r['status']}")
    assert (r["status"] == 1) == conv_o
    assert abs(r["Etot"] - st.energies.Etot) <= 1e-10 * abs(st.energies.Etot)
    sg = bt.state(0, want_D=False, want_U=False)
    same = np.allclose(sg["n"], st.n, atol=2e-3)
    mirrored = np.allclose(sg["n"][..., ::-1], st.n, atol=2e-3)
    assert same or mirrored
    ng = sg["n"] if same else sg["n"][..., ::-1]
    integer = st.n == np.round(st.n)
    assert np.array_equal(ng[integer], st.n[integer])
    assert abs(ng.sum() - 21) < 1e-9
    bt.close()


def test_per_problem_oda_options_and_device_side_maxiter():
    """One batch, two atoms, each with ITS algorithm and iteration cap (run.jl:74-76) == two batches with
    batch-wide options, bit for bit; the cap stops a problem on the device with status MAXITERS."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from bench import build_disc
    from atomickohnsham_b200.sweep import periodic_table_models
    disc = build_disc(0)
    models, lhs, nhs = periodic_table_models([18, 26])
    default = aks.ODA(scftol=1e-10, tinit=0.6)
    iron = aks.ODA(scftol=1e-12, tinit=0.15, frozen_t=True)
    both = Batch(models, disc, [default, iron], lh=lhs, nh=nhs, maxiter=[100, 37])
    fb, hb = both.run(100, history=True)
    both.close()
    a = Batch(models[:1], disc, default, lh=lhs[:1], nh=nhs[:1])
    fa, ha = a.run(100, history=True)
    a.close()
    f = Batch(models[1:], disc, iron, lh=lhs[1:], nh=nhs[1:])
    ff, hf = f.run(37, history=True)
    f.close()
    assert fb[0] == fa[0] and hb[0] == ha[0]
    assert fb[0]["status"] == 1
    assert fb[1]["status"] == 2 and fb[1]["niter"] == 37          # stopped by its own cap, on the device
    assert [h["Etot"] for h in hb[1]] == [h["Etot"] for h in hf[0]]
    assert all(h["t"] == 0.15 for h in hb[1][1:])                 # frozen_t for this problem only
    assert any(h["t"] != 0.6 for h in hb[0][1:])                  # the neighbour runs the adaptive line search


def test_run_loop_and_graph_replay_are_bit_transparent():
    """aks_scf_run (status words of a chunk of 4 steps read back while the next chunk runs, history written on the
    device, step replayed as a CUDA graph) == a host loop of aks_scf_step with plain stream launches, record for record."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from bench import build_disc
    from atomickohnsham_b200.sweep import periodic_table_models
    disc = build_disc(0)
    zs = list(range(1, 31))     # <= 32 problems: the graph replay is active (4 lanes)
    models, lhs, nhs = periodic_table_models(zs)
    alg = aks.ODA(scftol=1e-10, tinit=0.6)
    ctx = disc.ctx
    out = {}
    try:
        for mode in ("graph_run", "plain_steps", "graph_run_every_1"):
            ctx.configure(use_graphs=(mode != "plain_steps"), check_every=1 if mode.endswith("_1") else 4)
            bt = Batch(models, disc, alg, lh=lhs, nh=nhs)
            l0 = ctx.launches
            if mode == "plain_steps":
                hist = [[] for _ in zs]
                for _ in range(30):
                    for i, r in enumerate(bt.step()):
                        if not hist[i] or hist[i][-1]["niter"] != r["niter"]:
                            hist[i].append(r)
                final = [dict(h[-1]) for h in hist]
            else:
                final, hist = bt.run(30, history=True)
            out[mode] = (final, hist, ctx.launches - l0)
            bt.close()
    finally:
        ctx.configure(use_graphs=True, check_every=4)
    ref_final, ref_hist, _ = out["plain_steps"]
    for mode in ("graph_run", "graph_run_every_1"):
        final, hist, _ = out[mode]
        for i in range(len(zs)):
            assert hist[i] == ref_hist[i], (mode, zs[i])
            a, b = dict(final[i]), dict(ref_final[i])
            if b["status"] == 0:
                b["status"] = 2      # aks_scf_run marks what is still running after maxiter steps
            assert a == b, (mode, zs[i])
    assert out["graph_run"][2] > 0   # gpu_launches counts the kernels of every replay


def test_no_uncertified_eigenpair_on_the_sweep():
    """an eigenpair that is not converged + certified must never be used silently.  It now raises
    AKS_EEIG for its problem; on the C5 workload (86 atoms, 30 steps from the cold start: shell reordering,
    refills, near-degenerate d/f blocks) it never happens."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from bench import build_disc
    from atomickohnsham_b200.sweep import periodic_table_models
    disc = build_disc(0)
    models, lhs, nhs = periodic_table_models(range(1, 87))
    bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs)
    for it in range(30):
        recs = bt.step()
        assert all(r["status"] == 0 for r in recs), [(i + 1, r["status"]) for i, r in enumerate(recs) if r["status"]]
        assert not bt.eig_info()[2].any(), it
    bt.close()


def _c5_final(a):
    """The sweep's end state of one atom of the golden file: the better of the two passes (sweep.solve_sweep)."""
    p1, p2 = a["pass1"], a.get("pass2")
    if p2 is not None and (p2["status"] == "SUCCESS" or p2["stop"] < p1["stop"]):
        return p2, 2
    return p1, 1


def test_c5_sweep_against_oracle_golden():
    """North star "Target": the periodic-table LDA sweep Z = 1..86 at order 20 reproduces the reference's energies.
    tests/golden/c5_sweep.json is the oracle's run of the reference's own protocol (tools/make_golden_c5.py:
    examples/atomic density comparison/run.jl -- default ODA(0.4, 1e-10) + OptimizedAufbau(tol = 1e-2) for everyone,
    iron's frozen-t algorithm as a second pass for what ends at MAXITERS, the better end state kept).  The same
    protocol runs here through `sweep.solve_sweep` (two device batches) and all 86 atoms are compared:
      * 86 of 86 end in a defined state (SUCCESS, or MAXITERS on the noise floor of the line searches: stop < 1e-5);
      * Etot within 1e-10 relative of the oracle's for EVERY atom (the energy is stationary);
      * integer occupations bit-exact, fractional ones (two-fold blocks) to 1e-4;
      * atoms the oracle converges cleanly (monotone, <= 45 iterations) converge here, in pass 1;
      * atoms with a fractional two-fold block sit on the noise floor in both implementations (the reference's
        stopping criterion bounces in 1e-10 ... 1e-6 there, run.jl:33-37): whether 1e-10 is met within maxiter is a
        coin flip of rounding noise, so for them the status is reported, not asserted."""
    import atomickohnsham_b200 as aks
    from bench import build_disc
    from atomickohnsham_b200.sweep import periodic_table_models, solve_sweep
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "c5_sweep.json")))
    disc = build_disc(0)
    models, lhs, nhs = periodic_table_models(range(1, 87))
    p1, p2 = g["pass1_alg"], g["pass2_alg"]
    sols = solve_sweep(models, disc, lh=lhs, nh=nhs, maxiter=p1["maxiter"],
                       alg=aks.ODA(scftol=p1["scftol"], tinit=p1["tinit"], frozen_t=p1["frozen_t"],
                                   aufbau=aks.OptimizedAufbau(tol=p1["degen_tol"])),
                       remedy_alg=aks.ODA(scftol=p2["scftol"], tinit=p2["tinit"], frozen_t=p2["frozen_t"],
                                          aufbau=aks.OptimizedAufbau(tol=p2["degen_tol"])),
                       remedy_maxiter=p2["maxiter"], want_arrays=False)
    report, worst_dE = [], 0.0
    for Z in range(1, 87):
        sol, npass = sols[Z - 1]
        a = g["atoms"][str(Z)]
        ref, ref_pass = _c5_final(a)
        L, nh = lhs[Z - 1] + 1, nhs[Z - 1]
        n_g = sol.n[:L, :nh, 0]
        n_o, e_o = np.array(ref["n"]), np.array(ref["eps"])
        dE = abs(sol.energies.Etot - ref["Etot"]) / abs(ref["Etot"])
        worst_dE = max(worst_dE, dE)
        report.append((Z, npass, sol.success, sol.niter, sol.stopping_criteria, ref_pass, ref["status"], ref["niter"], ref["stop"], dE))
        # a defined state, never an error; on the noise floor at worst
        assert sol.success in ("SUCCESS", "MAXITERS"), (Z, sol.success)
        assert sol.stopping_criteria < 1e-5 and ref["stop"] < 1e-5, (Z, sol.stopping_criteria, ref["stop"])
        assert abs(n_g.sum() - Z) < 1e-9, Z
        assert dE <= 1e-10, (Z, sol.energies.Etot, ref["Etot"], dE)                          # NORTH-STAR TOLERANCE
        integer = (n_o == np.round(n_o)) & (n_g == np.round(n_g))
        assert np.array_equal(n_g[integer], n_o[integer]), Z                                  # bit-exact occupations
        assert np.allclose(n_g, n_o, atol=1e-4), (Z, n_g[n_g != n_o], n_o[n_g != n_o])
        occ = n_o > 1e-6
        # against LAPACK's eigenvalues (5e-9 absolute, test_lapack_error_against_referee) of a density converged to
        # 1e-10 ... 1e-6; the 1e-10 relative assertion on orbital energies is test_orbital_energies_against_dd_referee
        tol_e = 5e-8 * max(1.0, np.max(np.abs(e_o[occ]))) + 10.0 * max(ref["stop"], sol.stopping_criteria)
        assert np.max(np.abs(sol.eps[:L, :nh, 0][occ] - e_o[occ])) < tol_e, Z
        clean = a["pass1"]["status"] == "SUCCESS" and a["pass1"]["niter"] <= 45 and np.array_equal(n_o, np.round(n_o))
        if clean:
            assert npass == 1 and sol.success == "SUCCESS", (Z, npass, sol.success, sol.niter, sol.stopping_criteria)
    conv = sum(1 for r in report if r[2] == "SUCCESS")
    conv_o = sum(1 for a in g["atoms"].values() if _c5_final(a)[0]["status"] == "SUCCESS")
    print(f"[C5] 86/86 in a defined state; converged {conv}/86 here, {conv_o}/86 in the oracle; "
          f"worst |dEtot/Etot| over all 86 atoms {worst_dE:.1e}")
    for r in report:
        if r[2] != "SUCCESS" or r[6] != "SUCCESS":
            print("   Z=%d pass %d %s niter %d stop %.1e | oracle pass %d %s niter %d stop %.1e | dE/E %.1e" % r)


_RANK_SCRIPT = r"""
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
import atomickohnsham_b200 as aks
from atomickohnsham_b200.sweep import periodic_table_models, problem_cost, partition, solve_sweep
from helpers import package_tables
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
dev = rank % torch.cuda.device_count()
zs = list(range(1, 19))
models, lhs, nhs = periodic_table_models(zs)
basis, ops, quad = package_tables(["exp", 0, 300, 25, 1.5], 10, 400)
disc = aks.KSEDiscretization(basis, None, lh=max(lhs), nh=max(nhs), nspin=1, fem_integration_method=quad, femops=ops, device=dev)
mine = partition([problem_cost(z) for z in zs], world)[rank]
sols = solve_sweep([models[i] for i in mine], disc, lh=[lhs[i] for i in mine], nh=[nhs[i] for i in mine])
recs = [(zs[i], s.success, s.niter, s.energies.Etot, s.n.tolist()) for i, (s, _) in zip(mine, sols)]
allrecs = [None] * world
dist.all_gather_object(allrecs, recs)
if rank == 0:
    json.dump(sorted(sum(allrecs, [])), open({out!r}, "w"))
dist.destroy_process_group()
"""


def test_two_rank_partitioned_sweep_matches_one_batch(tmp_path):
    """The multi-GPU path of SURVEY 8(e): LPT partition, one PROCESS per rank with its own real batch, no data-path
    collective, host gather of the records -- equal, bit for bit, to all problems in one batch.  (Both ranks share
    cuda:0 when the box has one GPU.)"""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.sweep import periodic_table_models, solve_sweep
    from helpers import package_tables
    out = str(tmp_path / "gather.json")
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], check=True, env=env,
                   timeout=600)
    got = json.load(open(out))
    zs = list(range(1, 19))
    models, lhs, nhs = periodic_table_models(zs)
    basis, ops, quad = package_tables(["exp", 0, 300, 25, 1.5], 10, 400)
    disc = aks.KSEDiscretization(basis, None, lh=max(lhs), nh=max(nhs), nspin=1, fem_integration_method=quad, femops=ops)
    sols = solve_sweep(models, disc, lh=lhs, nh=nhs)
    want = sorted((z, s.success, s.niter, s.energies.Etot, s.n.tolist()) for z, (s, _) in zip(zs, sols))
    assert [tuple(g[:4]) for g in got] == [w[:4] for w in want]
    assert [g[4] for g in got] == [w[4] for w in want]


def test_groundstate_over_several_devices():
    """`groundstate(models, ...; devices = [...])`: one host thread and one batch per GPU; same solutions, same order."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.sweep import periodic_table_models
    from helpers import package_tables
    zs = list(range(1, 19))
    models, lhs, nhs = periodic_table_models(zs)
    basis, ops, quad = package_tables(["exp", 0, 300, 25, 1.5], 10, 400)
    disc = aks.KSEDiscretization(basis, None, lh=max(lhs), nh=max(nhs), nspin=1, fem_integration_method=quad, femops=ops)
    alg = aks.ODA(scftol=1e-10, tinit=0.6)
    one = aks.groundstate(models, disc, alg, lh=lhs, nh=nhs, want_arrays=False)
    two = aks.groundstate(models, disc, alg, lh=lhs, nh=nhs, want_arrays=False, devices=[0, 1])
    assert [s.energies.Etot for s in one] == [s.energies.Etot for s in two]
    assert [s.niter for s in one] == [s.niter for s in two]


def test_maximum_ionization_against_the_oracle_bisection():
    """SURVEY 8(f) rank 2 (examples/periodic table/process.jl:12-86): the batched N-bisection front end against the SAME
    bisection driven by the oracle, one atom and one trial at a time as the reference script does -- same bracket
    [Z, Z + 1.2], same stability test (HOMO <= 0 and SCF converged to 1e-6 within 50 iterations, process.jl:77-82), same
    ODA(0.4) / OptimizedAufbau(1e-2).  Every trial decision (bound / not bound) and the final N_max agree; the HOMO of
    every converged trial agrees to 5e-6 Ha (runs that stop at scftol = 1e-6 end at different iterates near the same
    fixed point: the levels carry an error of the order of the stopping criterion; measured worst 6e-7)."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.sweep import maximum_ionization
    from helpers import make_oracle, package_tables
    mesh = ["exp", 0, 300, 30, 1.3]
    basis, ops, quad = package_tables(mesh, 10, 1000)
    disc = aks.KSEDiscretization(basis, None, lh=1, nh=6, nspin=1, fem_integration_method=quad, femops=ops)
    zs = [1, 2, 3, 9]
    precision = 1
    trace = {}
    got = maximum_ionization(zs, disc, precision=precision, lh=1, nh=6, trace=trace)
    want = {}
    for z in zs:
        lo, hi = float(z), z + 1.2
        k = 0
        while hi - lo > 10.0 ** (-precision):
            mid = (hi + lo) / 2
            o = make_oracle({"Z": z, "N": mid, "lh": 1, "nh": 6, "ex": "lda_x", "ec": "lda_c_pw", "mesh": mesh, "p": 10,
                             "scftol": 1e-6, "tinit": 0.4, "degen_tol": 1e-2})
            st, _ = o.groundstate(maxiter=50)
            homo = float(st.eps[st.n != 0].max())
            stable = homo <= 0.0 and st.stop < 1e-6
            g_mid, g_homo, g_stable = trace[z][k]
            assert g_mid == mid and g_stable == stable, (z, k, mid, homo, g_homo, st.stop)
            if st.stop < 1e-6:
                assert abs(g_homo - homo) < 5e-6, (z, k, g_homo, homo)
            lo, hi = (mid, hi) if stable else (lo, mid)
            k += 1
        want[z] = round(mid, precision)   # process.jl:85
    assert got == want, (got, want)
    print("[maximum ionization vs oracle bisection]", got)


def test_postprocessing_on_the_device_matches_the_host_mirror():
    """SURVEY 8(f) rank 4 (src/solution/postprocess.jl:34-387): `aks_eval_grid` evaluates orbitals, density, Hartree and XC
    potentials and the effective potential on arbitrary radial points straight from the device state (D, U, W), against
    the host mirror `postprocess.py` fed with the arrays `aks_get_state` returns -- 1e-11 relative to the largest value of
    each curve (the device evaluates the generators by the double recurrence, the host in extended precision; D enters
    with cancellation between its entries).  One closed-shell LDA atom on the order-20 discretization and one LSDA atom."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200 import postprocess as pp
    from atomickohnsham_b200.solver import Batch, _solution
    from helpers import package_tables
    X = np.concatenate([np.geomspace(1e-4, 250.0, 301), [0.5, 1.0, 299.0, 300.0, 301.0]])

    def check(name, dev, host):
        ok = np.isfinite(host)
        scale = np.max(np.abs(host[ok]))
        assert np.max(np.abs(dev[ok] - host[ok])) <= 1e-11 * scale, (name, np.max(np.abs(dev[ok] - host[ok])) / scale)

    # (a) Neon, LDA, order 20
    basis, ops, quad = package_tables(["exp", 0, 300, 41, 1.5], 20, 1000)
    disc = aks.KSEDiscretization(basis, None, lh=1, nh=4, nspin=1, fem_integration_method=quad, femops=ops)
    bt = Batch([aks.LDA(Z=10, N=10)], disc, aks.ODA(scftol=1e-10, tinit=0.6))
    final, _ = bt.run(60)
    sol = _solution(bt, 0, final[0])
    for (n, l) in ((1, 0), (2, 0), (2, 1)):
        check(f"orbital {n}{l}", bt.eval_orbital(0, n, l, X), pp.eval_orbital(sol, n, l, X))
        check(f"orbital {n}{l} h10", bt.eval_orbital(0, n, l, X, h10=True), pp.eval_orbital(sol, n, l, X, h10=True))
    check("density", bt.eval_density(0, X), pp.eval_density(sol, X))
    check("hartree", bt.eval_hartree(0, X), pp.eval_hartree(sol, X))
    check("vxc", bt.eval_vxc(0, X), pp.eval_vxc(sol, X))
    check("veff l=1", bt.eval_effective_potential(0, 1, X), pp.eval_effective_potential(sol, 1, X))
    assert bt.eval_density(0, np.array([301.0]))[0] == 0.0   # outside the mesh
    # 4 pi int rho r^2 dr = N on a fine grid (trapezoid on the log grid)
    r = np.geomspace(1e-6, 300.0, 20001)
    rho = bt.eval_density(0, r)
    assert abs(np.trapezoid(4 * np.pi * rho * r ** 3, np.log(r)) - 10.0) < 1e-5
    bt.close()
    # (b) Lithium, LSDA, order 10
    basis, ops, quad = package_tables(["exp", 0, 100, 30, 1.4], 10, 200)
    disc = aks.KSEDiscretization(basis, None, lh=0, nh=3, nspin=2, fem_integration_method=quad, femops=ops)
    bt = Batch([aks.KSEModel(Z=3, N=3, ex=aks.BuiltinFunctional("lda_x", 2), ec=aks.BuiltinFunctional("lda_c_pw", 2))],
               disc, aks.ODA(scftol=1e-9, tinit=0.6))
    final, _ = bt.run(80)
    sol = _solution(bt, 0, final[0])
    Xs = X[X <= 100.0]
    for sg in (1, 2):
        check(f"orbital 1s sigma {sg}", bt.eval_orbital(0, 1, 0, Xs, sigma=sg), pp.eval_orbital(sol, 1, 0, Xs, sigma=sg))
        check(f"density sigma {sg}", bt.eval_density(0, Xs, sigma=sg), pp.eval_density(sol, Xs, sigma=sg))
        check(f"vxc sigma {sg}", bt.eval_vxc(0, Xs, sigma=sg), pp.eval_vxc(sol, Xs, sigma=sg))
        check(f"veff sigma {sg}", bt.eval_effective_potential(0, 0, Xs, sigma=sg), pp.eval_effective_potential(sol, 0, Xs, sigma=sg))
    check("density sum", bt.eval_density(0, Xs), pp.eval_density(sol, Xs))
    with pytest.raises(ValueError):
        bt.eval_vxc(0, Xs)
    bt.close()


def test_iteration_counts_lie_within_the_oracles_own_jitter():
    """SURVEY 8(c) protocol (iii): with the default (adaptive) ODA the iteration count is decided by rounding noise -- in the
    reference's algorithm itself.  `tools/oracle_jitter.py` measured that on the oracle: every case once on the package's
    tables and 8 times with the quadrature weights perturbed by at most 1 ulp (`tests/golden/iteration_count_jitter.json`:
    spreads of 3-5 iterations for the adaptive line search, none for `frozen_t` or without a functional, Etot spread
    <= 1.3e-15).  The device's count must lie within the oracle's own range widened by one iteration on either side, and
    must be EQUAL where the oracle shows no jitter; Etot agrees to 1e-12 in every case."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    from atomickohnsham_b200.solver import Batch
    from gpu_diag import make_gpu
    from helpers import load_golden
    fx = load_golden("iteration_count_jitter.json")
    report = {}
    for name, c in fx["cases"].items():
        model, disc, alg = make_gpu(c["setup"])
        bt = Batch([model], disc, alg)
        final, _ = bt.run(200)
        bt.close()
        f = final[0]
        lo, hi = c["range"]
        report[name] = (f["niter"], lo, hi)
        assert f["status"] == 1, (name, f)
        assert abs(f["Etot"] - c["Etot"]) <= 1e-12 * abs(c["Etot"]), (name, f["Etot"], c["Etot"])
        if lo == hi:
            assert f["niter"] == lo, (name, f["niter"], lo)
        else:
            assert lo - 1 <= f["niter"] <= hi + 1, (name, f["niter"], lo, hi)
    print("[iterations: device vs the oracle's range under 1-ulp perturbations]", report)
