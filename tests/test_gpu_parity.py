"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical
tables and identical inputs.  Protocol (DESIGN.md "parity"):
  step level   -- identical input D -> Hartree / Vxc matrices, energies, eigenpairs, one full step
  trajectory   -- whole SCF runs: Etot, occupations, orbital energies, iteration count
Tolerances are written next to each assertion.  LAPACK's dense generalized solver (what the
oracle and the reference call) loses ~eps*||M0^-1 H|| absolute on eigenvalues; the GPU pairs are
checked by their own residual and against LAPACK with that allowance."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
pytestmark = pytest.mark.gpu

from helpers import make_oracle  # noqa: E402

CASES = {
    "h_slater_c1": {"Z": 1, "N": 1, "mesh": ["exp", 0, 300, 20, 1.5], "p": 10, "lh": 0, "nh": 3, "ex": "lda_x", "ec": None, "scftol": 1e-11},
    "o_lda": {"Z": 8, "N": 8, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10},
    "ar_lda_c2": {"Z": 18, "N": 18, "mesh": ["exp", 0, 300, 41, 1.5], "p": 20, "lh": 2, "nh": 10, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10},
    "sc_rhf": {"Z": 21, "N": 21, "mesh": ["exp", 0, 2000, 40, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": None, "ec": None, "scftol": 1e-10},
    "h_lsda": {"Z": 1, "N": 1, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 0, "nh": 2, "ex": "lda_x", "ec": "lda_c_pw", "nspin": 2, "scftol": 1e-10, "max_degen": 1},
    # config C3: scandium LSDA on the order-20 basis (test/scandium.jl discretization, XC switched on)
    "sc_lsda_c3": {"Z": 21, "N": 21, "mesh": ["exp", 0, 2000, 40, 1.2], "p": 20, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "nspin": 2, "scftol": 1e-8},
    "na_slater_q2000": {"Z": 11, "N": 11, "mesh": ["exp", 0, 500, 20, 1.2], "p": 10, "lh": 2, "nh": 10, "npoints": 2000, "ex": "lda_x", "ec": None, "scftol": 1e-9, "degen_tol": 0.1},
}


def gpu_batch(setup, nprob=1):
    from gpu_diag import make_gpu
    from atomickohnsham_b200.solver import Batch
    model, disc, alg = make_gpu(setup)
    return Batch([model] * nprob, disc, alg), model, disc, alg


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.mark.parametrize("name", ["h_slater_c1", "o_lda", "ar_lda_c2", "sc_rhf", "h_lsda"])
def test_step_level_parity_identical_D(name):
    setup = CASES[name]
    o = make_oracle(setup)
    st = o.initial_state()
    for _ in range(3):
        o.scf_step(st)
    bt, *_ = gpu_batch(setup)
    e = st.energies
    bt.set_density(0, st.D, [e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc], niter=st.niter)
    Har_o, Vxc_o, _ = o.hamiltonian_parts(st.D)
    assert rel(bt.assemble_hartree(0), Har_o) < 1e-13            # assemble_hartree!
    if o.hartree != 0:
        # postcomputations! (oda/loop.jl:51-54): W = A \\ (F:D) of the current D, directly against the oracle's CHOLMOD-style solve
        W_o = o.assemble_hartree(st.D)[1]
        assert rel(bt.state(0, want_D=False, want_U=False)["W"], W_o) < 1e-12
    if o.xc.active:
        assert rel(bt.assemble_vxc(0), Vxc_o) < 1e-13            # assemble_exc!
        assert abs(bt.exc_energy(0) - o.exc_energy(st.D)) < 1e-13 * abs(o.exc_energy(st.D))
    assert abs(bt.hartree_energy(0) - o.hartree_energy(st.D)) < 1e-13 * abs(o.hartree_energy(st.D))
    eps_g, U_g, res = bt.eig_lowest(0)                            # find_orbital!
    U_o, eps_o = o.find_orbital(Har_o, Vxc_o)
    scale = max(1.0, float(np.max(np.abs(eps_o))))
    # eigenvector error = reduction error / relative gap: residual allowance 1e-10
    assert res.max() < 1e-10 * scale, "GPU eigen-residual"
    # LAPACK allowance: eps * ||M0^-1 H|| ~ 1e-8 absolute at order 20
    assert np.max(np.abs(eps_g - eps_o)) < 5e-8
    M0 = o.tb.M0
    for s in range(o.nspin):
        for l in range(o.L):
            G = U_g[:, :, l, s].T @ M0 @ U_g[:, :, l, s]
            assert np.max(np.abs(G - np.eye(G.shape[0]))) < 1e-12     # U^T M0 U = I
            ov = np.abs(np.diag(U_g[:, :, l, s].T @ M0 @ U_o[:, :, l, s]))
            assert ov.min() > 1 - 1e-8                                 # same eigenvectors
    # one full scf_step! from the identical state
    detail = {}
    o.scf_step(st, detail=detail)
    rec = bt.step()[0]
    if np.array_equal(detail["n"], np.round(detail["n"])):
        # density!(D, U, n): trial density of this step (skipped where a degenerate split depends on t);
        # the eigenvectors carry LAPACK's error on the oracle side
        Dg = bt.density(0)
        assert rel(Dg, np.asarray(detail["Dtrial"]).reshape(Dg.shape)) < 1e-7
    eo = st.energies
    for k, v in (("Etot", eo.Etot), ("Ekin", eo.Ekin), ("Ecou", eo.Ecou), ("Ehar", eo.Ehar), ("Eexc", eo.Eexc)):
        tol = 1e-10 if k == "Etot" else 5e-9    # components follow t, which sits on a flat minimum (dt ~ 1e-8)
        assert abs(rec[k] - v) <= tol * max(1.0, abs(v)), (k, rec[k], v)
    sg = bt.state(0)
    assert np.allclose(sg["n"], st.n, atol=1e-6)
    bt.close()


def test_c3_scandium_lsda_steps():
    """Config C3 (Sc, LSDA, order 20): first SCF steps against the oracle, step by step from identical
    states.  (The full run ends in a fractional 3d/4s occupation where ODA plateaus, see DESIGN.md.)"""
    setup = CASES["sc_lsda_c3"]
    o = make_oracle(setup)
    st = o.initial_state()
    bt, *_ = gpu_batch(setup)
    for it in range(3):
        e = st.energies
        bt.set_density(0, st.D, [e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc], niter=st.niter)
        o.scf_step(st)
        rec = bt.step()[0]
        eo = st.energies
        if it == 0:
            continue   # bare-Coulomb degeneracies at iteration 0 (see test_trajectory_parity)
        assert abs(rec["Etot"] - eo.Etot) <= 1e-10 * abs(eo.Etot), (it, rec["Etot"], eo.Etot)
        sg = bt.state(0, want_D=False, want_U=False)
        # fractional occupations of the degenerate block follow t, which sits on a flat minimum of E(t).
        # While D_up == D_down the up and down 3d levels are EXACTLY degenerate and E(split) has two
        # mirror-image minima (all 3d electrons up, or all down): which one Brent lands in is decided by
        # rounding, in the reference too.  Accept the spin-mirrored occupation.
        same = np.allclose(sg["n"], st.n, atol=2e-4)
        mirrored = np.allclose(sg["n"][..., ::-1], st.n, atol=2e-4)
        assert same or mirrored
        occ = st.n > 1e-6
        eps_g = sg["eps"] if same else sg["eps"][..., ::-1]
        assert np.max(np.abs(eps_g[occ] - st.eps[occ])) < 5e-8
    bt.close()


@pytest.mark.parametrize("name", [k for k in CASES if k != "sc_lsda_c3"])
def test_trajectory_parity(name):
    setup = CASES[name]
    so, ho = make_oracle(setup).groundstate()
    bt, *_ = gpu_batch(setup)
    final, hist = bt.run(100, history=True)
    f = final[0]
    assert f["status"] == 1
    # total energy within 1e-10 relative (north star)
    assert abs(f["Etot"] - so.energies.Etot) <= 1e-10 * abs(so.energies.Etot)
    # iteration 0 has no line search: every energy to 1e-10 relative (sum n*eps carries LAPACK's error).
    # Not for Z >= 11: the first Hamiltonian is bare-Coulomb, its n = 3 shell (3s = 3p = 3d) is EXACTLY
    # degenerate, and which two levels enter the aufbau's two-fold block is decided by rounding noise
    # of the eigensolver (LAPACK's 1e-9 in the oracle/reference, 1e-14 here) -- DESIGN.md "parity".
    if setup["Z"] <= 10:
        for k in ("Etot", "Ekin", "Ecou", "Ehar", "Eexc"):
            assert abs(hist[0][0][k] - ho[0][k]) <= 1e-10 * max(1.0, abs(ho[0][k])), k
    # iteration count: reported against the oracle; the Brent tail is chaotic
    assert abs(f["niter"] - so.niter) <= 8
    sg = bt.state(0, want_D=False, want_U=False)
    occ = so.n != 0
    # aufbau ordering / integer occupations bit-exact, fractional ones to the line-search accuracy
    assert np.array_equal(sg["n"] != 0, occ)
    integer = occ & (so.n == np.round(so.n))
    assert np.array_equal(sg["n"][integer], so.n[integer])
    assert np.allclose(sg["n"], so.n, atol=5e-5)
    # occupied orbital energies: first order in the density error left at the stopping rule (1e-7 .. 3e-7
    # for the open 3d shell of Sc, whose two-fold block sits on a flat E(t)) + LAPACK allowance in the oracle
    assert np.max(np.abs(sg["eps"][occ] - so.eps[occ]) / np.maximum(1.0, np.abs(so.eps[occ]))) < 1e-6
    bt.close()


def test_reference_goldens_through_the_c_abi():
    """The CUDA path against the reference's own committed values (not the oracle)."""
    from helpers import load_golden
    g = load_golden("scf_goldens.json")
    sc = g["scandium_rhf"]
    bt, *_ = gpu_batch(sc["setup"])
    final, _ = bt.run(100)
    assert abs(final[0]["Etot"] - sc["Etot"]) < sc["tol"]                 # test/scandium.jl:59, 1e-9
    n = bt.state(0, want_D=False, want_U=False)["n"]
    assert abs(n[2, 0, 0] / 5 - sc["n3d_over_5"]) < sc["n3d_tol"]
    bt.close()
    hs = g["hydrogen_slater"]
    bt, *_ = gpu_batch(hs["setup"])
    final, _ = bt.run(100)
    for k in ("Ekin", "Ecou", "Ehar", "Eexc", "Etot"):
        assert abs(final[0][k] - hs[k]) < hs["tol"]                      # test/hydrogen.jl:103-107
    bt.close()
    tr = load_golden("trajectories.json")["sodium_slater"]
    bt, *_ = gpu_batch(tr["setup"])
    final, hist = bt.run(100, history=True)
    # Sodium's first Hamiltonian is bare-Coulomb: 3s = 3p = 3d exactly, so which two levels form the
    # aufbau block at iteration 0 is decided by eigensolver rounding noise (see test_trajectory_parity);
    # the committed trajectory is therefore pinned through its converged values only.
    assert abs(final[0]["Etot"] - tr["results"]["Etot"]) < 1e-11 * abs(tr["results"]["Etot"])
    for k in ("Ekin", "Ecou", "Ehar", "Eexc"):
        assert abs(final[0][k] - tr["results"][k]) < 5e-7 * abs(tr["results"][k])   # scftol = 1e-9 run
    n = bt.state(0, want_D=False, want_U=False)["n"]
    assert n[0, 0, 0] == 2.0 and n[0, 1, 0] == 2.0 and n[1, 0, 0] == 6.0
    assert abs(n[0, 2, 0] - 0.999999999991612) < 1e-6 and abs(n[0, 2, 0] + n[1, 1, 0] - 1.0) < 1e-12
    assert any(h["t"] == 0.999999999991612 for h in hist[0])                 # Brent's boundary value
    bt.close()


@pytest.mark.parametrize("name", ["hydrogen_rhf", "hydrogen_slater", "hydrogen_bare", "readme_hydrogen"])
def test_hydrogen_goldens_through_the_c_abi(name):
    """test/hydrogen.jl:55-116,138-171 and README.md:57-84 against the CUDA path (not the oracle)."""
    from helpers import load_golden
    g = load_golden("scf_goldens.json")[name]
    bt, *_ = gpu_batch(g["setup"])
    final, _ = bt.run(g["setup"].get("maxiter", 100))
    f = final[0]
    assert f["status"] == 1 and f["stop"] < g["setup"]["scftol"]
    for k in ("Ekin", "Ecou", "Ehar", "Eexc", "Etot"):
        if k in g:
            assert abs(f[k] - g[k]) < g["tol"], k
    st = bt.state(0, want_D=False, want_U=False)
    if "eps_1s" in g:
        assert abs(st["eps"][0, 0, 0] - g["eps_1s"]) < g.get("eps_tol", 1e-9)
        assert st["n"][0, 0, 0] == 1.0
    if "eps" in g:
        for key, val in g["eps"].items():
            l, k = map(int, key.split(","))
            assert abs(st["eps"][l, k, 0] - val) < g["tol"]
    if "niter" in g:
        assert abs(f["niter"] - g["niter"]) <= 4      # chaotic Brent tail
    bt.close()


@pytest.mark.parametrize("name", ["oxygen_lda", "oxygen_x", "oxygen_none"])
def test_oxygen_reference_results_through_the_c_abi(name):
    """examples/functional comparison/{lda,lda_x,none}/results.txt: final energies, levels and occupations."""
    from helpers import load_golden
    g = load_golden("trajectories.json")[name]
    bt, *_ = gpu_batch(g["setup"])
    final, _ = bt.run(100)
    r = g["results"]
    assert final[0]["status"] == 1
    assert abs(final[0]["Etot"] - r["Etot"]) < 1e-9 * abs(r["Etot"])
    assert abs(final[0]["niter"] - r["niter"]) <= 6
    st = bt.state(0, want_D=False, want_U=False)
    gold = {lab: (e, n) for lab, e, n in r["orbitals"]}
    for lab, (l, k) in {"1s": (0, 0), "2s": (0, 1), "2p": (1, 0)}.items():
        assert abs(st["eps"][l, k, 0] - gold[lab][0]) < 2e-7 * max(1.0, abs(gold[lab][0])), lab
        assert st["n"][l, k, 0] == gold[lab][1]
    bt.close()


def test_batch_is_deterministic_and_problem_independent():
    """A problem's trajectory must not depend on its batch neighbours; runs are bit-reproducible."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from gpu_diag import make_gpu
    setup = CASES["o_lda"]
    _, disc, alg = make_gpu(setup)
    models = [aks.LDA(Z=Z, N=Z) for Z in (8, 6, 10, 8)]
    b1 = Batch(models, disc, alg)
    f1, _ = b1.run(100)
    b2 = Batch([models[0]], disc, alg)
    f2, _ = b2.run(100)
    assert f1[0] == f1[3] == f2[0]
    # aks_get_state_all == aks_get_state of every problem
    allst = b1.states()
    for i in range(4):
        one = b1.state(i, want_D=False, want_U=False)
        assert all(np.array_equal(allst[i][k], one[k]) for k in ("eps", "n", "W"))
    b1.reset()
    f3, _ = b1.run(100)
    assert f3 == f1
    b1.close()
    b2.close()


@pytest.mark.parametrize("kind", ["smeared", "frozen"])
def test_smeared_and_frozen_aufbau_parity(kind):
    """SmearedAufbau / FrozenAufbau (aufbau/smeared.jl, aufbau/frozen.jl) through the C ABI against the
    oracle: oxygen LDA, eight SCF steps, each from the oracle's state."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from gpu_diag import make_gpu
    setup = dict(CASES["o_lda"])
    if kind == "smeared":
        setup.update(aufbau_kind="smeared", temp=0.05)
        auf = aks.SmearedAufbau(Temp=0.05)
    else:
        auf = aks.FrozenAufbau({"1s": 2.0, "2s": 2.0, "2p": 3.0, "3s": 1.0})
        setup.update(aufbau_kind="frozen", nfix=auf.nfix(setup["lh"], setup["nh"], 1))
    o = make_oracle(setup)
    model, disc, _ = make_gpu(setup)
    bt = Batch([model], disc, aks.ODA(scftol=0.0, tinit=0.6, aufbau=auf))
    st = o.initial_state()
    for it in range(8):
        # every step starts from the ORACLE's state (aks_set_density): Brent's t sits on a flat minimum of E(t) and is
        # only determined to sqrt(eps) ~ 1e-8 by function values, so two free-running trajectories drift apart by
        # that much per step in D (not in the converged result); step-by-step parity is exact to rounding
        e = st.energies
        if it > 0:
            bt.set_density(0, st.D, [e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc], niter=st.niter)
        o.scf_step(st)
        rec = bt.step()[0]
        eo = st.energies
        assert abs(rec["Etot"] - eo.Etot) <= 1e-10 * abs(eo.Etot), (it, rec["Etot"], eo.Etot)
        assert abs(rec["t"] - st.t) <= 1e-5      # flat minimum of E(t): t is known to sqrt(eps)
    sg = bt.state(0, want_D=False, want_U=False)
    assert np.allclose(sg["n"], st.n, atol=1e-8)
    assert abs(sg["n"].sum() - 8.0) < 1e-8
    occ = st.n > 1e-8
    assert np.max(np.abs(sg["eps"][occ] - st.eps[occ])) < 5e-8
    bt.close()


def test_maximum_ionization_driver():
    """Batched N-bisection front end (examples/periodic table/process.jl:12-86): the returned N is stable
    (bound HOMO, converged) just below and unstable just above, within the bisection tolerance."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from atomickohnsham_b200.sweep import maximum_ionization
    from helpers import package_tables
    basis, ops, quad = package_tables(["exp", 0, 300, 30, 1.3], 10, 1000)
    disc = aks.KSEDiscretization(basis, None, lh=1, nh=6, nspin=1, fem_integration_method=quad, femops=ops)
    zs = [1, 2, 3, 9]
    nmax = maximum_ionization(zs, disc)
    alg = aks.ODA(scftol=1e-6, tinit=0.4, aufbau=aks.OptimizedAufbau(tol=1e-2))
    for z in zs:
        assert z <= nmax[z] <= z + 1.2
    # LDA binds only a fraction of the extra electron of H-, more of F-
    assert nmax[1] < 2.0 and nmax[9] > nmax[1] - 1 + 9 - 0.5
    below = [aks.LDA(Z=z, N=max(float(z), nmax[z] - 0.03)) for z in zs]
    bt = Batch(below, disc, alg, lh=[1] * 4, nh=[min(max(z, 2), 6) for z in zs])
    final, _ = bt.run(50)
    for i, z in enumerate(zs):
        st = bt.state(i, want_D=False, want_U=False)
        assert final[i]["status"] == 1 and st["eps"][st["n"] != 0].max() <= 0.0, z
    bt.close()


@pytest.mark.parametrize("case", ["frozen_t", "cation", "no_hartree"])
def test_deterministic_trajectories(case):
    """SURVEY 8(c) protocol (ii): without the Brent line search (frozen t), or where it is well conditioned,
    the run is deterministic -- same iteration count, Etot of every iteration within 1e-10."""
    from atomickohnsham_b200.solver import Batch
    from gpu_diag import make_gpu
    setup = dict(CASES["o_lda"])
    if case == "frozen_t":
        setup.update(frozen_t=True, tinit=0.5, scftol=1e-9)
    elif case == "cation":
        setup.update(Z=11, N=10, scftol=1e-9)             # Na+: closed shell
    else:
        setup.update(Z=4, N=4, hartree=0, ex=None, ec=None, scftol=1e-9)   # non-interacting: converges in 2 steps
    o = make_oracle(setup)
    so, ho = o.groundstate()
    model, disc, alg = make_gpu(setup)
    bt = Batch([model], disc, alg)
    final, hist = bt.run(100, history=True)
    f = final[0]
    assert f["status"] == 1 and so.stop < setup["scftol"]
    assert abs(f["niter"] - so.niter) <= (0 if case != "cation" else 2), (f["niter"], so.niter)
    for it in range(min(f["niter"], so.niter, 12)):
        assert abs(hist[0][it]["Etot"] - ho[it]["Etot"]) <= 1e-10 * abs(ho[it]["Etot"]), (it, hist[0][it]["Etot"], ho[it]["Etot"])
        if case == "frozen_t":
            assert hist[0][it]["t"] == ho[it]["t"] == 0.5
    assert abs(f["Etot"] - so.energies.Etot) <= 1e-10 * abs(so.energies.Etot)
    sg = bt.state(0, want_D=False, want_U=False)
    occ = so.n != 0
    assert np.array_equal(sg["n"][occ], so.n[occ])
    assert np.max(np.abs(sg["eps"][occ] - so.eps[occ])) < 5e-8
    bt.close()


@pytest.mark.parametrize("p,ncell,npoints,lh,nspin", [(3, 12, 100, 1, 1), (12, 25, 200, 1, 1), (24, 10, 301, 2, 1),
                                                      (29, 6, 128, 0, 1), (12, 25, 200, 1, 2), (24, 10, 300, 1, 2)])
def test_other_orders_and_grids(p, ncell, npoints, lh, nspin):
    """Every template instantiation and fallback: small / large orders (one or two register tiles per lane,
    k_reduce<8..28>, k_dens_cells<12..32>), node counts that are not a multiple of the TMA chunk, an odd node
    count (no 16-byte rows: plain staging instead of bulk copies), cell counts that do not fill a chunk."""
    from atomickohnsham_b200.solver import Batch
    from gpu_diag import make_gpu
    setup = {"Z": 4, "N": 4, "mesh": ["exp", 0, 60, ncell + 1, 1.3], "p": p, "npoints": npoints, "lh": lh, "nh": 3,
             "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10, "nspin": nspin}
    if nspin == 2:
        setup.update(Z=2, N=2)   # helium: lithium's bare-Coulomb 2s = 2p (x2 spins) start is decided by rounding
    o = make_oracle(setup)
    model, disc, alg = make_gpu(setup)
    bt = Batch([model, model, model], disc, alg)     # three problems: a partly filled 4-problem CTA
    st = o.initial_state()
    for it in range(4):
        o.scf_step(st)
        recs = bt.step()
        assert recs[0] == recs[1] == recs[2]
        assert abs(recs[0]["Etot"] - st.energies.Etot) <= 1e-10 * abs(st.energies.Etot), (it, recs[0]["Etot"], st.energies.Etot)
    sg = bt.state(1)
    M0 = disc.femops.M0
    for e in range(nspin):
        for l in range(lh + 1):
            U = sg["U"][:, :, l, e]
            assert np.max(np.abs(U.T @ M0 @ U - np.eye(3))) < 1e-10
    occ = st.n.reshape(sg["n"].shape) != 0
    assert np.max(np.abs(sg["eps"][occ] - st.eps.reshape(sg["eps"].shape)[occ])) < 1e-7
    bt.close()


def test_postprocessing_and_sanity_checks():
    """The consumers of the path's outputs (postprocess.jl, sanity.jl): hydrogen-like checks with closed forms
    and the reference's SanityChecks on a converged LDA atom."""
    import atomickohnsham_b200 as aks
    from helpers import package_tables
    basis, ops, quad = package_tables(["exp", 0, 60, 31, 1.3], 10, 400)
    disc = aks.KSEDiscretization(basis, None, lh=1, nh=3, nspin=1, fem_integration_method=quad, femops=ops)
    # one electron, no interaction: the exact 1s orbital 2 exp(-r), rho = exp(-2r)/pi
    h = aks.groundstate(aks.KSEModel(Z=1, N=1, hartree=0), disc, aks.ODA(scftol=1e-10, tinit=0.6))
    X = np.array([0.1, 0.5, 1.0, 2.0, 4.0])
    assert abs(h.energies.Etot + 0.5) < 1e-8
    u = aks.eval_orbital(h, "1s", X)                                  # R_1s(r) = 2 exp(-r), sign free
    assert np.max(np.abs(np.abs(u) - 2.0 * np.exp(-X))) < 1e-5
    rho = aks.eval_density(h, X)
    assert np.max(np.abs(rho - np.exp(-2 * X) / np.pi)) < 1e-6
    assert np.allclose(aks.eval_nuclear(h, X), -1.0 / X)
    assert np.allclose(aks.eval_hartree(h, X), 1.0 / 60.0)          # hartree = 0: boundary term only
    # neon, LDA: sanity checks as the reference reports them
    ne = aks.groundstate(aks.LDA(Z=10, N=10), disc, aks.ODA(scftol=1e-10, tinit=0.6))
    sc = aks.sanity_checks(ne)
    assert ne.success == "SUCCESS"
    assert abs(sc.energy_balance) < 1e-9 and abs(sc.electron_count_error) < 1e-12
    assert abs(sc.density_norm_error) < 1e-6 and sc.orthonormality_error < 1e-10
    assert 0.95 < sc.virial_ratio < 1.05
    # V_H -> N/r outside the atom; v_xc < 0; effective potential assembled from its parts
    far = np.array([20.0, 30.0])
    assert np.max(np.abs(aks.eval_hartree(ne, far) - 10.0 / far)) < 1e-6
    assert np.all(aks.eval_vxc(ne, X) < 0)
    veff = aks.eval_effective_potential(ne, 1, X)
    assert np.allclose(veff, -10.0 / X + 1.0 / X ** 2 + aks.eval_hartree(ne, X) + aks.eval_vxc(ne, X))


def test_c_abi_error_paths():
    """Errors never cross the ABI as exceptions: status codes + aks_last_error."""
    import ctypes as C
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200 import _capi
    from atomickohnsham_b200._capi import AKSError
    from atomickohnsham_b200.solver import Batch
    from gpu_diag import make_gpu
    from helpers import package_tables
    setup = dict(CASES["o_lda"])
    model, disc, alg = make_gpu(setup)
    # Double64 (dtype = 1) is built for nspin = 1 (tests/test_gpu_double64.py); spin-polarised Double64 is refused
    small = aks.P1IntLegendreBasis(aks.expmesh(0, 50, 7, s=1.3), ordermax=4)
    dd = aks.KSEDiscretization(small, None, lh=0, nh=2, nspin=2, dtype="double64")
    with pytest.raises(AKSError) as ei:
        dd.handle()
    assert ei.value.code == -6
    with pytest.raises(TypeError):      # Float64 tables cannot be passed off as Double64 ones
        basis, ops, quad = package_tables(setup["mesh"], setup["p"], 1000)
        aks.KSEDiscretization(basis, model, lh=2, nh=5, fem_integration_method=quad, femops=ops, dtype="double64")
    # bad arguments
    lib = _capi.lib
    assert lib.aks_scf_step(None, None) == -1
    bt = Batch([model], disc, alg)
    assert lib.aks_batch_set_oda(bt.h, 1e-10, 2.0, 0, 100, 1e-12, 1e-12, 2, 1e-3, 1) == -1      # tinit > 1
    assert b"bad option" in lib.aks_last_error(disc.ctx.h)
    assert lib.aks_batch_set_aufbau(bt.h, 1, 0.0, None) == -1                                    # Temp = 0
    assert lib.aks_get_state(bt.h, 7, None, None, None, None, None) == -1                        # no such problem
    bt.close()
    with pytest.raises(AKSError):
        Batch([aks.KSEModel(Z=8, N=-1.0)], disc, alg)
    # > 2-fold degeneracy = the reference's error() (optimized.jl:139): bare-Coulomb sodium has 3s = 3p = 3d
    na = aks.KSEModel(Z=11, N=11, ex=aks.NoFunctional(1), ec=aks.NoFunctional(1))
    bt = Batch([na, model], disc, aks.ODA(scftol=1e-10, tinit=0.6, aufbau=aks.OptimizedAufbau(max_degen=3)))
    recs = bt.step()
    assert recs[0]["status"] == -4 and recs[1]["status"] == 0        # the neighbour keeps running
    rc = lib.aks_scf_run(bt.h, 5, None, None)
    assert rc == -4
    bt.close()


def test_eigenpair_window_is_transparent():
    """aks_batch_set_eig_window: solving only the levels the aufbau can reach (certified on the device,
    refilled when the certificate fails) must give the iterations of the all-nh solve, and aks_get_state
    must still return all nh eigenpairs of the last Hamiltonian."""
    from bench import build_disc
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from atomickohnsham_b200.sweep import periodic_table_models
    disc = build_disc(0)
    zs = [1, 3, 8, 11, 18, 21, 24, 26, 29, 36, 46, 57, 64, 79, 86]
    models, lhs, nhs = periodic_table_models(zs)
    alg = aks.ODA(scftol=0.0, tinit=0.6)
    runs = {}
    for margin in (-1, 0, 1, 2):
        bt = Batch(models, disc, alg, lh=lhs, nh=nhs, eig_margin=margin)
        recs = [bt.step() for _ in range(12)]
        states = [bt.state(i, want_D=False) for i in range(len(zs))]
        runs[margin] = (recs, states, bt.eig_refills())
        bt.close()
    assert runs[-1][2].sum() == 0
    M0 = disc.femops.M0
    ref_recs, ref_states, _ = runs[-1]
    for margin in (0, 1, 2):
        recs, states, refills = runs[margin]
        for it in range(12):
            for i in range(len(zs)):
                a, b = recs[it][i], ref_recs[it][i]
                # open d shells (Fe at iteration 5): E(t) is flat over the two-fold degenerate block, so
                # eigenvectors that differ in the last digits move Etot by ~1e-9 relative
                assert abs(a["Etot"] - b["Etot"]) <= 2e-8 * abs(b["Etot"]), (margin, it, zs[i], a["Etot"], b["Etot"])
        for i in range(len(zs)):
            L, nh = lhs[i] + 1, nhs[i]
            ea, eb = states[i]["eps"][:L, :nh], ref_states[i]["eps"][:L, :nh]
            # open d / f shells sit on a flat energy surface (3d <-> 4s transfer): last-digit differences
            # between the runs grow to ~1e-4 in the levels while Etot agrees to 2e-8 (checked above)
            flat = zs[i] in (21, 24, 26, 29, 46, 57, 64, 79)
            assert np.max(np.abs(ea - eb)) < (1e-3 if flat else 1e-7 * max(1.0, np.max(np.abs(eb)))), (margin, zs[i])
            assert np.allclose(states[i]["n"], ref_states[i]["n"], atol=2e-2 if flat else 1e-5)
            for l in range(L):
                U = states[i]["U"][:, :nh, l, 0]
                G = U.T @ M0 @ U
                assert np.max(np.abs(G - np.eye(nh))) < 1e-10, (margin, zs[i], l)


def test_batches_come_and_go_bit_reproducibly():
    """Repeated create / step / destroy cycles (pooled memory, recycled streams) give bit-identical records,
    and two batches alive on one context equal the same problems in a single batch (stream lanes included)."""
    from bench import build_disc
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from atomickohnsham_b200.sweep import periodic_table_models
    disc = build_disc(0)
    models, lhs, nhs = periodic_table_models(range(1, 87))
    alg = aks.ODA(scftol=1e-10, tinit=0.6)
    ref = None
    for _ in range(8):
        bt = Batch(models, disc, alg, lh=lhs, nh=nhs)
        for _ in range(6):
            recs = bt.step()
        sig = tuple(r["Etot"] for r in recs)
        ref = ref or sig
        assert sig == ref
        bt.close()
    b1 = Batch(models[:40], disc, alg, lh=lhs[:40], nh=nhs[:40])
    b2 = Batch(models[40:], disc, alg, lh=lhs[40:], nh=nhs[40:])
    for _ in range(5):
        b1.step_async()
        b2.step_async()
    r12 = b1.step() + b2.step()
    assert tuple(r["Etot"] for r in r12) == ref
    b1.close()
    b2.close()


def test_sweep_full_size_properties():
    """C5 at full size: properties that need no oracle -- electron count, orthonormality,
    energy balance, aufbau monotonicity, virial-like sanity (reference SanityChecks)."""
    from bench import build_disc
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from atomickohnsham_b200.sweep import periodic_table_models
    disc = build_disc(0)
    models, lhs, nhs = periodic_table_models(range(1, 87))
    bt = Batch(models, disc, aks.ODA(scftol=1e-9, tinit=0.6), lh=lhs, nh=nhs)
    final, _ = bt.run(120)
    # Open d/f shells end in a fractional two-fold occupation where ODA's stopping criterion plateaus
    # at 1e-7..1e-2 -- the reference behaves the same (its own example gives iron a dedicated
    # frozen-t algorithm and 800 iterations, examples/atomic density comparison/run.jl:40-47; the
    # oracle reproduces the plateau for Z = 22 and 26).  Everything else must converge.
    dblock = set(range(21, 31)) | set(range(39, 49)) | set(range(57, 81))
    bad = [i + 1 for i, r in enumerate(final) if r["status"] != 1]
    assert all(z in dblock for z in bad), bad
    assert len(bad) <= 30 and all(r["status"] in (1, 2) for r in final)
    M0 = disc.femops.M0
    for i in (0, 17, 25, 53, 85):
        st = bt.state(i)
        r = final[i]
        assert abs(st["n"].sum() - models[i].N) < 1e-10
        assert abs(r["Etot"] - (r["Ekin"] + r["Ecou"] + r["Ehar"] + r["Eexc"])) < 1e-9 * abs(r["Etot"])
        assert abs(np.trace(st["D"][:, :, 0] @ M0) - models[i].N) < 1e-7
        for l in range(lhs[i] + 1):
            U = st["U"][:, :nhs[i], l, 0]
            G = U.T @ M0 @ U
            assert np.max(np.abs(G - np.eye(nhs[i]))) < 1e-10
        occ_e = st["eps"][st["n"] != 0]
        emp_e = st["eps"][:lhs[i] + 1, :nhs[i]][st["n"][:lhs[i] + 1, :nhs[i]] == 0]
        assert occ_e.max() <= emp_e.min() + 1e-3
    # literature LDA(PW92) total energies, Hartree (NIST atomic reference data, LDA): Ar, Kr
    assert abs(final[17]["Etot"] - (-525.946195)) < 0.02
    bt.close()
