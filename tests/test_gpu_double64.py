"""Double64 (T = Double64 of the reference) through the C ABI.

The oracle for this row is the reference itself: `tests/golden/sodium_double64.json` transcribes the
committed log of `examples/basic_example_doublefloat/` (sodium, Hartree only, scftol = 1e-20, 59 iterations,
36 digits per number; made by tools/make_golden_dd.py).  Tolerances (north star: 1e-20 relative in Double64):
Etot of EVERY iteration within 1e-24 relative (measured 1e-30), the energy components within
max(1e-24, 1e-28 / stop), the iteration count within +-3 of the reference's 59 (measured: 59 with the cold-start
eigensolver, 61 / 58 with the warm-started one at 1 / 4 lanes per cell; the last iterations are a coin flip, see below).

Why the components and the orbital energies of the END of the run cannot agree better than ~1e-15 absolute:
the reference's run does not stop because D has converged to 1e-20 but because its line search returns
t = 0 exactly (last line of the log: t = 0.0, stop = 1.6e-30 = |dE| alone).  The quadratic coefficient of the
line search, a = h0 + h1 - 2 h01, is a difference of O(70) numbers of size ||dD||^2; below ||dD|| ~ 1e-13 it
is rounding noise in ANY 106-bit arithmetic, so t (printed in the log) is noise from iteration ~50 on and the
final D is the iterate at ||dD|| ~ 2e-15.  Etot is stationary (error ~ ||dD||^2 = 1e-30), Ekin / Ecou / Ehar
and the orbital energies are first order in that residual.
"""
import mpmath as mp
import numpy as np
import pytest

from helpers import load_golden, make_mesh

pytestmark = pytest.mark.gpu

NAMES = ["stop", "Etot", "Ekin", "Ecou", "Ehar", "Eexc", "t"]


def _mpf(pair):
    return mp.mpf(float(pair[0])) + mp.mpf(float(pair[1]))


def _setup(g, **over):
    from atomickohnsham_b200 import KSEDiscretization, KSEModel, ODA, OptimizedAufbau, P1IntLegendreBasis
    from atomickohnsham_b200.solver import Batch
    s = dict(g["setup"])
    s.update(over)
    basis = P1IntLegendreBasis(make_mesh(s["mesh"]), ordermax=s["p"])
    model = KSEModel(Z=s["Z"], N=s["N"])
    disc = KSEDiscretization(basis, model, lh=s["lh"], nh=s["nh"], dtype="double64")
    alg = ODA(tinit=s["tinit"], scftol=s["scftol"],
              aufbau=OptimizedAufbau(max_degen=s["max_degen"], tol=s["degen_tol"]))
    return Batch([model], disc, alg), disc, model


def _mantissa_match(ours, ref_str, rel):
    """The reference prints t without its exponent unless it is 1.0 / 0.0 / 0.6: compare mantissas."""
    ref = mp.mpf(ref_str)
    if ref == 0:
        return abs(ours) < rel
    return any(abs(ours * mp.mpf(10) ** k - ref) <= rel * abs(ref) for k in range(0, 4))


def test_sodium_double64_trajectory_matches_reference_log():
    mp.mp.dps = 50
    g = load_golden("sodium_double64.json")
    bt, disc, _ = _setup(g)
    its = g["iterations"]
    worst = mp.mpf(0)
    status = 0
    k = 0
    for k in range(100):
        r = bt.step()[0]
        rd = bt.records_dd()[0]
        vals = {nm: _mpf(rd[i]) for i, nm in enumerate(NAMES)}
        # the leading doubles of the record are the Float64 view of the same numbers
        assert r["Etot"] == rd[1, 0] and r["stop"] == rd[0, 0]
        if k < len(its):
            ref = its[k]
            # see the module docstring: the components carry the rounding noise of the line-search parameter,
            # which grows like 1e-30 / stop; Etot (stationary) stays exact
            tol = max(mp.mpf("1e-24"), mp.mpf("1e-28") / mp.mpf(ref["stop"]))
            for nm in ("Ekin", "Ecou", "Ehar"):
                dev = abs(vals[nm] - mp.mpf(ref[nm])) / abs(mp.mpf(ref[nm]))
                assert dev < tol, (k, nm, mp.nstr(dev, 5))
            dev = abs(vals["Etot"] - mp.mpf(ref["Etot"])) / abs(mp.mpf(ref["Etot"]))
            assert dev < mp.mpf("1e-24"), (k, "Etot", mp.nstr(dev, 5))
            worst = max(worst, dev)
            if mp.mpf(ref["stop"]) > mp.mpf("1e-4"):
                assert abs(vals["stop"] - mp.mpf(ref["stop"])) < mp.mpf("1e-18") * abs(mp.mpf(ref["stop"])), k
                assert _mantissa_match(vals["t"], ref["t"], mp.mpf("1e-15")), (k, mp.nstr(vals["t"], 30), ref["t"])
        status = r["status"]
        if status != 0:
            break
    assert status == 1, status
    fin = g["final"]
    # the run ends when rounding noise makes the line search return t = 0 (module docstring): +-3 iterations
    assert abs((k + 1) - fin["niter"]) <= 3, (k + 1, fin["niter"])
    rd = bt.records_dd()[0]
    for i, nm, tol in ((1, "Etot", "1e-28"), (2, "Ekin", "1e-15"), (3, "Ecou", "1e-15"), (4, "Ehar", "1e-15")):
        dev = abs(_mpf(rd[i]) - mp.mpf(fin[nm])) / abs(mp.mpf(fin[nm]))
        assert dev < mp.mpf(tol), (nm, mp.nstr(dev, 5))
    assert _mpf(rd[0]) < mp.mpf("1e-20")
    # orbital energies and occupations of the report (bit-exact integer occupations)
    st = bt.state(0)
    lab = {"1s": (0, 0), "2s": (0, 1), "3s": (0, 2), "2p": (1, 0)}
    for name, ref in fin["orbitals"].items():
        l, kk = lab[name]
        e = _mpf(st["eps"][l, kk, 0])
        assert abs(e - mp.mpf(ref["eps"])) < mp.mpf("2e-14"), (name, mp.nstr(e, 34))   # first order in the final dD
        assert st["n"][l, kk, 0, 0] == float(ref["n"]) and st["n"][l, kk, 0, 1] == 0.0
    occ = st["n"][..., 0, 0]
    assert occ.sum() == 11.0 and np.count_nonzero(occ) == 4
    # orthonormality in the mass matrix and the eigen-residuals of the last Hamiltonian's occupied pairs,
    # evaluated on the host in double-double from the returned pairs: U^T M0 U = I to 1e-28
    from atomickohnsham_b200.dd import DD
    M0 = disc.femops.assemble_pairs(disc.femops.M0_loc)
    M = DD(M0[..., 0], M0[..., 1])
    U = st["U"]
    for l in range(3):
        V = DD(U[:, :4, l, 0, 0], U[:, :4, l, 0, 1])               # (Nh, 4)
        MV = (M[:, :, None] * V[None, :, :]).sum(axis=1)            # (Nh, 4)
        G = (V[:, :, None] * MV[:, None, :]).sum(axis=0)            # (4, 4)
        err = np.abs(G.hi - np.eye(4)) + np.abs(G.lo)
        assert err.max() < 1e-27, (l, err.max())
    # eigen-residuals in double-double on the host: H_l = Kin_l - Z M-1 + F.W + N/Rmax M0 assembled from the dd
    # tables and the returned W (the last step had t = 0, so W belongs to the Hamiltonian of the returned pairs)
    ops, bs = disc.femops, disc.basis
    Wd = DD(st["W"][:, 0], st["W"][:, 1])
    shift = DD(np.array(11.0)) / DD(np.array(float(bs.mesh.points[-1])))
    for l in range(3):
        Hh, Hl = np.zeros((bs.size, bs.size)), np.zeros((bs.size, bs.size))
        Hm = DD(Hh, Hl)
        for c in range(bs.ncells):
            Ib, Ig = np.asarray(bs.cells_to_indices[c]), np.asarray(bs.cells_to_generators[c])
            sub2, sub3 = np.ix_(Ig, Ig), np.ix_(Ig, Ig, Ig)
            pick = lambda t, sub: DD(t.hi[c][sub], t.lo[c][sub])
            FW = (pick(ops.F_loc, sub3) * Wd[Ib][None, None, :]).sum(axis=2)
            blk = (pick(ops.A_loc, sub2) + pick(ops.Mm2_loc, sub2) * float(l * (l + 1))) * 0.5 \
                + pick(ops.Mm1_loc, sub2) * (-11.0) + FW + pick(ops.M0_loc, sub2) * shift
            g = np.ix_(Ib, Ib)
            Hm[g] = Hm[g] + blk
        V = DD(U[:, :4, l, 0, 0], U[:, :4, l, 0, 1])
        E = DD(st["eps"][l, :4, 0, 0], st["eps"][l, :4, 0, 1])
        HV = (Hm[:, :, None] * V[None, :, :]).sum(axis=1)
        MV = (M[:, :, None] * V[None, :, :]).sum(axis=1)
        R = HV - MV * E[None, :]
        res = np.abs(R.hi).max()
        assert res < 1e-26 * max(1.0, np.abs(E.hi).max()), (l, res)
    # trace(D M0) = N
    D = st["D"]
    Dd = DD(D[:, :, 0, 0], D[:, :, 0, 1])
    tr = (Dd * M).sum(axis=0).sum(axis=0)
    assert abs((float(tr.hi) - 11.0) + float(tr.lo)) < 1e-26
    bt.close()


def test_double64_run_and_float64_agree_and_errors():
    """aks_scf_run on a Double64 batch; the Float64 path on the same discretization agrees to Float64
    accuracy; unsupported options are refused loudly."""
    from atomickohnsham_b200 import KSEDiscretization, KSEModel, ODA, OptimizedAufbau, P1IntLegendreBasis, LDA
    from atomickohnsham_b200 import _capi
    from atomickohnsham_b200.solver import Batch
    mp.mp.dps = 50
    g = load_golden("sodium_double64.json")
    bt, disc, model = _setup(g)
    final, hist = bt.run(100, history=True)
    assert final[0]["status"] == 1 and abs(final[0]["niter"] - g["final"]["niter"]) <= 3
    assert len(hist[0]) == final[0]["niter"]
    e_dd = _mpf(bt.records_dd()[0][1])
    assert abs(e_dd - mp.mpf(g["final"]["Etot"])) < mp.mpf("1e-24") * abs(e_dd)
    # the eigenpair window of the Double64 path is transparent: with every level solved in every step
    # (margin < 0) the run is the same bit for bit, and aks_get_state returns all nh eigenpairs either way
    rec_win, st_win = bt.records_dd().copy(), bt.state(0)
    bt.reset()
    bt.set_eig_window(-1)
    final_all, _ = bt.run(100)
    assert final_all[0]["niter"] == final[0]["niter"] and np.array_equal(bt.records_dd(), rec_win)
    st_all = bt.state(0)
    occ = st_all["n"][..., 0, 0] > 0
    assert np.array_equal(st_all["eps"][occ], st_win["eps"][occ]) and np.array_equal(st_all["n"], st_win["n"])
    unocc = ~occ
    dev = np.abs(st_all["eps"][unocc][:, 0] - st_win["eps"][unocc][:, 0])
    assert dev.max() < 1e-25 + 1e-28 * np.abs(st_all["eps"][unocc][:, 0]).max() or dev.max() < 1e-20
    bt.set_eig_window(0)
    # two problems in one batch (neutral atom and cation) are independent of each other
    ion = KSEModel(Z=11, N=10)
    bt2 = Batch([model, ion], disc, bt.alg)
    f2, _ = bt2.run(100)
    r2 = bt2.records_dd()
    assert abs(_mpf(r2[0][1]) - e_dd) < mp.mpf("1e-28") * abs(e_dd)
    assert f2[1]["status"] == 1 and _mpf(r2[1][1]) > e_dd
    bt2.close()
    # Float64 path, same mesh and basis
    s = g["setup"]
    basis = P1IntLegendreBasis(make_mesh(s["mesh"]), ordermax=s["p"])
    d64 = KSEDiscretization(basis, model, lh=s["lh"], nh=s["nh"])
    alg64 = ODA(tinit=0.6, scftol=1e-11, aufbau=OptimizedAufbau(max_degen=2, tol=1e-2))
    b64 = Batch([model], d64, alg64)
    f64, _ = b64.run(100)
    assert abs(f64[0]["Etot"] - float(e_dd)) < 1e-10 * abs(float(e_dd))
    b64.close()
    # a functional on a Double64 discretization built without Gauss tables is refused, as are the Float64-only hooks
    with pytest.raises(_capi.AKSError) as ei:
        Batch([LDA(Z=11, N=11)], disc, bt.alg)
    assert ei.value.code == _capi.AKS_EINVAL
    with pytest.raises(_capi.AKSError) as ei:
        bt.assemble_hartree(0)
    assert ei.value.code == _capi.AKS_ENOSYS
    bt.close()


def test_double64_other_atoms_against_float64():
    """Closed-shell and open-shell atoms on a small order-6 discretization: the Double64 result agrees with
    the Float64 path to Float64 accuracy and satisfies the virial-free identities in double-double
    (Etot = Ekin + Ecou + Ehar exactly as the reference assembles it)."""
    from atomickohnsham_b200 import KSEDiscretization, KSEModel, ODA, OptimizedAufbau, P1IntLegendreBasis
    from atomickohnsham_b200.solver import Batch
    mp.mp.dps = 50
    basis = P1IntLegendreBasis(make_mesh(["exp", 0, 100, 16, 1.5]), ordermax=6)
    models = [KSEModel(Z=z, N=z) for z in (2, 4, 10, 12)]
    ddisc = KSEDiscretization(basis, models[0], lh=1, nh=5, dtype="double64")
    fdisc = KSEDiscretization(basis, models[0], lh=1, nh=5)
    alg = ODA(tinit=0.6, scftol=1e-20, aufbau=OptimizedAufbau(max_degen=2, tol=1e-3))
    bdd = Batch(models, ddisc, alg)
    fdd, _ = bdd.run(200)
    rdd = bdd.records_dd()
    b64 = Batch(models, fdisc, ODA(tinit=0.6, scftol=1e-11, aufbau=OptimizedAufbau(max_degen=2, tol=1e-3)))
    f64, _ = b64.run(200)
    for i in range(len(models)):
        assert fdd[i]["status"] == 1, (i, fdd[i])
        e = [_mpf(rdd[i][q]) for q in range(1, 6)]
        assert abs(e[0] - (e[1] + e[2] + e[3] + e[4])) < mp.mpf("1e-28") * abs(e[0])
        assert abs(float(e[0]) - f64[i]["Etot"]) < 2e-10 * abs(f64[i]["Etot"]), (i, float(e[0]), f64[i]["Etot"])
    bdd.close()
    b64.close()
    # the C2 / C4 discretization (order 20, 40 cells, Nh = 799): neon and the fluoride anion, Hartree only; and
    # 52 cells (Nh = 1039), where the eigensolver's packed cell blocks no longer fit shared memory (global scratch)
    for npts in (41, 53):
        basis = P1IntLegendreBasis(make_mesh(["exp", 0, 300, npts, 1.5]), ordermax=20)
        models = [KSEModel(Z=10, N=10), KSEModel(Z=9, N=10)]
        ddisc = KSEDiscretization(basis, models[0], lh=1, nh=4, dtype="double64")
        fdisc = KSEDiscretization(basis, models[0], lh=1, nh=4)
        bdd = Batch(models, ddisc, alg)
        fdd, _ = bdd.run(200)
        rdd = bdd.records_dd()
        b64 = Batch(models, fdisc, ODA(tinit=0.6, scftol=1e-11, aufbau=OptimizedAufbau(max_degen=2, tol=1e-3)))
        f64, _ = b64.run(200)
        for i in range(len(models)):
            assert fdd[i]["status"] == 1, (npts, i, fdd[i])
            assert abs(float(_mpf(rdd[i][1])) - f64[i]["Etot"]) < 2e-10 * abs(f64[i]["Etot"]), (npts, i, fdd[i], f64[i])
        bdd.close()
        b64.close()


def test_double64_functionals_pointwise_against_mpmath():
    """evaluate_vrho! / evaluate_zk! in double-double (aks_dd_xc_eval) against the mpmath restatement of the
    reference's Double64 formulas (oracle/xc_dd.py): 1e-27 relative, zero at rho <= 0."""
    import ctypes as C
    from atomickohnsham_b200 import _capi
    from atomickohnsham_b200.dd import from_mp
    from oracle import xc_dd
    mp.mp.dps = 50
    ctx = _capi.default_context(0)
    rng = np.random.default_rng(20261020)
    vals = [mp.mpf(10) ** mp.mpf(float(e)) * (1 + mp.mpf(float(u)) * mp.mpf(2) ** -54)
            for e, u in zip(np.linspace(-30, 4, 120), rng.standard_normal(120))]
    vals += [mp.mpf(0), mp.mpf("-1e-12"), mp.mpf(1), mp.mpf("0.5")]
    rho = from_mp(vals).pairs()
    for ex, ec in ((1, 0), (0, 12), (1, 12)):
        v = np.zeros_like(rho)
        z = np.zeros_like(rho)
        ctx.check(_capi.lib.aks_dd_xc_eval(ctx.h, len(vals), _capi.dptr(rho), ex, ec, _capi.dptr(v), _capi.dptr(z)))
        for i, r in enumerate(vals):
            rr = _mpf(rho[i])
            for got, ref in ((_mpf(v[i]), xc_dd.vrho(rr, ex == 1, ec == 12)), (_mpf(z[i]), xc_dd.zk(rr, ex == 1, ec == 12))):
                if rr <= 0:
                    assert got == 0
                else:
                    assert abs(got - ref) <= mp.mpf("1e-27") * abs(ref), (ex, ec, mp.nstr(rr, 5), mp.nstr(got, 30), mp.nstr(ref, 30))


def test_double64_lda_against_float64_path():
    """LDA (lda_x + lda_c_pw) and exchange-only atoms / an anion in Double64 against the Float64 path (pinned to
    the oracle and the reference's goldens by the other GPU tests) on the same mesh and Gauss rule: Etot within
    1e-11 relative, the same occupations, Etot = Ekin + Ecou + Ehar + Eexc in double-double."""
    from atomickohnsham_b200 import (KSEDiscretization, KSEModel, LDA, ODA, OptimizedAufbau, P1IntLegendreBasis,
                                     Slater)
    from atomickohnsham_b200.dd import GaussLegendreDD
    from atomickohnsham_b200.quadrature import GaussLegendre
    from atomickohnsham_b200.solver import Batch
    mp.mp.dps = 50
    basis = P1IntLegendreBasis(make_mesh(["exp", 0, 60, 21, 1.5]), ordermax=8)
    models = [LDA(Z=2, N=2), LDA(Z=10, N=10), Slater(Z=4, N=4), LDA(Z=9, N=10), LDA(Z=6, N=6)]
    ddisc = KSEDiscretization(basis, models[0], lh=1, nh=4, dtype="double64",
                              fem_integration_method=GaussLegendreDD(basis, 300))
    fdisc = KSEDiscretization(basis, models[0], lh=1, nh=4, fem_integration_method=GaussLegendre(basis, 300))
    aufbau = OptimizedAufbau(max_degen=2, tol=1e-3)
    bdd = Batch(models, ddisc, ODA(tinit=0.6, scftol=1e-18, aufbau=aufbau))
    fdd, _ = bdd.run(150)
    rdd = bdd.records_dd()
    b64 = Batch(models, fdisc, ODA(tinit=0.6, scftol=1e-11, aufbau=aufbau))
    f64, _ = b64.run(150)
    for i in range(len(models)):
        assert fdd[i]["status"] == 1, (i, fdd[i])
        e = [_mpf(rdd[i][q]) for q in range(1, 6)]
        assert abs(e[0] - (e[1] + e[2] + e[3] + e[4])) < mp.mpf("1e-28") * abs(e[0])
        assert abs(float(e[0]) - f64[i]["Etot"]) < 1e-11 * abs(f64[i]["Etot"]), (i, float(e[0]), f64[i]["Etot"])
        for q, nm in ((1, "Ekin"), (2, "Ecou"), (3, "Ehar"), (4, "Eexc")):
            assert abs(float(e[q]) - f64[i][nm]) < 1e-7 * abs(f64[i][nm]), (i, nm)
        sd, sf = bdd.state(i, want_D=False, want_U=False), b64.state(i, want_D=False, want_U=False)
        assert np.allclose(sd["n"][..., 0, 0], sf["n"][..., 0], atol=1e-6)
        # bound occupied levels (the LDA anion's outermost level is a box state above zero and sits in a
        # two-fold block that both paths resolve only slowly)
        occ = (sf["n"][..., 0] > 0) & (sf["eps"][..., 0] < -1e-3)
        assert np.allclose(sd["eps"][..., 0, 0][occ], sf["eps"][..., 0][occ], rtol=0, atol=1e-7)
    bdd.close()
    b64.close()


def test_double64_reset_state_all_and_frozen_t():
    """aks_batch_reset reproduces a Double64 run bit for bit; aks_get_state_all returns the pairs of every problem;
    frozen_t = true (oda/loop.jl:86-96) keeps t at tinit and converges to the same Etot as the line search."""
    import ctypes as C
    from atomickohnsham_b200 import KSEDiscretization, KSEModel, ODA, OptimizedAufbau, P1IntLegendreBasis
    from atomickohnsham_b200 import _capi
    from atomickohnsham_b200.solver import Batch
    mp.mp.dps = 50
    basis = P1IntLegendreBasis(make_mesh(["exp", 0, 80, 15, 1.4]), ordermax=6)
    models = [KSEModel(Z=2, N=2), KSEModel(Z=10, N=10), KSEModel(Z=3, N=2)]
    disc = KSEDiscretization(basis, models[0], lh=1, nh=4, dtype="double64")
    aufbau = OptimizedAufbau(max_degen=2, tol=1e-3)
    bt = Batch(models, disc, ODA(tinit=0.6, scftol=1e-22, aufbau=aufbau), lh=[0, 1, 0], nh=[2, 4, 3])
    f1, _ = bt.run(200)
    r1 = bt.records_dd().copy()
    s1 = [bt.state(i) for i in range(3)]
    bt.reset()
    f2, _ = bt.run(200)
    r2 = bt.records_dd()
    assert [f["niter"] for f in f1] == [f["niter"] for f in f2] and np.array_equal(r1, r2)
    for i in range(3):
        s2 = bt.state(i)
        for k in ("eps", "n", "W", "D", "U"):
            assert np.array_equal(s1[i][k], s2[k]), (i, k)
    # every problem at once
    L, nh, Nh = 2, 4, basis.size
    eps = np.zeros((3, L * nh, 2))
    n = np.zeros((3, L * nh, 2))
    W = np.zeros((3, Nh, 2))
    bt.ctx.check(_capi.lib.aks_get_state_all(bt.h, _capi.dptr(eps), _capi.dptr(n), _capi.dptr(W)))
    for i in range(3):
        assert np.array_equal(eps[i].reshape((L, nh, 2), order="F")[..., 0], s1[i]["eps"][..., 0, 0]) or \
            np.array_equal(eps[i][:, 0].reshape((L, nh), order="F"), s1[i]["eps"][..., 0, 0])
        assert np.array_equal(W[i], s1[i]["W"])
        assert n[i][:, 0].sum() == models[i].N
    # per-problem channel / level counts: helium was solved with lh = 0, nh = 2 only
    assert np.count_nonzero(s1[0]["eps"][1, :, 0, 0]) == 0 and np.count_nonzero(s1[0]["eps"][0, 2:, 0, 0]) == 0
    bt.close()
    # frozen t
    # (the two-electron systems: plain damping with t = 0.6 converges linearly there)
    bf = Batch([models[0], models[2]], disc, ODA(tinit=0.6, scftol=1e-24, frozen_t=True, aufbau=aufbau), lh=[0, 0], nh=[2, 3])
    ff, _ = bf.run(600)
    rf = bf.records_dd()
    for j, i in enumerate((0, 2)):
        assert ff[j]["status"] == 1 and ff[j]["t"] == 0.6, ff[j]
        e_ls, e_fr = _mpf(r1[i][1]), _mpf(rf[j][1])
        # the line-search run ends when rounding noise makes t = 0 (module docstring): its Etot carries ~||dD||^2
        assert abs(e_ls - e_fr) < mp.mpf("1e-20") * abs(e_ls), (i, mp.nstr(e_ls, 34), mp.nstr(e_fr, 34))
    bf.close()


@pytest.mark.parametrize("kind", ["smeared", "frozen"])
def test_double64_smeared_and_frozen_aufbau_against_float64(kind):
    """SmearedAufbau / FrozenAufbau (aufbau/smeared.jl:68-101, aufbau/frozen.jl:84-88) on a Double64 batch against the
    Float64 path (pinned to the oracle by test_smeared_and_frozen_aufbau_parity)."""
    from atomickohnsham_b200 import (FrozenAufbau, KSEDiscretization, KSEModel, ODA, P1IntLegendreBasis, SmearedAufbau)
    from atomickohnsham_b200.solver import Batch
    mp.mp.dps = 50
    basis = P1IntLegendreBasis(make_mesh(["exp", 0, 60, 15, 1.5]), ordermax=6)
    models = [KSEModel(Z=8, N=8), KSEModel(Z=11, N=11)]
    auf = SmearedAufbau(Temp=0.05) if kind == "smeared" else FrozenAufbau({"1s": 2.0, "2s": 2.0, "2p": 3.0, "3s": 1.0})
    if kind == "frozen":
        models = [KSEModel(Z=8, N=8)]
    ddisc = KSEDiscretization(basis, models[0], lh=1, nh=4, dtype="double64")
    fdisc = KSEDiscretization(basis, models[0], lh=1, nh=4)
    bdd = Batch(models, ddisc, ODA(tinit=0.6, scftol=1e-18, aufbau=auf))
    b64 = Batch(models, fdisc, ODA(tinit=0.6, scftol=1e-11, aufbau=auf))
    fdd, _ = bdd.run(200)
    f64, _ = b64.run(200)
    rdd = bdd.records_dd()
    for i in range(len(models)):
        assert fdd[i]["status"] == 1 and f64[i]["status"] == 1, (fdd[i], f64[i])
        assert abs(float(_mpf(rdd[i][1])) - f64[i]["Etot"]) < 1e-10 * abs(f64[i]["Etot"]), (i, fdd[i], f64[i])
        nd = bdd.state(i, want_D=False, want_U=False)["n"]
        nf = b64.state(i, want_D=False, want_U=False)["n"]
        assert np.allclose(nd[..., 0, 0] + nd[..., 0, 1], nf[..., 0], rtol=0, atol=1e-8)
        assert abs((nd[..., 0, 0].sum() + nd[..., 0, 1].sum()) - models[i].N) < 1e-12
    bdd.close()
    b64.close()


# ---------------------------------------------------------------------------------------------------------------
# Double64 WITH a functional (config C4 runs lda_x + lda_c_pw): step oracle in mpmath, oracle/scf_mp.py
def _mp_case(Z, *, frozen_t, tinit="0.6"):
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.dd import GaussLegendreDD, build_femops_dd
    from atomickohnsham_b200.solver import Batch
    from oracle.scf_mp import OracleMP
    mp.mp.dps = 50
    basis = aks.P1IntLegendreBasis(aks.expmesh(0, 30, 9, s=1.5), ordermax=6)     # 8 cells, Nh = 47
    ops, quad = build_femops_dd(basis), GaussLegendreDD(basis, 64)
    model = aks.LDA(Z=Z, N=Z)
    disc = aks.KSEDiscretization(basis, model, lh=1, nh=3, dtype="double64", fem_integration_method=quad, femops=ops)
    bt = Batch([model], disc, aks.ODA(tinit=float(tinit), scftol=1e-25, frozen_t=frozen_t,
                                      aufbau=aks.OptimizedAufbau(max_degen=2, tol=1e-3)))
    # tinit crosses the C ABI as a Float64 (ODA(tinit = 0.6) is the Float64 number 0.6 in the reference too)
    o = OracleMP(basis, ops, quad, Z=Z, N=Z, lh=1, nh=3, tinit=mp.mpf(float(tinit)), frozen_t=frozen_t, scftol="1e-25")
    return bt, o


def _gpu_step(bt):
    bt.step()
    rd = bt.records_dd()[0]
    st = bt.state(0, want_D=False, want_U=False)
    rec = {k: _mpf(rd[i]) for i, k in enumerate(NAMES)}
    eps = [[_mpf(st["eps"][l, k, 0]) for k in range(3)] for l in range(2)]
    n = [[_mpf(st["n"][l, k, 0]) for k in range(3)] for l in range(2)]
    return rec, eps, n


def _rel(a, b):
    return abs(a - b) / max(abs(b), mp.mpf("1e-300"))


@pytest.mark.parametrize("Z", [2, 10])
def test_double64_with_functional_frozen_t_against_mpmath_step_oracle(Z):
    """Closes the parity gap of row a20: the Double64 path WITH lda_x + lda_c_pw against the mpmath restatement of the
    reference's step (oracle/scf_mp.py) on identical double-double tables.  With frozen_t the trajectory is a
    deterministic function of the tables (no Brent on a flat minimum), so EVERY quantity of every step is pinned:
    all energies, t, stop, all orbital energies and the occupations, to 1e-24 relative (north star: 1e-20).
    Exercises ddk_xc, ddk_H (Vxc and Hartree blocks), ddk_eig, ddk_aufbau, ddk_rho_*, ddk_decide (Exc on the grid,
    energy bookkeeping), ddk_mix, ddk_finish."""
    bt, o = _mp_case(Z, frozen_t=True, tinit="0.5")
    st = o.initial_state()
    worst = mp.mpf(0)
    for it in range(4):
        o.scf_step(st)
        rec, eps, n = _gpu_step(bt)
        for k in ("Etot", "Ekin", "Ecou", "Ehar", "Eexc"):
            worst = max(worst, _rel(rec[k], st["E"][k]))
            assert _rel(rec[k], st["E"][k]) < mp.mpf("1e-24"), (it, k, mp.nstr(rec[k], 34), mp.nstr(st["E"][k], 34))
        assert _rel(rec["stop"], st["stop"]) < mp.mpf("1e-22"), (it, mp.nstr(rec["stop"], 30), mp.nstr(st["stop"], 30))
        assert rec["t"] == (mp.mpf("0.5") if it > 0 else rec["t"])
        for l in range(2):
            for k in range(3):
                assert n[l][k] == st["n"][l][k], (it, l, k)                           # occupations: exact
                # orbital energies: every level, occupied or not (relative to the spectrum's scale for |eps| << 1)
                tol = mp.mpf("1e-24") * max(abs(st["eps"][l][k]), mp.mpf(1))
                assert abs(eps[l][k] - st["eps"][l][k]) < tol, (it, l, k, mp.nstr(eps[l][k], 34), mp.nstr(st["eps"][l][k], 34))
    print(f"[dd + LDA, frozen_t] Z={Z}: 4 steps, worst relative energy difference {mp.nstr(worst, 3)}")
    bt.close()


def test_double64_with_functional_oda_brent_against_mpmath_step_oracle():
    """The same with the adaptive line search (Brent in double-double, tolerances 1e4 eps(Double64)).  Step 0 (Harris-like
    energy of the first trial density) and the orbital energies of step 1 (eigenvalues of H[D_0] with Vxc[D_0]) do not
    depend on any line search: 1e-24.  From step 1 on the mixing parameter is the minimiser of an energy that is flat to
    rounding (t is determined to ~sqrt(eps) = 1e-16 by function values, in the reference's Double64 arithmetic too): Etot
    stays pinned (stationary in t), t and everything first order in t to 1e-12.  Step 2 of neon ends on the boundary
    t -> 1, where E(t) is NOT stationary and Brent runs into `maxiter_ls = 100` before its 1e-27 tolerance (golden
    section steps shrink the bracket by 0.62: 0.62^100 = 1e-21), so the last t -- and with it Etot -- is a path-dependent
    1 - O(1e-21) in ANY implementation of the reference's algorithm: Etot to the north star's 1e-20 there."""
    bt, o = _mp_case(10, frozen_t=False)
    st = o.initial_state()
    for it in range(3):
        o.scf_step(st)
        rec, eps, n = _gpu_step(bt)
        assert _rel(rec["Etot"], st["E"]["Etot"]) < (mp.mpf("1e-24") if it <= 1 else mp.mpf("1e-20")), \
            (it, mp.nstr(rec["Etot"], 34), mp.nstr(st["E"]["Etot"], 34))
        tol = mp.mpf("1e-24") if it == 0 else mp.mpf("1e-12")
        for k in ("Ekin", "Ecou", "Ehar", "Eexc"):
            assert _rel(rec[k], st["E"][k]) < tol, (it, k, mp.nstr(rec[k], 34), mp.nstr(st["E"][k], 34))
        assert abs(rec["t"] - st["t"]) < (mp.mpf("1e-30") if it == 0 else mp.mpf("1e-12")), (it, mp.nstr(rec["t"], 30), mp.nstr(st["t"], 30))
        etol = mp.mpf("1e-24") if it <= 1 else mp.mpf("1e-11")
        for l in range(2):
            for k in range(3):
                assert n[l][k] == st["n"][l][k]
                assert abs(eps[l][k] - st["eps"][l][k]) < etol * max(abs(st["eps"][l][k]), mp.mpf(1)), (it, l, k)
    bt.close()


def test_double64_states_all_and_occupied_labels():
    """Batch.states() of a Double64 batch (aks_get_state_all writes (hi, lo) pairs: twice the Float64 size),
    KSESolution.occupied and sweep.maximum_ionization on a Double64 discretization."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch, groundstate
    from atomickohnsham_b200.sweep import maximum_ionization
    basis = aks.P1IntLegendreBasis(aks.expmesh(0, 30, 9, s=1.5), ordermax=6)
    models = [aks.KSEModel(Z=2, N=2), aks.KSEModel(Z=4, N=4)]
    disc = aks.KSEDiscretization(basis, models[0], lh=1, nh=3, dtype="double64")
    bt = Batch(models, disc, aks.ODA(tinit=0.6, scftol=1e-20))
    for _ in range(3):
        bt.step()
    sts = bt.states()
    for i in range(2):
        one = bt.state(i, want_D=False, want_U=False)
        assert sts[i]["eps"].shape == (2, 3, 1, 2) and sts[i]["W"].shape == (disc.Nh, 2)
        assert np.array_equal(sts[i]["eps"], one["eps"]) and np.array_equal(sts[i]["n"], one["n"])
        assert np.array_equal(sts[i]["W"], one["W"])
    bt.close()
    sol = groundstate(models[0], disc, aks.ODA(tinit=0.6, scftol=1e-20), maxiter=60)
    assert [o[0] for o in sol.occupied] == ["1s"] and abs(sol.occupied[0][2] - 2.0) < 1e-30
    nmax = maximum_ionization([2], disc, ex=None, ec=None, width=0.4, precision=1, maxiter=30)
    assert 2.0 <= nmax[2] <= 2.4


def test_double64_tables_built_on_the_device_match_the_host_mirror():
    """the Double64 tables (`GaussLegendre{Double64}`, `init_cache!`) built by the library on the
    device (`aks_dd_gauss_legendre`, `aks_dd_build_operators`) against the mpmath / NumPy double-double mirror of dd.py
    (itself checked against mpmath quadrature in tests/test_dd_host.py): 1e-30 relative per table at order 20; and a
    Double64 run on device-built tables reproduces the reference's committed sodium log like the host-built ones."""
    import time
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200 import _capi
    from atomickohnsham_b200.dd import GaussLegendreDD, build_femops_dd, gauss_legendre_dd
    ctx = _capi.default_context(0)

    def rel(a, b):
        d = (a.hi - b.hi) + (a.lo - b.lo)
        return float(np.max(np.abs(d)) / np.max(np.abs(b.hi)))

    for n in (7, 140, 1000):
        xd, wd = gauss_legendre_dd(n, ctx=ctx)
        xh, wh = gauss_legendre_dd(n)
        assert rel(xd, xh) < 1e-30 and rel(wd, wh) < 5e-30, (n, rel(xd, xh), rel(wd, wh))
    basis = aks.P1IntLegendreBasis(aks.expmesh(0, 300, 41, s=1.5), ordermax=20)
    t0 = time.perf_counter()
    dev = build_femops_dd(basis, ctx=ctx)
    qd = GaussLegendreDD(basis, 1000, ctx=ctx)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    host = build_femops_dd(basis)
    qh = GaussLegendreDD(basis, 1000)
    t_host = time.perf_counter() - t0
    for name in ("M0_loc", "A_loc", "Mm1_loc", "Mm2_loc", "F_loc"):
        assert rel(getattr(dev, name), getattr(host, name)) < 1e-30, (name, rel(getattr(dev, name), getattr(host, name)))
    for name in ("x", "w", "y", "wy", "Pgenx", "Pgeny"):
        assert rel(getattr(qd, name), getattr(qh, name)) < 1e-28, (name, rel(getattr(qd, name), getattr(qh, name)))
    print(f"[dd tables, order 20, 40 cells, Q = 1000] device {t_dev:.3f} s, host mirror {t_host:.2f} s")
    assert t_dev < 0.5
    # the reference's Double64 sodium run on device-built tables (default of KSEDiscretization(dtype="double64"))
    g = load_golden("sodium_double64.json")
    bt, disc, _ = _setup(g)
    for it in range(3):
        bt.step()
        rd = bt.records_dd()[0]
        assert abs(_mpf(rd[1]) - mp.mpf(g["iterations"][it]["Etot"])) < mp.mpf("1e-24") * abs(mp.mpf(g["iterations"][it]["Etot"]))
    bt.close()
