"""CPU tests of the Double64 host side: double-double NumPy arithmetic and the dd FEM tables
(atomickohnsham_b200/dd.py) against mpmath, and the committed Double64 golden of the reference."""
import mpmath as mp
import numpy as np

from helpers import load_golden, make_mesh


def test_dd_numpy_arithmetic_against_mpmath():
    from atomickohnsham_b200.dd import DD, from_mp
    mp.mp.dps = 60
    rng = np.random.default_rng(20261020)
    a = [mp.mpf(float(x)) * mp.mpf(10) ** int(e) + mp.mpf(float(y)) * mp.mpf(2) ** -60
         for x, y, e in zip(rng.standard_normal(50), rng.standard_normal(50), rng.integers(-5, 5, 50))]
    b = [mp.mpf(float(x)) + mp.mpf(float(y)) * mp.mpf(2) ** -58 for x, y in zip(rng.standard_normal(50), rng.standard_normal(50))]
    A, B = from_mp(a), from_mp(b)
    for got, ref in ((A + B, [x + y for x, y in zip(a, b)]), (A * B, [x * y for x, y in zip(a, b)]),
                     (A - B, [x - y for x, y in zip(a, b)])):
        for g, r, x, y in zip(got.to_mp(), ref, a, b):
            assert abs(g - r) <= mp.mpf(2) ** -102 * max(abs(x), abs(y), abs(r)), (g, r)
    s = A.sum(axis=0).to_mp()[0]
    assert abs(s - sum(a)) < mp.mpf(2) ** -100 * sum(abs(x) for x in a)
    assert A.pairs().shape == (50, 2) and isinstance(A[3:5], DD)


def test_dd_tables_against_mpmath_quadrature_and_float64_tables():
    from atomickohnsham_b200.basis import P1IntLegendreBasis
    from atomickohnsham_b200.dd import build_femops_dd, pair_to_mp
    from atomickohnsham_b200.femops import build_femops
    mp.mp.dps = 45
    mesh = make_mesh(["exp", 0, 50, 7, 1.3])
    basis = P1IntLegendreBasis(mesh, ordermax=5)
    ops = build_femops_dd(basis)
    o64 = build_femops(basis)
    for nm in ("M0_loc", "A_loc", "Mm1_loc", "Mm2_loc", "F_loc"):
        a, b = getattr(ops, nm), getattr(o64, nm)
        assert np.max(np.abs(a.hi - b)) <= 1e-13 * max(1.0, np.max(np.abs(b))), nm
        assert np.all(np.abs(a.lo) <= 2.0 ** -52 * np.abs(a.hi) + 1e-300), nm

    def gen(g, x):
        if g == 0:
            return 1 + x
        if g == 1:
            return 1 - x
        k = g - 1
        return (mp.legendre(k + 1, x) - mp.legendre(k - 1, x)) / (2 * k + 1)

    def dgen(g, x):
        return mp.mpf(1) if g == 0 else (mp.mpf(-1) if g == 1 else mp.legendre(g - 1, x))

    pts = [mp.mpf(float(v)) for v in mesh.points]
    for (c, i, j, m) in [(1, 0, 1, 2), (3, 4, 5, 3), (0, 0, 2, 2), (5, 1, 3, 5), (2, 0, 0, 0)]:
        a, b = pts[c], pts[c + 1]
        h2 = (b - a) / 2
        mid = a + h2
        ref = dict(
            F=h2 * mp.quad(lambda x: gen(i, x) * gen(j, x) * gen(m, x) / (h2 * x + mid), [-1, 0, 1]),
            Mm1=h2 * mp.quad(lambda x: gen(i, x) * gen(j, x) / (h2 * x + mid), [-1, 0, 1]),
            Mm2=h2 * mp.quad(lambda x: gen(i, x) * gen(j, x) / (h2 * x + mid) ** 2, [-1, 0, 1]),
            M0=h2 * mp.quad(lambda x: gen(i, x) * gen(j, x), [-1, 0, 1]),
            A=mp.quad(lambda x: dgen(i, x) * dgen(j, x), [-1, 0, 1]) / h2)
        got = dict(F=pair_to_mp(ops.F_loc.hi[c, i, j, m], ops.F_loc.lo[c, i, j, m]),
                   **{k: pair_to_mp(getattr(ops, k + "_loc").hi[c, i, j], getattr(ops, k + "_loc").lo[c, i, j])
                      for k in ("Mm1", "Mm2", "M0", "A")})
        for k in ref:
            assert abs(got[k] - ref[k]) <= mp.mpf("1e-29") * max(abs(ref[k]), 1 / h2, mp.mpf(1)), (c, i, j, m, k)
    # global views: pairs, symmetric, COO keys of the Float64 builder
    P = ops.assemble_pairs(ops.Mm1_loc)
    assert P.shape == (basis.size, basis.size, 2) and np.array_equal(P[..., 0], P[..., 0].T)
    assert np.max(np.abs(P[..., 0] - o64.Mm1)) < 1e-13
    I, J, M, V = ops.F_coo_pairs()
    i64, j64, m64, v64 = o64.F_coo()
    d1 = {(a, b, c): x for a, b, c, x in zip(i64, j64, m64, v64)}
    d2 = {(a, b, c): x for a, b, c, x in zip(I, J, M, V[:, 0])}
    assert set(d1) == set(d2) and max(abs(d1[k] - d2[k]) for k in d1) < 1e-13


def test_double64_golden_fixture_is_consistent():
    """The transcribed reference log: 59 iterations, Etot = Ekin + Ecou + Ehar + Eexc to 1e-30, monotone
    energies, final report == last iteration."""
    mp.mp.dps = 50
    g = load_golden("sodium_double64.json")
    its = g["iterations"]
    assert len(its) == 59 == g["final"]["niter"]
    prev = None
    for r in its:
        e = {k: mp.mpf(r[k]) for k in ("Etot", "Ekin", "Ecou", "Ehar", "Eexc")}
        assert abs(e["Etot"] - (e["Ekin"] + e["Ecou"] + e["Ehar"] + e["Eexc"])) < mp.mpf("1e-30") * abs(e["Etot"])
        if prev is not None:
            assert e["Etot"] <= prev + mp.mpf("1e-28")
        prev = e["Etot"]
    assert mp.mpf(its[-1]["Etot"]) == mp.mpf(g["final"]["Etot"])
    assert mp.mpf(its[-1]["stop"]) < mp.mpf(g["setup"]["scftol"])


def test_xc_dd_oracle_matches_float64_oracle():
    """The mpmath restatement of the Double64 functionals (oracle/xc_dd.py) against the Float64 oracle
    (oracle/xc.py, pinned to the reference's goldens) on Float64 densities."""
    from oracle import xc as xc64
    from oracle import xc_dd
    mp.mp.dps = 40
    rho = np.concatenate([10.0 ** np.linspace(-28, 3, 40), [0.0, -1e-9]])
    vx, zx = xc64.slater_vrho(rho), xc64.slater_zk(rho)
    X = xc64.XC("lda_x", "lda_c_pw", 1)
    v_all, z_all = X.vrho(rho), X.zk(rho)
    for i, r in enumerate(rho):
        assert abs(float(xc_dd.vrho(r, True, False)) - vx[i]) <= 4e-15 * abs(vx[i])
        assert abs(float(xc_dd.zk(r, True, False)) - zx[i]) <= 4e-15 * abs(zx[i])
        # log(1 + 1/Q1) of PW92 is evaluated by the Float64 oracle (and by the reference) as written: its relative
        # error grows like eps * Q1 ~ eps * rs^2 towards low densities
        if r > 0:
            rs = (3.0 / (4.0 * np.pi * r)) ** (1.0 / 3.0)
            tol = 1e-13 + 1e-16 * (1.0 + rs * rs)
            if tol < 1e-3:
                assert abs(float(xc_dd.vrho(r, True, True)) - v_all[i]) <= tol * abs(v_all[i])
                assert abs(float(xc_dd.zk(r, True, True)) - z_all[i]) <= tol * abs(z_all[i])
        else:
            assert xc_dd.vrho(r, True, True) == 0 and xc_dd.zk(r, True, True) == 0 and v_all[i] == 0


def test_mpmath_step_oracle_matches_float64_oracle():
    """oracle/scf_mp.py (checker of the Double64 path with a functional) pinned to oracle/scf.py, which is pinned to the
    reference's goldens: two SCF steps of helium (LDA) on the same small discretization; rounded to Float64 the
    50-digit restatement must reproduce the Float64 one (Etot 1e-12, orbital energies 1e-10, t 1e-8: the Float64 line
    search only determines t to sqrt(eps))."""
    import mpmath as mp
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.dd import GaussLegendreDD, build_femops_dd
    from helpers import make_oracle
    from oracle.scf_mp import OracleMP
    basis = aks.P1IntLegendreBasis(aks.expmesh(0, 30, 9, s=1.5), ordermax=6)
    o = OracleMP(basis, build_femops_dd(basis), GaussLegendreDD(basis, 64), Z=2, N=2, lh=1, nh=3)
    of = make_oracle({"Z": 2, "N": 2, "mesh": ["exp", 0, 30, 9, 1.5], "p": 6, "npoints": 64, "lh": 1, "nh": 3,
                      "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-12})
    st, sf = o.initial_state(), of.initial_state()
    for it in range(2):
        o.scf_step(st)
        of.scf_step(sf)
        assert abs(float(st["E"]["Etot"]) - sf.energies.Etot) < 1e-12 * abs(sf.energies.Etot), it
        assert abs(float(st["t"]) - sf.t) < 1e-8
        for l in range(2):
            for k in range(3):
                assert abs(float(st["eps"][l][k]) - sf.eps[l, k, 0]) < 1e-10 * max(1.0, abs(sf.eps[l, k, 0])), (it, l, k)
                assert float(st["n"][l][k]) == sf.n[l, k, 0]
