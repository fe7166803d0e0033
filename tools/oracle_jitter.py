"""Iteration-count jitter of the reference's algorithm under 1-ulp table perturbations (SURVEY 8c protocol iii).

The default ODA run ends on a flat objective (Brent on function values: t is determined to ~sqrt(eps)), so the number
of iterations to `scftol` is decided by rounding noise.  This script measures that spread on the ORACLE itself: every
case is run once on the package's tables and K times on tables whose quadrature weights and whose stiffness and 1/r
matrices (symmetrically) were multiplied entrywise by (1 + s * 2^-52), s = -1, 0, +1 at random (seeded) -- a perturbation
no larger than the difference between two summation orders.  Output: tests/golden/iteration_count_jitter.json (copy under profiles/); the GPU test compares the device.s
counts for the same cases.  CPU only.

  python tools/oracle_jitter.py [--reps 8]
"""
import argparse, json, os, sys
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

CASES = {
    "o_lda":       {"Z": 8, "N": 8, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10},
    "na_cation":   {"Z": 11, "N": 10, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-9},
    "na_slater":   {"Z": 11, "N": 11, "mesh": ["exp", 0, 500, 20, 1.2], "p": 10, "lh": 2, "nh": 10, "npoints": 2000, "ex": "lda_x", "ec": None, "scftol": 1e-9, "degen_tol": 0.1},
    "ne_lda":      {"Z": 10, "N": 10, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 1, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10},
    "sc_rhf":      {"Z": 21, "N": 21, "mesh": ["exp", 0, 2000, 40, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": None, "ec": None, "scftol": 1e-10},
    "o_frozen_t":  {"Z": 8, "N": 8, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-9, "frozen_t": True, "tinit": 0.5},
}


def run(setup, seed):
    from helpers import oracle_tables_from_package, package_tables
    from oracle.scf import ODAOptions, Oracle
    s = dict(setup)
    tb = oracle_tables_from_package(*package_tables(s["mesh"], s["p"], s.get("npoints", 1000)))
    if seed is not None:   # perturb BEFORE the oracle precomputes anything from the tables
        rng = np.random.default_rng(seed)
        for name in ("w", "wy"):
            v = np.array(getattr(tb, name), dtype=np.float64, copy=True)
            v *= 1.0 + rng.integers(-1, 2, size=v.shape) * 2.0 ** -52
            setattr(tb, name, v)
        for name in ("A", "Mm1"):   # operators every run uses (also without a functional): symmetric 1-ulp perturbation
            m = np.array(getattr(tb, name), dtype=np.float64, copy=True)
            f = np.triu(rng.integers(-1, 2, size=m.shape)).astype(np.float64)
            f = f + np.triu(f, 1).T
            setattr(tb, name, m * (1.0 + f * 2.0 ** -52))
    opt = ODAOptions(scftol=s.get("scftol", 1e-10), tinit=s.get("tinit", 0.6), frozen_t=s.get("frozen_t", False),
                     max_degen=s.get("max_degen", 2), degen_tol=s.get("degen_tol", 1e-3), aufbau_kind="optimized", temp=0.0, nfix=None)
    o = Oracle(tb, Z=s["Z"], N=s["N"], hartree=s.get("hartree", 1), ex=s.get("ex"), ec=s.get("ec"), nspin=s.get("nspin", 1),
               lh=s["lh"], nh=s["nh"], options=opt, eig="generalized")
    st, hist = o.groundstate(maxiter=200)
    return int(st.niter), float(st.stop), float(st.energies.Etot)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "iteration_count_jitter.json"))
    args = ap.parse_args()
    out = {"what": "iterations of the oracle to scftol on the package's tables and on tables with the quadrature weights "
                   "and the A, M_-1 matrices perturbed by at most 1 ulp (seeds 0..reps-1); Etot spread of the converged runs", "cases": {}}
    for name, setup in CASES.items():
        base = run(setup, None)
        pert = [run(setup, s) for s in range(args.reps)]
        its = [p[0] for p in pert]
        es = [p[2] for p in pert] + [base[2]]
        out["cases"][name] = {"setup": setup, "oracle_niter": base[0], "oracle_niter_perturbed": its,
                              "range": [min(its + [base[0]]), max(its + [base[0]])],
                              "Etot": base[2], "Etot_rel_spread": (max(es) - min(es)) / abs(base[2])}
        print(name, "oracle", base[0], "perturbed", its, "Etot spread %.1e" % out["cases"][name]["Etot_rel_spread"], flush=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
