"""Oracle run of config C5 (periodic-table sweep Z = 1..86, LDA, order 20) -> tests/golden/c5_sweep.json.

The protocol is the reference's own (examples/atomic density comparison/run.jl:40-76, applied to the whole
table instead of its six atoms):

  pass 1   every atom with the reference's default algorithm (run.jl:39,70): ODA(tinit = 0.4, scftol = 1e-10),
           OptimizedAufbau(tol = 1e-2), maxiter = 100, (lh, nh) = basis_size(Z)  (run.jl:58)
  pass 2   every atom that ended pass 1 with MAXITERS is solved again FROM SCRATCH with the algorithm the
           reference gives iron (run.jl:40-47): ODA(tinit = 0.15, frozen_t = true, scftol = 1e-12), maxiter = 800.
           As the reference notes for iron, such a run may plateau and report MAXITERS: that is a defined end
           state too (status, stop and Etot are recorded).

Discretization = C2 (expmesh(0,300,41;s=1.5), ordermax 20, 1000 Gauss points per cell), stable tables of the
package (parity is always checked on identical tables).  One atom per process, one BLAS thread each.

  python tools/make_golden_c5.py [--procs 8] [--zs 1-86] [--pass2-maxiter 800]
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

C2 = {"mesh": ["exp", 0, 300, 41, 1.5], "p": 20, "npoints": 1000}
PASS1 = {"tinit": 0.4, "scftol": 1e-10, "frozen_t": False, "maxiter": 100, "degen_tol": 1e-2}
PASS2 = {"tinit": 0.15, "scftol": 1e-12, "frozen_t": True, "maxiter": 800, "degen_tol": 1e-2}


def run_atom(job):
    Z, alg = job
    import numpy as np
    from helpers import make_oracle
    from atomickohnsham_b200.sweep import basis_size
    from oracle.scf import DegeneracyError
    lh, nh = basis_size(Z)
    setup = {"Z": Z, "N": Z, "lh": lh, "nh": nh, "ex": "lda_x", "ec": "lda_c_pw", **C2,
             "scftol": alg["scftol"], "tinit": alg["tinit"], "frozen_t": alg["frozen_t"], "degen_tol": alg["degen_tol"]}
    o = make_oracle(setup)
    t0 = time.perf_counter()
    err = None
    try:
        st, hist = o.groundstate(maxiter=alg["maxiter"])
    except DegeneracyError as e:   # optimized.jl:139
        return {"Z": Z, "status": "EDEGEN3", "error": str(e), "seconds": time.perf_counter() - t0}
    e = st.energies
    conv = st.stop < alg["scftol"]
    return {"Z": Z, "lh": lh, "nh": nh, "status": "SUCCESS" if conv else "MAXITERS", "niter": st.niter,
            "stop": st.stop, "Etot": e.Etot, "Ekin": e.Ekin, "Ecou": e.Ecou, "Ehar": e.Ehar, "Eexc": e.Eexc,
            "t": st.t, "eps": st.eps[:, :, 0].tolist(), "n": st.n[:, :, 0].tolist(),
            "stop_history": [h["stop"] for h in hist], "Etot_history": [h["Etot"] for h in hist],
            "seconds": time.perf_counter() - t0, "error": err}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--zs", default="1-86")
    ap.add_argument("--pass2-maxiter", type=int, default=PASS2["maxiter"])
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "c5_sweep.json"))
    args = ap.parse_args()
    lo, hi = (int(v) for v in args.zs.split("-"))
    Zs = list(range(lo, hi + 1))
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    p2 = dict(PASS2, maxiter=args.pass2_maxiter)
    out = {"config": C2, "pass1_alg": PASS1, "pass2_alg": p2,
           "source": "oracle/scf.py (NumPy/LAPACK restatement of the reference; Julia cannot run here), "
                     "protocol of examples/atomic density comparison/run.jl:40-76", "atoms": {}}
    t0 = time.perf_counter()
    with ctx.Pool(args.procs) as pool:
        # heavy atoms first: the pool drains evenly
        for r in pool.imap_unordered(run_atom, [(Z, PASS1) for Z in sorted(Zs, reverse=True)]):
            out["atoms"][str(r["Z"])] = {"pass1": r}
            print(f"pass1 Z={r['Z']:3d} {r['status']:9s} niter={r.get('niter')} stop={r.get('stop')} "
                  f"Etot={r.get('Etot')} [{r['seconds']:.0f}s]", flush=True)
        json.dump(out, open(args.out + ".partial", "w"))
        todo = [Z for Z in Zs if out["atoms"][str(Z)]["pass1"]["status"] == "MAXITERS"]
        print("pass 2 for", todo, flush=True)
        for r in pool.imap_unordered(run_atom, [(Z, p2) for Z in sorted(todo, reverse=True)]):
            out["atoms"][str(r["Z"])]["pass2"] = r
            print(f"pass2 Z={r['Z']:3d} {r['status']:9s} niter={r.get('niter')} stop={r.get('stop')} "
                  f"Etot={r.get('Etot')} [{r['seconds']:.0f}s]", flush=True)
            json.dump(out, open(args.out + ".partial", "w"))
    out["wall_seconds"] = time.perf_counter() - t0
    out["atoms"] = {str(Z): out["atoms"][str(Z)] for Z in Zs}
    json.dump(out, open(args.out, "w"), indent=0)
    os.remove(args.out + ".partial")
    n1 = sum(1 for a in out["atoms"].values() if a["pass1"]["status"] == "SUCCESS")
    n2 = sum(1 for a in out["atoms"].values() if a.get("pass2", {}).get("status") == "SUCCESS")
    print(f"done: pass1 converged {n1}/{len(Zs)}, pass2 converged {n2}/{len(todo)}, {out['wall_seconds']:.0f}s")


if __name__ == "__main__":
    main()
