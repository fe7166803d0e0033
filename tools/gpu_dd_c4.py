"""C4 of SURVEY.md section 8: Double64 negative-ion study -- anions Z = 1..20 with N = Z + 1 on the order-20
discretization (expmesh(0, 300, 41; s = 1.5), Q = G = 1000), builtin lda_x + lda_c_pw, scftol = 1e-20.
Prints per-step device time and, per anion, status / iterations / Etot / outermost orbital energy in
double-double next to the Float64 path on the same tables.  JSON summary -> gpurun_out/dd_c4.json."""
import json
import os
import sys
import time

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomickohnsham_b200 import KSEDiscretization, LDA, ODA, OptimizedAufbau, P1IntLegendreBasis, expmesh  # noqa: E402
from atomickohnsham_b200.solver import Batch  # noqa: E402

mp.mp.dps = 40
Zs = list(range(1, int(sys.argv[1]) + 1 if len(sys.argv) > 1 else 21))
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
basis = P1IntLegendreBasis(expmesh(0, 300, 41, s=1.5), ordermax=20)
models = [LDA(Z=z, N=z + 1) for z in Zs]
lh = [min(-(-z // 20), 3) for z in Zs]   # lh of SURVEY section 8 (L = lh + 1 channels)
nh = [min(max(z, 2), 10) for z in Zs]
t0 = time.time()
ddisc = KSEDiscretization(basis, models[0], lh=max(lh), nh=max(nh), dtype="double64")
ddisc.handle()
t_tab = time.time() - t0
fdisc = KSEDiscretization(basis, models[0], lh=max(lh), nh=max(nh))
auf = OptimizedAufbau(max_degen=2, tol=1e-3)
bdd = Batch(models, ddisc, ODA(tinit=0.6, scftol=1e-20, aufbau=auf), lh=lh, nh=nh)
b64 = Batch(models, fdisc, ODA(tinit=0.6, scftol=1e-10, aufbau=auf), lh=lh, nh=nh)
print(f"tables (double-double, host) {t_tab:.2f} s", flush=True)
steps = []
homo_hist = [[] for _ in Zs]     # outermost occupied orbital energy of every anion, iteration by iteration


def homo_of(i):
    st = bdd.state(i, want_D=False, want_U=False)
    occ = st["n"][..., 0, 0] > 0
    eps = st["eps"][..., 0, :]
    return max((mp.mpf(float(eps[l, k, 0])) + mp.mpf(float(eps[l, k, 1])) for l in range(eps.shape[0]) for k in range(eps.shape[1]) if occ[l, k]))


for it in range(maxiter):
    t1 = time.time()
    rec = bdd.step()
    steps.append(time.time() - t1)
    if it >= maxiter - 3 or all(r["status"] != 0 for r in rec):
        for i in range(len(Zs)):
            homo_hist[i].append(homo_of(i))
    nrun = sum(1 for r in rec if r["status"] == 0)
    if it < 5 or it % 10 == 0 or nrun == 0:
        cold = bdd.eig_refills()
        print(f"it {it:3d} {steps[-1] * 1e3:8.2f} ms running {nrun}  cold-path eigenpairs {int(cold.sum())} (per anion: {cold.tolist()})", flush=True)
    if nrun == 0:
        break
f64, _ = b64.run(maxiter, raise_on_problem_error=False)
rd = bdd.records_dd()
out = []
for i, z in enumerate(Zs):
    st = bdd.state(i, want_D=False, want_U=False)
    occ = st["n"][..., 0, 0] > 0
    eps = st["eps"][..., 0, :]
    homo = max((mp.mpf(float(eps[l, k, 0])) + mp.mpf(float(eps[l, k, 1])) for l in range(eps.shape[0]) for k in range(eps.shape[1]) if occ[l, k]))
    etot = mp.mpf(rd[i, 1, 0]) + mp.mpf(rd[i, 1, 1])
    # achieved precision of the outermost orbital energy: its change over the last recorded iterations
    hh = homo_hist[i]
    dh = max((abs(hh[j] - hh[-1]) for j in range(len(hh) - 1)), default=mp.mpf(0))
    row = dict(Z=z, N=z + 1, status=rec[i]["status"], niter=rec[i]["niter"], stop=float(rd[i, 0, 0]), Etot=mp.nstr(etot, 32),
               eps_homo=mp.nstr(homo, 32), eps_homo_change_last_iterations=mp.nstr(dh, 3), bound=bool(homo < 0), Etot_f64=f64[i]["Etot"], status_f64=f64[i]["status"], niter_f64=f64[i]["niter"],
               rel_dev_vs_f64=float(abs(etot - mp.mpf(f64[i]["Etot"])) / abs(etot)))
    out.append(row)
    print(row, flush=True)
summary = dict(config="C4: anions Z=1..%d, N=Z+1, p=20, 40 cells, Q=G=1000, lda_x+lda_c_pw, Double64, scftol=1e-20" % Zs[-1],
               host_tables_s=t_tab, ms_per_step_median=float(np.median(steps) * 1e3), steps=len(steps),
               atom_iterations_per_s=float(len(Zs) / np.median(steps)), problems=out)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(summary, open("gpurun_out/dd_c4.json", "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != "problems"}))
