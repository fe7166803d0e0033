"""Diagnose AKS_EEIG: which eigenpairs fail the certificate in the maximum-ionization workload (fractional anions)."""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import basis_size
from helpers import package_tables

basis, ops, quad = package_tables(["exp", 0, 300, 30, 1.3], 10, 1000)
disc = aks.KSEDiscretization(basis, None, lh=1, nh=6, nspin=1, fem_integration_method=quad, femops=ops)
alg = aks.ODA(scftol=1e-6, tinit=0.4, aufbau=aks.OptimizedAufbau(tol=1e-2))
zs = [1, 2, 3, 9]
lo = {z: float(z) for z in zs}
hi = {z: z + 1.2 for z in zs}
rnd = 0
while max(hi[z] - lo[z] for z in zs) > 1e-2:
    rnd += 1
    mid = {z: (hi[z] + lo[z]) / 2 for z in zs}
    models = [aks.LDA(Z=z, N=mid[z]) for z in zs]
    lhs = [min(min(math.ceil(mid[z] / 20), 3), 1) for z in zs]
    nhs = [min(basis_size(z)[1], 6) for z in zs]
    bt = Batch(models, disc, alg, lh=lhs, nh=nhs)
    failed = False
    for it in range(50):
        recs = bt.step()
        cnt, its, bad = bt.eig_info()
        for i in range(4):
            if recs[i]["status"] < 0:
                failed = True
                st = bt.state(i, want_D=False, want_U=False)
                print(f"round {rnd} Z={zs[i]} N={mid[zs[i]]} it {it} status {recs[i]['status']} stop {recs[i]['stop']:.2e} lh {lhs[i]} nh {nhs[i]}")
                print("   flagged (l,k):", np.argwhere(bad[i, 0]).tolist(), "counts", cnt[i, 0][bad[i, 0]].tolist(), "its", its[i, 0][bad[i, 0]].tolist())
                print("   counts all:", cnt[i, 0].tolist(), " its all:", its[i, 0].tolist())
                print("   eps:", np.array2string(st["eps"][:, :, 0], precision=14))
                print("   n  :", np.array2string(st["n"][:, :, 0], precision=6))
        if failed or all(r["status"] != 0 for r in recs):
            break
    states = bt.states()
    for i, z in enumerate(zs):
        st = states[i]
        occ = st["n"] != 0
        homo = float(st["eps"][occ].max())
        stable = homo <= 0.0 and recs[i]["status"] == 1
        if stable: lo[z] = mid[z]
        else: hi[z] = mid[z]
    bt.close()
    print(f"round {rnd}: mid {mid} status {[r['status'] for r in recs]}", flush=True)
