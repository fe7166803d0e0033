"""Which eigenpairs of the Float64 path fail the certificate on the C4 anions (order 20)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
Zs = list(range(1, 21))
models = [aks.LDA(Z=z, N=z + 1) for z in Zs]
lh = [min(-(-z // 20), 3) for z in Zs]
nh = [min(max(z, 2), 10) for z in Zs]
disc = build_disc(0)
bt = Batch(models, disc, aks.ODA(tinit=0.6, scftol=1e-10, aufbau=aks.OptimizedAufbau(max_degen=2, tol=1e-3)), lh=lh, nh=nh)
seen = set()
for it in range(120):
    recs = bt.step()
    cnt, its, bad = bt.eig_info()
    for i, z in enumerate(Zs):
        if recs[i]["status"] < 0 and z not in seen:
            seen.add(z)
            st = bt.state(i, want_D=False, want_U=False)
            print(f"Z={z} it {it} status {recs[i]['status']} stop {recs[i]['stop']:.2e}")
            for l, k in np.argwhere(bad[i, 0]).tolist():
                e = st["eps"][l, :, 0]
                print(f"   flagged l={l} k={k} counts {cnt[i,0,l,k]} its {its[i,0,l,k]}  eps[k-1..k+1] = {[repr(float(v)) for v in e[max(0,k-1):k+2]]}  gaps {np.diff(e[max(0,k-1):k+2]).tolist()}")
            print("   n:", st["n"][:, :, 0][st["n"][:, :, 0] != 0].tolist())
print("failed:", sorted(seen), " final statuses:", [r["status"] for r in recs])
