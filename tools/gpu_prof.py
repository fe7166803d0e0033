"""Eigensolver statistics + failing-atom report for the C5 sweep (run on the GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models

disc = build_disc(0)
models, lhs, nhs = periodic_table_models(range(1, 87))
os.environ["AKS_PHASE_TIMERS"] = "1"
bt = Batch(models, disc, aks.ODA(scftol=1e-10, tinit=0.6), lh=lhs, nh=nhs)
for it in range(12):
    recs = bt.step()
    cnt, its, bad = bt.eig_info()
    act = np.zeros_like(cnt, dtype=bool)
    for i in range(86):
        act[i, :, :lhs[i] + 1, :nhs[i]] = True
    ph = bt.phase_ms()
    print(f"it {it}: eig {ph[1]:.2f} ms  counts mean {cnt[act].mean():.1f} max {cnt[act].max()}  rqi its mean {its[act].mean():.2f} max {its[act].max()}  notconv {bad[act].sum()}  phases {[round(x,3) for x in ph[:7]]}", flush=True)
    if it in (3, 8):
        print("   counts histogram:", np.bincount(cnt[act].ravel())[:40].tolist())
        print("   its histogram   :", np.bincount(its[act].ravel())[:16].tolist())
final, hist = bt.run(150, history=True)
bad = [(i + 1, r["status"], r["niter"], r["stop"]) for i, r in enumerate(final) if r["status"] != 1]
print("not converged:", bad)
for Z, *_ in bad[:6]:
    h = hist[Z - 1]
    print(f"Z={Z} tail:", [(r["niter"], f"{r['stop']:.2e}", f"{r['t']:.4f}", f"{r['Etot']:.10f}") for r in h[-6:]])
    st = bt.state(Z - 1, want_D=False, want_U=False)
    print("   n:", st["n"][:, :, 0][st["n"][:, :, 0] != 0].tolist(), " eps occ:", st["eps"][:, :, 0][st["n"][:, :, 0] != 0].tolist())
