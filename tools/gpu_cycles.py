"""Cycle breakdown of k_eig (library must be built with -DAKS_EIG_TIMING)."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200 import _capi
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
Zs = list(range(1, 87))
disc = build_disc(0)
models, lhs, nhs = periodic_table_models(Zs)
os.environ["AKS_PHASE_TIMERS"] = "1"
bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs, eig_margin=-1)
for it in range(14):
    bt.step()
    if it in (0, 3, 8, 13):
        cnt, its, bad = bt.eig_info()
        buf = np.zeros(len(Zs) * 4 * 10 * 8, dtype=np.int64)
        _capi.lib.aks_debug_eig_cycles(bt.h, buf.ctypes.data_as(C.POINTER(C.c_int64)))
        cyc = buf.reshape(len(Zs), 1, 4, 10, 8)
        act = np.zeros(cnt.shape, dtype=bool)
        for i in range(len(Zs)):
            act[i, :, :lhs[i] + 1, :nhs[i]] = True
        nf = (cnt + its)[act]
        g = lambda j: cyc[..., j][act]
        print(f"it {it}: eig phase {bt.phase_ms()[1]:.2f} ms; counts {cnt[act].mean():.1f} its {its[act].mean():.2f}; per item cycles: setup {g(0).mean():.0f} factor {g(1).mean():.0f} "
              f"(per call {g(1).sum()/nf.sum():.0f}) solve {g(2).mean():.0f} (per call {g(2).sum()/max(its[act].sum(),1):.0f}) apply+dots {g(3).mean():.0f} (per rqi {g(3).sum()/max(its[act].sum(),1):.0f}) "
              f"[factor split: cells {g(6).sum()/nf.sum():.0f} hats {g(7).sum()/nf.sum():.0f} per call] final {g(4).mean():.0f} total {g(5).mean():.0f} max {g(5).max()};  sum/296 = {g(5).sum()/296/1.9e6:.2f} ms", flush=True)
