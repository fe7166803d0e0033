"""Which eigenpairs are the slowest CTAs of k_eig?  (library built with -DAKS_EIG_TIMING)"""
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
margin = int(sys.argv[1]) if len(sys.argv) > 1 else 1
bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs, eig_margin=margin)
prev_rf = np.zeros(len(Zs), int)
for it in range(16):
    buf0 = np.zeros(len(Zs) * 4 * 10 * 8, dtype=np.int64)
    bt.step()
    rf = bt.eig_refills()
    if it in (6, 9, 12, 15):
        cnt, its, bad = bt.eig_info()
        buf = np.zeros(len(Zs) * 4 * 10 * 8, dtype=np.int64)
        _capi.lib.aks_debug_eig_cycles(bt.h, buf.ctypes.data_as(C.POINTER(C.c_int64)))
        cyc = buf.reshape(len(Zs), 4, 10, 8)
        tot = cyc[..., 5]
        order = np.dstack(np.unravel_index(np.argsort(-tot, axis=None), tot.shape))[0][:12]
        print(f"it {it}: refilled atoms this step: {[Zs[i] for i in np.nonzero(rf - prev_rf)[0]]}")
        for (i, l, k) in order:
            print(f"   Z={Zs[i]:2d} l={l} k={k}: total {tot[i,l,k]:7d} cyc  setup {cyc[i,l,k,0]:6d} factor {cyc[i,l,k,1]:7d} solve {cyc[i,l,k,2]:6d} "
                  f"apply {cyc[i,l,k,3]:6d}  counts {cnt[i,0,l,k]} rqi {its[i,0,l,k]} bad {bad[i,0,l,k]}")
        nz = tot[tot > 0]
        print(f"   levels timed {nz.size}, mean {nz.mean():.0f}, median {np.median(nz):.0f}, p90 {np.percentile(nz,90):.0f}, max {nz.max()}")
    prev_rf = rf.copy()
