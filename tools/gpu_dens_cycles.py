"""Cycle breakdown of k_dens_cells_tc per CTA (library built with -DAKS_DENS_TIMING).

  nvcc $(python -c "from atomickohnsham_b200.build import NVCC_FLAGS; print(' '.join(NVCC_FLAGS))") -DAKS_DENS_TIMING \
       -o tools/micro/libaks_dens_timing.so atomickohnsham_b200/csrc/aks_b200.cu
  AKS_B200_LIB=tools/micro/libaks_dens_timing.so python tools/gpu_dens_cycles.py
"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["AKS_STREAM_LANES"] = "1"
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200 import _capi
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
Zs = list(range(1, 87))
disc = build_disc(0)
models, lhs, nhs = periodic_table_models(Zs)
bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs)
for it in range(10):
    bt.step()
buf = np.zeros(len(Zs) * 4 * 10 * 8, dtype=np.int64)
_capi.lib.aks_debug_eig_cycles(bt.h, buf.ctypes.data_as(C.POINTER(C.c_int64)))
st = buf[:64 * 10 * 8].reshape(64, 10, 8)   # 10 CTAs of 4 cells per problem
d = np.diff(st[:, :, :3], axis=2)
names = ["list + barriers", "B fragments + node groups (warp 0)"]
for zsel, label in ((slice(0, 36), "Z = 1..36 (one orbital tile)"), (slice(36, 64), "Z = 37..64 (two tiles)")):
    tot = np.mean(st[zsel, :, 2] - st[zsel, :, 0])
    print(f"k_dens_cells_tc per CTA, {label}: total {tot:.0f} cycles")
    for k, nm in enumerate(names):
        print(f"  {nm:38s} mean {d[zsel, :, k].mean():8.0f}  max {d[zsel, :, k].max():8.0f}   {100 * d[zsel, :, k].mean() / tot:5.1f} %")
