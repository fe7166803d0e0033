"""Cycle breakdown of k_potential per CTA (library built with -DAKS_POT_TIMING: see the command in the docstring).

  nvcc $(python -c "from atomickohnsham_b200.build import NVCC_FLAGS; print(' '.join(NVCC_FLAGS))") -DAKS_POT_TIMING \
       -o tools/micro/libaks_pot_timing.so atomickohnsham_b200/csrc/aks_b200.cu
  AKS_B200_LIB=tools/micro/libaks_pot_timing.so [AKS_POTENTIAL_TC=0] python tools/gpu_pot_cycles.py
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
ncta = 40 * ((len(Zs) + 3) // 4)
st = buf[:ncta * 8].reshape(ncta, 8)
d = np.diff(st[:, :6], axis=1)
tc = os.environ.get("AKS_POTENTIAL_TC", "1") != "0"
if tc:   # k_potential_tc: 4 stamps
    names = ["setup", "contraction (DMMA)", "tiles -> Kxc"]
    last = 3
else:    # k_potential: 6 stamps
    names = ["load f", "contraction (DFMA)", "reduction", "hartree F.W", "H assembly"]
    last = 5
tot = np.mean(st[:, last] - st[:, 0])
print(f"{'k_potential_tc' if tc else 'k_potential'} per CTA ({ncta} CTAs): total {tot:.0f} cycles")
for k, nm in enumerate(names):
    print(f"  {nm:20s} mean {d[:, k].mean():8.0f}  max {d[:, k].max():8.0f}   {100 * d[:, k].mean() / tot:5.1f} %")
