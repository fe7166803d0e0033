"""Cycle stamps of k_linesearch (library built with -DAKS_LS_TIMING): load grids | orbital sums + corr | degenerate search /
trial Exc | ODA search."""
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
disc = build_disc(0)
for label, Zs, alg in (("Ar, ODA", [18], aks.ODA(scftol=0.0, tinit=0.4, aufbau=aks.OptimizedAufbau(tol=1e-2))),
                       ("Fe, frozen_t", [26], aks.ODA(scftol=0.0, tinit=0.15, frozen_t=True, aufbau=aks.OptimizedAufbau(tol=1e-2))),
                       ("86 atoms, ODA", list(range(1, 87)), aks.ODA(scftol=0.0, tinit=0.6))):
    models, lhs, nhs = periodic_table_models(Zs)
    bt = Batch(models, disc, alg, lh=lhs, nh=nhs)
    for _ in range(30):
        bt.step()
    buf = np.zeros(len(Zs) * 4 * 10 * 8, dtype=np.int64)
    _capi.lib.aks_debug_eig_cycles(bt.h, buf.ctypes.data_as(C.POINTER(C.c_int64)))
    st = buf[:len(Zs) * 8].reshape(len(Zs), 8)
    d = np.diff(st[:, :5], axis=1)
    print(label, "cycles per phase (mean / max over atoms): load grids %d/%d  sums+corr %d/%d  degenerate search or trial Exc %d/%d  ODA search %d/%d  total %d/%d"
          % (d[:, 0].mean(), d[:, 0].max(), d[:, 1].mean(), d[:, 1].max(), d[:, 2].mean(), d[:, 2].max(), d[:, 3].mean(), d[:, 3].max(),
             (st[:, 4] - st[:, 0]).mean(), (st[:, 4] - st[:, 0]).max()), flush=True)
    bt.close()
