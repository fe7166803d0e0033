"""GPU diagnostic: Double64 LDA runs next to the Float64 path."""
import os
import sys
import time

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomickohnsham_b200 import KSEDiscretization, LDA, ODA, OptimizedAufbau, P1IntLegendreBasis, expmesh  # noqa: E402
from atomickohnsham_b200.dd import GaussLegendreDD  # noqa: E402
from atomickohnsham_b200.quadrature import GaussLegendre  # noqa: E402
from atomickohnsham_b200.solver import Batch  # noqa: E402

mp.mp.dps = 40
p, C1, Q = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 21, 300)
basis = P1IntLegendreBasis(expmesh(0, 60 if p < 20 else 300, C1, s=1.5), ordermax=p)
models = [LDA(Z=2, N=2), LDA(Z=10, N=10), LDA(Z=9, N=10)]
t0 = time.time()
dd_disc = KSEDiscretization(basis, models[0], lh=1, nh=4, dtype="double64", fem_integration_method=GaussLegendreDD(basis, Q))
f_disc = KSEDiscretization(basis, models[0], lh=1, nh=4, fem_integration_method=GaussLegendre(basis, Q))
auf = OptimizedAufbau(max_degen=2, tol=1e-3)
bdd = Batch(models, dd_disc, ODA(tinit=0.6, scftol=1e-18, aufbau=auf))
b64 = Batch(models, f_disc, ODA(tinit=0.6, scftol=1e-11, aufbau=auf))
print("setup", round(time.time() - t0, 2), flush=True)
for it in range(int(sys.argv[4]) if len(sys.argv) > 4 else 60):
    t1 = time.time()
    rd = bdd.step()
    dt = time.time() - t1
    rf = b64.step()
    print(it, f"{dt * 1e3:8.2f} ms", " | ".join(f"{a['status']} {a['Etot']:.15g} {a['stop']:.2e} t={a['t']:.6f} / {b['Etot']:.15g} {b['stop']:.1e} t={b['t']:.6f}"
                                                  for a, b in zip(rd, rf)), flush=True)
    if all(a["status"] != 0 for a in rd):
        break
r = bdd.records_dd()
for i in range(len(models)):
    print(i, [mp.nstr(mp.mpf(r[i, q, 0]) + mp.mpf(r[i, q, 1]), 32) for q in range(7)])
