"""ncu target: a few Double64 steps of C4 (20 anions, order 20, LDA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import atomickohnsham_b200 as aks  # noqa: E402
from atomickohnsham_b200.solver import Batch  # noqa: E402

Zs = list(range(1, 21))
basis = aks.P1IntLegendreBasis(aks.expmesh(0, 300, 41, s=1.5), ordermax=20)
models = [aks.LDA(Z=z, N=z + 1) for z in Zs]
disc = aks.KSEDiscretization(basis, models[0], lh=2, nh=10, dtype="double64")
bt = Batch(models, disc, aks.ODA(tinit=0.6, scftol=0.0, aufbau=aks.OptimizedAufbau(max_degen=2, tol=1e-3)),
           lh=[2] * 20, nh=[min(max(z, 2), 10) for z in Zs])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    bt.step()
print("done")
