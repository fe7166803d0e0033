"""Small Double64 runs for compute-sanitizer (memcheck / racecheck): sodium (Hartree only) and an LDA batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import atomickohnsham_b200 as aks  # noqa: E402
from atomickohnsham_b200.dd import GaussLegendreDD  # noqa: E402
from atomickohnsham_b200.solver import Batch  # noqa: E402

basis = aks.P1IntLegendreBasis(aks.expmesh(0, 60, 11, s=1.5), ordermax=5)
models = [aks.LDA(Z=2, N=2), aks.LDA(Z=4, N=4), aks.KSEModel(Z=11, N=11)]
disc = aks.KSEDiscretization(basis, models[0], lh=1, nh=3, dtype="double64", fem_integration_method=GaussLegendreDD(basis, 40))
bt = Batch(models, disc, aks.ODA(tinit=0.6, scftol=1e-18, aufbau=aks.OptimizedAufbau(max_degen=2, tol=1e-2)))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    rec = bt.step()
print([r["Etot"] for r in rec])
st = bt.state(1)
print(st["eps"][..., 0, 0])
bt.close()
