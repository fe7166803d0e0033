"""Per-kernel device time of a step of SMALL batches (one lane, plain launches): where the single-atom latency goes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
disc = build_disc(0)
os.environ["AKS_PHASE_TIMERS"] = "1"
for label, Zs, alg in (("Ar alone, ODA", [18], aks.ODA(scftol=0.0, tinit=0.4, aufbau=aks.OptimizedAufbau(tol=1e-2))),
                       ("Fe alone, frozen_t", [26], aks.ODA(scftol=0.0, tinit=0.15, frozen_t=True, aufbau=aks.OptimizedAufbau(tol=1e-2))),
                       ("21 d/f atoms, frozen_t", [22, 23, 24, 25, 26, 27, 28, 41, 44, 45, 58, 59, 60, 61, 62, 64, 65, 66, 67, 68, 78],
                        aks.ODA(scftol=0.0, tinit=0.15, frozen_t=True, aufbau=aks.OptimizedAufbau(tol=1e-2)))):
    models, lhs, nhs = periodic_table_models(Zs)
    bt = Batch(models, disc, alg, lh=lhs, nh=nhs)
    for _ in range(30):
        bt.step()
    acc = {}
    for _ in range(8):
        bt.step()
        for k, v in bt.kernel_ms():
            acc[k] = acc.get(k, 0.0) + v / 8
    print(label, "sum %.4f ms:" % sum(acc.values()), " ".join("%s=%.4f" % kv for kv in sorted(acc.items(), key=lambda kv: -kv[1])), flush=True)
    bt.close()
