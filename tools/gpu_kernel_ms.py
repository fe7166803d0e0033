"""Per-kernel device time of a steady-state 86-atom step (AKS_PHASE_TIMERS: one stream, plain launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
disc = build_disc(0)
models, lhs, nhs = periodic_table_models(range(1, 87))
os.environ["AKS_PHASE_TIMERS"] = "1"
bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs)
for _ in range(10):
    bt.step()
acc = {}
for _ in range(8):
    bt.step()
    for k, v in bt.kernel_ms():
        acc[k] = acc.get(k, 0.0) + v / 8
print(os.path.basename(os.environ.get("AKS_B200_LIB", "libaks_b200.so")), "sum %.4f ms:" % sum(acc.values()),
      " ".join("%s=%.4f" % kv for kv in sorted(acc.items(), key=lambda kv: -kv[1])))
