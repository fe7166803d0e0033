"""Device time of every SCF step of a cold-started 86-atom batch (8 lanes), and the per-kernel split (one lane) at a few steps:
where the first 25 iterations (what bench.py's `value` covers with --warmup 5 --steps 20) differ from the settled ones."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
disc = build_disc(0); disc.handle()
stream = torch.cuda.ExternalStream(disc.ctx.stream, device=torch.device("cuda", 0))
models, lhs, nhs = periodic_table_models(range(1, 87))
bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(61)]
ev[0].record(stream)
for i in range(60):
    bt.step_async()
    ev[i + 1].record(stream)
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(60)]
print("ms per step:", " ".join("%.2f" % v for v in ms))
print("mean 5..24: %.3f   mean 40..59: %.3f" % (sum(ms[5:25]) / 20, sum(ms[40:60]) / 20))
bt.close()
os.environ["AKS_PHASE_TIMERS"] = "1"
bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs)
for i in range(45):
    bt.step()
    if i in (6, 12, 20, 44):
        k = dict(bt.kernel_ms())
        top = sorted(k.items(), key=lambda kv: -kv[1])[:12]
        print("step %2d (one lane) sum %.3f:" % (i, sum(k.values())), " ".join("%s=%.3f" % kv for kv in top))
