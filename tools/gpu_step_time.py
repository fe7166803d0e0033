"""Device time per 86-atom SCF step (fixed work, scftol = 0) -- the number bench.py reports as ms_per_step, alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
nsweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
disc = build_disc(0); disc.handle()
ctx = disc.ctx
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
models, lhs, nhs = periodic_table_models(range(1, 87))
models, lhs, nhs = models * nsweeps, lhs * nsweeps, nhs * nsweeps
out = []
for rep in range(3):
    bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs)
    for _ in range(10):
        bt.step_async()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(100):
        bt.step_async()
    e1.record(stream)
    torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / 100)
    bt.close()
print("lanes=%s lib=%s atoms=%d: ms/step %s  -> %.0f atom-iterations/s" % (os.environ.get("AKS_STREAM_LANES", "8"), os.path.basename(os.environ.get("AKS_B200_LIB", "libaks_b200.so")),
      len(models), " ".join("%.4f" % v for v in out), len(models) / (min(out) * 1e-3)))
