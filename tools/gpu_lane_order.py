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
def run(order, label):
    m = [models[i] for i in order]; l = [lhs[i] for i in order]; k = [nhs[i] for i in order]
    out = []
    for rep in range(2):
        bt = Batch(m, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=l, nh=k)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for _ in range(5): bt.step_async()
        ev[0].record(stream)
        for _ in range(20): bt.step_async()
        ev[1].record(stream)
        for _ in range(100): bt.step_async()
        ev[2].record(stream)
        torch.cuda.synchronize()
        out.append((ev[0].elapsed_time(ev[1]) / 20, ev[1].elapsed_time(ev[2]) / 100))
        bt.close()
    print(label, "steps 5-24: %.4f %.4f ms   steps 25-124: %.4f %.4f ms" % (out[0][0], out[1][0], out[0][1], out[1][1]), flush=True)
nat = list(range(86))
run(nat, "natural order Z=1..86      ")
# interleave: heaviest first dealt round-robin into 8 lanes, lanes concatenated
cost = [(lhs[i] + 1) * nhs[i] + 4.0 for i in nat]
srt = sorted(nat, key=lambda i: -cost[i])
buckets = [[] for _ in range(8)]
for j, i in enumerate(srt):
    buckets[j % 8 if (j // 8) % 2 == 0 else 7 - j % 8].append(i)
inter = [i for b in buckets for i in b]
run(inter, "interleaved (snake over 8)  ")
run(list(reversed(nat)), "reversed                    ")
