"""Is a problem's trajectory independent of the batch it runs in?  One batch of 86 vs LPT halves, step by step."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models, partition
disc = build_disc(0)
models, lhs, nhs = periodic_table_models(range(1, 87))
alg = aks.ODA(scftol=1e-10, tinit=0.4, aufbau=aks.OptimizedAufbau(tol=1e-2))
costs = [(lhs[i] + 1) * nhs[i] + 4.0 for i in range(86)]
parts = partition(costs, 2)
full = Batch(models, disc, alg, lh=lhs, nh=nhs)
halves = [Batch([models[i] for i in p], disc, alg, lh=[lhs[i] for i in p], nh=[nhs[i] for i in p]) for p in parts]
first = {}
for it in range(100):
    rf = full.step()
    rh = [h.step() for h in halves]
    for p, rr in zip(parts, rh):
        for j, i in enumerate(p):
            if i not in first and (rf[i]["Etot"] != rr[j]["Etot"] or rf[i]["stop"] != rr[j]["stop"] or rf[i]["t"] != rr[j]["t"]):
                first[i] = (it, rf[i]["Etot"], rr[j]["Etot"], rf[i]["t"], rr[j]["t"], rf[i]["stop"], rr[j]["stop"])
print("problems whose records differ between the two batchings:", len(first))
for i in sorted(first)[:12]:
    print("  Z=%d first at step %d: Etot %r vs %r, t %r vs %r, stop %.3e vs %.3e" % ((i + 1,) + first[i]))
print("steps of first divergence:", sorted(set(v[0] for v in first.values())))
