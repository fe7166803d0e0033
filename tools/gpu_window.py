"""Diagnostic: eigenpair window vs all-nh solve on a subset of the sweep (run on the GPU box)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models

disc = build_disc(0)
zs = list(range(1, 87))
models, lhs, nhs = periodic_table_models(zs)
alg = aks.ODA(scftol=0.0, tinit=0.6)
nst = int(sys.argv[1]) if len(sys.argv) > 1 else 12
runs = {}
for margin in (-1, 0, 1, 2):
    bt = Batch(models, disc, alg, lh=lhs, nh=nhs, eig_margin=margin)
    recs, rf = [], []
    for _ in range(nst):
        recs.append(bt.step())
        rf.append(bt.eig_refills().copy())
    runs[margin] = (recs, np.array(rf))
    bt.close()
ref = runs[-1][0]
for margin in (0, 1, 2):
    recs, rf = runs[margin]
    print("margin", margin, "refills per step:", np.diff(np.vstack([np.zeros(len(zs), int), rf]), axis=0).sum(axis=1))
    print("  atoms refilled most:", [(zs[i], int(rf[-1][i])) for i in np.argsort(-rf[-1])[:15]])
    for it in range(nst):
        bad = [(zs[i], recs[it][i]["Etot"] - ref[it][i]["Etot"]) for i in range(len(zs))
               if abs(recs[it][i]["Etot"] - ref[it][i]["Etot"]) > 1e-9 * abs(ref[it][i]["Etot"])]
        if bad:
            print("  it", it, "Etot mismatches:", bad[:8])
