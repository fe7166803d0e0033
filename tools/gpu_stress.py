import sys, os, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
disc = build_disc(0); disc.handle()
models, lhs, nhs = periodic_table_models(range(1, 87))
alg = aks.ODA(scftol=1e-10, tinit=0.6)
free0 = torch.cuda.mem_get_info()[0]
ref = None
t0 = time.time()
for rep in range(60):
    bt = Batch(models, disc, alg, lh=lhs, nh=nhs)
    for _ in range(6):
        recs = bt.step()
    sig = tuple(r["Etot"] for r in recs)
    if ref is None: ref = sig
    assert sig == ref, "not reproducible across batches"
    bt.close()
free1 = torch.cuda.mem_get_info()[0]
print("60 create/step/destroy cycles ok, bit-identical; time", round(time.time()-t0,2), "s; device memory held by the pool:", round((free0-free1)/2**20), "MiB")
# two batches alive at once on the same ctx
b1 = Batch(models[:40], disc, alg, lh=lhs[:40], nh=nhs[:40]); b2 = Batch(models[40:], disc, alg, lh=lhs[40:], nh=nhs[40:])
for _ in range(6):
    b1.step_async(); b2.step_async()
r1 = b1.step(); r2 = b2.step()
bt = Batch(models, disc, alg, lh=lhs, nh=nhs)
for _ in range(7): r = bt.step()
assert [x["Etot"] for x in r1 + r2] == [x["Etot"] for x in r], "split batches differ from the whole"
print("two concurrent batches == one batch, bit-identical")
