"""Where the end-to-end time of bench.py's e2e leg goes (Batch(), every step, state read-back)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
disc = build_disc(0); disc.handle()
models, lhs, nhs = periodic_table_models(range(1, 87))
alg = aks.ODA(scftol=0.0, tinit=0.6)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); bt = Batch(models, disc, alg, lh=lhs, nh=nhs); t1 = time.perf_counter()
    ts = []
    for _ in range(20):
        a = time.perf_counter(); bt.step(); ts.append(time.perf_counter() - a)
    t2 = time.perf_counter()
    outs = [bt.state(i, want_D=False, want_U=False) for i in range(86)]
    t3 = time.perf_counter()
    bt.close(); t4 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.2f} ms, steps {1e3*(t2-t1):.2f} ms ({' '.join(f'{1e3*x:.2f}' for x in ts)}), state {1e3*(t3-t2):.2f} ms, close {1e3*(t4-t3):.2f} ms, total {1e3*(t4-t0):.2f}")
