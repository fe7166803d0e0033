"""Where the end-to-end time goes (Batch(), aks_scf_run, state read-back), CUDA graph on / off, batch sizes 1 / 11 / 86."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
disc = build_disc(0); disc.handle()
ctx = disc.ctx
alg = aks.ODA(scftol=0.0, tinit=0.6)
for zs in ([18], list(range(30, 41)), list(range(1, 87))):
    models, lhs, nhs = periodic_table_models(zs)
    for graphs in (True, False):
        ctx.configure(use_graphs=graphs)
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); bt = Batch(models, disc, alg, lh=lhs, nh=nhs); t1 = time.perf_counter()
            bt.step(); t1b = time.perf_counter()      # first step: graph capture + instantiate when graphs are on
            final, hist = bt.run(40, history=True); t2 = time.perf_counter()
            outs = bt.states(); t3 = time.perf_counter()
            bt.close(); t4 = time.perf_counter()
            if rep == 2:
                print(f"atoms {len(zs):3d} graphs {int(graphs)}: create {1e3*(t1-t0):.2f} ms, first step {1e3*(t1b-t1):.2f} ms, run(40) {1e3*(t2-t1b):.2f} ms "
                      f"= {1e3*(t2-t1b)/40:.3f} ms/step, states {1e3*(t3-t2):.2f} ms, close {1e3*(t4-t3):.2f} ms", flush=True)
ctx.configure(use_graphs=True)
