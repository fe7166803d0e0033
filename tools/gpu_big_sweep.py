"""The 1014-problem list (atoms x ions x functionals) of bench.py to convergence on one GPU, three times, and the split of one
run into batch creation / run loop / state copy (what `sweeps_to_convergence.atoms_x_ions_x_functionals` times)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import build_disc, sweep_lists
from atomickohnsham_b200.sweep import solve_sweep
from atomickohnsham_b200.solver import Batch
import atomickohnsham_b200 as aks
disc = build_disc(0)
(c5m, c5l, c5k), (bigm, bigl, bigk) = sweep_lists()
solve_sweep(bigm[:8], disc, lh=bigl[:8], nh=bigk[:8], remedy_maxiter=0, want_arrays=False)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sols = solve_sweep(bigm, disc, lh=bigl, nh=bigk, remedy_maxiter=0, want_arrays=False)
    torch.cuda.synchronize()
    print("sweep", rep, time.perf_counter() - t0, flush=True)
alg = aks.ODA(scftol=1e-10, tinit=0.4, aufbau=aks.OptimizedAufbau(tol=1e-2))
for rep in range(2):
    t0 = time.perf_counter(); bt = Batch(bigm, disc, alg, lh=bigl, nh=bigk); t1 = time.perf_counter()
    final, _ = bt.run(100, history=False, raise_on_problem_error=False); t2 = time.perf_counter()
    st = bt.states(); t3 = time.perf_counter(); bt.close(); t4 = time.perf_counter()
    print(f"create {t1-t0:.3f} run {t2-t1:.3f} states {t3-t2:.3f} close {t4-t3:.3f}", flush=True)
