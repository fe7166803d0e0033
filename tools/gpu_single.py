"""SURVEY 8(d) metric (1): SCF iterations/s of ONE atom (batch 1) for configs C1-C3, and the time of a full
run to convergence.  Run on the GPU box; prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from gpu_diag import make_gpu
from atomickohnsham_b200.solver import Batch
import atomickohnsham_b200 as aks

CONFIGS = {
    "C1_H_slater_p10": {"Z": 1, "N": 1, "mesh": ["exp", 0, 300, 20, 1.5], "p": 10, "lh": 0, "nh": 3, "ex": "lda_x", "ec": None, "scftol": 1e-11},
    "C2_Ar_lda_p20": {"Z": 18, "N": 18, "mesh": ["exp", 0, 300, 41, 1.5], "p": 20, "lh": 2, "nh": 10, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10},
    "C3_Sc_lsda_p20": {"Z": 21, "N": 21, "mesh": ["exp", 0, 2000, 40, 1.2], "p": 20, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "nspin": 2, "scftol": 1e-10},
}
out = {}
for name, setup in CONFIGS.items():
    model, disc, alg = make_gpu(setup)
    fixed = aks.ODA(scftol=0.0, tinit=0.6, aufbau=alg.aufbau)
    bt = Batch([model], disc, fixed)
    for _ in range(5):
        bt.step_async()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 40
    for _ in range(K):
        bt.step_async()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    bt.close()
    bt = Batch([model], disc, alg)
    t0 = time.perf_counter()
    final, _ = bt.run(100)
    trun = time.perf_counter() - t0
    out[name] = {"ms_per_iteration": dt / K * 1e3, "iterations_per_s": K / dt, "run_to_convergence_s": trun,
                 "niter": final[0]["niter"], "status": final[0]["status"], "Etot": final[0]["Etot"]}
    bt.close()
print(json.dumps(out))
