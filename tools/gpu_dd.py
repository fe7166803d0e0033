"""GPU diagnostic: the reference's Double64 sodium run (tests/golden/sodium_double64.json) through the C ABI,
iteration by iteration, with the relative deviation of every energy from the reference's log."""
import json
import os
import sys
import time

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atomickohnsham_b200 import KSEDiscretization, KSEModel, ODA, OptimizedAufbau, P1IntLegendreBasis, expmesh  # noqa: E402
from atomickohnsham_b200.solver import Batch  # noqa: E402

mp.mp.dps = 50
g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "sodium_double64.json")))
s = g["setup"]
t0 = time.time()
mesh = expmesh(s["mesh"][1], s["mesh"][2], s["mesh"][3], s=s["mesh"][4])
basis = P1IntLegendreBasis(mesh, ordermax=s["p"])
model = KSEModel(Z=s["Z"], N=s["N"])
disc = KSEDiscretization(basis, model, lh=s["lh"], nh=s["nh"], dtype="double64")
alg = ODA(tinit=s["tinit"], scftol=s["scftol"], aufbau=OptimizedAufbau(max_degen=s["max_degen"], tol=s["degen_tol"]))
bt = Batch([model], disc, alg)
print("setup", round(time.time() - t0, 2), "s", flush=True)
names = ["stop", "Etot", "Ekin", "Ecou", "Ehar", "Eexc", "t"]
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 100):
    t1 = time.time()
    r = bt.step()[0]
    dt = time.time() - t1
    rd = bt.records_dd()[0]
    vals = {nm: mp.mpf(rd[i, 0]) + mp.mpf(rd[i, 1]) for i, nm in enumerate(names)}
    line = f"it {it:2d} st {r['status']} {dt * 1e3:7.2f} ms stop {mp.nstr(vals['stop'], 6)} Etot {mp.nstr(vals['Etot'], 34)} t {mp.nstr(vals['t'], 12)}"
    if it < len(g["iterations"]):
        ref = g["iterations"][it]
        dev = " ".join(f"{nm}:{mp.nstr(abs(vals[nm] - mp.mpf(ref[nm])) / abs(mp.mpf(ref[nm])), 3)}"
                       for nm in ("Etot", "Ekin", "Ecou", "Ehar"))
        line += f" | ref t {ref['t'][:14]} stop {ref['stop'][:10]} | rel {dev}"
    print(line, flush=True)
    if r["status"] != 0:
        break
st = bt.state(0)
eps = st["eps"]
n = st["n"]
for l in range(s["lh"] + 1):
    for k in range(3):
        print("eps", l, k, mp.nstr(mp.mpf(eps[l, k, 0, 0]) + mp.mpf(eps[l, k, 0, 1]), 34), "n", n[l, k, 0, 0])
print("final ref", g["final"])
