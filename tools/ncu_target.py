"""Small profiling target: a few SCF steps of a handful of atoms on the C2 discretization."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import build_disc
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from atomickohnsham_b200.sweep import periodic_table_models
Zs = list(range(1, 87)) if (len(sys.argv) > 2 and sys.argv[2] == "all") else [int(z) for z in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["18", "36", "54", "86"])]
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
disc = build_disc(0)
models, lhs, nhs = periodic_table_models(Zs)
bt = Batch(models, disc, aks.ODA(scftol=0.0, tinit=0.6), lh=lhs, nh=nhs)
for it in range(nsteps):
    r = bt.step()
print("done", [x["Etot"] for x in r], "launches", disc.ctx.launches)
