import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import atomickohnsham_b200 as aks
from atomickohnsham_b200.solver import Batch
from gpu_diag import make_gpu
from helpers import make_oracle
setup = {"Z": 8, "N": 8, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10,
         "aufbau_kind": "smeared", "temp": 0.05}
o = make_oracle(setup)
model, disc, _ = make_gpu(setup)
bt = Batch([model], disc, aks.ODA(scftol=0.0, tinit=0.6, aufbau=aks.SmearedAufbau(Temp=0.05)))
st = o.initial_state()
for it in range(4):
    o.scf_step(st)
    rec = bt.step()[0]
    cnt, its, bad = bt.eig_info()
    cl = bt.eig_clusters()
    sg = bt.state(0, want_D=False, want_U=False)
    print(it, "dE", rec["Etot"] - st.energies.Etot, "clusters", np.argwhere(cl[0, 0]).tolist(), "bad", np.argwhere(bad[0, 0]).tolist())
    print("   counts", cnt[0, 0].tolist(), "its", its[0, 0].tolist())
    print("   deps", np.array2string(sg["eps"][:, :, 0] - st.eps[:, :, 0], precision=2))
    print("   dn", np.array2string(sg["n"][:, :, 0] - st.n[:, :, 0], precision=2))
