import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from helpers import make_oracle
from gpu_diag import make_gpu, CASES
from atomickohnsham_b200.solver import Batch
np.set_printoptions(linewidth=200, precision=10)
for name in sys.argv[1:]:
    setup = CASES[name]
    o = make_oracle(setup); st = o.initial_state(); o.scf_step(st)
    model, disc, alg = make_gpu(setup)
    bt = Batch([model], disc, alg)
    rec = bt.step()[0]
    sg = bt.state(0, want_D=False, want_U=False)
    cnt, its, bad = bt.eig_info()
    print(name, "Etot gpu", rec["Etot"], "oracle", st.energies.Etot)
    print("eps gpu\n", sg["eps"][:, :, 0]); print("eps oracle\n", st.eps[:, :, 0])
    print("count rounds\n", cnt[0, 0]); print("rqi its\n", its[0, 0]); print("flags\n", bad[0, 0].astype(int))
    print("n gpu", sg["n"][:, :, 0].tolist()); print("n oracle", st.n[:, :, 0].tolist())
