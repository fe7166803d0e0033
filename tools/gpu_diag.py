"""GPU diagnostic: step-level and trajectory-level comparison CUDA path vs oracle.
Run on the GPU box:  python tools/gpu_diag.py [case ...]  (prints a report)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_oracle, package_tables, make_mesh  # noqa: E402

import atomickohnsham_b200 as aks  # noqa: E402
from atomickohnsham_b200.solver import Batch  # noqa: E402

CASES = {
    "h_slater": {"Z": 1, "N": 1, "mesh": ["exp", 0, 300, 20, 1.5], "p": 10, "lh": 0, "nh": 3, "ex": "lda_x", "ec": None, "scftol": 1e-11},
    "o_lda": {"Z": 8, "N": 8, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10},
    "ar_lda": {"Z": 18, "N": 18, "mesh": ["exp", 0, 300, 41, 1.5], "p": 20, "lh": 2, "nh": 10, "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-10},
    "sc_rhf": {"Z": 21, "N": 21, "mesh": ["exp", 0, 2000, 40, 1.2], "p": 10, "lh": 2, "nh": 5, "ex": None, "ec": None, "scftol": 1e-10},
    "h_lsda": {"Z": 1, "N": 1, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 0, "nh": 2, "ex": "lda_x", "ec": "lda_c_pw", "nspin": 2, "scftol": 1e-10, "max_degen": 1},
}


def make_gpu(setup):
    basis, ops, quad = package_tables(setup["mesh"], setup["p"], setup.get("npoints", 1000))
    ns = setup.get("nspin", 1)
    mk = lambda name: aks.NoFunctional(ns) if name is None else aks.BuiltinFunctional(name, ns)
    model = aks.KSEModel(Z=setup["Z"], N=setup["N"], hartree=setup.get("hartree", 1), ex=mk(setup.get("ex")), ec=mk(setup.get("ec")))
    disc = aks.KSEDiscretization(basis, model, lh=setup["lh"], nh=setup["nh"], fem_integration_method=quad, femops=ops)
    alg = aks.ODA(scftol=setup.get("scftol", 1e-10), tinit=setup.get("tinit", 0.6), frozen_t=setup.get("frozen_t", False),
                  aufbau=aks.OptimizedAufbau(max_degen=setup.get("max_degen", 2), tol=setup.get("degen_tol", 1e-3)))
    return model, disc, alg


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def run_case(name, nsteps_oracle=3):
    setup = CASES[name]
    print(f"\n===== {name} =====", flush=True)
    t0 = time.time()
    o = make_oracle(setup)
    st = o.initial_state()
    for _ in range(nsteps_oracle):
        o.scf_step(st)
    print(f"oracle {nsteps_oracle} steps: Etot={st.energies.Etot!r} stop={st.stop:.3e}  ({time.time()-t0:.1f}s)", flush=True)
    model, disc, alg = make_gpu(setup)
    bt = Batch([model], disc, alg)
    e = st.energies
    bt.set_density(0, st.D, [e.Etot, e.Ekin, e.Ecou, e.Ehar, e.Eexc], niter=st.niter)
    Har_o, Vxc_o, W_o = o.hamiltonian_parts(st.D)
    Har_g = bt.assemble_hartree(0)
    Vxc_g = bt.assemble_vxc(0)
    print("Hartree matrix rel err:", rel(Har_g, Har_o), " |asym|:", float(np.max(np.abs(Har_g - Har_g.T))))
    if o.xc.active:
        print("Vxc matrix rel err   :", rel(Vxc_g, Vxc_o))
    print("Ehar  gpu/oracle:", bt.hartree_energy(0), o.hartree_energy(st.D))
    print("Eexc  gpu/oracle:", bt.exc_energy(0), o.exc_energy(st.D))
    t1 = time.time()
    eps_g, U_g, res_g = bt.eig_lowest(0)
    print(f"eig_lowest took {time.time()-t1:.3f}s")
    U_o, eps_o = o.find_orbital(Har_o, Vxc_o)
    print("eps gpu:\n", eps_g[:, :, 0])
    print("eps rel diff vs LAPACK (per channel max):", [float(np.max(np.abs(eps_g[l] - eps_o[l]) / np.abs(eps_o[l]))) for l in range(o.L)])
    print("max residual ||Hu-eMu||:", float(res_g.max()))
    # orthonormality of GPU vectors
    M0 = o.tb.M0
    for l in range(o.L):
        G = U_g[:, :, l, 0].T @ M0 @ U_g[:, :, l, 0]
        print(f"  l={l} |U^T M U - I|max = {np.max(np.abs(G - np.eye(G.shape[0]))):.2e}  overlap with oracle min = {np.min(np.abs(np.diag(U_g[:, :, l, 0].T @ M0 @ U_o[:, :, l, 0]))):.12f}")
    # one full step from the identical state
    det = {}
    st2 = o.scf_step(st, detail=det)
    rec = bt.step()[0]
    print("step record gpu   :", {k: rec[k] for k in ("niter", "status", "stop", "Etot", "Ekin", "Ecou", "Ehar", "Eexc", "t")})
    eo = st2.energies
    print("step record oracle:", dict(niter=st2.niter, stop=st2.stop, Etot=eo.Etot, Ekin=eo.Ekin, Ecou=eo.Ecou, Ehar=eo.Ehar, Eexc=eo.Eexc, t=st2.t))
    sg = bt.state(0)
    print("n gpu   :", sg["n"][:, :, 0].tolist())
    print("n oracle:", st2.n[:, :, 0].tolist())
    print("D rel err after step:", rel(sg["D"], st2.D))
    # full run from scratch
    bt.reset()
    t2 = time.time()
    final, hist = bt.run(100, history=True)
    tg = time.time() - t2
    t3 = time.time()
    so, ho = make_oracle(setup).groundstate()
    to = time.time() - t3
    f = final[0]
    print(f"RUN gpu   : niter={f['niter']} status={f['status']} Etot={f['Etot']!r} stop={f['stop']:.3e}  ({tg:.2f}s)")
    print(f"RUN oracle: niter={so.niter} Etot={so.energies.Etot!r} stop={so.stop:.3e}  ({to:.2f}s)")
    sg = bt.state(0)
    print("final eps gpu   :", sg["eps"][:, :4, 0].tolist())
    print("final eps oracle:", so.eps[:, :4, 0].tolist())
    print("final n gpu   :", sg["n"][:, :4, 0].tolist())
    print("final n oracle:", so.n[:, :4, 0].tolist())
    for i in range(min(len(hist[0]), len(ho), 40)):
        g, h = hist[0][i], ho[i]
        print(f"  it {i:2d} Etot g={g['Etot']:.12f} o={h['Etot']:.12f} d={g['Etot']-h['Etot']:+.2e}  t g={g['t']:.12f} o={h['t']:.12f}  stop g={g['stop']:.3e} o={h['stop']:.3e}")
    bt.close()


if __name__ == "__main__":
    from atomickohnsham_b200 import _capi
    ctx = _capi.default_context(0)
    print("FP64 DFMA peak (TFLOP/s):", ctx.fp64_peak_tflops())
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        try:
            run_case(nm)
        except Exception as ex:  # keep going: one report per call
            import traceback
            traceback.print_exc()
            print("CASE FAILED:", nm, ex)
