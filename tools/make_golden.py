"""Extract the reference's committed golden values into tests/golden/*.json.

Run in the build container only (reads /root/reference, which does not exist on
the GPU box).  Nothing is computed here: values are copied from
  * test/reference_data/fem_matrices.jld2   (HDF5; doubles located by scanning)
  * test/hydrogen.jl, test/scandium.jl, README.md   (asserted constants)
  * examples/*/log.txt, results.txt                  (committed trajectories)
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def fem_matrices():
    b = open(f"{REF}/test/reference_data/fem_matrices.jld2", "rb").read()
    a = np.frombuffer(b[: len(b) // 8 * 8], dtype="<f8")
    # offsets found by scanning for runs of finite doubles
    blocks = {"M0": (672, 9), "A": (856, 9), "T0": (1048, 27), "M2": (1376, 9),
              "M1": (1560, 9), "T1": (1752, 27), "T2": (2088, 27)}
    out = {"basis": {"mesh": "polynomialmesh(0,1,3;s=0.9)", "ordermax": 2},
           "source": "test/reference_data/fem_matrices.jld2 (test/fem.jl:147-174), column-major"}
    for k, (off, n) in blocks.items():
        out[k] = [float(v) for v in a[off // 8: off // 8 + n]]
    assert abs(out["M0"][0] - 4 / 3) < 1e-15
    return out


def parse_log(path):
    recs = []
    for line in open(path):
        if line.startswith("iter ="):
            kv = dict(re.findall(r"(\w+) = ([-+0-9.eE]+)", line))
            recs.append({k: (int(v) if k == "iter" else float(v)) for k, v in kv.items()})
    return recs


def parse_results(path):
    txt = open(path).read()
    out = {"niter": int(re.search(r"niter\s+= (\d+)", txt).group(1))}
    for k, lab in (("Ekin", "Ekin"), ("Ecou", "Ecou"), ("Ehar", "Ehar"), ("Eexc", "Eexc"), ("Etot", "Etot")):
        m = re.search(rf"{lab} \([^)]*\)\s+= ([-+0-9.eE]+)", txt)
        out[k] = float(m.group(1))
    out["orbitals"] = [(m.group(1), float(m.group(2)), float(m.group(3))) for m in
                       re.finditer(r"^\s+(\d\w) : ε = ([-+0-9.eE]+)\s+n = ([-+0-9.eE]+)", txt, re.M)]
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    json.dump(fem_matrices(), open(f"{OUT}/fem_matrices.json", "w"), indent=1)
    scf = {
        "hydrogen_rhf": {  # test/hydrogen.jl:55-84
            "setup": {"Z": 1, "N": 1, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 0, "nh": 1,
                      "ex": None, "ec": None, "max_degen": 1, "scftol": 1e-10},
            "Ekin": 0.2439648717067077, "Ecou": -0.6856721626901766, "Ehar": 0.19774242379804155,
            "Etot": -0.24396486718542498, "tol": 1e-7, "eps_1s": -0.046222, "eps_tol": 1e-5},
        "hydrogen_slater": {  # test/hydrogen.jl:87-116 (Libxc lda_x)
            "setup": {"Z": 1, "N": 1, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 0, "nh": 1,
                      "ex": "lda_x", "ec": None, "max_degen": 1, "scftol": 1e-10},
            "Ekin": 0.4065340797823483, "Ecou": -0.9000752049269123, "Ehar": 0.2749225031615249,
            "Eexc": -0.18791545723754685, "Etot": -0.40653407922058604, "tol": 1e-7,
            "eps_1s": -0.194250, "eps_tol": 1e-6},
        "hydrogen_bare": {  # test/hydrogen.jl:138-171
            "setup": {"Z": 1, "N": 1, "hartree": 0, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10,
                      "lh": 1, "nh": 2, "ex": None, "ec": None, "max_degen": 1, "scftol": 1e-12,
                      "maxiter": 200},
            "Etot": -0.5, "eps": {"0,0": -0.5, "0,1": -0.125, "1,0": -0.125, "1,1": -1 / 18},
            "tol": 1e-8},
        "scandium_rhf": {  # test/scandium.jl:38-107
            "setup": {"Z": 21, "N": 21, "mesh": ["exp", 0, 2000, 40, 1.2], "p": 10, "lh": 2, "nh": 5,
                      "ex": None, "ec": None, "max_degen": 2, "scftol": 1e-10},
            "Etot": -722.5838224454205, "tol": 1e-9,
            "eps": {"1s": -154.35865, "2s": -15.78538, "2p": -12.74151, "3s": -1.69002,
                    "3p": -0.96964, "4s": -0.08646, "4p": -0.00262, "3d": -0.00262},
            "eps_tol": 1e-5, "n3d_over_5": 0.0056, "n3d_tol": 1e-4},
        "readme_hydrogen": {  # README.md:57-84 (config C1)
            "setup": {"Z": 1, "N": 1, "mesh": ["exp", 0, 300, 20, 1.5], "p": 10, "lh": 0, "nh": 3,
                      "ex": "lda_x", "ec": None, "max_degen": 2, "scftol": 1e-11},
            "niter": 19, "stop": 9.029719505180509e-12, "eps_1s": -1.942500619576336e-01},
    }
    json.dump(scf, open(f"{OUT}/scf_goldens.json", "w"), indent=1)
    traj = {
        "sodium_slater": {  # examples/basic_example/{run.jl,log.txt,results.txt}
            "setup": {"Z": 11, "N": 11, "mesh": ["exp", 0, 500, 20, 1.2], "p": 10, "lh": 2, "nh": 10,
                      "npoints": 2000, "ex": "lda_x", "ec": None, "max_degen": 2, "degen_tol": 0.1,
                      "scftol": 1e-9},
            "log": parse_log(f"{REF}/examples/basic_example/log.txt"),
            "results": parse_results(f"{REF}/examples/basic_example/results.txt")},
    }
    for name, sub, ex, ec in (("oxygen_lda", "lda", "lda_x", "lda_c_pw"), ("oxygen_x", "lda_x_only", "lda_x", None),
                              ("oxygen_none", "no_functional", None, None)):
        d = f"{REF}/examples/functional comparison/{sub}"
        traj[name] = {"setup": {"Z": 8, "N": 8, "mesh": ["exp", 0, 1000, 30, 1.2], "p": 10, "lh": 2, "nh": 5,
                                "ex": ex, "ec": ec, "max_degen": 2, "scftol": 1e-10},
                      "log": parse_log(f"{d}/log.txt"), "results": parse_results(f"{d}/results.txt")}
    json.dump(traj, open(f"{OUT}/trajectories.json", "w"), indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
