"""Transcribe the reference's committed Double64 run into tests/golden/sodium_double64.json.

Source (read-only, only available in the build container): the per-iteration log and the final report of
`examples/basic_example_doublefloat/` (sodium, Hartree only, T = Double64, scftol = 1e-20) -- outputs of the
reference itself, all digits kept as strings.  Run from the repo root:  python tools/make_golden_dd.py
"""
import json
import os
import re

REF = "/root/reference/examples/basic_example_doublefloat"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "sodium_double64.json")


def main():
    its = []
    for line in open(os.path.join(REF, "log.txt")):
        if not line.startswith("iter ="):
            continue
        rec = dict(re.findall(r"(\w+) = (\S+)", line))
        its.append({k: rec[k] for k in ("iter", "stop", "Etot", "Ekin", "Ecou", "Ehar", "Eexc", "t")})
    rep = open(os.path.join(REF, "results.txt")).read()
    grab = lambda pat: re.search(pat, rep).group(1)
    final = dict(
        niter=int(grab(r"niter\s+= (\d+)")), stop=grab(r"Final stopping criterion = (\S+)"),
        Ekin=grab(r"Ekin \(kinetic\)\s+= (\S+)"), Ecou=grab(r"Ecou \(nuclear attraction\)\s+= (\S+)"),
        Ehar=grab(r"Ehar \(Hartree\)\s+= (\S+)"), Etot=grab(r"Etot \(total\)\s+= (\S+)"),
        orbitals={m[0]: dict(eps=m[1], n=m[2]) for m in re.findall(r"(\d[spdf]) : ε = (\S+)\s+n = (\S+)", rep)})
    setup = dict(Z=11, N=11, hartree=1, mesh=["exp", 0, 500, 20, 1.2], p=10, lh=2, nh=10, scftol=1e-20, tinit=0.6,
                 degen_tol=1e-2, max_degen=2, maxiter=100,
                 source="examples/basic_example_doublefloat/{run.jl,log.txt,results.txt}",
                 note="t is printed by the reference without its exponent when it is not 1.0 / 0.0 / 0.6 "
                      "(a DoubleFloats show quirk): compare mantissas")
    json.dump(dict(setup=setup, iterations=its, final=final), open(OUT, "w"), indent=1, ensure_ascii=False)
    print(OUT, len(its), "iterations")


if __name__ == "__main__":
    main()
