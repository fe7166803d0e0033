"""SASS instruction mix per kernel of the built library (`cuobjdump -sass`): the mnemonics that show which hardware paths a
kernel uses -- DMMA = FP64 tensor-core tile updates (mma.sync m8n8k4), UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier
operations, DFMA / DMUL / DADD = FP64 pipe, LDL / STL = local memory (spills), LDGSTS = cp.async.

  python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "atomickohnsham_b200", "libaks_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["DFMA", "DMUL", "DADD", "DSETP", "MUFU", "DMMA", "UBLKCP", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "SHFL", "IMAD", "BRA"]
kern, rows = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        rows[kern] = dict.fromkeys(["instr"] + cols, 0)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        rows[kern]["instr"] += 1
        base = op.split(".")[0]
        if base in rows[kern]:
            rows[kern][base] += 1
print(__doc__.split("\n\n")[0].replace("\n", " "))
print()
print("%-44s %7s " % ("kernel", "instr") + " ".join("%6s" % c for c in cols))
for k, r in sorted(rows.items(), key=lambda kv: -kv[1]["instr"]):
    print("%-44s %7d " % (k[:44], r["instr"]) + " ".join("%6d" % r[c] for c in cols))
