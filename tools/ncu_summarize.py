"""Summary of an `ncu --set full` capture of ONE step: one row per launch (a kernel that runs twice in a step -- the window
pass and the refill pass of the eigensolver -- appears as `name` and `name (refill pass)`) and the DRAM bytes per launch
that bench.py copies into roofline.traffic.

  ncu -i gpurun_out/r02_ncu_full.ncu-rep --page raw --csv > /tmp/ncu_raw.csv
  python tools/ncu_summarize.py /tmp/ncu_raw.csv profiles/r02_ncu_full_summary.csv profiles/r02_traffic.json "header text"
"""
import csv, json, re, sys

raw, out_csv, out_json = sys.argv[1:4]
header = sys.argv[4] if len(sys.argv) > 4 else ""
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
stall = [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(r, n):
    v = float(r[col[n]].replace(",", ""))
    u = units[col[n]]
    return v * SCALE.get(u, 1.0)


last = {}
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).strip()
    if name.startswith("void "):
        name = name[5:]
    if name in last and (name.startswith("k_eig") or name.startswith("k_aufbau")):
        name += " (refill pass)"
    if name in last:
        continue   # a launch of the neighbouring step in the capture: the first one is kept
    last[name] = r
with open(out_csv, "w") as f:
    if header:
        f.write("# " + header + "\n")
    f.write("kernel," + ",".join(want) + ",top_stalls\n")
    for name, r in last.items():
        st = sorted(((float(r[col[n]]), n.split("stalled_")[1].split("_per_")[0]) for n in stall if r[col[n]] not in ("", "n/a")), reverse=True)[:4]
        f.write('"%s",' % name + ",".join("%g" % num(r, n) for n in want) + "," + " ".join("%s=%.2f" % (b, a) for a, b in st) + "\n")
traffic = {"source": out_csv + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, 86 atoms, one lane)"}
for name, r in last.items():
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    traffic[re.sub(r"<.*", "", name)] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                                         "ncu_duration_us": num(r, "gpu__time_duration.sum")}
json.dump(traffic, open(out_json, "w"), indent=1)
print("wrote", out_csv, out_json, len(last), "kernels")
