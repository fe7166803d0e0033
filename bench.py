"""bench.py -- SCF-iteration throughput of the B200 hot path (and the CPU reference arm).

Workload (config.workload): ONE periodic-table sweep per GPU on the Argon/C2 discretization of
BASELINE.json (expmesh(0,300,41;s=1.5) = 40 elements, ordermax 20, Nh = 799, 1000 Gauss points per
element, LDA = lda_x + lda_c_pw, Float64): 86 independent atoms Z = 1..86 with the reference's
Z-dependent (lh, nh) rule.  A "step" is one `scf_step!` (assembly + eigensolve + aufbau + density +
line search + mixing) of every atom of the batch; `scftol` is set to 0 for the timed steps so
that every atom does every step (fixed work).  metric = atom-SCF-iterations per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

C2 = {"mesh": ["exp", 0, 300, 41, 1.5], "p": 20, "npoints": 1000, "lh": 3, "nh": 10}
METRIC, UNIT = "scf_iterations_per_s", "atom-iterations/s"
# rank -> (exchange, correlation, ionic charge): independent atoms, ions and functionals
RANK_SWEEPS = [("lda_x", "lda_c_pw", 0), ("lda_x", "lda_c_pw", 1), ("lda_x", None, 0), ("lda_x", None, 1),
               ("lda_x", "lda_c_pw", 2), ("lda_x", None, 2), ("lda_x", "lda_c_pw", 3), ("lda_x", None, 3)]


_RESULT_LINE = []   # the single JSON record, printed by main() on the real stdout


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  nvidia-smi needs
    ~0.1 s to start, longer than a short timed region, so it is started before the warm-up and its rows are
    filtered by their own timestamps to the window [mark_begin, mark_end]."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark_begin(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for r in self.rows:
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
                parsed.append((ts, float(r[1]), float(r[2]), [v.lower().startswith("active") for v in r[3:7]]))
            except Exception:
                pass
        pad = datetime.timedelta(milliseconds=15)   # sampling period: keep the samples that bracket the region
        inside = [x for x in parsed if self.t0 and self.t1 and self.t0 - pad <= x[0] <= self.t1 + pad]
        used = inside or parsed
        sm = sorted(x[1] for x in used)
        reasons = sorted({nm for x in used for nm, on in zip(names, x[3]) if on})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((x[2] for x in used), default=None),
                "reasons": reasons, "samples": len(used), "samples_in_timed_region": len(inside)}


def build_disc(device):
    from helpers import package_tables
    import atomickohnsham_b200 as aks
    basis, ops, quad = package_tables(C2["mesh"], C2["p"], C2["npoints"])
    disc = aks.KSEDiscretization(basis, None, lh=C2["lh"], nh=C2["nh"], nspin=1, fem_integration_method=quad,
                                 femops=ops, device=device)
    return disc


def b200_arm(args):
    import torch
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    from atomickohnsham_b200.sweep import periodic_table_models

    disc = build_disc(local)
    ctx = disc.ctx
    disc.handle()
    ex, ec, charge = RANK_SWEEPS[rank % len(RANK_SWEEPS)]
    models, lhs, nhs = periodic_table_models(range(1, 87), ex=ex, ec=ec, charge=charge)
    if args.sweeps > 1:
        models, lhs, nhs = models * args.sweeps, lhs * args.sweeps, nhs * args.sweeps
    B = len(models)
    alg_fixed = aks.ODA(scftol=0.0, tinit=0.6, aufbau=aks.OptimizedAufbau(tol=1e-3))
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(local)

    # ---------------- device-resident throughput (`value`)
    bt = Batch(models, disc, alg_fixed, lh=lhs, nh=nhs, eig_margin=args.eig_margin)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        bt.step_async()
    launches0 = ctx.launches
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        bt.step_async()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    refills = int(bt.eig_refills().sum())
    recs = bt.step()  # sanity: energies finite
    assert all(r["status"] >= 0 for r in recs), "a problem failed"
    units = torch.tensor([float(B * args.steps), ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        import torch.distributed as dist
        tot = units.clone()
        dist.all_reduce(tot[0:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(tot[1:2], op=dist.ReduceOp.MAX)
        units = tot
    total_units, max_ms = float(units[0]), float(units[1])
    value = total_units / (max_ms * 1e-3)
    per_rank_ms = [ms]
    if world > 1:   # load imbalance between the ranks' sweeps
        import torch.distributed as dist
        allms = [torch.zeros(1, dtype=torch.float64, device=f"cuda:{local}") for _ in range(world)]
        dist.all_gather(allms, torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}"))
        per_rank_ms = [float(x) for x in allms]

    # ---------------- per-phase device time + roofline of the dominant kernel (rank 0)
    roof, phases = None, None
    if rank == 0:
        os.environ["AKS_PHASE_TIMERS"] = "1"
        bt2 = Batch(models, disc, alg_fixed, lh=lhs, nh=nhs, eig_margin=args.eig_margin)
        del os.environ["AKS_PHASE_TIMERS"]
        for _ in range(max(args.warmup, 3)):
            bt2.step()
        acc = [0.0] * 7
        mixd = potk = xck = 0.0
        nrep = 5
        for _ in range(nrep):
            bt2.step()
            ph = bt2.phase_ms()
            acc = [a + b for a, b in zip(acc, ph[:7])]
            mixd += ph[7]
            potk += ph[8]
            xck += ph[9]
        names = ["potential", "eigensolve", "aufbau+orbital_energies", "density", "poisson", "linesearch", "mix+stop"]
        phases = {n: a / nrep for n, a in zip(names, acc)}
        phases["potential: k_xc"] = xck / nrep
        phases["potential: k_potential"] = potk / nrep
        bt2.close()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = ctx.fp64_peak_tflops()
        n, Q, Cc, Nh = C2["p"] + 1, C2["npoints"], 40, 799
        dom = max(names, key=lambda nm: phases[nm])
        # k_potential: 2 n^2 Q flop per cell and spin (second half of W_xc) + C Q XC points
        pot_flops = B * Cc * 2.0 * n * n * Q
        # k_mix_dense: read D + write D, 16 B per entry of the lower triangle (D is symmetric; SURVEY 8d
        # W_dens bytes, in-place form)
        mix_bytes = B * 16.0 * Nh * (Nh + 1) / 2
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        roof_all = {
            "potential": {"bound": "fp64", "achieved": pot_flops / (potk / nrep * 1e-3) / 1e12, "kernel_ms": potk / nrep,
                          "peak": fp64_peak, "unit": "TFLOP/s"},
            "mix+stop": {"bound": "hbm", "achieved": mix_bytes / (mixd / nrep * 1e-3) / 1e9, "kernel_ms": mixd / nrep,
                         "peak": hbm_peak, "unit": "GB/s"},
        }
        # DRAM bytes per launch of the same kernels from the committed ncu --set full capture (same workload)
        traffic = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            traffic = {"potential": tj["k_potential<1, 4>"]["dram_bytes_per_launch"] * (B / 86.0),
                       "mix+stop": tj["k_mix_dense"]["dram_bytes_per_launch"] * (B / 86.0)}
        except Exception:
            pass
        for k, v in roof_all.items():
            v["frac"] = v["achieved"] / v["peak"]
            v["traffic"] = traffic.get(k)
        roof = dict(roof_all["potential"])
        roof["kernel"] = "k_potential"
        roof["peak_source"] = "aks_measure_fp64_peak (DFMA chains, same run); MEASURED_PEAKS.json has no FP64 figure"
        roof["dominant_phase"] = dom
        roof["algorithmic_per_launch"] = pot_flops
        roof["traffic_source"] = "profiles/r01_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch)"
        # the largest PHASE is the eigensolve; SURVEY 8(d) counts it as the banded two-stage reduction
        # W_eig = L s 12 Nh^2 b flop per atom (b = 39; the reference's dense route is L s (16/3) Nh^3).  The
        # per-cell reduction + certified RQI used here does ~1000x fewer flops and is latency bound, so the
        # "equivalent" rate below can exceed the FP64 peak: it measures the algorithmic saving, not the pipe.
        nchan = sum(l + 1 for l in lhs)
        w_eig_banded = nchan * 12.0 * Nh * Nh * 39
        w_eig_dense = nchan * (16.0 / 3.0) * Nh ** 3
        eig_s = phases["eigensolve"] * 1e-3
        roof["other"] = {"k_mix_dense": roof_all["mix+stop"],
                         "eigensolve_phase": {"ms": phases["eigensolve"], "channels": nchan,
                                              "survey_banded_flops": w_eig_banded,
                                              "equivalent_tflops_vs_banded_count": w_eig_banded / eig_s / 1e12,
                                              "equivalent_tflops_vs_reference_dense_count": w_eig_dense / eig_s / 1e12,
                                              "bound": "latency (sequential pivots); see DESIGN.md 3.2"},
                         "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}

    # ---------------- end to end through the public API with host buffers (`e2e`)
    e2e_times = []
    for _rep in range(3):                                          # median of three repetitions
        barrier()
        t0 = time.perf_counter()
        bt3 = Batch(models, disc, alg_fixed, lh=lhs, nh=nhs, eig_margin=args.eig_margin)   # H2D: model parameters
        for _ in range(args.steps):
            bt3.step()                                             # D2H: one record per atom, every step
        outs = bt3.states()                                        # D2H: eps, n, W of every atom
        torch.cuda.synchronize(local)
        e2e_times.append(time.perf_counter() - t0)
        bt3.close()
    t_e2e = sorted(e2e_times)[1]
    assert len(outs) == B
    h2d = B * (3 * 8 + 4 * 4)
    d2h_total = args.steps * B * 72 + B * (3 * 4 * 10 * 8 + Nh_bytes())   # records + eps, both occupation slots, W
    e2e_units = torch.tensor([float(B * args.steps), t_e2e], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(e2e_units[0:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(e2e_units[1:2], op=dist.ReduceOp.MAX)
    e2e_val = float(e2e_units[0]) / float(e2e_units[1])

    # ---------------- full sweep to convergence (atoms/s), informational
    sweep = None
    if rank == 0:
        alg = aks.ODA(scftol=1e-10, tinit=0.6, aufbau=aks.OptimizedAufbau(tol=1e-3))
        btc = Batch(models, disc, alg, lh=lhs, nh=nhs, eig_margin=args.eig_margin)
        t0 = time.perf_counter()
        final, _ = btc.run(100)
        dt = time.perf_counter() - t0
        sweep = {"atoms": B, "seconds": dt, "atoms_per_s": B / dt,
                 "converged": sum(1 for r in final if r["status"] == 1),
                 "iterations_total": sum(r["niter"] for r in final),
                 "iterations_max": max(r["niter"] for r in final)}
        btc.close()

    # ---------------- Double64 path, informational: 20 anions on the same
    # order-20 discretization with double-double tables and kernels
    dd64 = None
    if rank == 0 and world == 1 and not args.no_dd:
        dd64 = double64_leg(local)

    # ---------------- CPU baseline (rank 0, N = 1 only): the oracle on one host core
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_single(budget_s=15.0)

    bt.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5 periodic-table sweep Z=1..86 (one sweep per GPU; rank r: functional/charge "
                                   "variant r) on the C2 Argon discretization: expmesh(0,300,41;s=1.5), ordermax=20, "
                                   "Nh=799, 1000 Gauss points/element, LDA, lh=min(ceil(Z/20),3), nh=clamp(Z,2,10)",
                       "atoms_per_gpu": B, "eig_window_margin": args.eig_margin,
                       "eig_window_refills_in_run": refills, "scftol": "0 during timed steps (fixed work per step)",
                       "l2": "state per GPU (dense D: %.0f MB) exceeds the 126 MB L2; no explicit flush" % (B * 8 * 799 * 799 / 1e6),
                       "tables": "resident before the timed regions (setup is outside the reference's loop too)"},
            "clocks": clocks, "gpu_launches": launches, "per_rank_ms_per_step": [x / args.steps for x in per_rank_ms],
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h_total / args.steps,
                    "what": "Batch() + K x aks_scf_step with host records + aks_get_state_all (eps, n, W of every atom); median of 3",
                    "seconds_each": e2e_times},
            "roofline": roof, "phases_ms": phases,
            "phases_note": "per-phase times are measured with the batch on ONE stream (AKS_PHASE_TIMERS); the timed steps "
                           "run the batch as 8 concurrent stream lanes, so the phases add up to more than ms_per_step", "cpu_baseline": cpu, "sweep_to_convergence": sweep,
            "double64": dd64,
        }
        _RESULT_LINE.append(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def double64_leg(device, steps=8, warmup=4):
    """C4 of SURVEY.md section 8: anions Z = 1..20, N = Z + 1, order 20, 40 cells, Q = G = 1000, builtin
    lda_x + lda_c_pw, T = Double64 (dtype = 1: double-double tables and kernels).  Per-step time of iterations
    warmup .. warmup+steps of a cold start, host-synchronised through the iteration records."""
    import atomickohnsham_b200 as aks
    from atomickohnsham_b200.solver import Batch
    Zs = list(range(1, 21))
    basis = aks.P1IntLegendreBasis(aks.expmesh(0, 300, 41, s=1.5), ordermax=20)
    models = [aks.LDA(Z=z, N=z + 1) for z in Zs]
    lh = [min(-(-z // 20), 3) + 1 for z in Zs]
    nh = [min(max(z, 2), 10) for z in Zs]
    t0 = time.perf_counter()
    disc = aks.KSEDiscretization(basis, models[0], lh=max(lh), nh=max(nh), dtype="double64", device=device)
    disc.handle()
    t_tab = time.perf_counter() - t0
    bt = Batch(models, disc, aks.ODA(tinit=0.6, scftol=0.0, aufbau=aks.OptimizedAufbau(max_degen=2, tol=1e-3)), lh=lh, nh=nh)
    for _ in range(warmup):
        bt.step()
    l0 = disc.ctx.launches
    t0 = time.perf_counter()
    for _ in range(steps):
        rec = bt.step()
    dt = time.perf_counter() - t0
    out = {"workload": "C4: anions Z=1..20, N=Z+1, ordermax=20, 40 cells, Q=G=1000, lda_x+lda_c_pw, Double64",
           "dtype": "double-double (2 x f64)", "atoms": len(Zs), "steps": steps, "warmup": warmup,
           "ms_per_step": dt / steps * 1e3, "atom_iterations_per_s": len(Zs) * steps / dt,
           "gpu_launches": disc.ctx.launches - l0, "host_tables_s": t_tab,
           "all_finite": all(r["Etot"] == r["Etot"] and abs(r["Etot"]) < 1e300 for r in rec)}
    bt.close()
    disc.close()
    return out


def Nh_bytes():
    return 799 * 8


# ------------------------------------------------------------------------------------------- CPU side
def _oracle_for(Z):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from helpers import make_oracle
    from atomickohnsham_b200.sweep import basis_size
    lh, nh = basis_size(Z)
    setup = {"Z": Z, "N": Z, "mesh": C2["mesh"], "p": C2["p"], "npoints": C2["npoints"], "lh": lh, "nh": nh,
             "ex": "lda_x", "ec": "lda_c_pw", "scftol": 1e-300}
    return make_oracle(setup)


def cpu_baseline_single(budget_s=15.0):
    """Oracle (CPU restatement of the reference; Julia is unavailable) on ONE core: Argon, C2.  Also the
    per-iteration assembly (a3-a9) and eigensolve (a10) times, the split SURVEY 8(d) asks to report."""
    o = _oracle_for(18)
    st = o.initial_state()
    o.scf_step(st)  # warm-up (includes first-use costs)
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < budget_s:
        o.scf_step(st)
        k += 1
    dt = time.perf_counter() - t0
    ta = time.perf_counter()
    Har, Vxc, _ = o.hamiltonian_parts(st.D)
    tb = time.perf_counter()
    o.find_orbital(Har, Vxc)
    tc = time.perf_counter()
    return {"value": k / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{k} SCF iterations of Argon (Z=18, C2 discretization) with the numpy/LAPACK oracle, 1 thread",
            "ms_per_iteration": dt / k * 1e3, "assembly_ms": (tb - ta) * 1e3, "eigensolve_ms": (tc - tb) * 1e3}


_W = {}


def _worker_init(Z):
    _W["o"] = _oracle_for(Z)
    _W["st"] = _W["o"].initial_state()
    _W["o"].scf_step(_W["st"])
    return Z


def _worker_step(_):
    _W["o"].scf_step(_W["st"])
    return 1


def reference_arm(args):
    """The reference's CPU implementation of the path = the oracle (Julia cannot run here), one
    atom per host core, every step = one SCF iteration of each core's atom."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    ncore = os.cpu_count() or 1
    Zs = [18, 36, 54, 86, 10, 26, 47, 79][:ncore] if ncore <= 8 else [(7 * i) % 86 + 1 for i in range(ncore)]
    ctx = mp.get_context("spawn")
    pools = [ctx.Pool(1) for _ in Zs]
    for p, Z in zip(pools, Zs):
        p.apply_async(_worker_init, (Z,)).get()
    for _ in range(args.warmup):
        for r in [p.apply_async(_worker_step, (0,)) for p in pools]:
            r.get()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for r in [p.apply_async(_worker_step, (0,)) for p in pools]:
            r.get()
    dt = time.perf_counter() - t0
    for p in pools:
        p.terminate()
    val = len(Zs) * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "same C2 discretization and LDA model; one atom per host core, Z in %s" % Zs},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": len(Zs), "kind": "port",
                             "sample": f"{args.steps} steps x {len(Zs)} atoms, numpy/LAPACK oracle (Julia reference cannot run: no julia in the image)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _RESULT_LINE.append(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps", type=int, default=1, help="periodic-table sweeps per GPU (default 1)")
    ap.add_argument("--eig-margin", type=int, default=0, help="eigenpair window margin (-1: all nh every step)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-dd", action="store_true", help="skip the informational Double64 (C4) leg")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON record): anything libraries print there (e.g. NCCL's version
    # banner under torchrun) is sent to stderr instead.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            reference_arm(args)
        else:
            b200_arm(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[0], flush=True)


if __name__ == "__main__":
    main()
