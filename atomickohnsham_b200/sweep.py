"""Independent-problem sweeps: the only axis along which this path shards.

A single atom (Nh <= ~800) is far too small to split; the reference itself only ever
parallelises over atoms (`Threads.@threads for z`, `examples/periodic table/process.jl:92`).
`basis_size` is the Z-dependent (lh, nh) rule of `examples/atomic density comparison/run.jl:58`;
`partition` is a static longest-processing-time split of the problem list over ranks, with
no data-path collective -- ranks only gather O(100 B) records per problem at the end.
"""
from __future__ import annotations

import math

from .model import BuiltinFunctional, KSEModel, NoFunctional


def basis_size(Z: int):
    """(lh, nh) = (min(ceil(Z/20), 3), clamp(Z, 2, 10))."""
    return min(math.ceil(Z / 20), 3), min(max(int(Z), 2), 10)


def periodic_table_models(Zs=range(1, 87), *, ex="lda_x", ec="lda_c_pw", charge=0, nspin=1):
    """Neutral atoms (or ions with N = Z - charge) for a sweep; returns (models, lh[], nh[])."""
    mk = lambda name: NoFunctional(nspin) if name is None else BuiltinFunctional(name, nspin)
    models, lhs, nhs = [], [], []
    for Z in Zs:
        N = Z - charge
        if N < 1:
            continue
        models.append(KSEModel(Z=Z, N=N, ex=mk(ex), ec=mk(ec)))
        lh, nh = basis_size(Z)
        lhs.append(lh)
        nhs.append(nh)
    return models, lhs, nhs


def problem_cost(Z: int) -> float:
    """Relative cost of one SCF iteration of atom Z: eigenpairs solved + cells assembled."""
    lh, nh = basis_size(Z)
    return (lh + 1) * nh + 4.0


def partition(costs, world: int):
    """Static LPT partition: returns `world` index lists with balanced total cost."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: loads[j])
        parts[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in parts]


def lane_balanced_order(costs, nlanes: int = 8):
    """Order of a batch's problems that balances the library's stream lanes.

    A batch is split into `nlanes` CONTIGUOUS problem ranges, each on its own stream; a list sorted by
    nuclear charge puts all the light atoms into the first lane and all the heavy ones into the last.  This deals the
    problems, heaviest first, in a snake over the lanes and concatenates the lanes: measured on the 86-atom sweep
    1.24 -> 1.20 ms per step over iterations 5-24 and 1.090 -> 1.078 ms once settled (`tools/gpu_lane_order.py`).  Results
    of a problem do not depend on its position in the batch (bit for bit: `test_batch_independence_and_reproducibility`).
    Returns the permutation (list of indices into `costs`)."""
    srt = sorted(range(len(costs)), key=lambda i: -costs[i])
    buckets = [[] for _ in range(nlanes)]
    for j, i in enumerate(srt):
        buckets[j % nlanes if (j // nlanes) % 2 == 0 else nlanes - 1 - j % nlanes].append(i)
    return [i for b in buckets for i in b]


def solve_sweep(models, discretization, *, lh=None, nh=None, alg=None, maxiter=100, remedy_alg=None,
                remedy_maxiter=800, devices=None, want_arrays=False):
    """A sweep of independent problems solved with the reference's own two-stage protocol
    (`examples/atomic density comparison/run.jl:40-47,74-76`, there hard-wired to iron; here decided by the outcome):

      pass 1  every problem with the reference's default algorithm `alg` (run.jl:39: ODA(tinit = 0.4,
              scftol = 1e-10) with OptimizedAufbau(tol = 1e-2), run.jl:70), `maxiter` steps (run.jl:41: 100);
      pass 2  every problem that ended pass 1 with MAXITERS (open d / f shells: the adaptive line search keeps
              oscillating between two fillings of a near-degenerate block) is solved again FROM SCRATCH with the
              algorithm the reference gives iron, `remedy_alg` = ODA(tinit = 0.15, frozen_t = true, scftol = 1e-12),
              for up to `remedy_maxiter` steps.  As the reference notes for iron, such a run may still plateau and
              report MAXITERS (run.jl:43-47); for some open f shells the frozen step even cycles between fillings.
              The sweep therefore keeps, per problem, the BETTER of its two end states: pass 2 if it converged or
              stopped at a smaller stopping criterion than pass 1, else pass 1.  `remedy_maxiter = 0` skips pass 2.

    A two-fold degenerate block makes the reference's algorithm minimise an energy that is flat to rounding in the
    mixing parameters (Brent on function values: t is only determined to ~sqrt(eps)), so its stopping criterion
    bounces in 1e-10 ... 1e-6 and whether `scftol = 1e-10` is met within `maxiter` is decided by rounding noise -- in
    the reference too (run.jl:33-37: "stopping criterion bounces between 1e-9 and 1e-7 for hundreds of iterations").
    Such an end state (MAXITERS with a stopping criterion below ~1e-6) has Etot converged to ~1e-14 relative.

    Both passes are single device batches (`groundstate` of a list; partitioned over `devices` when given).
    Returns a list of (KSESolution, pass_number) in the order of `models`.
    """
    from .algorithms import ODA, OptimizedAufbau
    from .solver import groundstate
    alg = alg or ODA(scftol=1e-10, tinit=0.4, aufbau=OptimizedAufbau(tol=1e-2))
    remedy_alg = remedy_alg or ODA(scftol=1e-12, tinit=0.15, frozen_t=True, aufbau=alg.aufbau)
    models = list(models)
    first = groundstate(models, discretization, alg, maxiter=maxiter, lh=lh, nh=nh, want_arrays=want_arrays,
                        devices=devices)
    out = [(sol, 1) for sol in first]
    todo = [i for i, sol in enumerate(first) if sol.success == "MAXITERS"] if remedy_maxiter else []
    if todo:
        sub = lambda v: None if v is None else [v[i] for i in todo]
        second = groundstate([models[i] for i in todo], discretization, remedy_alg, maxiter=remedy_maxiter,
                             lh=sub(lh), nh=sub(nh), want_arrays=want_arrays, devices=devices)
        if not isinstance(second, list):
            second = [second]
        for i, sol in zip(todo, second):
            if sol.success == "SUCCESS" or sol.stopping_criteria < first[i].stopping_criteria:
                out[i] = (sol, 2)
    return out


def maximum_ionization(Zs, discretization, *, nspin=1, ex="lda_x", ec="lda_c_pw", width=1.2, precision=2,
                       scftol=1e-6, maxiter=50, tinit=0.4, degen_tol=1e-2, lh=None, nh=None, trace=None):
    """Largest electron number N in [Z, Z + width] each nucleus Z can bind, by bisection on N.

    Batched front end of the driver `compute_maximum_ionization` (`examples/periodic table/process.jl:12-86`):
    the same stability test per trial (the highest occupied level must be negative and the SCF must have
    reached `scftol`, `process.jl:77-82`), the same bracket, tolerance `10^-precision`, `ODA(0.4)`,
    `scftol = 1e-6`, `maxiter = 50`, `degen_tol = 1e-2` and `(lh, nh)` rule -- but every bisection round is
    ONE device batch over all Z on a shared discretization instead of a thread per atom (the reference script
    also re-meshes per atom from a first neutral run; that stays with the caller, who owns the discretization).
    Returns {Z: N_max} rounded to `precision` digits (the last trial N, process.jl:85).  `trace`, a dict, receives per Z the
    list of trials (N, HOMO energy, bound?) in bisection order.
    """
    from .algorithms import ODA, OptimizedAufbau
    from .solver import Batch

    Zs = [int(z) for z in Zs]
    mk = lambda name: NoFunctional(nspin) if name is None else BuiltinFunctional(name, nspin)
    lo = {z: float(z) for z in Zs}
    hi = {z: z + float(width) for z in Zs}
    tol = 10.0 ** (-precision)
    alg = ODA(scftol=scftol, tinit=tinit, aufbau=OptimizedAufbau(tol=degen_tol))
    mid = dict(lo)
    while max(hi[z] - lo[z] for z in Zs) > tol:
        mid = {z: (hi[z] + lo[z]) / 2 for z in Zs}
        models = [KSEModel(Z=z, N=mid[z], ex=mk(ex), ec=mk(ec)) for z in Zs]
        lhs = [min(math.ceil(mid[z] / 20), 3) if lh is None else lh for z in Zs]
        nhs = [basis_size(z)[1] if nh is None else nh for z in Zs]
        lhs = [min(v, discretization.lh) for v in lhs]
        nhs = [min(v, discretization.nh) for v in nhs]
        bt = Batch(models, discretization, alg, lh=lhs, nh=nhs)
        final, _ = bt.run(maxiter, raise_on_problem_error=False)   # a failed trial counts as not bound
        states = bt.states()
        for i, z in enumerate(Zs):
            st = states[i]
            nn, ee = st["n"], st["eps"]
            if discretization.dtype == "double64":   # (hi, lo) pairs: the test below only needs the leading parts
                nn, ee = nn[..., 0], ee[..., 0]
            occ = nn != 0
            homo = float(ee[occ].max())
            stable = homo <= 0.0 and final[i]["status"] == 1 and final[i]["stop"] <= scftol
            if trace is not None:
                trace.setdefault(z, []).append((mid[z], homo, bool(stable)))
            if stable:
                lo[z] = mid[z]
            else:
                hi[z] = mid[z]
        bt.close()
    return {z: round(mid[z], precision) for z in Zs}
