"""groundstate / KSESolver / KSESolution on top of the CUDA library.

Mirrors `groundstate(model, discretization, alg; maxiter=100, callback=nothing)`,
`KSESolver`, `solve!`, `register!`, `LogBook`, `CallbackSet`, `LogFileCallback`
(`src/solver/{groundstate,solver,logbook,callback,logging}.jl`) and the fields of
`KSESolution` the hot path produces (`src/solution/types.jl:74-140`).  The SCF loop body
`scf_step!` is one `aks_scf_step` call on the device for the whole batch; the host only
keeps the reference's bookkeeping (logbook, callbacks, status strings).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import AKS_ENOTCONV, AKS_MAXITERS, AKS_OK, AKS_RUNNING, AKS_SUCCESS, IterRecord
from .algorithms import ODA, FrozenAufbau, OptimizedAufbau, SmearedAufbau
from .discretization import KSEDiscretization
from .model import KSEModel

RECORD_DTYPE = np.dtype([("niter", np.int32), ("status", np.int32), ("stop", np.float64), ("Etot", np.float64),
                         ("Ekin", np.float64), ("Ecou", np.float64), ("Ehar", np.float64), ("Eexc", np.float64),
                         ("t", np.float64)])   # aks_iter_record

ATOMIC_NAMES = ("Hydrogen Helium Lithium Beryllium Boron Carbon Nitrogen Oxygen Fluorine Neon Sodium "
                "Magnesium Aluminium Silicon Phosphorus Sulfur Chlorine Argon Potassium Calcium Scandium "
                "Titanium Vanadium Chromium Manganese Iron Cobalt Nickel Copper Zinc Gallium Germanium "
                "Arsenic Selenium Bromine Krypton Rubidium Strontium Yttrium Zirconium Niobium Molybdenum "
                "Technetium Ruthenium Rhodium Palladium Silver Cadmium Indium Tin Antimony Tellurium Iodine "
                "Xenon Caesium Barium Lanthanum Cerium Praseodymium Neodymium Promethium Samarium Europium "
                "Gadolinium Terbium Dysprosium Holmium Erbium Thulium Ytterbium Lutetium Hafnium Tantalum "
                "Tungsten Rhenium Osmium Iridium Platinum Gold Mercury Thallium Lead Bismuth Polonium "
                "Astatine Radon").split()


@dataclass
class Energies:
    Etot: float = 0.0
    Ekin: float = 0.0
    Ecou: float = 0.0
    Ehar: float = 0.0
    Eexc: float = 0.0


@dataclass
class LogBook:
    stopping_criteria: list = field(default_factory=list)
    Etot: list = field(default_factory=list)


class CallbackSet:
    def __init__(self, callbacks=()):
        self.callbacks = tuple(callbacks)

    def __call__(self, solver):
        for cb in self.callbacks:
            cb(solver)


class LogFileCallback:
    """One line per iteration: iter, stop, Etot, Ekin, Ecou, Ehar, Eexc, t (logging.jl:26-37)."""

    def __init__(self, io):
        self.io = io

    def __call__(self, solver):
        e = solver.energies
        self.io.write(f"iter = {solver.niter}   stop = {solver.stopping_criteria!r}   Etot = {e.Etot!r}   "
                      f"Ekin = {e.Ekin!r}   Ecou = {e.Ecou!r}   Ehar = {e.Ehar!r}   Eexc = {e.Eexc!r}   "
                      f"t = {solver.t!r}\n")


def solution_name(Z, N):
    """groundstate.jl:24-34."""
    if float(Z).is_integer():
        d = Z - N
        base = ATOMIC_NAMES[int(Z) - 1] if 1 <= int(Z) <= len(ATOMIC_NAMES) else f"Z={Z}"
        if d == 0:
            return base
        return f"{base}{abs(d)}{'+' if d > 0 else '-'}"
    return f"Z={Z}, N={N}"


class Batch:
    """A batch of models sharing one discretization, resident on the device (aks_batch)."""

    def __init__(self, models, discretization: KSEDiscretization, alg, *, lh=None, nh=None, eig_margin=None,
                 maxiter=None):
        """`alg` is one ODA for the whole batch or a sequence with one ODA per problem (the reference chooses the
        algorithm per atom, examples/atomic density comparison/run.jl:74); the occupation scheme and the line
        search tolerances are those of the first.  `maxiter` (int or one per problem) caps each problem's
        iteration count ON THE DEVICE (status MAXITERS), as `groundstate(...; maxiter)` does per call."""
        self.models = list(models)
        self.disc = discretization
        nprob = len(self.models)
        self.algs = None
        if not isinstance(alg, ODA):
            self.algs = list(alg)
            if len(self.algs) != nprob:
                raise ValueError("one ODA per problem expected")
            if discretization.dtype == "double64":
                raise NotImplementedError("per-problem ODA options are not built for Double64 batches")
            alg = self.algs[0]
        self.alg = alg
        if maxiter is not None and not np.isscalar(maxiter) and len(maxiter) != nprob:
            raise ValueError("one maxiter per problem expected")
        self.maxiter = maxiter
        for m in self.models:
            if m.nspin != discretization.nspin:
                raise ValueError("all models of a batch must have the discretization's spin count")
        f64 = lambda v: np.ascontiguousarray(v, dtype=np.float64)
        i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        self._Z = f64([m.Z for m in self.models])
        self._N = f64([m.N for m in self.models])
        self._har = f64([m.hartree for m in self.models])
        self._ex = i32([m.exchange.xc_id for m in self.models])
        self._ec = i32([m.correlation.xc_id for m in self.models])
        self._lh = None if lh is None else i32(lh)
        self._nh = None if nh is None else i32(nh)
        self.ctx = discretization.ctx
        h = C.c_void_p()
        self.ctx.check(_capi.lib.aks_batch_create(
            discretization.handle(), nprob, _capi.dptr(self._Z), _capi.dptr(self._N), _capi.dptr(self._har),
            _capi.i32ptr(self._ex), _capi.i32ptr(self._ec), _capi.i32ptr(self._lh), _capi.i32ptr(self._nh),
            C.byref(h)))
        self.h = h
        self.nprob = nprob
        a = alg.aufbau
        opt = a if isinstance(a, OptimizedAufbau) else OptimizedAufbau()
        abstol_ls, reltol_ls = alg.abstol_ls, alg.reltol_ls
        if discretization.dtype == "double64":
            # ODA's defaults are 1e4 eps(typeof(tinit)) (oda/types.jl:39-40): eps(Double64) for a Double64 run
            f64_default, dd_default = 1e4 * np.finfo(np.float64).eps, 4.930380657631324e-28
            abstol_ls = dd_default if abstol_ls == f64_default else abstol_ls
            reltol_ls = dd_default if reltol_ls == f64_default else reltol_ls
        self.ctx.check(_capi.lib.aks_batch_set_oda(h, alg.scftol, alg.tinit, int(alg.frozen_t), alg.maxiter_ls,
                                                   abstol_ls, reltol_ls, opt.max_degen, opt.tol,
                                                   int(opt.handle_degen_spin_1iter)))
        if isinstance(a, SmearedAufbau):
            self.ctx.check(_capi.lib.aks_batch_set_aufbau(h, 1, a.Temp, None))
        elif isinstance(a, FrozenAufbau):
            # one FrozenAufbauCache per problem, in the discretization's layout (column-major)
            one = a.nfix(discretization.lh, discretization.nh, discretization.nspin)
            self._nfix = np.ascontiguousarray(
                np.concatenate([one.reshape(-1, order="F")] * nprob), dtype=np.float64)
            self.ctx.check(_capi.lib.aks_batch_set_aufbau(h, 2, 0.0, _capi.dptr(self._nfix)))
        if eig_margin is not None:
            self.set_eig_window(eig_margin)
        if self.algs is not None or maxiter is not None:
            for i in range(nprob):
                a_i = self.algs[i] if self.algs is not None else alg
                m_i = 0 if maxiter is None else int(maxiter if np.isscalar(maxiter) else maxiter[i])
                self.set_oda_problem(i, a_i.scftol, a_i.tinit, a_i.frozen_t, m_i)
        self._rec = (IterRecord * nprob)()

    def set_oda_problem(self, prob, scftol, tinit, frozen_t=False, maxiter=0):
        """Algorithm of ONE problem (aks_batch_set_oda_problem): ODA(scftol, tinit, frozen_t) and the cap on its
        iteration count, enforced on the device (0: none)."""
        self.ctx.check(_capi.lib.aks_batch_set_oda_problem(self.h, int(prob), float(scftol), float(tinit),
                                                           int(bool(frozen_t)), int(maxiter)))

    def set_eig_window(self, margin):
        """Eigenpair window (aks_batch_set_eig_window): levels solved per channel and step = occupied +
        margin, certified on the device; margin < 0 solves all nh eigenpairs in every step."""
        self.ctx.check(_capi.lib.aks_batch_set_eig_window(self.h, int(margin)))

    def eig_refills(self):
        """Per problem: number of steps whose eigenpair window had to be refilled (diagnostic)."""
        out = np.zeros(self.nprob, dtype=np.int32)
        self.ctx.check(_capi.lib.aks_debug_eig_refills(self.h, _capi.i32ptr(out)))
        return out

    def reset(self):
        self.ctx.check(_capi.lib.aks_batch_reset(self.h))

    def step(self):
        """One scf_step! for every running problem; returns the list of records."""
        self.ctx.check(_capi.lib.aks_scf_step(self.h, self._rec))
        return [r.asdict() for r in self._rec]

    def step_async(self):
        """Enqueue one scf_step! on the ctx stream without reading anything back."""
        self.ctx.check(_capi.lib.aks_scf_step(self.h, None))

    def run(self, maxiter=100, history=False, raise_on_problem_error=True):
        """solve!: returns (final records, history or None).  history=True: per problem the list of its per-iteration
        records (dicts); history="array": the raw structured array (maxiter, nprob) the library filled, where slot k of
        a problem holds its record after iteration k + 1 (its last record repeated once it has stopped).
        A problem that ends in an error (negative status: AKS_EDEGEN3 = the reference's `error()` of optimized.jl:139, AKS_EEIG, AKS_ENAN) raises, unless
        `raise_on_problem_error=False` (sweeps: the other problems of the batch are unaffected, the status is in the
        record)."""
        hist = (IterRecord * (maxiter * self.nprob))() if history else None
        rc = _capi.lib.aks_scf_run(self.h, maxiter, hist, self._rec)
        ok = (AKS_OK, AKS_ENOTCONV)
        if not raise_on_problem_error and rc in (_capi.AKS_EDEGEN3, _capi.AKS_EEIG, _capi.AKS_ENAN):
            ok = ok + (rc,)
        self.ctx.check(rc, ok=ok)
        final = [r.asdict() for r in self._rec]
        if not history:
            return final, None
        # the history buffer as a structured array (maxiter, nprob): no per-record Python work
        arr = np.frombuffer(hist, dtype=RECORD_DTYPE).reshape(maxiter, self.nprob)
        if history == "array":
            return final, arr
        out = []
        for i in range(self.nprob):
            col = arr[:, i]
            # iteration k of this problem sits in slot k (slots after its last iteration repeat the last record)
            keep = col["niter"] == np.arange(1, maxiter + 1)
            out.append([dict(zip(RECORD_DTYPE.names, row.tolist())) for row in col[keep]])
        return final, out

    def records_dd(self):
        """Double64 batches: the last record of every problem as (hi, lo) pairs, shape (nprob, 7, 2) with the
        columns stop, Etot, Ekin, Ecou, Ehar, Eexc, t (aks_get_records_dd)."""
        out = np.zeros((self.nprob, 7, 2))
        self.ctx.check(_capi.lib.aks_get_records_dd(self.h, _capi.dptr(out)))
        return out

    def _state_dd(self, prob, want_D, want_U):
        """Double64 batches: every array carries a trailing axis (hi, lo)."""
        d = self.disc
        Nh, L, nh = d.Nh, d.lh + 1, d.nh
        D = np.zeros((Nh, Nh, 2)) if want_D else None
        U = np.zeros((L, nh, Nh, 2)) if want_U else None
        eps, n, W = np.zeros((L * nh, 2)), np.zeros((L * nh, 2)), np.zeros((Nh, 2))
        self.ctx.check(_capi.lib.aks_get_state(self.h, prob, _capi.dptr(D), _capi.dptr(U), _capi.dptr(eps),
                                               _capi.dptr(n), _capi.dptr(W)))
        col = lambda a: np.stack([a[:, 0].reshape((L, nh, 1), order="F"), a[:, 1].reshape((L, nh, 1), order="F")], -1)
        out = dict(eps=col(eps), n=col(n), W=W)
        if want_D:
            out["D"] = D[:, :, None, :]
        if want_U:
            out["U"] = np.transpose(U, (2, 1, 0, 3))[:, :, :, None, :]
        return out

    def state(self, prob=0, *, want_D=True, want_U=True):
        d = self.disc
        if d.dtype == "double64":
            return self._state_dd(prob, want_D, want_U)
        s, Nh, L, nh = d.nspin, d.Nh, d.lh + 1, d.nh
        D = np.zeros((s, Nh, Nh)) if want_D else None
        U = np.zeros((s, L, nh, Nh)) if want_U else None
        eps = np.zeros(s * L * nh)
        n = np.zeros(s * L * nh)
        W = np.zeros(Nh)
        self.ctx.check(_capi.lib.aks_get_state(self.h, prob, _capi.dptr(D), _capi.dptr(U), _capi.dptr(eps),
                                               _capi.dptr(n), _capi.dptr(W)))
        # Julia layouts -> numpy [l, k, sigma] / [i, k, l, sigma] / [i, j, sigma]
        eps = eps.reshape((L, nh, s), order="F")
        n = n.reshape((L, nh, s), order="F")
        out = dict(eps=eps, n=n, W=W)
        if want_D:
            out["D"] = np.transpose(D, (1, 2, 0))
        if want_U:
            out["U"] = np.transpose(U, (3, 2, 1, 0))
        return out

    def states(self):
        """eps, n, W of every problem in one call (aks_get_state_all): list of dicts as `state(i)` returns
        without D and U.  Double64 batches: every array carries a trailing (hi, lo) axis, as `state(i)` does."""
        d = self.disc
        s, Nh, L, nh = d.nspin, d.Nh, d.lh + 1, d.nh
        per = s * L * nh
        w = 2 if d.dtype == "double64" else 1   # Double64 numbers cross the ABI as (hi, lo) pairs
        eps = np.zeros(self.nprob * per * w)
        n = np.zeros(self.nprob * per * w)
        W = np.zeros(self.nprob * Nh * w)
        self.ctx.check(_capi.lib.aks_get_state_all(self.h, _capi.dptr(eps), _capi.dptr(n), _capi.dptr(W)))
        if w == 2:
            col = lambda a: np.stack([a[:, 0].reshape((L, nh, 1), order="F"), a[:, 1].reshape((L, nh, 1), order="F")], -1)
            eps, n, W = eps.reshape(self.nprob, per, 2), n.reshape(self.nprob, per, 2), W.reshape(self.nprob, Nh, 2)
            return [dict(eps=col(eps[i]), n=col(n[i]), W=W[i]) for i in range(self.nprob)]
        return [dict(eps=eps[i * per:(i + 1) * per].reshape((L, nh, s), order="F"),
                     n=n[i * per:(i + 1) * per].reshape((L, nh, s), order="F"),
                     W=W[i * Nh:(i + 1) * Nh]) for i in range(self.nprob)]

    def set_density(self, prob, D, energies_prev=None, niter=1):
        """Restart / step-level parity: D is (Nh, Nh, nspin) (or (Nh, Nh))."""
        D = np.asarray(D, dtype=np.float64)
        if D.ndim == 2:
            D = D[:, :, None]
        Dd = np.ascontiguousarray(np.transpose(D, (2, 0, 1)))
        e = None if energies_prev is None else np.ascontiguousarray(energies_prev, dtype=np.float64)
        self.ctx.check(_capi.lib.aks_set_density(self.h, prob, _capi.dptr(Dd), _capi.dptr(e), niter))

    # ---- step-level hooks ------------------------------------------------------------
    def density(self, prob=0):
        """density!(D, U, n): the dense trial density of the last step, (Nh, Nh, nspin)."""
        d = self.disc
        D = np.zeros((d.nspin, d.Nh, d.Nh))
        self.ctx.check(_capi.lib.aks_density(self.h, prob, _capi.dptr(D)))
        return np.transpose(D, (1, 2, 0))

    def assemble_hartree(self, prob=0):
        out = np.zeros((self.disc.Nh, self.disc.Nh))
        self.ctx.check(_capi.lib.aks_assemble_hartree(self.h, prob, _capi.dptr(out)))
        return out.T.copy()

    def assemble_vxc(self, prob=0):
        s, Nh = self.disc.nspin, self.disc.Nh
        out = np.zeros((s, Nh, Nh))
        self.ctx.check(_capi.lib.aks_assemble_vxc(self.h, prob, _capi.dptr(out)))
        return np.transpose(out, (2, 1, 0)).copy()

    def eig_lowest(self, prob=0):
        d = self.disc
        s, Nh, L, nh = d.nspin, d.Nh, d.lh + 1, d.nh
        eps = np.zeros(s * L * nh)
        res = np.zeros(s * L * nh)
        U = np.zeros((s, L, nh, Nh))
        self.ctx.check(_capi.lib.aks_eig_lowest(self.h, prob, _capi.dptr(eps), _capi.dptr(U), _capi.dptr(res)))
        return (eps.reshape((L, nh, s), order="F"), np.transpose(U, (3, 2, 1, 0)),
                res.reshape((L, nh, s), order="F"))

    # --- post-processing on the device (src/solution/postprocess.jl:34-387; host mirror: postprocess.py) -----------------
    def _eval_grid(self, prob, kind, X, l=0, k=0, sigma=0):
        X = np.ascontiguousarray(np.atleast_1d(np.asarray(X, dtype=np.float64)))
        out = np.zeros(X.shape[0])
        self.ctx.check(_capi.lib.aks_eval_grid(self.h, int(prob), int(kind), int(l), int(k), int(sigma), X.shape[0],
                                                _capi.dptr(X), _capi.dptr(out)))
        return out

    def eval_orbital(self, prob, n, l, X, sigma=None, *, h10=False):
        """u_{nl}(r)/r (or u itself, h10=True) of problem `prob` on X, from the device state (`eval_orbital`,
        postprocess.jl:34-62; n is the principal quantum number, sigma 1 or 2 for spin-polarised batches)."""
        if not 0 <= l <= n - 1:
            raise ValueError("Wrong number quantum. You should have 0 <= l <= n-1.")
        if self.disc.nspin == 2 and sigma is None:
            raise ValueError("The discretization is spin-polarized. Please give a spin sigma.")
        return self._eval_grid(prob, 5 if h10 else 0, X, l=l, k=n - l - 1, sigma=(sigma or 1) - 1)

    def eval_density(self, prob, X, sigma=None):
        """rho(X) = phi^T D phi / (4 pi X^2), spin-summed unless sigma (1 or 2) is given (`eval_density`, postprocess.jl:81-182)."""
        return self._eval_grid(prob, 1, X, sigma=-1 if sigma is None else sigma - 1)

    def eval_hartree(self, prob, X):
        """V_H(X) = W(X)/X + N/Rmax (`eval_hartree`, postprocess.jl:207-222)."""
        return self._eval_grid(prob, 2, X)

    def eval_vxc(self, prob, X, sigma=None):
        """v_xc[rho](X) of the problem's builtin LDA functionals (`eval_vxc`, postprocess.jl:328-352)."""
        if self.disc.nspin == 2 and sigma is None:
            raise ValueError("The discretization is spin-polarized. Please give a spin sigma.")
        return self._eval_grid(prob, 3, X, sigma=(sigma or 1) - 1)

    def eval_effective_potential(self, prob, l, X, sigma=None):
        """-Z/r + l(l+1)/(2r^2) + hartree V_H + v_xc (`eval_effective_potential`, postprocess.jl:378-387)."""
        if self.disc.nspin == 2 and sigma is None:
            raise ValueError("The discretization is spin-polarized. Please give a spin sigma.")
        return self._eval_grid(prob, 4, X, l=l, sigma=(sigma or 1) - 1)

    def exc_energy(self, prob=0):
        v = C.c_double()
        self.ctx.check(_capi.lib.aks_exc_energy(self.h, prob, C.byref(v)))
        return v.value

    def hartree_energy(self, prob=0):
        v = C.c_double()
        self.ctx.check(_capi.lib.aks_hartree_energy(self.h, prob, C.byref(v)))
        return v.value

    def eig_info(self):
        """(counts, rqi_its, not_converged) arrays of shape (nprob, nspin, L, nh) for the last step (word layout:
        counts << 8 | 0x80 not converged / certified | 0x40 member of a cluster of near-degenerate levels | RQI its)."""
        d = self.disc
        info = np.zeros(self.nprob * d.nspin * (d.lh + 1) * d.nh, dtype=np.int32)
        self.ctx.check(_capi.lib.aks_debug_eig_info(self.h, _capi.i32ptr(info)))
        info = info.reshape(self.nprob, d.nspin, d.lh + 1, d.nh)
        return info >> 8, info & 0x3f, (info & 0x80) != 0

    def eig_clusters(self):
        """Boolean (nprob, nspin, L, nh): levels the last step certified as members of a cluster of near-degenerate
        levels of one channel (info bit 0x40, k_eig)."""
        d = self.disc
        info = np.zeros(self.nprob * d.nspin * (d.lh + 1) * d.nh, dtype=np.int32)
        self.ctx.check(_capi.lib.aks_debug_eig_info(self.h, _capi.i32ptr(info)))
        return (info.reshape(self.nprob, d.nspin, d.lh + 1, d.nh) & 0x40) != 0

    def phase_ms(self):
        out = (C.c_double * 10)()
        rc = _capi.lib.aks_last_phase_ms(self.h, out)
        return list(out) if rc == AKS_OK else None

    def kernel_ms(self):
        """[(kernel name, device ms)] of the last step, launch order (AKS_PHASE_TIMERS=1 batches only)."""
        cap, stride = 64, 48
        ms = (C.c_double * cap)()
        names = C.create_string_buffer(cap * stride)
        cnt = C.c_int32()
        rc = _capi.lib.aks_last_kernel_ms(self.h, cap, ms, names, stride, C.byref(cnt))
        if rc != AKS_OK:
            return None
        raw = names.raw
        return [(raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode(), ms[i]) for i in range(min(cnt.value, cap))]

    def close(self):
        if self.h:
            _capi.lib.aks_batch_destroy(self.h)
            self.h = None


class KSESolver:
    """Single-problem solver state with the reference's fields (solver.jl:41-88)."""

    def __init__(self, model: KSEModel, discretization: KSEDiscretization, alg: ODA, *, maxiter=100,
                 callback=None):
        self.model, self.discretization, self.alg = model, discretization, alg
        self.maxiter = int(maxiter)
        self.callback = callback
        self.niter = 0
        self.stopping_criteria = 0.0
        self.energies = Energies()
        self.logbook = LogBook()
        self.t = alg.tinit
        self.status = AKS_RUNNING
        self.batch = Batch([model], discretization, alg)

    def scf_step(self):
        r = self.batch.step()[0]
        self.stopping_criteria = r["stop"]
        self.energies = Energies(r["Etot"], r["Ekin"], r["Ecou"], r["Ehar"], r["Eexc"])
        self.t = r["t"]
        self.status = r["status"]
        return r["stop"]

    def solve(self):
        """solve! (groundstate.jl:53-62)."""
        while self.niter < self.maxiter:
            self.scf_step()
            self.logbook.stopping_criteria.append(self.stopping_criteria)
            self.logbook.Etot.append(self.energies.Etot)
            if self.callback is not None:
                self.callback(self)
            self.niter += 1
            if self.status < 0:
                raise _capi.AKSError(self.status, "SCF step failed")
            if self.stopping_criteria < self.alg.scftol:
                break
        return self


@dataclass
class KSESolution:
    """The outputs of the hot path in the reference's naming (solution/types.jl:74-140)."""
    name: str
    success: str
    niter: int
    stopping_criteria: float
    energies: Energies
    ϵ: np.ndarray
    n: np.ndarray
    U: np.ndarray | None
    D: np.ndarray | None
    W: np.ndarray
    logbook: LogBook | None = None
    context: tuple | None = None   # (model, discretization), as `sol.context` of the reference

    @property
    def eps(self):
        return self.ϵ

    @property
    def occupied(self):
        """[(label, eps, n)] sorted by energy (solution/types.jl `occupied`)."""
        out = []
        nn, ee = self.n, self.ϵ
        if nn.ndim == 4:   # Double64 solution: trailing (hi, lo) axis; labels and sorting need the leading part only
            nn, ee = nn[..., 0] + nn[..., 1], ee[..., 0] + ee[..., 1]
        L, nh, s = nn.shape
        for sg in range(s):
            for l in range(L):
                for k in range(nh):
                    if nn[l, k, sg] != 0:
                        lab = f"{k + l + 1}{'spdfgh'[l]}" + ("" if s == 1 else "↑↓"[sg])
                        out.append((lab, float(ee[l, k, sg]), float(nn[l, k, sg])))
        return sorted(out, key=lambda v: v[1])


def _solution(batch: Batch, i: int, rec: dict, logbook=None, want_arrays=True):
    st = batch.state(i, want_D=want_arrays, want_U=want_arrays)
    m = batch.models[i]
    succ = "SUCCESS" if rec["status"] == AKS_SUCCESS else ("MAXITERS" if rec["status"] in (AKS_MAXITERS, AKS_RUNNING)
                                                            else f"ERROR({rec['status']})")
    return KSESolution(solution_name(m.Z, m.N), succ, rec["niter"], rec["stop"],
                       Energies(rec["Etot"], rec["Ekin"], rec["Ecou"], rec["Ehar"], rec["Eexc"]),
                       st["eps"], st["n"], st.get("U"), st.get("D"), st["W"], logbook, (m, batch.disc))


def _run_batch(models, discretization, alg, maxiter, lh, nh, want_arrays, history=True):
    """One device batch to convergence -> list of KSESolution (the body of the batched `groundstate`)."""
    per_problem = not np.isscalar(maxiter)
    # problems are dealt over the library's stream lanes by cost (sweep.lane_balanced_order); the solutions come back in
    # the caller's order
    n = len(models)
    order = list(range(n))
    if n >= 16 and not isinstance(alg, (list, tuple)):   # (a per-problem alg list takes its shared options from entry 0)
        from .sweep import lane_balanced_order
        dl, dn = discretization.lh, discretization.nh
        cost = [((lh[i] if lh is not None and not np.isscalar(lh) else (dl if lh is None else lh)) + 1.0)
                * (nh[i] if nh is not None and not np.isscalar(nh) else (dn if nh is None else nh)) + 4.0 for i in range(n)]
        order = lane_balanced_order(cost)
    pick = lambda v: v if (v is None or np.isscalar(v)) else [v[i] for i in order]
    algp = [alg[i] for i in order] if isinstance(alg, (list, tuple)) else alg
    batch = Batch([models[i] for i in order], discretization, algp, lh=pick(lh), nh=pick(nh),
                  maxiter=pick(maxiter) if per_problem else None)
    final, hist = batch.run(int(max(maxiter)) if per_problem else int(maxiter), history=history,
                            raise_on_problem_error=len(models) == 1)
    sols = [None] * n
    for j, rec in enumerate(final):
        lb = LogBook([h["stop"] for h in hist[j]], [h["Etot"] for h in hist[j]]) if hist else None
        sols[order[j]] = _solution(batch, j, rec, lb, want_arrays)
    batch.close()
    return sols


def groundstate(model, discretization: KSEDiscretization, alg, *, maxiter=100, callback=None,
                lh=None, nh=None, want_arrays=True, devices=None):
    """Ground state of one model (-> KSESolution) or of a list of models (-> list), on the B200.

    With a callback the loop runs one `aks_scf_step` per iteration so that `register!` and
    the callback fire exactly as in the reference; without one the whole loop is a single
    `aks_scf_run`.

    For a list of models, `alg` and `maxiter` may be sequences with one entry per model (the reference picks
    both per atom, examples/atomic density comparison/run.jl:74-76), and `devices = [0, 1, ...]` partitions the
    list over several GPUs of the node: the problems are independent (the only parallel axis of the reference,
    `Threads.@threads for z`, examples/periodic table/process.jl:92), so each device runs its share as one batch
    from its own host thread, with its own copy of the tables and no device-to-device traffic; the solutions
    come back in the order of `model`.
    """
    if isinstance(model, KSEModel):
        if callback is not None:
            solver = KSESolver(model, discretization, alg, maxiter=maxiter, callback=callback).solve()
            rec = dict(niter=solver.niter, status=solver.status, stop=solver.stopping_criteria,
                       **solver.energies.__dict__)
            sol = _solution(solver.batch, 0, rec, solver.logbook, want_arrays)
            solver.batch.close()
            return sol
        models, single = [model], True
    else:
        models, single = list(model), False
    if devices is not None and len(devices) > 1 and len(models) > 1:
        return _groundstate_multi(models, discretization, alg, maxiter, lh, nh, want_arrays, list(devices))
    if devices is not None and len(devices) == 1:
        discretization = discretization.on_device(int(devices[0]))
    sols = _run_batch(models, discretization, alg, maxiter, lh, nh, want_arrays)
    return sols[0] if single else sols


def _groundstate_multi(models, discretization, alg, maxiter, lh, nh, want_arrays, devices):
    """Partitioned sweep: static LPT split (sweep.partition) by a per-problem cost, one host thread per GPU."""
    import threading
    from .sweep import partition
    n = len(models)
    L = [discretization.lh] * n if lh is None else list(lh)
    K = [discretization.nh] * n if nh is None else list(nh)
    costs = [(L[i] + 1) * K[i] + 4.0 for i in range(n)]
    parts = [p for p in partition(costs, len(devices)) if p]
    pick = lambda v, idx: v if (v is None or np.isscalar(v) or isinstance(v, ODA)) else [v[i] for i in idx]
    out, errs = [None] * n, []

    def work(dev, idx):
        try:
            disc = discretization.on_device(int(dev))
            sols = _run_batch([models[i] for i in idx], disc, pick(alg, idx), pick(maxiter, idx), pick(lh, idx),
                              pick(nh, idx), want_arrays)
            for i, sol in zip(idx, sols):
                out[i] = sol
        except Exception as e:   # surfaced in the caller's thread
            errs.append(e)

    threads = [threading.Thread(target=work, args=(dev, idx)) for dev, idx in zip(devices, parts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return out
