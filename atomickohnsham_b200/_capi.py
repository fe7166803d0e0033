"""ctypes binding of include/aks_b200.h.

The product path has no CPU fallback: importing this module raises if the CUDA
library has not been built (`python -m atomickohnsham_b200.build`), and every
call raises `AKSError` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AKS_B200_LIB: an instrumented build of the same library (tools/: -DAKS_EIG_TIMING cycle counters)
LIB_PATH = os.environ.get("AKS_B200_LIB") or os.path.join(_HERE, "libaks_b200.so")

AKS_OK, AKS_EINVAL, AKS_ECUDA, AKS_ENOTCONV, AKS_EDEGEN3, AKS_ENAN, AKS_ENOSYS, AKS_EEIG = 0, -1, -2, -3, -4, -5, -6, -7
AKS_RUNNING, AKS_SUCCESS, AKS_MAXITERS = 0, 1, 2
AKS_XC_NONE, AKS_LDA_X, AKS_LDA_C_PW = 0, 1, 12
STATUS_NAMES = {AKS_EINVAL: "AKS_EINVAL", AKS_ECUDA: "AKS_ECUDA", AKS_ENOTCONV: "AKS_ENOTCONV",
                AKS_EDEGEN3: "AKS_EDEGEN3", AKS_ENAN: "AKS_ENAN", AKS_ENOSYS: "AKS_ENOSYS", AKS_EEIG: "AKS_EEIG"}


class AKSError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")


class IterRecord(C.Structure):
    _fields_ = [("niter", C.c_int32), ("status", C.c_int32), ("stop", C.c_double),
                ("Etot", C.c_double), ("Ekin", C.c_double), ("Ecou", C.c_double),
                ("Ehar", C.c_double), ("Eexc", C.c_double), ("t", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


class DiscDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("p", C.c_int32), ("C", C.c_int32), ("Nh", C.c_int32),
                ("Q", C.c_int32), ("G", C.c_int32), ("lh", C.c_int32), ("nh", C.c_int32),
                ("nspin", C.c_int32),
                ("mesh", _dp), ("invshifts", _dp), ("x", _dp), ("w", _dp), ("Pgenx", _dp),
                ("y", _dp), ("wy", _dp), ("ycell", _ip), ("Pgeny", _dp),
                ("M0", _dp), ("A", _dp), ("Mm1", _dp), ("Mm2", _dp),
                ("nF", C.c_int64), ("Fi", _ip), ("Fj", _ip), ("Fm", _ip), ("Fv", _dp),
                ("cell_len", _ip), ("cell_indices", _ip), ("cell_generators", _ip)]


EXPORTS = ["aks_ctx_create", "aks_ctx_destroy", "aks_last_error", "aks_ctx_stream", "aks_ctx_configure",
           "aks_disc_create", "aks_disc_destroy", "aks_batch_create", "aks_batch_destroy",
           "aks_batch_set_oda", "aks_batch_set_oda_problem", "aks_batch_set_aufbau", "aks_batch_set_eig_window", "aks_batch_reset", "aks_scf_step", "aks_scf_run", "aks_get_state", "aks_get_state_all",
           "aks_set_density", "aks_assemble_hartree", "aks_assemble_vxc", "aks_eig_lowest",
           "aks_density", "aks_eval_grid", "aks_exc_energy", "aks_hartree_energy", "aks_launch_count", "aks_measure_fp64_peak", "aks_measure_dmma_peak",
           "aks_last_phase_ms", "aks_last_kernel_ms", "aks_debug_eig_info", "aks_debug_eig_cycles", "aks_debug_eig_refills",
           "aks_get_records_dd", "aks_dd_xc_eval", "aks_dd_gauss_legendre", "aks_dd_build_operators"]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m atomickohnsham_b200.build` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.aks_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.aks_ctx_destroy.argtypes = [vp]
    lib.aks_ctx_destroy.restype = None
    lib.aks_last_error.argtypes = [vp]
    lib.aks_last_error.restype = C.c_char_p
    lib.aks_ctx_stream.argtypes = [vp]
    lib.aks_ctx_stream.restype = vp
    lib.aks_ctx_configure.argtypes = [vp, C.c_int32, C.c_int32]
    lib.aks_disc_create.argtypes = [vp, C.POINTER(DiscDesc), C.POINTER(vp)]
    lib.aks_disc_destroy.argtypes = [vp]
    lib.aks_disc_destroy.restype = None
    lib.aks_batch_create.argtypes = [vp, C.c_int32, _dp, _dp, _dp, _i32p, _i32p, _i32p, _i32p, C.POINTER(vp)]
    lib.aks_batch_destroy.argtypes = [vp]
    lib.aks_batch_destroy.restype = None
    lib.aks_batch_set_oda.argtypes = [vp, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_double,
                                      C.c_double, C.c_int32, C.c_double, C.c_int32]
    lib.aks_batch_set_eig_window.argtypes = [vp, C.c_int32]
    lib.aks_batch_set_oda_problem.argtypes = [vp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32]
    lib.aks_batch_set_aufbau.argtypes = [vp, C.c_int32, C.c_double, _dp]
    lib.aks_debug_eig_refills.argtypes = [vp, _i32p]
    lib.aks_batch_reset.argtypes = [vp]
    lib.aks_scf_step.argtypes = [vp, C.POINTER(IterRecord)]
    lib.aks_scf_run.argtypes = [vp, C.c_int32, C.POINTER(IterRecord), C.POINTER(IterRecord)]
    lib.aks_get_state.argtypes = [vp, C.c_int32, _dp, _dp, _dp, _dp, _dp]
    lib.aks_get_state_all.argtypes = [vp, _dp, _dp, _dp]
    lib.aks_set_density.argtypes = [vp, C.c_int32, _dp, _dp, C.c_int32]
    lib.aks_assemble_hartree.argtypes = [vp, C.c_int32, _dp]
    lib.aks_density.argtypes = [vp, C.c_int32, _dp]
    lib.aks_assemble_vxc.argtypes = [vp, C.c_int32, _dp]
    lib.aks_eig_lowest.argtypes = [vp, C.c_int32, _dp, _dp, _dp]
    lib.aks_eval_grid.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _dp, _dp]
    lib.aks_exc_energy.argtypes = [vp, C.c_int32, _dp]
    lib.aks_hartree_energy.argtypes = [vp, C.c_int32, _dp]
    lib.aks_launch_count.argtypes = [vp]
    lib.aks_launch_count.restype = C.c_int64
    lib.aks_measure_fp64_peak.argtypes = [vp, _dp]
    lib.aks_measure_dmma_peak.argtypes = [vp, _dp]
    lib.aks_last_phase_ms.argtypes = [vp, _dp]
    lib.aks_last_kernel_ms.argtypes = [vp, C.c_int32, _dp, C.c_char_p, C.c_int32, _i32p]
    lib.aks_debug_eig_info.argtypes = [vp, _i32p]
    lib.aks_debug_eig_cycles.argtypes = [vp, _ip]
    lib.aks_get_records_dd.argtypes = [vp, _dp]
    lib.aks_dd_xc_eval.argtypes = [vp, C.c_int32, _dp, C.c_int32, C.c_int32, _dp, _dp]
    lib.aks_dd_gauss_legendre.argtypes = [vp, C.c_int32, _dp, _dp]
    lib.aks_dd_build_operators.argtypes = [vp, C.c_int32, C.c_int32, _dp, _i32p, C.c_int32, _dp, _dp, _dp, _dp, _dp]
    return lib


lib = _load()


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def i32ptr(a):
    return None if a is None else a.ctypes.data_as(_i32p)


class Context:
    """One CUDA device + stream (aks_ctx)."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.aks_ctx_create(device, C.byref(h))
        self.h = h
        if rc != AKS_OK:
            msg = lib.aks_last_error(h).decode() if h else ""
            raise AKSError(rc, msg)

    def check(self, rc, ok=(AKS_OK,)):
        if rc not in ok:
            raise AKSError(rc, lib.aks_last_error(self.h).decode())
        return rc

    def configure(self, *, use_graphs=None, check_every=None):
        """Execution options (aks_ctx_configure): CUDA-graph replay of the step, status polling period of run()."""
        self.check(lib.aks_ctx_configure(self.h, -1 if use_graphs is None else int(bool(use_graphs)),
                                         0 if check_every is None else int(check_every)))

    @property
    def launches(self) -> int:
        return int(lib.aks_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(lib.aks_ctx_stream(self.h) or 0)

    def fp64_peak_tflops(self) -> float:
        v = C.c_double()
        self.check(lib.aks_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def dmma_peak_tflops(self) -> float:
        v = C.c_double()
        self.check(lib.aks_measure_dmma_peak(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            lib.aks_ctx_destroy(self.h)
            self.h = None


_DEFAULT_CTX = {}


def default_context(device: int = 0) -> Context:
    if device not in _DEFAULT_CTX:
        _DEFAULT_CTX[device] = Context(device)
    return _DEFAULT_CTX[device]
