"""KSEDiscretization: tables of one radial discretization, host- and device-side.

Mirrors `KSEDiscretization(basis, model; lh, nh=10, fem_integration_method=GaussLegendre(basis))`
+ `init_cache!` (`src/discretization/discretization.jl:270-341`).  The density-independent
tables are built once on the host (`femops.py`, `quadrature.py`) and handed to the CUDA
library through `aks_disc_create`; everything density-dependent lives on the device.
One discretization serves any number of models (a batch) with the same spin count.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .basis import FEMBasis
from .femops import FEMOperators, build_femops
from .quadrature import GaussLegendre


class KSEDiscretization:
    def __init__(self, basis: FEMBasis, model=None, *, lh: int, nh: int = 10, nspin: int | None = None,
                 fem_integration_method: GaussLegendre | None = None, femops: FEMOperators | None = None,
                 device: int = 0, dtype: str = "float64", table_builder: str = "device"):
        self.basis = basis
        if dtype not in ("float64", "double64"):
            raise ValueError("dtype must be 'float64' or 'double64'")
        self.dtype = dtype   # 'double64' = T = Double64 of the reference: double-double tables and kernels
        self.lh, self.nh = int(lh), int(nh)
        self.nspin = int(nspin if nspin is not None else (model.nspin if model is not None else 1))
        self.Nh = basis.size
        self.Rmax = basis.mesh.last
        if dtype == "double64":
            # Gauss tables in double-double are only needed (and only built by default) when the model has
            # an exchange-correlation functional; the reference's own Double64 example runs without one
            from .dd import FEMOperatorsDD, GaussLegendreDD, build_femops_dd
            if (femops is not None and not isinstance(femops, FEMOperatorsDD)) or \
                    (fem_integration_method is not None and not isinstance(fem_integration_method, GaussLegendreDD)):
                raise TypeError("dtype='double64' needs FEMOperatorsDD / GaussLegendreDD tables (atomickohnsham_b200.dd)")
            # tables the caller does not bring are built by the library on the device (aks_dd_build_operators,
            # aks_dd_gauss_legendre) or, table_builder="host", by the mpmath / NumPy mirror (dd.py)
            if table_builder not in ("device", "host"):
                raise ValueError("table_builder must be 'device' or 'host'")
            tctx = _capi.default_context(device) if table_builder == "device" and (
                femops is None or (fem_integration_method is None and model is not None and model.has_exchcorr)) else None
            if fem_integration_method is None and model is not None and model.has_exchcorr:
                fem_integration_method = GaussLegendreDD(basis, ctx=tctx)
            self.fem_integration_method = fem_integration_method
            self.femops = femops or build_femops_dd(basis, ctx=tctx)
        else:
            self.fem_integration_method = fem_integration_method or GaussLegendre(basis)
            self.femops = femops or build_femops(basis)
        self.device = device
        self._handle = None
        self._ctx = None
        self._keep = None

    def on_device(self, device: int) -> "KSEDiscretization":
        """The same discretization (host tables shared, nothing recomputed) bound to another GPU: every device
        of a partitioned sweep holds its own copy of the tables (a few MB), there is no peer access."""
        if device == self.device:
            return self
        return KSEDiscretization(self.basis, None, lh=self.lh, nh=self.nh, nspin=self.nspin,
                                 fem_integration_method=self.fem_integration_method, femops=self.femops,
                                 device=device, dtype=self.dtype)

    # ---- device side -------------------------------------------------------------------
    @property
    def ctx(self) -> _capi.Context:
        if self._ctx is None:
            self._ctx = _capi.default_context(self.device)
        return self._ctx

    def handle(self):
        """aks_disc* of this discretization (created on first use)."""
        if self._handle is not None:
            return self._handle
        if self.dtype == "double64":
            return self._handle_dd()
        b, ops, q = self.basis, self.femops, self.fem_integration_method
        n, Cn = b.nloc, b.ncells
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        cell_len = i64([len(v) for v in b.cells_to_indices])
        cidx = np.zeros((Cn, n), dtype=np.int64)
        cgen = np.zeros((Cn, n), dtype=np.int64)
        for c in range(Cn):
            L = len(b.cells_to_indices[c])
            cidx[c, :L] = np.asarray(b.cells_to_indices[c]) + 1
            cgen[c, :L] = np.asarray(b.cells_to_generators[c]) + 1
        Fi, Fj, Fm, Fv = ops.F_coo()
        keep = dict(
            mesh=f64(b.mesh.points), invshifts=f64(b.invshifts.reshape(-1)), x=f64(q.x), w=f64(q.w),
            Pgenx=f64(q.Pgenx.T.reshape(-1)),          # column-major (p+1, Q)
            y=f64(q.y), wy=f64(q.wy), ycell=i64(q.ycell + 1), Pgeny=f64(q.Pgeny.T.reshape(-1)),
            M0=f64(ops.M0.T.reshape(-1)), A=f64(ops.A.T.reshape(-1)), Mm1=f64(ops.Mm1.T.reshape(-1)),
            Mm2=f64(ops.Mm2.T.reshape(-1)), Fi=i64(Fi + 1), Fj=i64(Fj + 1), Fm=i64(Fm + 1), Fv=f64(Fv),
            cell_len=cell_len, cidx=i64(cidx.reshape(-1)), cgen=i64(cgen.reshape(-1)))
        d = _capi.DiscDesc()
        d.dtype, d.p, d.C, d.Nh = (1 if self.dtype == "double64" else 0), b.ordermax, Cn, b.size
        d.Q, d.G, d.lh, d.nh, d.nspin = q.npoints, q.npoints, self.lh, self.nh, self.nspin
        for name in ("mesh", "invshifts", "x", "w", "Pgenx", "y", "wy", "Pgeny", "M0", "A", "Mm1", "Mm2", "Fv"):
            setattr(d, name, _capi.dptr(keep[name]))
        d.ycell = _capi.iptr(keep["ycell"])
        d.nF = len(Fv)
        d.Fi, d.Fj, d.Fm = _capi.iptr(keep["Fi"]), _capi.iptr(keep["Fj"]), _capi.iptr(keep["Fm"])
        d.cell_len = _capi.iptr(keep["cell_len"])
        d.cell_indices = _capi.iptr(keep["cidx"])
        d.cell_generators = _capi.iptr(keep["cgen"])
        h = C.c_void_p()
        self.ctx.check(_capi.lib.aks_disc_create(self.ctx.h, C.byref(d), C.byref(h)))
        self._handle = h
        return h

    def _handle_dd(self):
        """aks_disc* with dtype = 1: every floating-point table is an array of interleaved (hi, lo) pairs."""
        b, ops = self.basis, self.femops
        n, Cn = b.nloc, b.ncells
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        cidx = np.zeros((Cn, n), dtype=np.int64)
        cgen = np.zeros((Cn, n), dtype=np.int64)
        for c in range(Cn):
            L = len(b.cells_to_indices[c])
            cidx[c, :L] = np.asarray(b.cells_to_indices[c]) + 1
            cgen[c, :L] = np.asarray(b.cells_to_generators[c]) + 1
        Fi, Fj, Fm, Fv = ops.F_coo_pairs()
        pairs_of = lambda v: np.ascontiguousarray(np.stack([np.asarray(v, dtype=np.float64),
                                                            np.zeros(len(v))], axis=-1))
        # global matrices are symmetric: row- and column-major coincide
        keep = dict(mesh=pairs_of(b.mesh.points), M0=ops.assemble_pairs(ops.M0_loc), A=ops.assemble_pairs(ops.A_loc),
                    Mm1=ops.assemble_pairs(ops.Mm1_loc), Mm2=ops.assemble_pairs(ops.Mm2_loc),
                    Fi=i64(Fi + 1), Fj=i64(Fj + 1), Fm=i64(Fm + 1), Fv=Fv,
                    cell_len=i64([len(v) for v in b.cells_to_indices]), cidx=i64(cidx.reshape(-1)),
                    cgen=i64(cgen.reshape(-1)))
        d = _capi.DiscDesc()
        d.dtype, d.p, d.C, d.Nh = 1, b.ordermax, Cn, b.size
        d.Q, d.G, d.lh, d.nh, d.nspin = 0, 0, self.lh, self.nh, self.nspin
        q = self.fem_integration_method
        names = ["mesh", "M0", "A", "Mm1", "Mm2", "Fv"]
        if q is not None:
            colmajor = lambda t: np.ascontiguousarray(np.transpose(t.pairs(), (1, 0, 2)))   # (p+1, Q) column-major
            keep.update(x=q.x.pairs(), w=q.w.pairs(), y=q.y.pairs(), wy=q.wy.pairs(), Pgenx=colmajor(q.Pgenx),
                        Pgeny=colmajor(q.Pgeny), invshifts=q.invshifts.pairs(), ycell=i64(q.ycell + 1))
            d.Q = d.G = q.npoints
            d.ycell = _capi.iptr(keep["ycell"])
            names += ["x", "w", "y", "wy", "Pgenx", "Pgeny", "invshifts"]
        for name in names:
            setattr(d, name, _capi.dptr(keep[name]))
        self._keep = keep
        d.nF = len(Fi)
        d.Fi, d.Fj, d.Fm = _capi.iptr(keep["Fi"]), _capi.iptr(keep["Fj"]), _capi.iptr(keep["Fm"])
        d.cell_len = _capi.iptr(keep["cell_len"])
        d.cell_indices = _capi.iptr(keep["cidx"])
        d.cell_generators = _capi.iptr(keep["cgen"])
        h = C.c_void_p()
        self.ctx.check(_capi.lib.aks_disc_create(self.ctx.h, C.byref(d), C.byref(h)))
        self._handle = h
        return h

    def close(self):
        if self._handle is not None:
            _capi.lib.aks_disc_destroy(self._handle)
            self._handle = None
