"""CPU check of the device table builders (csrc/k_dd_tables.cuh, compiled for the host by tools/micro/dd_tables_host.cu)
against the host mirror dd.py: Gauss rule and local operator blocks in double-double."""
import ctypes as C, os, sys, time
import numpy as np
import mpmath as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import atomickohnsham_b200 as aks
from atomickohnsham_b200.dd import DD, build_femops_dd, gauss_legendre_dd
lib = C.CDLL(os.path.join(ROOT, "tools", "micro", "dd_tables_host.so"))
dp = C.POINTER(C.c_double)
mp.mp.dps = 50


def reldiff(a_pairs, b: DD):
    a = a_pairs.reshape(b.hi.shape + (2,))
    d = (a[..., 0] - b.hi) + (a[..., 1] - b.lo)
    return float(np.max(np.abs(d)) / max(np.max(np.abs(b.hi)), 1e-300))


for n in (7, 64, 140, 1000):
    x, w = np.zeros((n, 2)), np.zeros((n, 2))
    t0 = time.time()
    lib.ddt_host_gauss(n, x.ctypes.data_as(dp), w.ctypes.data_as(dp))
    t1 = time.time()
    xr, wr = gauss_legendre_dd(n)
    print(f"gauss n={n}: x {reldiff(x, xr):.1e} w {reldiff(w, wr):.1e}  ({t1 - t0:.3f} s host C++, {time.time() - t1:.2f} s dd.py)")
for mesh, p in ((aks.expmesh(0, 30, 9, s=1.5), 6), (aks.expmesh(0, 300, 41, s=1.5), 20)):
    basis = aks.P1IntLegendreBasis(mesh, ordermax=p)
    Cn, n = basis.ncells, p + 1
    present = np.zeros((Cn, n), dtype=np.int32)
    for c in range(Cn):
        present[c, np.asarray(basis.cells_to_generators[c])] = 1
    pts = np.ascontiguousarray(basis.mesh.points, dtype=np.float64)
    outs = [np.zeros((Cn, n, n, 2)) for _ in range(4)] + [np.zeros((Cn, n, n, n, 2))]
    t0 = time.time()
    lib.ddt_host_femops(p, Cn, pts.ctypes.data_as(dp), present.ctypes.data_as(C.POINTER(C.c_int)), 3 * p + 80, *[o.ctypes.data_as(dp) for o in outs])
    t1 = time.time()
    ref = build_femops_dd(basis)
    t2 = time.time()
    print(f"femops p={p} C={Cn}: " + " ".join(f"{nm} {reldiff(o, getattr(ref, nm + '_loc')):.1e}" for nm, o in zip(("M0", "A", "Mm1", "Mm2", "F"), outs)) + f"  ({t1 - t0:.2f} s host C++ one thread, {t2 - t1:.2f} s dd.py)")
