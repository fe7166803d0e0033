// The device table builders of csrc/k_dd_tables.cuh run on the HOST (same __host__ __device__ code), as a shared
// library for tools/check_dd_tables_host.py: lets the CPU-only container check them against dd.py / mpmath.
#include <vector>
#include "../../atomickohnsham_b200/csrc/k_dd_tables.cuh"

extern "C" void ddt_host_gauss(int n, double* x, double* w) {
  for (int i = 0; i < n; ++i) {
    dd X, W;
    ddt_gauss(n, i, &X, &W);
    x[2 * i] = X.hi; x[2 * i + 1] = X.lo; w[2 * i] = W.hi; w[2 * i + 1] = W.lo;
  }
}
extern "C" void ddt_host_femops(int p, int C, const double* mesh, const int* present, int nq, double* M0, double* A,
                                double* Mm1, double* Mm2, double* F) {
  const int n = p + 1;
  std::vector<dd> x(nq), w(nq), G((size_t)n * nq), dG((size_t)n * nq), wr(nq), wr2(nq), gri(nq);
  for (int q = 0; q < nq; ++q) ddt_gauss(nq, q, &x[q], &w[q]);
  for (int q = 0; q < nq; ++q) {
    dd g[DDT_PMAX + 1], dg[DDT_PMAX + 1];
    ddt_generators(p, x[q], g, dg);
    for (int i = 0; i < n; ++i) { G[(size_t)i * nq + q] = g[i]; dG[(size_t)i * nq + q] = dg[i]; }
  }
  for (int c = 0; c < C; ++c) {
    for (int q = 0; q < nq; ++q) ddt_cell_node(mesh[c], mesh[c + 1], q, x.data(), w.data(), wr.data(), wr2.data(), gri.data());
    DdtCell cell;
    cell.p = p; cell.n = n; cell.nq = nq; cell.G = G.data(); cell.dG = dG.data(); cell.w = w.data();
    cell.wr = wr.data(); cell.wr2 = wr2.data(); cell.gri = gri.data();
    cell.h2 = (dd_make(mesh[c + 1]) - dd_make(mesh[c])) * 0.5;
    cell.ih2 = dd_make(1.0) / cell.h2;
    cell.first = mesh[c] == 0.0;
    ddt_cell_entries(cell, present + (size_t)c * n, 0, 1, (dd*)M0 + (size_t)c * n * n, (dd*)A + (size_t)c * n * n,
                     (dd*)Mm1 + (size_t)c * n * n, (dd*)Mm2 + (size_t)c * n * n, (dd*)F + (size_t)c * n * n * n);
  }
}
