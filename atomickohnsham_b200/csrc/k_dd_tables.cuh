// Double64 table generation on the device ("table generation / setup").
//
// Replaces, for T = Double64, what the reference builds once per discretization on the host:
//   GaussLegendre(basis, npoints): nodes and weights by QuadGK.gauss(T, n)          src/fem/integration methods.jl:46-70
//   init_cache!: M0, A, M-1, M-2 and the mass tensor F                                src/discretization/discretization.jl:320-341
//                (exact / singular local matrices: src/fem/local matrix.jl:5-123, src/fem/matrices.jl:297-343)
// The host mirror (atomickohnsham_b200/dd.py) does the same in mpmath + NumPy double-double and needs 4.5 s for the
// operators and 1.2 s for a 1000-point rule at order 20 -- 40x the time of a whole 86-atom sweep.  Here: every local
// block is a Gauss quadrature in double-double arithmetic (dd.cuh), the polynomial parts exact (3p + 80 nodes), the
// 1/r and 1/r^2 weights converged geometrically; one thread per entry, nodes summed in the host mirror's order.
// Every routine is __host__ __device__: tools/micro/dd_tables_host.cu runs the same code on the CPU.
#pragma once
#include "dd.cuh"

#define DDT_PMAX 30   // ordermax <= 29

// P_0 .. P_nmax at x (Bonnet recurrence)
DD_HD void ddt_legendre(int nmax, dd x, dd* P) {
  P[0] = dd_make(1.0);
  if (nmax >= 1) P[1] = x;
  for (int m = 2; m <= nmax; ++m) P[m] = (x * P[m - 1] * (double)(2 * m - 1) - P[m - 2] * (double)(m - 1)) / (double)m;
}
// P_n(x), P_{n-1}(x) for large n without storing the sequence
DD_HD void ddt_legendre_pair(int n, dd x, dd* pn, dd* pn1) {
  dd a = dd_make(1.0), b = x;   // P_0, P_1
  for (int m = 2; m <= n; ++m) {
    const dd c = (x * b * (double)(2 * m - 1) - a * (double)(m - 1)) / (double)m;
    a = b; b = c;
  }
  *pn = b; *pn1 = a;
}

// node i (ascending) and weight of the n-point Gauss-Legendre rule on [-1, 1]: Newton on P_n in Float64 from the
// Chebyshev-like guess, then three steps in double-double (QuadGK.gauss(T, n) of the reference)
DD_HD void ddt_gauss(int n, int i, dd* x_out, dd* w_out) {
  const int k = (i >= n / 2) ? (n - 1 - i) : i;        // mirror: compute the node with x >= 0, descending index k
  if ((n & 1) && i == n / 2) {                         // middle node of an odd rule: x = 0 exactly
    dd pn, pn1;
    ddt_legendre_pair(n, dd_make(0.0), &pn, &pn1);
    const dd dp = (dd_make(0.0) - pn1) * (double)n / dd_make(-1.0);
    *x_out = dd_make(0.0);
    *w_out = dd_make(2.0) / (dp * dp);
    return;
  }
  double x = cos(3.14159265358979323846 * ((double)k + 0.75) / ((double)n + 0.5));
  for (int it = 0; it < 8; ++it) {
    double a = 1.0, b = x;
    for (int m = 2; m <= n; ++m) { const double c = ((2 * m - 1) * x * b - (m - 1) * a) / m; a = b; b = c; }
    const double dp = n * (x * b - a) / (x * x - 1.0);
    x -= b / dp;
  }
  dd X = dd_make(x);
  dd pn, pn1, dp;
  for (int it = 0; it < 3; ++it) {
    ddt_legendre_pair(n, X, &pn, &pn1);
    dp = (X * pn - pn1) * (double)n / (X * X - 1.0);
    X = X - pn / dp;
  }
  ddt_legendre_pair(n, X, &pn, &pn1);
  dp = (X * pn - pn1) * (double)n / (X * X - 1.0);
  const dd w = dd_make(2.0) / ((dd_make(1.0) - X * X) * dp * dp);
  *x_out = (i >= n / 2) ? X : -X;
  *w_out = w;
}

// generator values g_0 .. g_p and derivatives at the reference point x (src/fem/generators.jl:34-57):
// g_0 = 1 + x, g_1 = 1 - x, g_{k+1} = (P_{k+1} - P_{k-1}) / (2k + 1); derivatives +1, -1, P_k
DD_HD void ddt_generators(int p, dd x, dd* G, dd* dG) {
  dd P[DDT_PMAX + 2];
  ddt_legendre(p > 1 ? p : 1, x, P);
  G[0] = dd_make(1.0) + x;
  G[1] = dd_make(1.0) - x;
  if (dG) { dG[0] = dd_make(1.0); dG[1] = dd_make(-1.0); }
  for (int k = 1; k < p; ++k) {
    G[k + 1] = (P[k + 1] - P[k - 1]) / (double)(2 * k + 1);
    if (dG) dG[k + 1] = P[k];
  }
}

#ifdef __CUDACC__
__global__ void ddk_tab_gauss(int n, dd* x, dd* w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ddt_gauss(n, i, &x[i], &w[i]);
}
// G [n][nq], dG [n][nq] at the nodes x
__global__ void ddk_tab_genvals(int p, int nq, const dd* x, dd* G, dd* dG) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  dd g[DDT_PMAX + 1], dg[DDT_PMAX + 1];
  ddt_generators(p, x[q], g, dg);
  for (int i = 0; i <= p; ++i) { G[(size_t)i * nq + q] = g[i]; dG[(size_t)i * nq + q] = dg[i]; }
}
#endif

// ---- one entry of a local block; q-sums in ascending node order, operation order of dd.py::build_femops_dd
struct DdtCell {
  int p, n, nq;
  const dd *G, *dG, *w;    // [n][nq], [n][nq], [nq]
  const dd *wr, *wr2;      // [nq]  w / r and w / r^2 at the nodes of this cell (cells away from the origin)
  const dd *gri;           // [nq]  1 / (h2 (1 + x)): first cell, where the 1/r weight is divided out of the generators
  dd h2, ih2;              // (b - a) / 2 and its reciprocal
  bool first;              // a == 0 (local matrix.jl:36-75)
};
// per-node weights of a cell: thread-parallel in the kernel, a plain loop on the host
DD_HD void ddt_cell_node(double a, double b, int q, const dd* x, const dd* w, dd* wr, dd* wr2, dd* gri) {
  const dd A = dd_make(a), h2 = (dd_make(b) - A) * 0.5, mid = A + h2;
  if (a == 0.0) {
    gri[q] = dd_make(1.0) / (h2 * (dd_make(1.0) + x[q]));
    wr[q] = dd_make(0.0); wr2[q] = dd_make(0.0);
  } else {
    const dd r = h2 * x[q] + mid;
    wr[q] = w[q] / r;
    wr2[q] = w[q] / (r * r);
    gri[q] = dd_make(0.0);
  }
}
// generator g divided by r at node q in the first cell; the hat 1 + x divides exactly: (1 + x) / (h2 (1 + x)) = 1 / h2
DD_HD dd ddt_gr(const DdtCell& c, int g, int q) {
  if (g == 0) return c.ih2;
  return c.G[(size_t)g * c.nq + q] * c.gri[q];
}
DD_HD dd ddt_khat(const DdtCell& c, int i, int j, bool deriv) {
  const dd* A = deriv ? c.dG : c.G;
  dd acc = dd_make(0.0);
  for (int q = 0; q < c.nq; ++q) acc = acc + (A[(size_t)i * c.nq + q] * A[(size_t)j * c.nq + q]) * c.w[q];
  return acc;
}
// M_-1 entry (before the h2 factor): sum_q w/r g_i g_j; first cell: sum_q w (g_i / r) g_j, symmetrised by the caller
DD_HD dd ddt_m1_raw(const DdtCell& c, int i, int j) {
  dd acc = dd_make(0.0);
  for (int q = 0; q < c.nq; ++q) {
    if (c.first) acc = acc + (ddt_gr(c, i, q) * c.G[(size_t)j * c.nq + q]) * c.w[q];
    else acc = acc + (c.G[(size_t)i * c.nq + q] * c.G[(size_t)j * c.nq + q]) * c.wr[q];
  }
  return acc;
}
DD_HD dd ddt_m2_raw(const DdtCell& c, int i, int j) {
  dd acc = dd_make(0.0);
  for (int q = 0; q < c.nq; ++q) {
    if (c.first) acc = acc + (ddt_gr(c, i, q) * ddt_gr(c, j, q)) * c.w[q];
    else acc = acc + (c.G[(size_t)i * c.nq + q] * c.G[(size_t)j * c.nq + q]) * c.wr2[q];
  }
  return acc;
}
// mass tensor entry F_ijm = int phi_i phi_j phi_m / r (before the h2 factor)
DD_HD dd ddt_f_raw(const DdtCell& c, int i, int j, int m) {
  dd acc = dd_make(0.0);
  for (int q = 0; q < c.nq; ++q) {
    const dd gg = c.G[(size_t)j * c.nq + q] * c.G[(size_t)m * c.nq + q];
    if (c.first) acc = acc + ddt_gr(c, i, q) * (gg * c.w[q]);
    else acc = acc + c.G[(size_t)i * c.nq + q] * (gg * c.wr[q]);
  }
  return acc;
}
// all blocks of one cell, entries e0, e0 + stride, ... (the kernel deals them out to threads; the host runs stride 1)
DD_HD void ddt_cell_entries(const DdtCell& cell, const int* pr, int e0, int stride, dd* M0, dd* A, dd* Mm1, dd* Mm2, dd* F) {
  const int n = cell.n;
  const dd zero = dd_make(0.0);
  for (int e = e0; e < n * n; e += stride) {
    const int i = e / n, j = e - i * n;
    const bool on = pr[i] && pr[j];
    M0[e] = on ? ddt_khat(cell, i, j, false) * cell.h2 : zero;
    A[e] = on ? ddt_khat(cell, i, j, true) * cell.ih2 : zero;
    dd m1 = zero, m2 = zero;
    if (on) {
      m1 = ddt_m1_raw(cell, i, j);
      if (cell.first) m1 = (m1 + ddt_m1_raw(cell, j, i)) * 0.5;
      m1 = m1 * cell.h2;
      m2 = ddt_m2_raw(cell, i, j) * cell.h2;
    }
    Mm1[e] = m1;
    Mm2[e] = m2;
  }
  for (int e = e0; e < n * n * n; e += stride) {
    const int i = e / (n * n), r = e - i * n * n, j = r / n, m = r - j * n;
    const bool on = pr[i] && pr[j] && pr[m];
    F[e] = on ? ddt_f_raw(cell, i, j, m) * cell.h2 : zero;
  }
}

#ifdef __CUDACC__
// one CTA per cell; entries dealt out to the threads.  mesh: Float64 points [C+1]; present [C][n]: 1 if the cell has
// generator g.  Outputs [C][n][n] and [C][n][n][n] in generator order, zero where a generator is absent.
// scratch: [C][3][nq] per-node weights.
__global__ void __launch_bounds__(256) ddk_tab_cells(int p, int C, int nq, const double* mesh, const int* present,
                                                      const dd* x, const dd* w, const dd* G, const dd* dG, dd* scratch,
                                                      dd* M0, dd* A, dd* Mm1, dd* Mm2, dd* F) {
  const int c = blockIdx.x, n = p + 1;
  dd* wr = scratch + (size_t)c * 3 * nq;
  dd* wr2 = wr + nq;
  dd* gri = wr2 + nq;
  for (int q = threadIdx.x; q < nq; q += blockDim.x) ddt_cell_node(mesh[c], mesh[c + 1], q, x, w, wr, wr2, gri);
  __syncthreads();
  DdtCell cell;
  cell.p = p; cell.n = n; cell.nq = nq; cell.G = G; cell.dG = dG; cell.w = w; cell.wr = wr; cell.wr2 = wr2; cell.gri = gri;
  cell.h2 = (dd_make(mesh[c + 1]) - dd_make(mesh[c])) * 0.5;
  cell.ih2 = dd_make(1.0) / cell.h2;
  cell.first = mesh[c] == 0.0;
  ddt_cell_entries(cell, present + (size_t)c * n, threadIdx.x, blockDim.x, M0 + (size_t)c * n * n, A + (size_t)c * n * n,
                   Mm1 + (size_t)c * n * n, Mm2 + (size_t)c * n * n, F + (size_t)c * n * n * n);
}
#endif
