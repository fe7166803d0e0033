// Double64 (double-double) instantiation of the SCF step -- SURVEY.md section 8 row a20: the reference runs the
// same generic Julia code with T = Double64 (src/discretization/discretization.jl:289,
// examples/basic_example_doublefloat/run.jl).  Here the step is a short kernel sequence in dd arithmetic
// (dd.cuh); one launch sequence advances every problem of the batch:
//
//   ddk_H        assemble_hartree! (F.W + N/Rmax M0, assemble.jl:63-83) + assemble_hamiltonian! (:334-358)
//   ddk_eig      find_orbital! (oda/loop.jl:65-77): per eigenpair, Sturm bisection with FLOAT64 inertia counts
//                (static condensation of the bubbles of every cell + LDL^T of the hat tridiagonal) to isolate
//                level k, then inverse / Rayleigh-quotient iteration in dd on the same block structure;
//                orbital kinetic / Coulomb energies ride along (energies.jl:47-87)
//   ddk_aufbau   aufbau! (aufbau/optimized.jl:99-222): sort, greedy fill, the two extreme fillings of a
//                two-fold degenerate last block
//   ddk_orbB     tensor_matrix_dict! per orbital: b_o = F:(u_o u_o^T) (free allocation.jl:17-36)
//   ddk_poisson  w_o = A\b_o with the prefactored stiffness operator (assemble.jl:71, energies.jl:105-137)
//   ddk_decide   resolve_twofold_degeneracy! (optimized.jl:228-306), compute_total_energy (energies.jl:7-39),
//                line_search_density! / line_search_energy (oda/loop.jl:79-116, line_search.jl:45-80): every
//                Hartree energy is a quadratic form of the occupations with the Gram matrix g_ab = b_a.w_b
//   ddk_mix      D = t D_new + (1-t) D_prev (loop.jl:100) with density! (density.jl:15-41) folded in, ||dD||_F
//   ddk_finish   stopping rule (loop.jl:43-49), record, state mixing
//
// Exchange-correlation in dd is not built: the reference's own Double64 example runs without a functional
// (its Gauss tables are Float64-only, run.jl:9-20); batches on a dd discretization must have ex = ec = none.
#pragma once
#include "dd.cuh"

#define DD_MAXN 30     // p + 1 <= 30
#define DD_NT 64       // threads of the per-cell kernels: one thread per cell, C <= 64
#define DD_MAXORB 96

struct DDDiscDev {
  int p, n, nb, C, Nh, L, nh;
  dd Rmax;
  const dd* A_loc;    // [C][n][n] cell-owned blocks, generator order
  const dd* M0_loc;
  const dd* Mm1_loc;
  const dd* Mm2_loc;
  const dd* F_loc;    // [C][n (m)][n*n (ij)]
  const int* l2g;     // [C][n]
  dd* Afac;           // [C][n][n]  condensed factors of the stiffness blocks (ddk_setup)
  dd* Atri;           // [2][64]    inverse pivots / multipliers of the hat system
};

struct DDBatchDev {
  int B, L, nh;
  const double *Z, *N, *hartree;
  const int *Lp, *nhp;
  double scftol, tinit, degen_tol;
  int frozen_t, max_degen;
  dd* D;        // [B][Nh][Nh]
  dd* Bv;       // [B][Nh]   F:D of the mixed density
  dd* Wv;       // [B][Nh]   A^-1 Bv
  dd* E;        // [B][5]    Etot, Ekin, Ecou, Ehar, Eexc of the mixed density
  dd* t;        // [B]
  int* niter;   // [B]
  int* status;  // [B]
  int* warm;    // [B]
  dd* H;        // [B][L][C][n*n]
  dd* eps;      // [B][L][nh]
  dd* U;        // [B][L][nh][Nh]
  dd* vtmp;     // [B][L][nh][Nh]
  dd* ekin_orb; // [B][L][nh]
  dd* ecou_orb; // [B][L][nh]
  dd* occ;      // [B][L][nh]   occupations of the trial density
  int* eig_fail;// [B]
  dd* scratch;  // [B][L][nh][C][n*n]
  int* norb;    // [B]
  int* orb_lk;  // [B][DD_MAXORB]   l << 8 | k
  dd* orb_n0;   // [B][DD_MAXORB]   first extreme filling (the filling itself when not degenerate)
  dd* orb_n1;   // [B][DD_MAXORB]   second extreme filling
  dd* orb_n;    // [B][DD_MAXORB]   occupations of the trial density
  int* degen;   // [B]
  dd* bloc;     // [B][DD_MAXORB][C][n]
  dd* borb;     // [B][DD_MAXORB][Nh]
  dd* worb;     // [B][DD_MAXORB][Nh]
  dd* Bnew;     // [B][Nh]
  dd* Wnew;     // [B][Nh]
  dd* Enew;     // [B][5]   energies after the line search (committed by ddk_finish)
  dd* tmix;     // [B]      weight of the trial density in the mix (1 in iteration 0)
  dd* tnew;     // [B]      ODA t reported for this step
  dd* norm_part;// [B][nparts]
  int nparts;
  dd* rec;      // [B][7]   stop, Etot, Ekin, Ecou, Ehar, Eexc, t
  aks_iter_record* rec64;  // [B]
};

// ------------------------------------------------------------------ cell-local linear algebra
// K: n x n row-major symmetric block in generator order (0: right hat, 1: left hat, 2..: bubbles).  On exit
// the lower triangle of the bubble block holds L (unit lower) with the INVERSE pivots on the diagonal, rows
// 0 / 1 hold X_g = K_bb^{-1} K_bg, and (s00, s10, s11) is the Schur complement on the two hats; neg counts
// the negative pivots (Sylvester: eigenvalues below the shift that belong to this cell's elimination).
template <class R>
__device__ __forceinline__ void dd_bb_solve(const R* __restrict__ K, int n, R* x) {
  const int nb = n - 2;
  for (int i = 0; i < nb; ++i) {
    R a = x[i];
    const R* Ki = K + (2 + i) * n + 2;
    for (int k = 0; k < i; ++k) a = a - Ki[k] * x[k];
    x[i] = a;
  }
  for (int i = 0; i < nb; ++i) x[i] = x[i] * K[(2 + i) * n + 2 + i];
  for (int i = nb - 1; i >= 0; --i) {
    R a = x[i];
    for (int k = i + 1; k < nb; ++k) a = a - K[(2 + k) * n + 2 + i] * x[k];
    x[i] = a;
  }
}

template <class R>
__device__ void dd_cell_factor(R* __restrict__ K, int n, R& s00, R& s10, R& s11, int& neg) {
  const int nb = n - 2;
  R w[DD_MAXN], dl[DD_MAXN];
  neg = 0;
  for (int j = 0; j < nb; ++j) {
    R* Kj = K + (2 + j) * n + 2;
    R d = Kj[j];
    for (int k = 0; k < j; ++k) {
      w[k] = Kj[k] * dl[k];
      d = d - Kj[k] * w[k];
    }
    if (Sc<R>::hi(d) < 0.0) ++neg;
    if (Sc<R>::hi(d) == 0.0) d = Sc<R>::fromd(1e-100);
    dl[j] = d;
    const R di = Sc<R>::inv(d);
    Kj[j] = di;
    for (int i = j + 1; i < nb; ++i) {
      R* Ki = K + (2 + i) * n + 2;
      R a = Ki[j];
      for (int k = 0; k < j; ++k) a = a - Ki[k] * w[k];
      Ki[j] = a * di;
    }
  }
  R* x = w;
  for (int i = 0; i < nb; ++i) x[i] = K[(2 + i) * n + 0];
  dd_bb_solve<R>(K, n, x);
  s00 = K[0];
  s10 = K[n];
  for (int i = 0; i < nb; ++i) {
    s00 = s00 - K[(2 + i) * n + 0] * x[i];
    s10 = s10 - K[(2 + i) * n + 1] * x[i];
  }
  for (int i = 0; i < nb; ++i) K[2 + i] = x[i];
  for (int i = 0; i < nb; ++i) x[i] = K[(2 + i) * n + 1];
  dd_bb_solve<R>(K, n, x);
  s11 = K[n + 1];
  for (int i = 0; i < nb; ++i) s11 = s11 - K[(2 + i) * n + 1] * x[i];
  for (int i = 0; i < nb; ++i) K[n + 2 + i] = x[i];
}

// shared work arrays of the per-cell kernels (R = double reuses the same storage)
struct DDShared {
  dd s00[DD_NT], s10[DD_NT], s11[DD_NT], trid[DD_NT], tril[DD_NT], hr[DD_NT], hl[DD_NT], xh[DD_NT], red[DD_NT];
  int neg[DD_NT];
  int total;
};

// LDL^T of the hat tridiagonal from the per-cell Schur complements (thread 0); returns its negative pivots
template <class R>
__device__ int dd_tri_factor(const R* s00, const R* s10, const R* s11, R* trid, R* tril, int C) {
  int neg = 0;
  R dprev = Sc<R>::fromd(1.0);
  for (int j = 0; j < C - 1; ++j) {
    R d = s00[j] + s11[j + 1];
    R l = Sc<R>::fromd(0.0);
    if (j > 0) {
      const R e = s10[j];   // cell j couples its left hat (j-1) with its right hat (j)
      l = e * trid[j - 1];
      d = d - l * e;
    }
    if (Sc<R>::hi(d) < 0.0) ++neg;
    if (Sc<R>::hi(d) == 0.0) d = Sc<R>::fromd(1e-100);
    tril[j] = l;
    trid[j] = Sc<R>::inv(d);
    dprev = d;
  }
  (void)dprev;
  return neg;
}

// (Hc - sigma Mc) of every cell into the CTA's scratch, factor; all threads return the inertia count
template <class R>
__device__ int dd_factor_shift(const dd* __restrict__ H, const dd* __restrict__ M, R sigma, void* scratch,
                               DDShared& S, int C, int n) {
  const int c = threadIdx.x;
  R* s00 = (R*)S.s00; R* s10 = (R*)S.s10; R* s11 = (R*)S.s11;
  __syncthreads();
  if (c < C) {
    R* K = (R*)scratch + (size_t)c * n * n;
    const dd* Hc = H + (size_t)c * n * n;
    const dd* Mc = M + (size_t)c * n * n;
    for (int e = 0; e < n * n; ++e) K[e] = Sc<R>::from(Hc[e]) - sigma * Sc<R>::from(Mc[e]);
    R a, b, d;
    int neg;
    dd_cell_factor<R>(K, n, a, b, d, neg);
    s00[c] = a; s10[c] = b; s11[c] = d;
    S.neg[c] = neg;
  }
  __syncthreads();
  if (c == 0) {
    int tot = dd_tri_factor<R>(s00, s10, s11, (R*)S.trid, (R*)S.tril, C);
    for (int j = 0; j < C; ++j) tot += S.neg[j];
    S.total = tot;
  }
  __syncthreads();
  return S.total;
}

// x = K^{-1} b for the factored operator in `fac` ([C][n][n] dd) and the hat factors in S.trid / S.tril;
// b, x: global vectors in the reference numbering (may alias)
__device__ void dd_solve(const dd* __restrict__ fac, DDShared& S, const dd* b, dd* x, const int* __restrict__ l2g,
                         int C, int n) {
  const int c = threadIdx.x, nb = n - 2;
  dd y[DD_MAXN];
  const dd* K = fac + (size_t)c * n * n;
  int base = 0;
  if (c < C) {
    base = l2g[c * n + 2];
    for (int i = 0; i < nb; ++i) y[i] = b[base + i];
    dd_bb_solve<dd>(K, n, y);
    dd r0 = dd_make(0.0), r1 = dd_make(0.0);
    for (int i = 0; i < nb; ++i) {
      r0 = r0 + K[(2 + i) * n + 0] * y[i];
      r1 = r1 + K[(2 + i) * n + 1] * y[i];
    }
    S.hr[c] = r0;
    S.hl[c] = r1;
  }
  __syncthreads();
  if (c == 0) {
    const int m = C - 1;
    for (int j = 0; j < m; ++j) {
      dd rh = b[l2g[j * n + 0]] - S.hr[j] - S.hl[j + 1];
      if (j > 0) rh = rh - S.tril[j] * S.xh[j - 1];
      S.xh[j] = rh;
    }
    for (int j = m - 1; j >= 0; --j) {
      dd v = S.xh[j] * S.trid[j];
      if (j < m - 1) v = v - S.tril[j + 1] * S.xh[j + 1];
      S.xh[j] = v;
    }
  }
  __syncthreads();
  if (c < C) {
    const dd xr = (c < C - 1) ? S.xh[c] : dd_make(0.0);
    const dd xl = (c > 0) ? S.xh[c - 1] : dd_make(0.0);
    for (int i = 0; i < nb; ++i) x[base + i] = y[i] - K[2 + i] * xr - K[n + 2 + i] * xl;
    if (c < C - 1) x[l2g[c * n + 0]] = xr;
  }
  __syncthreads();
}

// local part of a global vector on cell c (generator order; absent hats are zero)
__device__ __forceinline__ void dd_load_local(const dd* x, const int* __restrict__ l2g, int c, int C, int n, dd* v) {
  v[0] = (c < C - 1) ? x[l2g[c * n + 0]] : dd_make(0.0);
  v[1] = (c > 0) ? x[l2g[c * n + 1]] : dd_make(0.0);
  const int base = l2g[c * n + 2];
  for (int i = 0; i < n - 2; ++i) v[2 + i] = x[base + i];
}
__device__ __forceinline__ dd dd_quad_local(const dd* __restrict__ K, const dd* v, int n) {
  dd acc = dd_make(0.0);
  for (int i = 0; i < n; ++i) {
    dd r = dd_make(0.0);
    for (int j = 0; j < n; ++j) r = r + K[i * n + j] * v[j];
    acc = acc + r * v[i];
  }
  return acc;
}
// sum over the CTA in a fixed order; every thread receives the same value
__device__ __forceinline__ dd dd_block_sum(dd v, DDShared& S, int nact) {
  __syncthreads();
  S.red[threadIdx.x] = v;
  __syncthreads();
  dd s = dd_make(0.0);
  for (int i = 0; i < nact; ++i) s = s + S.red[i];
  return s;
}
// y = Op x, Op given by cell-owned blocks
__device__ void dd_matvec(const dd* __restrict__ Op, DDShared& S, const dd* x, dd* y, const int* __restrict__ l2g,
                          int C, int n) {
  const int c = threadIdx.x;
  __syncthreads();
  if (c < C) {
    dd v[DD_MAXN];
    dd_load_local(x, l2g, c, C, n, v);
    const dd* K = Op + (size_t)c * n * n;
    const int base = l2g[c * n + 2];
    for (int i = 0; i < n; ++i) {
      dd r = dd_make(0.0);
      for (int j = 0; j < n; ++j) r = r + K[i * n + j] * v[j];
      if (i == 0) S.hr[c] = r;
      else if (i == 1) S.hl[c] = r;
      else y[base + i - 2] = r;
    }
  }
  __syncthreads();
  if (c < C - 1) y[l2g[c * n + 0]] = S.hr[c] + S.hl[c + 1];
  __syncthreads();
}

// ------------------------------------------------------------------ setup: prefactor the stiffness operator
__global__ void __launch_bounds__(DD_NT) ddk_setup(DDDiscDev dv) {
  __shared__ DDShared S;
  const int c = threadIdx.x, n = dv.n, C = dv.C;
  if (c < C) {
    dd* K = dv.Afac + (size_t)c * n * n;
    for (int e = 0; e < n * n; ++e) K[e] = dv.A_loc[(size_t)c * n * n + e];
    dd a, b, d;
    int neg;
    dd_cell_factor<dd>(K, n, a, b, d, neg);
    S.s00[c] = a; S.s10[c] = b; S.s11[c] = d;
  }
  __syncthreads();
  if (c == 0) {
    dd_tri_factor<dd>(S.s00, S.s10, S.s11, S.trid, S.tril, C);
    for (int j = 0; j < C; ++j) { dv.Atri[j] = S.trid[j]; dv.Atri[DD_NT + j] = S.tril[j]; }
  }
}

// ------------------------------------------------------------------ Hamiltonian blocks
__global__ void __launch_bounds__(128) ddk_H(DDDiscDev dv, DDBatchDev bt) {
  const int c = blockIdx.x, b = blockIdx.y, n = dv.n, nn = n * n;
  if (bt.status[b] != 0) return;
  __shared__ dd Wl[DD_MAXN];
  const double har = bt.hartree[b];
  if (threadIdx.x < n) {
    const int g = dv.l2g[c * n + threadIdx.x];
    Wl[threadIdx.x] = (g >= 0 && har != 0.0) ? bt.Wv[(size_t)b * dv.Nh + g] : dd_make(0.0);
  }
  __syncthreads();
  const dd shift = dd_make(bt.N[b]) / dv.Rmax;
  const int Lp = bt.Lp[b];
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    dd fw = dd_make(0.0);
    if (har != 0.0) {
      for (int m = 0; m < n; ++m) fw = fw + dv.F_loc[((size_t)c * n + m) * nn + e] * Wl[m];
      fw = (fw + shift * dv.M0_loc[(size_t)c * nn + e]) * har;
    }
    const dd a = dv.A_loc[(size_t)c * nn + e], m1 = dv.Mm1_loc[(size_t)c * nn + e], m2 = dv.Mm2_loc[(size_t)c * nn + e];
    const dd cou = m1 * (-bt.Z[b]);
    for (int l = 0; l < Lp; ++l) {
      const dd kin = (a + m2 * (double)(l * (l + 1))) * 0.5;
      bt.H[(((size_t)b * bt.L + l) * dv.C + c) * nn + e] = kin + cou + fw;
    }
  }
}

// ------------------------------------------------------------------ eigenpairs
__global__ void __launch_bounds__(DD_NT) ddk_eig(DDDiscDev dv, DDBatchDev bt) {
  __shared__ DDShared S;
  const int k = blockIdx.x, l = blockIdx.y, b = blockIdx.z;
  if (bt.status[b] != 0) return;
  if (l >= bt.Lp[b] || k >= bt.nhp[b]) return;
  const int n = dv.n, nn = n * n, C = dv.C, Nh = dv.Nh, c = threadIdx.x;
  const size_t chan = (size_t)b * bt.L + l, pair = chan * bt.nh + k;
  const dd* H = bt.H + chan * C * nn;
  const dd* M = dv.M0_loc;
  dd* scratch = bt.scratch + pair * C * nn;
  dd* x = bt.U + pair * Nh;
  dd* rhs = bt.vtmp + pair * Nh;
  const int* l2g = dv.l2g;

  // ---- phase A: isolate eigenvalue number k with Float64 inertia counts
  const double Zb = bt.Z[b];
  double lo = -0.6 * Zb * Zb - 1.0, hi = 0.5;
  for (int it = 0; it < 12; ++it) {
    if (dd_factor_shift<double>(H, M, lo, scratch, S, C, n) <= k) break;
    lo = 2.0 * lo;
  }
  for (int it = 0; it < 80; ++it) {
    if (dd_factor_shift<double>(H, M, hi, scratch, S, C, n) >= k + 1) break;
    lo = hi;
    hi = 2.0 * hi + 1.0;
  }
  for (int it = 0; it < 120; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (hi - lo <= 1e-10 * fmax(1.0, fabs(mid)) || mid <= lo || mid >= hi) break;
    if (dd_factor_shift<double>(H, M, mid, scratch, S, C, n) >= k + 1) hi = mid; else lo = mid;
  }
  const double mid = 0.5 * (lo + hi);

  // ---- phase B: inverse iteration at the bracket, then Rayleigh-quotient iteration, in dd
  if (!bt.warm[b]) {
    if (c < C) {
      const int base = l2g[c * n + 2];
      for (int i = 0; i < n - 2; ++i) {
        unsigned h = (unsigned)(base + i) * 2654435761u + (unsigned)(k * 97 + l * 13 + 1) * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        x[base + i] = dd_make(0.25 + (double)(h & 0xffff) / 65536.0);
      }
      if (c < C - 1) x[l2g[c * n + 0]] = dd_make(1.0);
    }
  }
  __syncthreads();
  dd_factor_shift<dd>(H, M, dd_make(mid), scratch, S, C, n);
  dd rho = dd_make(mid);
  dd v[DD_MAXN];
  for (int rep = 0; rep < 2; ++rep) {
    dd_matvec(M, S, x, rhs, l2g, C, n);
    dd_solve(scratch, S, rhs, x, l2g, C, n);
    dd part = dd_make(0.0);
    if (c < C) { dd_load_local(x, l2g, c, C, n, v); part = dd_quad_local(M + (size_t)c * nn, v, n); }
    const dd nrm = dd_block_sum(part, S, C);
    const dd sc = dd_inv(dd_sqrt(nrm));
    for (int i = c; i < Nh; i += DD_NT) x[i] = x[i] * sc;
    __syncthreads();
  }
  {
    dd part = dd_make(0.0);
    if (c < C) { dd_load_local(x, l2g, c, C, n, v); part = dd_quad_local(H + (size_t)c * nn, v, n); }
    rho = dd_block_sum(part, S, C);
  }
  int fail = 1;
  for (int it = 0; it < 8; ++it) {
    dd_factor_shift<dd>(H, M, rho, scratch, S, C, n);
    dd_matvec(M, S, x, rhs, l2g, C, n);
    dd_solve(scratch, S, rhs, x, l2g, C, n);
    dd pm = dd_make(0.0), ph = dd_make(0.0);
    if (c < C) {
      dd_load_local(x, l2g, c, C, n, v);
      pm = dd_quad_local(M + (size_t)c * nn, v, n);
      ph = dd_quad_local(H + (size_t)c * nn, v, n);
    }
    const dd nrm = dd_block_sum(pm, S, C);
    const dd hq = dd_block_sum(ph, S, C);
    const dd sc = dd_inv(dd_sqrt(nrm));
    for (int i = c; i < Nh; i += DD_NT) x[i] = x[i] * sc;
    __syncthreads();
    const dd rnew = hq / nrm;
    const double delta = fabs((rnew - rho).hi);
    rho = rnew;
    if (delta <= 1e-25 * fmax(1.0, fabs(rho.hi))) { fail = 0; break; }
  }
  // the pair must be the one the bracket isolated
  if (!(rho.hi >= lo - 1e-6 * fmax(1.0, fabs(mid)) && rho.hi <= hi + 1e-6 * fmax(1.0, fabs(mid)))) fail = 1;
  // sign convention: positive weight on the first bubble of the first cell with a non-zero entry
  // (irrelevant for D; keeps warm starts and outputs reproducible)
  dd pk = dd_make(0.0), pc = dd_make(0.0);
  if (c < C) {
    dd_load_local(x, l2g, c, C, n, v);
    // Kin_l = (A + l(l+1) Mm2)/2, Coulomb = -Z Mm1 (assemble.jl:11,25)
    const dd qa = dd_quad_local(dv.A_loc + (size_t)c * nn, v, n);
    const dd q2 = (l > 0) ? dd_quad_local(dv.Mm2_loc + (size_t)c * nn, v, n) : dd_make(0.0);
    const dd q1 = dd_quad_local(dv.Mm1_loc + (size_t)c * nn, v, n);
    pk = (qa + q2 * (double)(l * (l + 1))) * 0.5;
    pc = q1 * (-Zb);
  }
  const dd ekin = dd_block_sum(pk, S, C);
  const dd ecou = dd_block_sum(pc, S, C);
  if (c == 0) {
    bt.eps[pair] = rho;
    bt.ekin_orb[pair] = ekin;
    bt.ecou_orb[pair] = ecou;
    if (fail) atomicAdd(&bt.eig_fail[b], 1);
  }
}

// ------------------------------------------------------------------ aufbau (one thread per problem)
__global__ void ddk_aufbau(DDDiscDev dv, DDBatchDev bt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bt.B || bt.status[b] != 0) return;
  const int Lp = bt.Lp[b], nhp = bt.nhp[b], ne = Lp * nhp;
  int order[DD_MAXORB];
  dd ef[DD_MAXORB], nf0[DD_MAXORB], nf1[DD_MAXORB];
  // flat index idx = l + L k (convert_index, discretization.jl:379-392)
  for (int idx = 0; idx < ne; ++idx) {
    const int l = idx % Lp, k = idx / Lp;
    ef[idx] = bt.eps[((size_t)b * bt.L + l) * bt.nh + k];
    nf0[idx] = dd_make(0.0);
    nf1[idx] = dd_make(0.0);
  }
  for (int i = 0; i < ne; ++i) {   // stable insertion sort
    int j = i;
    while (j > 0 && ef[i] < ef[order[j - 1]]) { order[j] = order[j - 1]; --j; }
    order[j] = i;
  }
  dd left = dd_make(bt.N[b]);
  int i = 0, degen = 0, err = 0;
  const dd tol = dd_make(bt.degen_tol);
  while (left > dd_make(0.0) && i < ne) {
    int blk[8];
    int len = 1;
    blk[0] = order[i];
    int j = i + 1;
    while (j < ne && dd_abs(ef[order[j]] - ef[blk[0]]) < tol && len < bt.max_degen && len < 8) { blk[len++] = order[j]; ++j; }
    i += len;
    dd tot = dd_make(0.0);
    for (int q = 0; q < len; ++q) tot = tot + (double)(4 * (blk[q] % Lp) + 2);
    if (left >= tot) {
      for (int q = 0; q < len; ++q) { nf0[blk[q]] = dd_make((double)(4 * (blk[q] % Lp) + 2)); nf1[blk[q]] = nf0[blk[q]]; }
      left = left - tot;
    } else if (len == 1) {
      nf0[blk[0]] = left; nf1[blk[0]] = left;
      break;
    } else if (len == 2) {
      const dd d0 = dd_make((double)(4 * (blk[0] % Lp) + 2)), d1 = dd_make((double)(4 * (blk[1] % Lp) + 2));
      const dd n10 = (d0 >= left) ? left : d0;
      const dd n20 = left - n10;
      const dd n21 = (d1 >= left) ? left : d1;
      const dd n11 = left - n21;
      nf0[blk[0]] = n10; nf0[blk[1]] = n20;
      nf1[blk[0]] = n11; nf1[blk[1]] = n21;
      degen = 1;
      break;
    } else {
      err = 1;   // reference: error("degeneracy > 2 not implemented"), optimized.jl:139
      break;
    }
  }
  // orbitals that carry weight in either filling, density! order (k outer, l inner, density.jl:15-41)
  int cnt = 0;
  for (int idx = 0; idx < ne; ++idx) {
    if (nf0[idx].hi != 0.0 || nf1[idx].hi != 0.0) {
      const int l = idx % Lp, k = idx / Lp;
      bt.orb_lk[(size_t)b * DD_MAXORB + cnt] = (l << 8) | k;
      bt.orb_n0[(size_t)b * DD_MAXORB + cnt] = nf0[idx];
      bt.orb_n1[(size_t)b * DD_MAXORB + cnt] = nf1[idx];
      ++cnt;
    }
  }
  bt.norb[b] = cnt;
  bt.degen[b] = degen;
  const int st = err ? AKS_EDEGEN3 : (bt.eig_fail[b] ? AKS_ENAN : 0);
  if (st != 0) {   // errors raised inside a step reach the iteration record
    bt.status[b] = st;
    bt.rec64[b].status = st;
  }
}

// ------------------------------------------------------------------ b_o = F:(u_o u_o^T) per cell
__global__ void __launch_bounds__(32) ddk_orbB(DDDiscDev dv, DDBatchDev bt) {
  const int c = blockIdx.x, o = blockIdx.y, b = blockIdx.z, n = dv.n, nn = n * n;
  if (bt.status[b] != 0 || o >= bt.norb[b]) return;
  __shared__ dd v[DD_MAXN];
  const int lk = bt.orb_lk[(size_t)b * DD_MAXORB + o], l = lk >> 8, k = lk & 0xff;
  const dd* u = bt.U + (((size_t)b * bt.L + l) * bt.nh + k) * dv.Nh;
  if (threadIdx.x == 0) dd_load_local(u, dv.l2g, c, dv.C, n, v);
  __syncthreads();
  const int m = threadIdx.x;
  if (m < n) {
    const dd* F = dv.F_loc + ((size_t)c * n + m) * nn;
    dd acc = dd_make(0.0);
    for (int i = 0; i < n; ++i) {
      dd r = dd_make(0.0);
      for (int j = 0; j < n; ++j) r = r + F[i * n + j] * v[j];
      acc = acc + r * v[i];
    }
    bt.bloc[(((size_t)b * DD_MAXORB + o) * dv.C + c) * n + m] = acc;
  }
}

// ------------------------------------------------------------------ w_o = A^{-1} b_o
__global__ void __launch_bounds__(DD_NT) ddk_poisson(DDDiscDev dv, DDBatchDev bt) {
  __shared__ DDShared S;
  const int o = blockIdx.x, b = blockIdx.y, n = dv.n, C = dv.C, c = threadIdx.x;
  if (bt.status[b] != 0 || o >= bt.norb[b]) return;
  dd* bo = bt.borb + ((size_t)b * DD_MAXORB + o) * dv.Nh;
  dd* wo = bt.worb + ((size_t)b * DD_MAXORB + o) * dv.Nh;
  const dd* bl = bt.bloc + (((size_t)b * DD_MAXORB + o) * C) * n;
  if (c < C) {
    const int base = dv.l2g[c * n + 2];
    for (int i = 0; i < n - 2; ++i) bo[base + i] = bl[c * n + 2 + i];
    if (c < C - 1) bo[dv.l2g[c * n + 0]] = bl[c * n + 0] + bl[(c + 1) * n + 1];
  }
  for (int j = c; j < C; j += DD_NT) { S.trid[j] = dv.Atri[j]; S.tril[j] = dv.Atri[DD_NT + j]; }
  __syncthreads();
  dd_solve(dv.Afac, S, bo, wo, dv.l2g, C, n);
}

// ------------------------------------------------------------------ occupations, energies, ODA parameter
// line_search_energy without an exchange-correlation term (line_search.jl:45-80): the energy along the
// segment is the quadratic a t^2 + b t + c
__device__ __forceinline__ void dd_line_search(dd k0, dd k1, dd c0, dd c1, dd h0, dd h1, dd h01, dd& t, dd& emin) {
  const dd A0 = k0 + c0, A1 = k1 + c1;
  const dd a = h0 + h1 - h01 * 2.0;
  const dd bq = (A0 - A1) + (h01 - h1) * 2.0;
  const dd cq = A1 + h1;
  if (a > dd_make(0.0)) {
    t = -bq / (a * 2.0);
    t = dd_min(dd_max(t, dd_make(0.0)), dd_make(1.0));
  } else {
    t = (cq <= cq + bq + a) ? dd_make(0.0) : dd_make(1.0);
  }
  emin = cq + bq * t + a * t * t;
}

__global__ void __launch_bounds__(128) ddk_decide(DDDiscDev dv, DDBatchDev bt) {
  const int b = blockIdx.x, Nh = dv.Nh, tid = threadIdx.x;
  if (bt.status[b] != 0) return;
  __shared__ dd g[32 * 32];      // Gram matrix b_a . w_b of the involved orbitals
  __shared__ dd hp[32];          // b_a . W_prev
  __shared__ dd nfin[DD_MAXORB];
  const int no = bt.norb[b];
  const dd* borb = bt.borb + (size_t)b * DD_MAXORB * Nh;
  const dd* worb = bt.worb + (size_t)b * DD_MAXORB * Nh;
  const dd* Wp = bt.Wv + (size_t)b * Nh;
  const bool har_on = bt.hartree[b] != 0.0;
  if (no > 32) { if (tid == 0) bt.status[b] = AKS_EINVAL; return; }
  for (int pq = tid; pq < no * no + no; pq += blockDim.x) {
    dd acc = dd_make(0.0);
    if (har_on) {
      if (pq < no * no) {
        const dd* x = borb + (size_t)(pq / no) * Nh;
        const dd* y = worb + (size_t)(pq % no) * Nh;
        for (int i = 0; i < Nh; ++i) acc = acc + x[i] * y[i];
      } else {
        const dd* x = borb + (size_t)(pq - no * no) * Nh;
        for (int i = 0; i < Nh; ++i) acc = acc + x[i] * Wp[i];
      }
    }
    if (pq < no * no) g[pq] = acc; else hp[pq - no * no] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    const dd* n0 = bt.orb_n0 + (size_t)b * DD_MAXORB;
    const dd* n1 = bt.orb_n1 + (size_t)b * DD_MAXORB;
    const int* lk = bt.orb_lk + (size_t)b * DD_MAXORB;
    const dd Nb = dd_make(bt.N[b]);
    const dd self = har_on ? (Nb * Nb) / dv.Rmax : dd_make(0.0);   // N^2 / Rmax
    dd kin[2], cou[2], har[2], seps[2];
    for (int s = 0; s < 2; ++s) {
      const dd* ns = s ? n1 : n0;
      kin[s] = cou[s] = seps[s] = dd_make(0.0);
      dd q = dd_make(0.0);
      for (int a = 0; a < no; ++a) {
        const size_t pair = ((size_t)b * bt.L + (lk[a] >> 8)) * bt.nh + (lk[a] & 0xff);
        kin[s] = kin[s] + ns[a] * bt.ekin_orb[pair];
        cou[s] = cou[s] + ns[a] * bt.ecou_orb[pair];
        seps[s] = seps[s] + ns[a] * bt.eps[pair];
        dd r = dd_make(0.0);
        for (int c2 = 0; c2 < no; ++c2) r = r + g[a * no + c2] * ns[c2];
        q = q + r * ns[a];
      }
      har[s] = har_on ? (q + self) * 0.5 : dd_make(0.0);
    }
    dd e[5];   // Etot, Ekin, Ecou, Ehar, Eexc of the trial density
    if (bt.degen[b]) {
      // resolve_twofold_degeneracy!: line search between the two extreme fillings
      dd q01 = dd_make(0.0);
      for (int a = 0; a < no; ++a) {
        dd r = dd_make(0.0);
        for (int c2 = 0; c2 < no; ++c2) r = r + g[a * no + c2] * n0[c2];   // B[D2] . W[D1]
        q01 = q01 + r * n1[a];
      }
      const dd har01 = har_on ? (q01 + self) * 0.5 : dd_make(0.0);
      dd td, emin;
      dd_line_search(kin[0], kin[1], cou[0], cou[1], har[0], har[1], har01, td, emin);
      const dd um = dd_make(1.0) - td;
      for (int a = 0; a < no; ++a) nfin[a] = td * n0[a] + um * n1[a];
      e[0] = emin;
      e[1] = td * kin[0] + um * kin[1];
      e[2] = td * cou[0] + um * cou[1];
      e[3] = td * td * har[0] + um * um * har[1] + td * um * har01 * 2.0;
      e[4] = e[0] - e[1] - e[2] - e[3];
    } else {
      for (int a = 0; a < no; ++a) nfin[a] = n0[a];
      e[0] = seps[0] - har[0];   // compute_total_energy without a functional: sum n eps - Ehar
      e[1] = kin[0]; e[2] = cou[0]; e[3] = har[0]; e[4] = dd_make(0.0);
    }
    // ODA step (oda/loop.jl:79-116)
    dd t = bt.t[b], tmix = dd_make(1.0);
    dd en[5];
    for (int q = 0; q < 5; ++q) en[q] = e[q];
    if (bt.niter[b] > 0) {
      const dd* ep = bt.E + (size_t)b * 5;
      dd q01 = dd_make(0.0);
      for (int a = 0; a < no; ++a) q01 = q01 + hp[a] * nfin[a];
      const dd har01 = har_on ? (q01 + self) * 0.5 : dd_make(0.0);
      dd etot = dd_make(0.0);
      if (!bt.frozen_t) dd_line_search(e[1], ep[1], e[2], ep[2], e[3], ep[3], har01, t, etot);
      const dd um = dd_make(1.0) - t;
      en[1] = t * e[1] + um * ep[1];
      en[2] = t * e[2] + um * ep[2];
      en[3] = t * t * e[3] + um * um * ep[3] + t * um * har01 * 2.0;
      if (bt.frozen_t) { en[4] = dd_make(0.0); en[0] = en[1] + en[2] + en[3]; }
      else { en[0] = etot; en[4] = etot - en[1] - en[2] - en[3]; }
      tmix = t;
    }
    for (int q = 0; q < 5; ++q) bt.Enew[(size_t)b * 5 + q] = en[q];
    bt.tnew[b] = t;
    bt.tmix[b] = tmix;
    // occupations of the trial density, (l, k) layout
    for (int q = 0; q < bt.L * bt.nh; ++q) bt.occ[(size_t)b * bt.L * bt.nh + q] = dd_make(0.0);
    for (int a = 0; a < no; ++a) {
      bt.orb_n[(size_t)b * DD_MAXORB + a] = nfin[a];
      bt.occ[((size_t)b * bt.L + (lk[a] >> 8)) * bt.nh + (lk[a] & 0xff)] = nfin[a];
    }
  }
  __syncthreads();
  // B, W of the trial density are linear in the occupations
  for (int i = tid; i < Nh; i += blockDim.x) {
    dd bs = dd_make(0.0), ws = dd_make(0.0);
    for (int a = 0; a < no; ++a) {
      bs = bs + borb[(size_t)a * Nh + i] * nfin[a];
      ws = ws + worb[(size_t)a * Nh + i] * nfin[a];
    }
    bt.Bnew[(size_t)b * Nh + i] = bs;
    bt.Wnew[(size_t)b * Nh + i] = ws;
  }
}

// ------------------------------------------------------------------ D = t sum n u u^T + (1-t) D, ||dD||_F^2 partials
__global__ void __launch_bounds__(256) ddk_mix(DDDiscDev dv, DDBatchDev bt) {
  const int b = blockIdx.y, Nh = dv.Nh;
  if (bt.status[b] != 0) return;
  __shared__ dd red[256];
  const int no = bt.norb[b];
  const dd t = bt.tmix[b], um = dd_make(1.0) - t;
  dd* D = bt.D + (size_t)b * Nh * Nh;
  dd part = dd_make(0.0);
  const size_t total = (size_t)Nh * Nh;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / Nh), j = (int)(e % Nh);
    dd acc = dd_make(0.0);
    for (int a = 0; a < no; ++a) {
      const int lk = bt.orb_lk[(size_t)b * DD_MAXORB + a];
      const dd* u = bt.U + (((size_t)b * bt.L + (lk >> 8)) * bt.nh + (lk & 0xff)) * Nh;
      acc = acc + (u[i] * u[j]) * bt.orb_n[(size_t)b * DD_MAXORB + a];
    }
    const dd old = D[e];
    const dd nw = t * acc + um * old;
    const dd df = nw - old;
    part = part + df * df;
    D[e] = nw;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    dd s = dd_make(0.0);
    for (int i = 0; i < (int)blockDim.x; ++i) s = s + red[i];
    bt.norm_part[(size_t)b * bt.nparts + blockIdx.x] = s;
  }
}

// ------------------------------------------------------------------ stopping rule, record, state update
__global__ void __launch_bounds__(128) ddk_finish(DDDiscDev dv, DDBatchDev bt) {
  const int b = blockIdx.x, Nh = dv.Nh, tid = threadIdx.x;
  if (bt.status[b] != 0) return;
  const dd t = bt.tmix[b], um = dd_make(1.0) - t;
  for (int i = tid; i < Nh; i += blockDim.x) {
    const size_t e = (size_t)b * Nh + i;
    bt.Bv[e] = t * bt.Bnew[e] + um * bt.Bv[e];
    bt.Wv[e] = t * bt.Wnew[e] + um * bt.Wv[e];
  }
  __syncthreads();
  if (tid == 0) {
    dd s = dd_make(0.0);
    for (int i = 0; i < bt.nparts; ++i) s = s + bt.norm_part[(size_t)b * bt.nparts + i];
    const dd nrm = dd_sqrt(s);
    dd* E = bt.E + (size_t)b * 5;
    const dd* En = bt.Enew + (size_t)b * 5;
    const dd de = dd_abs(En[0] - E[0]);
    const dd stop = dd_max(nrm, de);
    for (int q = 0; q < 5; ++q) E[q] = En[q];
    bt.t[b] = bt.tnew[b];
    bt.warm[b] = 1;
    const int it = bt.niter[b] + 1;
    bt.niter[b] = it;
    int st = 0;
    bool fin = dd_isfinite(stop);
    for (int q = 0; q < 5; ++q) fin = fin && dd_isfinite(En[q]);
    if (!fin) st = AKS_ENAN;
    else if (stop < dd_make(bt.scftol)) st = AKS_SUCCESS;
    bt.status[b] = st;
    dd* r = bt.rec + (size_t)b * 7;
    r[0] = stop;
    for (int q = 0; q < 5; ++q) r[1 + q] = En[q];
    r[6] = bt.tnew[b];
    aks_iter_record& o = bt.rec64[b];
    o.niter = it; o.status = st; o.stop = stop.hi;
    o.Etot = En[0].hi; o.Ekin = En[1].hi; o.Ecou = En[2].hi; o.Ehar = En[3].hi; o.Eexc = En[4].hi;
    o.t = bt.tnew[b].hi;
  }
}
