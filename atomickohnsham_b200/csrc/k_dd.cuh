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
//   ddk_xc       evaluate_vrho! in dd (models.jl:97-105; SlaterXa.jl:11-12, PerdewWang.jl:15-43) at the Gauss nodes
//   ddk_rho*     optimized_eval_density! / eval_density! (assemble.jl:91-125, density.jl:97-126) from the orbitals,
//                for both extreme fillings, and <Vxc[D_prev], D_new> (energies.jl:15-24)
// With a functional, ddk_H adds B^T diag(w v) B (assemble.jl:277-319, local matrix.jl:129-155) and ddk_decide runs
// Optim's Brent in dd on a t^2 + b t + c + Exc[t rho_A + (1-t) rho_B] (rho is linear in D).  The number-type
// mixture of the reference's Double64 run is kept: (3/pi)^(1/3), 4 pi, 1/(4 pi) and the
// PW92 parameters are Float64 values, rho^(1/3) of Slater exchange is rho^Double64(Float64(1/3)), the rational
// powers of PW92 are exact.  Spin-unpolarised functionals only (nspin = 1).
#pragma once
#include <cooperative_groups.h>

#include "dd.cuh"

#define DD_MAXN 30     // p + 1 <= 30
#define DD_NT 64       // threads of the per-cell kernels: one thread per cell, C <= 64
#define DD_MAXORB 96
#define DD_MAXINV 40    // orbitals that may carry weight in a step (every level of 4 channels x 10 with smearing)

struct DDDiscDev {
  int p, n, nb, C, Nh, L, nh;
  dd Rmax;
  const dd* A_loc;    // [C][n][n] cell-owned blocks, generator order
  const dd* M0_loc;
  const dd* Mm1_loc;
  const dd* Mm2_loc;
  const dd* F_loc;    // [C][n (m)][n*n (ij)]
  const int* l2g;     // [C][n]
  // Gauss tables (NULL / 0 when the discretization was created without them: no functional possible)
  int Q, G;
  const dd* Pg;       // [n][Q]  generator g at Gauss node q
  const dd* xq;       // [Q]
  const dd* wq;       // [Q]
  const dd* inv1;     // [C]  h/2
  const dd* inv2;     // [C]  midpoint
  const dd* y;        // [G]
  const dd* wy;       // [G]
  const int* ycell;   // [G]  0-based
  const dd* Pgy;      // [n][G]
  dd* Afac;           // [npk][C]   packed, cell-interleaved condensed factors of the stiffness blocks (ddk_setup)
  dd* Atri;           // [2][64]    inverse pivots / multipliers of the hat system
};

struct DDBatchDev {
  int B, L, nh;
  const double *Z, *N, *hartree;
  const int *Lp, *nhp, *ex, *ec;
  double scftol, tinit, degen_tol, abstol_ls, reltol_ls;
  int frozen_t, max_degen, maxiter_ls;
  dd* rho_q;      // [B][C][Q]     mixed density at the Gauss nodes
  dd* rho_y;      // [B][G]        mixed density on the energy grid
  dd* fxc;        // [B][C][Q]     w_q vxc(rho(q))
  dd* rho_q_new;  // [B][2][C][Q]  the two extreme fillings of the trial density
  dd* rho_y_new;  // [B][2][G]
  dd* corr_part;  // [B][2][C]     <Vxc[D_prev], D_new> per cell
  dd* td;         // [B]           weight of filling 0 in the trial density
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
  int* kact;    // [B][L]  eigenpair window: levels per channel the first pass of the next step solves
  int* kdone;   // [B][L]  levels per channel that are current for the last Hamiltonian
  int* redo;    // [B]     the window could not be certified: solve the remaining levels, repeat the aufbau
  int eig_window;  // 0: every step solves all nh levels of every channel
  int aufbau_kind; // 0 OptimizedAufbau, 1 SmearedAufbau (temp), 2 FrozenAufbau (nfix)
  double temp;     // smearing width in Hartree (aufbau/smeared.jl)
  double* nfix;    // [B][L][nh] prescribed occupations (aufbau/frozen.jl)
  int* eig_code;// [B][L][nh]  why the warm path was left: 1 RQI not converged, 2 index not certified (debug)
  long long* eig_cyc; // [B][L][nh][6] cycles of the last step: dd factorisations, dd solves, mat-vecs, -, total,
                      // number of dd factorisations (thread 0 of the CTA; debug aid, AKS_DD_EIG_CYCLES)
  int* eig_cold;// [B]   eigenpairs of the last step that went through the cold (bisection) path
  dd* scratch;  // [B][L][nh][npk][C]  packed cell blocks of the eigensolver when they do not fit shared memory
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
// Packed, cell-interleaved storage of one shifted cell block (generator order 0: right hat, 1: left hat, 2..:
// bubbles): element e of cell c lives at K[e * ks] with K = base + c and ks = C, so that the threads of a warp
// (one cell each) touch consecutive addresses -- conflict free in shared memory, coalesced in global memory.
//   e = i(i+1)/2 + j            bubble block, lower triangle (i >= j); after the factorisation: L, with the
//                               INVERSE pivots on the diagonal
//   e = T + g nb + i            column g of the hat-bubble coupling K_bg        (T = nb(nb+1)/2)
//   e = T + 2nb + g nb + i      X_g = K_bb^{-1} K_bg (written by the factorisation)
//   e = T + 4nb + {0,1,2}       K_00, K_10, K_11
__host__ __device__ inline int dd_npk(int nb) { return nb * (nb + 1) / 2 + 4 * nb + 3; }

// init - sum_k term(k) with four independent accumulators: the threads run one cell each at two warps per SM, so
// instruction-level parallelism is what hides the latency of the dependent double-double chains
template <class R, class F>
__device__ __forceinline__ R dd_dot_sub(R init, int cnt, F term) {
  R a0 = init, a1 = Sc<R>::fromd(0.0), a2 = Sc<R>::fromd(0.0), a3 = Sc<R>::fromd(0.0);
  int k = 0;
  for (; k + 3 < cnt; k += 4) {
    a0 = a0 - term(k);
    a1 = a1 - term(k + 1);
    a2 = a2 - term(k + 2);
    a3 = a3 - term(k + 3);
  }
  for (; k < cnt; ++k) a0 = a0 - term(k);
  return (a0 + a1) + (a2 + a3);
}

template <class R>
__device__ __forceinline__ void dd_bb_solve(const R* __restrict__ K, int ks, int nb, R* x) {
  for (int i = 0; i < nb; ++i) {
    const R* Ki = K + (size_t)(i * (i + 1) / 2) * ks;
    x[i] = dd_dot_sub<R>(x[i], i, [&](int k) { return Ki[(size_t)k * ks] * x[k]; });
  }
  for (int i = 0; i < nb; ++i) x[i] = x[i] * K[(size_t)(i * (i + 1) / 2 + i) * ks];
  for (int i = nb - 1; i >= 0; --i)
    x[i] = dd_dot_sub<R>(x[i], nb - 1 - i, [&](int q) {
      const int k = i + 1 + q;
      return K[(size_t)(k * (k + 1) / 2 + i) * ks] * x[k];
    });
}

// (s00, s10, s11): Schur complement on the two hats; neg: negative pivots (Sylvester: eigenvalues below the
// shift that belong to this cell's elimination).  LN lanes of one warp work on a cell (sub = lane within the
// group): every lane forms the pivot of column j (redundantly, identically), the rows below it are dealt out to
// the lanes; with LN > 1 lane 0 returns s00, s10 and lane 1 returns s11 (each solves one hat column).
// lanes of this warp that belong to a cell (< C) when LN lanes work on each cell: the participation mask of the
// warp-level synchronisations and shuffles (computed, not __activemask(): convergence is not guaranteed)
template <int LN>
__device__ __forceinline__ unsigned dd_group_mask(int C) {
  int nact = C * LN - (int)(threadIdx.x / 32) * 32;
  nact = nact < 0 ? 0 : (nact > 32 ? 32 : nact);
  return nact == 32 ? 0xffffffffu : ((1u << nact) - 1u);
}
__device__ __forceinline__ dd dd_shfl_xor(unsigned m, dd v, int lane) {
  return dd_make(__shfl_xor_sync(m, v.hi, lane), __shfl_xor_sync(m, v.lo, lane));
}
template <class R, int LN>
__device__ void dd_cell_factor(R* __restrict__ K, int ks, int nb, int sub, unsigned gm, R& s00, R& s10, R& s11,
                               int& neg) {
  R w[DD_MAXN], dl[DD_MAXN];
  const int T = nb * (nb + 1) / 2;
  neg = 0;
  for (int j = 0; j < nb; ++j) {
    R* Kj = K + (size_t)(j * (j + 1) / 2) * ks;
    for (int k = 0; k < j; ++k) w[k] = Kj[(size_t)k * ks] * dl[k];
    R d = dd_dot_sub<R>(Kj[(size_t)j * ks], j, [&](int k) { return Kj[(size_t)k * ks] * w[k]; });
    if (Sc<R>::hi(d) < 0.0) ++neg;
    if (Sc<R>::hi(d) == 0.0) d = Sc<R>::fromd(1e-100);
    dl[j] = d;
    const R di = Sc<R>::inv(d);
    for (int i = j + 1 + sub; i < nb; i += LN) {
      R* Ki = K + (size_t)(i * (i + 1) / 2) * ks;
      const R a = dd_dot_sub<R>(Ki[(size_t)j * ks], j, [&](int k) { return Ki[(size_t)k * ks] * w[k]; });
      Ki[(size_t)j * ks] = a * di;
    }
    if (LN > 1) __syncwarp(gm);           // every lane has read the pivot entry and row j
    if (sub == 0) Kj[(size_t)j * ks] = di;
    if (LN > 1) __syncwarp(gm);
  }
  R* x = w;
  const R* c0 = K + (size_t)T * ks;
  const R* c1 = K + (size_t)(T + nb) * ks;
  R* x0 = K + (size_t)(T + 2 * nb) * ks;
  R* x1 = K + (size_t)(T + 3 * nb) * ks;
  const R* hh = K + (size_t)(T + 4 * nb) * ks;
  s00 = s10 = s11 = Sc<R>::fromd(0.0);
  if (LN == 1 || sub == 0) {
    for (int i = 0; i < nb; ++i) x[i] = c0[(size_t)i * ks];
    dd_bb_solve<R>(K, ks, nb, x);
    s00 = hh[0];
    s10 = hh[ks];
    for (int i = 0; i < nb; ++i) {
      s00 = s00 - c0[(size_t)i * ks] * x[i];
      s10 = s10 - c1[(size_t)i * ks] * x[i];
      x0[(size_t)i * ks] = x[i];
    }
  }
  if (LN == 1 || sub == 1) {
    for (int i = 0; i < nb; ++i) x[i] = c1[(size_t)i * ks];
    dd_bb_solve<R>(K, ks, nb, x);
    s11 = hh[2 * (size_t)ks];
    for (int i = 0; i < nb; ++i) {
      s11 = s11 - c1[(size_t)i * ks] * x[i];
      x1[(size_t)i * ks] = x[i];
    }
  }
  if (LN > 1) __syncwarp(gm);
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
  }
  return neg;
}

// packed (Op_c - sigma M_c) of cell c from the full blocks (M == nullptr: Op_c alone); the entries are dealt out
// to the LN lanes of the cell
template <class R, int LN>
__device__ __forceinline__ void dd_fill_packed(const dd* __restrict__ Hc, const dd* __restrict__ Mc, R sigma, R* K, int ks,
                                               int n, int sub) {
  const int nb = n - 2, T = nb * (nb + 1) / 2;
  auto val = [&](int e) { return Mc ? Sc<R>::from(Hc[e]) - sigma * Sc<R>::from(Mc[e]) : Sc<R>::from(Hc[e]); };
  int cnt = 0;
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j <= i; ++j)
      if (LN == 1 || (cnt++ % LN) == sub) K[(size_t)(i * (i + 1) / 2 + j) * ks] = val((2 + i) * n + 2 + j);
  for (int g = 0; g < 2; ++g)
    for (int i = 0; i < nb; ++i)
      if (LN == 1 || (cnt++ % LN) == sub) K[(size_t)(T + g * nb + i) * ks] = val((2 + i) * n + g);
  if (sub == 0) {
    K[(size_t)(T + 4 * nb) * ks] = val(0);
    K[(size_t)(T + 4 * nb + 1) * ks] = val(n);
    K[(size_t)(T + 4 * nb + 2) * ks] = val(n + 1);
  }
}

// (Hc - sigma Mc) of every cell into the CTA's scratch, factor; all threads return the inertia count
template <class R, int LN>
__device__ int dd_factor_shift(const dd* __restrict__ H, const dd* __restrict__ M, R sigma, void* scratch,
                               DDShared& S, int C, int n) {
  const int c = threadIdx.x / LN, sub = threadIdx.x % LN;
  R* s00 = (R*)S.s00; R* s10 = (R*)S.s10; R* s11 = (R*)S.s11;
  __syncthreads();
  if (c < C) {
    const unsigned gm = dd_group_mask<LN>(C);
    R* K = (R*)scratch + c;
    dd_fill_packed<R, LN>(H + (size_t)c * n * n, M + (size_t)c * n * n, sigma, K, C, n, sub);
    if (LN > 1) __syncwarp(gm);
    R a, b, d;
    int neg;
    dd_cell_factor<R, LN>(K, C, n - 2, sub, gm, a, b, d, neg);
    if (sub == 0) { s00[c] = a; s10[c] = b; S.neg[c] = neg; }
    if (LN == 1 || sub == 1) s11[c] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = dd_tri_factor<R>(s00, s10, s11, (R*)S.trid, (R*)S.tril, C);
    for (int j = 0; j < C; ++j) tot += S.neg[j];
    S.total = tot;
  }
  __syncthreads();
  return S.total;
}

// x = K^{-1} b for the factored operator in `fac` (packed, interleaved, dd) and the hat factors in S.trid /
// S.tril; b, x: global vectors in the reference numbering (may alias).  Lane 0 of every cell does the (sequential)
// substitutions.
template <int LN>
__device__ void dd_solve(const dd* __restrict__ fac, DDShared& S, const dd* b, dd* x, const int* __restrict__ l2g,
                         int C, int n) {
  const int c = threadIdx.x / LN, sub = threadIdx.x % LN, nb = n - 2, T = nb * (nb + 1) / 2;
  const size_t ks = C;
  dd y[DD_MAXN];
  const dd* K = fac + c;
  int base = 0;
  const bool mine = (c < C && sub == 0);
  if (mine) {
    base = l2g[c * n + 2];
    for (int i = 0; i < nb; ++i) y[i] = b[base + i];
    dd_bb_solve<dd>(K, C, nb, y);
    dd r0 = dd_make(0.0), r1 = dd_make(0.0);
    for (int i = 0; i < nb; ++i) {
      r0 = r0 + K[(T + i) * ks] * y[i];
      r1 = r1 + K[(T + nb + i) * ks] * y[i];
    }
    S.hr[c] = r0;
    S.hl[c] = r1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
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
  if (mine) {
    const dd xr = (c < C - 1) ? S.xh[c] : dd_make(0.0);
    const dd xl = (c > 0) ? S.xh[c - 1] : dd_make(0.0);
    for (int i = 0; i < nb; ++i) x[base + i] = y[i] - K[(T + 2 * nb + i) * ks] * xr - K[(T + 3 * nb + i) * ks] * xl;
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
// v^T K v of one cell: the rows are dealt out to the LN lanes of the cell, every lane returns the total
template <int LN>
__device__ __forceinline__ dd dd_quad_local(const dd* __restrict__ K, const dd* v, int n, int C) {
  const int sub = threadIdx.x % LN;
  const unsigned gm = dd_group_mask<LN>(C);
  dd acc = dd_make(0.0);
  for (int i = sub; i < n; i += LN) {
    dd r = dd_make(0.0);
    for (int j = 0; j < n; ++j) r = r + K[i * n + j] * v[j];
    acc = acc + r * v[i];
  }
  for (int o = 1; o < LN; o <<= 1) acc = acc + dd_shfl_xor(gm, acc, o);
  return acc;
}
// sum of one value per cell over the CTA in a fixed order; every thread receives the same value
template <int LN>
__device__ __forceinline__ dd dd_block_sum(dd v, DDShared& S, int nact) {
  __syncthreads();
  if (threadIdx.x % LN == 0 && (int)(threadIdx.x / LN) < DD_NT) S.red[threadIdx.x / LN] = v;
  __syncthreads();
  dd s = dd_make(0.0);
  for (int i = 0; i < nact; ++i) s = s + S.red[i];
  return s;
}
// y = Op x, Op given by cell-owned blocks
template <int LN>
__device__ void dd_matvec(const dd* __restrict__ Op, DDShared& S, const dd* x, dd* y, const int* __restrict__ l2g,
                          int C, int n) {
  const int c = threadIdx.x / LN, sub = threadIdx.x % LN;
  __syncthreads();
  if (c < C) {
    dd v[DD_MAXN];
    dd_load_local(x, l2g, c, C, n, v);
    const dd* K = Op + (size_t)c * n * n;
    const int base = l2g[c * n + 2];
    for (int i = sub; i < n; i += LN) {
      dd r = dd_make(0.0);
      for (int j = 0; j < n; ++j) r = r + K[i * n + j] * v[j];
      if (i == 0) S.hr[c] = r;
      else if (i == 1) S.hl[c] = r;
      else y[base + i - 2] = r;
    }
  }
  __syncthreads();
  if (sub == 0 && c < C - 1) y[l2g[c * n + 0]] = S.hr[c] + S.hl[c + 1];
  __syncthreads();
}

// ------------------------------------------------------------------ setup: prefactor the stiffness operator
__global__ void __launch_bounds__(DD_NT) ddk_setup(DDDiscDev dv) {
  __shared__ DDShared S;
  const int c = threadIdx.x, n = dv.n, C = dv.C;
  if (c < C) {
    dd* K = dv.Afac + c;
    dd_fill_packed<dd, 1>(dv.A_loc + (size_t)c * n * n, nullptr, dd_make(0.0), K, C, n, 0);
    dd a, b, d;
    int neg;
    dd_cell_factor<dd, 1>(K, C, n - 2, 0, 0u, a, b, d, neg);
    S.s00[c] = a; S.s10[c] = b; S.s11[c] = d;
  }
  __syncthreads();
  if (c == 0) {
    dd_tri_factor<dd>(S.s00, S.s10, S.s11, S.trid, S.tril, C);
    for (int j = 0; j < C; ++j) { dv.Atri[j] = S.trid[j]; dv.Atri[DD_NT + j] = S.tril[j]; }
  }
}

// ------------------------------------------------------------------ LDA exchange-correlation in dd
// Number-type mixture of the reference's Double64 run: Float64 constants, the Slater
// exponent is Float64(1/3) promoted to Double64, PW92's rational powers are exact.
#define DD_CX 0.9847450218426965        // Float64 (3/pi)^(1/3)          (SlaterXa.jl:12)
#define DD_CX34 (-0.7385587663820223)   // Float64 -3/4 * (3/pi)^(1/3)   (SlaterXa.jl:11)
#define DD_THIRD_F64 0.3333333333333333 // Float64 1/3, the exponent of rho^(1/3)
#define DD_FOURPI 12.566370614359172    // Float64 4 pi
#define DD_INV4PI 0.07957747154594767   // Float64 1/(4 pi)              (assemble.jl:123)

// G(rs) and dG/drs of PW92 with the zeta = 0 parameters (PerdewWang.jl:15-30)
__device__ void dd_pw_G(dd rs, dd& G, dd& dG, bool want_d) {
  const double A = 0.031091, a1 = 0.21370, b1 = 7.5957, b2 = 3.5876, b3 = 1.6382, b4 = 0.49294;
  const dd sq = dd_sqrt(rs);
  const dd Q0 = (rs * a1 + 1.0) * (-2.0 * A);
  const dd Q1 = (sq * b1 + rs * b2 + (rs * sq) * b3 + (rs * rs) * b4) * (2.0 * A);
  const dd lg = dd_log1p(dd_inv(Q1));
  G = Q0 * lg;
  if (want_d) {
    const double dQ0 = -2.0 * A * a1;
    const dd dQ1 = (dd_make(b1) / (sq * 2.0) + b2 + sq * (1.5 * b3) + rs * (2.0 * b4)) * (2.0 * A);
    dG = lg * dQ0 - (Q0 * dQ1) / (Q1 * (Q1 + 1.0));
  }
}
__device__ dd dd_rs_of_rho(dd rho) { return dd_cbrt(dd_make(3.0) / (rho * DD_FOURPI)); }   // PerdewWang.jl:34
// vxc = vx + vc at rho >= 0 (clamped by the caller, BuiltinFunctional.jl:112); zero at rho = 0
__device__ dd dd_xc_vrho(dd rho, int ex, int ec) {
  dd v = dd_make(0.0);
  if (rho.hi <= 0.0) return v;
  if (ec == 12) {
    const dd rs = dd_rs_of_rho(rho);
    dd G, dG;
    dd_pw_G(rs, G, dG, true);
    v = v + (G - (rs / 3.0) * dG);
  }
  if (ex == 1) v = v - dd_pow(rho, dd_make(DD_THIRD_F64)) * DD_CX;
  return v;
}
// exc per particle
__device__ dd dd_xc_zk(dd rho, int ex, int ec) {
  dd e = dd_make(0.0);
  if (rho.hi <= 0.0) return e;
  if (ec == 12) {
    dd G, dG;
    dd_pw_G(dd_rs_of_rho(rho), G, dG, false);
    e = e + G;
  }
  if (ex == 1) e = e + dd_pow(rho, dd_make(DD_THIRD_F64)) * DD_CX34;
  return e;
}
// test hook: pointwise vrho / zk of an array of densities
__global__ void ddk_xc_eval(const dd* rho, int npts, int ex, int ec, dd* vrho, dd* zk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const dd r = dd_max(rho[i], dd_make(0.0));
  vrho[i] = dd_xc_vrho(r, ex, ec);
  zk[i] = dd_xc_zk(r, ex, ec);
}

// f = w_q vxc(rho(q)) at every Gauss node of every cell
__global__ void __launch_bounds__(128) ddk_xc(DDDiscDev dv, DDBatchDev bt) {
  const int c = blockIdx.x, b = blockIdx.y;
  if (bt.status[b] != 0) return;
  const int ex = bt.ex[b], ec = bt.ec[b];
  if (ex == 0 && ec == 0) return;
  const size_t off = ((size_t)b * dv.C + c) * dv.Q;
  for (int q = threadIdx.x; q < dv.Q; q += blockDim.x)
    bt.fxc[off + q] = dv.wq[q] * dd_xc_vrho(dd_max(bt.rho_q[off + q], dd_make(0.0)), ex, ec);
}

// rho of the two extreme fillings at the Gauss nodes of cell c, from the orbitals, and <Vxc[D_prev], D_new>
__global__ void __launch_bounds__(128) ddk_rho_nodes(DDDiscDev dv, DDBatchDev bt) {
  const int c = blockIdx.x, b = blockIdx.y, n = dv.n, Q = dv.Q;
  if (bt.status[b] != 0) return;
  if (bt.ex[b] == 0 && bt.ec[b] == 0) return;
  __shared__ dd ul[DD_MAXINV * DD_MAXN];
  __shared__ dd red0[128], red1[128];
  const int no = min(bt.norb[b], DD_MAXINV);
  for (int idx = threadIdx.x; idx < no * n; idx += blockDim.x) {
    const int o = idx / n, g = idx - o * n;
    const int lk = bt.orb_lk[(size_t)b * DD_MAXORB + o], gi = dv.l2g[c * n + g];
    const dd* u = bt.U + (((size_t)b * bt.L + (lk >> 8)) * bt.nh + (lk & 0xff)) * dv.Nh;
    ul[idx] = (gi >= 0) ? u[gi] : dd_make(0.0);
  }
  __syncthreads();
  const dd* n0 = bt.orb_n0 + (size_t)b * DD_MAXORB;
  const dd* n1 = bt.orb_n1 + (size_t)b * DD_MAXORB;
  const dd* f = bt.fxc + ((size_t)b * dv.C + c) * Q;
  dd* out0 = bt.rho_q_new + (((size_t)b * 2 + 0) * dv.C + c) * Q;
  dd* out1 = bt.rho_q_new + (((size_t)b * 2 + 1) * dv.C + c) * Q;
  dd c0 = dd_make(0.0), c1 = dd_make(0.0);
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    dd a0 = dd_make(0.0), a1 = dd_make(0.0);
    for (int o = 0; o < no; ++o) {
      dd psi = dd_make(0.0);
      for (int g = 0; g < n; ++g) psi = psi + ul[o * n + g] * dv.Pg[(size_t)g * Q + q];
      const dd p2 = psi * psi;
      a0 = a0 + p2 * n0[o];
      a1 = a1 + p2 * n1[o];
    }
    const dd r = dv.inv1[c] * dv.xq[q] + dv.inv2[c];
    const dd r2 = r * r;
    out0[q] = (a0 / r2) * DD_INV4PI;     // divide, then multiply (assemble.jl:121-123)
    out1[q] = (a1 / r2) * DD_INV4PI;
    c0 = c0 + f[q] * a0;
    c1 = c1 + f[q] * a1;
  }
  red0[threadIdx.x] = c0;
  red1[threadIdx.x] = c1;
  __syncthreads();
  if (threadIdx.x == 0) {
    dd s0 = dd_make(0.0), s1 = dd_make(0.0);
    for (int i = 0; i < 128; ++i) { s0 = s0 + red0[i]; s1 = s1 + red1[i]; }
    bt.corr_part[((size_t)b * 2 + 0) * dv.C + c] = s0 * dv.inv1[c];
    bt.corr_part[((size_t)b * 2 + 1) * dv.C + c] = s1 * dv.inv1[c];
  }
}
// the same on the global energy grid (eval_density!, density.jl:97-126)
__global__ void __launch_bounds__(128) ddk_rho_grid(DDDiscDev dv, DDBatchDev bt) {
  const int b = blockIdx.y, n = dv.n, G = dv.G;
  if (bt.status[b] != 0) return;
  if (bt.ex[b] == 0 && bt.ec[b] == 0) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= G) return;
  const int c = dv.ycell[q], no = min(bt.norb[b], DD_MAXINV);
  dd a0 = dd_make(0.0), a1 = dd_make(0.0);
  for (int o = 0; o < no; ++o) {
    const int lk = bt.orb_lk[(size_t)b * DD_MAXORB + o];
    const dd* u = bt.U + (((size_t)b * bt.L + (lk >> 8)) * bt.nh + (lk & 0xff)) * dv.Nh;
    dd psi = dd_make(0.0);
    for (int g = 0; g < n; ++g) {
      const int gi = dv.l2g[c * n + g];
      if (gi >= 0) psi = psi + u[gi] * dv.Pgy[(size_t)g * G + q];
    }
    const dd p2 = psi * psi;
    a0 = a0 + p2 * bt.orb_n0[(size_t)b * DD_MAXORB + o];
    a1 = a1 + p2 * bt.orb_n1[(size_t)b * DD_MAXORB + o];
  }
  const dd den = (dv.y[q] * dv.y[q]) * DD_FOURPI;
  bt.rho_y_new[((size_t)b * 2 + 0) * G + q] = a0 / den;
  bt.rho_y_new[((size_t)b * 2 + 1) * G + q] = a1 / den;
}

// ------------------------------------------------------------------ Hamiltonian blocks
__global__ void __launch_bounds__(256) ddk_H(DDDiscDev dv, DDBatchDev bt) {
  const int c = blockIdx.x, b = blockIdx.y, n = dv.n, nn = n * n;
  if (bt.status[b] != 0) return;
  __shared__ dd Wl[DD_MAXN];
  __shared__ dd Vs[DD_MAXN * DD_MAXN];
  const double har = bt.hartree[b];
  const bool has_xc = bt.ex[b] != 0 || bt.ec[b] != 0;
  if (threadIdx.x < n) {
    const int g = dv.l2g[c * n + threadIdx.x];
    Wl[threadIdx.x] = (g >= 0 && har != 0.0) ? bt.Wv[(size_t)b * dv.Nh + g] : dd_make(0.0);
  }
  // exchange-correlation block of the cell: K_ij = h/2 sum_q (w v)_q g_i(x_q) g_j(x_q), lower triangle
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    if (j > i) continue;
    dd acc = dd_make(0.0);
    if (has_xc && dv.l2g[c * n + i] >= 0 && dv.l2g[c * n + j] >= 0) {
      const dd* f = bt.fxc + ((size_t)b * dv.C + c) * dv.Q;
      const dd* gi = dv.Pg + (size_t)i * dv.Q;
      const dd* gj = dv.Pg + (size_t)j * dv.Q;
      acc = -dd_dot_sub<dd>(dd_make(0.0), dv.Q, [&](int q) { return (gi[q] * gj[q]) * f[q]; });
      acc = acc * dv.inv1[c];
    }
    Vs[i * n + j] = acc;
    Vs[j * n + i] = acc;
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
    fw = fw + Vs[e];
    const dd a = dv.A_loc[(size_t)c * nn + e], m1 = dv.Mm1_loc[(size_t)c * nn + e], m2 = dv.Mm2_loc[(size_t)c * nn + e];
    const dd cou = m1 * (-bt.Z[b]);
    for (int l = 0; l < Lp; ++l) {
      const dd kin = (a + m2 * (double)(l * (l + 1))) * 0.5;
      bt.H[(((size_t)b * bt.L + l) * dv.C + c) * nn + e] = kin + cou + fw;
    }
  }
}

// ------------------------------------------------------------------ eigenpairs
// use_smem: the CTA's packed cell blocks live in dynamic shared memory (C * dd_npk(nb) dd) instead of global scratch
extern __shared__ dd dd_dyn_smem[];
#define DD_EIG_LN 4     // lanes per cell in ddk_eig (8 measured slower: 18.0 vs 16.5 ms per C4 step)
template <int LN>
// pass 0: levels below the eigenpair window kact of every running problem (the others get a sentinel);
// pass 1: the remaining levels of the problems whose window was not certified (ddk_aufbau);
// pass 2: the remaining levels of every problem (aks_get_state returns all nh eigenpairs of the last Hamiltonian)
#define DD_EPS_MISSING 1e300
__global__ void __launch_bounds__(LN * DD_NT) ddk_eig(DDDiscDev dv, DDBatchDev bt, int use_smem, int pass) {
  __shared__ DDShared S;
  const int k = blockIdx.x, l = blockIdx.y, b = blockIdx.z;
  if (pass == 2 ? bt.status[b] < 0 : bt.status[b] != 0) return;
  if (l >= bt.Lp[b] || k >= bt.nhp[b]) return;
  {
    const int ch = b * bt.L + l;
    if (pass == 0) {
      if (k >= bt.kact[ch]) {
        if (threadIdx.x == 0) bt.eps[(size_t)ch * bt.nh + k] = dd_make(DD_EPS_MISSING);
        return;
      }
    } else {
      if (pass == 1 && !bt.redo[b]) return;
      if (k < bt.kdone[ch]) return;
    }
  }
  const int n = dv.n, nn = n * n, C = dv.C, Nh = dv.Nh, c = threadIdx.x / LN, tid = threadIdx.x;
  const size_t chan = (size_t)b * bt.L + l, pair = chan * bt.nh + k;
  const dd* H = bt.H + chan * C * nn;
  const dd* M = dv.M0_loc;
  dd* scratch = use_smem ? dd_dyn_smem : bt.scratch + pair * C * dd_npk(n - 2);
  dd* x = bt.U + pair * Nh;
  dd* rhs = bt.vtmp + pair * Nh;
  const int* l2g = dv.l2g;

  long long cyc[6] = {0, 0, 0, 0, 0, 0};
  const long long t_begin = clock64();
#define DD_TIME(slot, ...) do { const long long t0_ = clock64(); __VA_ARGS__; cyc[slot] += clock64() - t0_; } while (0)
  const double Zb = bt.Z[b];
  dd rho = dd_make(0.0);
  dd v[DD_MAXN];
  int fail = 1;
  double lo = 0.0, hi = 0.0, mid = 0.0;
  bool have_bracket = false;
  // ---- warm path: once the SCF has settled (last stop < 0.3), Rayleigh-quotient iteration in dd straight from
  // the previous eigenvector; the result is accepted only if two Float64 inertia counts certify it as level k
  const bool try_warm = bt.warm[b] && fabs(bt.rec[(size_t)b * 7].hi) < 0.3;
  if (tid == 0) bt.eig_code[pair] = 0;
  if (try_warm) {
    dd pm = dd_make(0.0), ph = dd_make(0.0);
    if (c < C) {
      dd_load_local(x, l2g, c, C, n, v);
      pm = dd_quad_local<LN>(M + (size_t)c * nn, v, n, C);
      ph = dd_quad_local<LN>(H + (size_t)c * nn, v, n, C);
    }
    const dd nrm0 = dd_block_sum<LN>(pm, S, C);
    rho = dd_block_sum<LN>(ph, S, C) / nrm0;
    // (a level that was outside the eigenpair window so far has no previous vector: straight to the cold path)
    const int warm_its = (nrm0.hi > 0.0 && isfinite(rho.hi)) ? 6 : 0;
    for (int it = 0; it < warm_its; ++it) {
      DD_TIME(0, dd_factor_shift<dd, LN>(H, M, rho, scratch, S, C, n)); cyc[5] += 1;
      DD_TIME(2, dd_matvec<LN>(M, S, x, rhs, l2g, C, n));
      DD_TIME(1, dd_solve<LN>(scratch, S, rhs, x, l2g, C, n));
      pm = dd_make(0.0); ph = dd_make(0.0);
      if (c < C) {
        dd_load_local(x, l2g, c, C, n, v);
        pm = dd_quad_local<LN>(M + (size_t)c * nn, v, n, C);
        ph = dd_quad_local<LN>(H + (size_t)c * nn, v, n, C);
      }
      const dd nrm = dd_block_sum<LN>(pm, S, C);
      const dd hq = dd_block_sum<LN>(ph, S, C);
      const dd sc = dd_inv(dd_sqrt(nrm));
      for (int i = tid; i < Nh; i += LN * DD_NT) x[i] = x[i] * sc;
      __syncthreads();
      const dd rnew = hq / nrm;
      const double delta = fabs((rnew - rho).hi);
      rho = rnew;
      if (delta <= 1e-25 * fmax(1.0, fabs(rho.hi))) { fail = 0; break; }
    }
    if (!fail) {
      const double m = 1e-9 * fmax(1.0, fabs(rho.hi));
      lo = rho.hi - m; hi = rho.hi + m; mid = rho.hi;
      const int cl = dd_factor_shift<double, LN>(H, M, lo, scratch, S, C, n);
      const int ch = dd_factor_shift<double, LN>(H, M, hi, scratch, S, C, n);
      if (cl != k || ch != k + 1) {
        fail = 1;
        // a neighbour sits inside the window (e.g. a localised resonance crossing a box state within 1e-9): the
        // window still brackets level k, bisect from there
        have_bracket = (cl <= k && ch >= k + 1);
        if (tid == 0) bt.eig_code[pair] = 2 | (cl << 8) | (ch << 16);
      }
    } else if (tid == 0) bt.eig_code[pair] = 1;
  }
  if (fail) {
  if (tid == 0) atomicAdd(&bt.eig_cold[b], 1);
  // ---- phase A: isolate eigenvalue number k with Float64 inertia counts
  if (!have_bracket) {
    lo = -0.6 * Zb * Zb - 1.0; hi = 0.5;
    for (int it = 0; it < 12; ++it) {
      if (dd_factor_shift<double, LN>(H, M, lo, scratch, S, C, n) <= k) break;
      lo = 2.0 * lo;
    }
    for (int it = 0; it < 80; ++it) {
      if (dd_factor_shift<double, LN>(H, M, hi, scratch, S, C, n) >= k + 1) break;
      lo = hi;
      hi = 2.0 * hi + 1.0;
    }
  }
  for (int it = 0; it < 120; ++it) {
    const double md = 0.5 * (lo + hi);
    if (hi - lo <= 1e-10 * fmax(1.0, fabs(md)) || md <= lo || md >= hi) break;
    if (dd_factor_shift<double, LN>(H, M, md, scratch, S, C, n) >= k + 1) hi = md; else lo = md;
  }
  mid = 0.5 * (lo + hi);

  // ---- phase B: inverse iteration at the bracket, then Rayleigh-quotient iteration, in dd
  // (a failed warm attempt may have left x on a neighbouring level: always restart from the hashed vector)
  {
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
  DD_TIME(0, dd_factor_shift<dd, LN>(H, M, dd_make(mid), scratch, S, C, n)); cyc[5] += 1;
  rho = dd_make(mid);
  for (int rep = 0; rep < 3; ++rep) {
    DD_TIME(2, dd_matvec<LN>(M, S, x, rhs, l2g, C, n));
    DD_TIME(1, dd_solve<LN>(scratch, S, rhs, x, l2g, C, n));
    dd part = dd_make(0.0);
    if (c < C) { dd_load_local(x, l2g, c, C, n, v); part = dd_quad_local<LN>(M + (size_t)c * nn, v, n, C); }
    const dd nrm = dd_block_sum<LN>(part, S, C);
    const dd sc = dd_inv(dd_sqrt(nrm));
    for (int i = tid; i < Nh; i += LN * DD_NT) x[i] = x[i] * sc;
    __syncthreads();
  }
  {
    dd part = dd_make(0.0);
    if (c < C) { dd_load_local(x, l2g, c, C, n, v); part = dd_quad_local<LN>(H + (size_t)c * nn, v, n, C); }
    rho = dd_block_sum<LN>(part, S, C);
  }
  fail = 1;
  for (int it = 0; it < 8; ++it) {
    DD_TIME(0, dd_factor_shift<dd, LN>(H, M, rho, scratch, S, C, n)); cyc[5] += 1;
    DD_TIME(2, dd_matvec<LN>(M, S, x, rhs, l2g, C, n));
    DD_TIME(1, dd_solve<LN>(scratch, S, rhs, x, l2g, C, n));
    dd pm = dd_make(0.0), ph = dd_make(0.0);
    if (c < C) {
      dd_load_local(x, l2g, c, C, n, v);
      pm = dd_quad_local<LN>(M + (size_t)c * nn, v, n, C);
      ph = dd_quad_local<LN>(H + (size_t)c * nn, v, n, C);
    }
    const dd nrm = dd_block_sum<LN>(pm, S, C);
    const dd hq = dd_block_sum<LN>(ph, S, C);
    const dd sc = dd_inv(dd_sqrt(nrm));
    for (int i = tid; i < Nh; i += LN * DD_NT) x[i] = x[i] * sc;
    __syncthreads();
    const dd rnew = hq / nrm;
    const double delta = fabs((rnew - rho).hi);
    rho = rnew;
    if (delta <= 1e-25 * fmax(1.0, fabs(rho.hi))) { fail = 0; break; }
  }
  }
  // the pair must be the one the bracket isolated
  if (!(rho.hi >= lo - 1e-6 * fmax(1.0, fabs(mid)) && rho.hi <= hi + 1e-6 * fmax(1.0, fabs(mid)))) fail = 1;
  // sign convention: positive weight on the first bubble of the first cell with a non-zero entry
  // (irrelevant for D; keeps warm starts and outputs reproducible)
  dd pk = dd_make(0.0), pc = dd_make(0.0);
  if (c < C) {
    dd_load_local(x, l2g, c, C, n, v);
    // Kin_l = (A + l(l+1) Mm2)/2, Coulomb = -Z Mm1 (assemble.jl:11,25)
    const dd qa = dd_quad_local<LN>(dv.A_loc + (size_t)c * nn, v, n, C);
    const dd q2 = (l > 0) ? dd_quad_local<LN>(dv.Mm2_loc + (size_t)c * nn, v, n, C) : dd_make(0.0);
    const dd q1 = dd_quad_local<LN>(dv.Mm1_loc + (size_t)c * nn, v, n, C);
    pk = (qa + q2 * (double)(l * (l + 1))) * 0.5;
    pc = q1 * (-Zb);
  }
  const dd ekin = dd_block_sum<LN>(pk, S, C);
  const dd ecou = dd_block_sum<LN>(pc, S, C);
  if (tid == 0) {
    bt.eps[pair] = rho;
    bt.ekin_orb[pair] = ekin;
    bt.ecou_orb[pair] = ecou;
    if (fail) atomicAdd(&bt.eig_fail[b], 1);
    cyc[4] = clock64() - t_begin;
    for (int q = 0; q < 6; ++q) bt.eig_cyc[pair * 6 + q] = cyc[q];
  }
}

// ------------------------------------------------------------------ aufbau (one thread per problem)
__global__ void ddk_aufbau(DDDiscDev dv, DDBatchDev bt, int pass) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bt.B || bt.status[b] != 0) return;
  if (pass == 1 && !bt.redo[b]) return;
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
  if (bt.aufbau_kind != 0) {
    // SmearedAufbau (aufbau/smeared.jl:68-101): n_i = g_i / (1 + exp((eps_i - mu)/T)), mu by 100 bisection steps on
    // [min eps - 50 T, max eps + 50 T];  FrozenAufbau (aufbau/frozen.jl:84-88): n = nfix.  Both run without the
    // eigenpair window (every level is solved in every step).
    if (bt.aufbau_kind == 1) {
      const dd T = dd_make(bt.temp), Np = dd_make(bt.N[b]);
      dd lo = ef[order[0]] - T * 50.0, hi = ef[order[ne - 1]] + T * 50.0;
      for (int itb = 0; itb < 100; ++itb) {
        const dd mid = (lo + hi) * 0.5;
        dd occ = dd_make(0.0);
        for (int idx = 0; idx < ne; ++idx) {
          const dd xa = (ef[idx] - mid) / T;
          if (xa.hi < 700.0) occ = occ + dd_make((double)(4 * (idx % Lp) + 2)) / (dd_exp(xa) + 1.0);
        }
        if (occ < Np) lo = mid; else hi = mid;
      }
      const dd mu = (lo + hi) * 0.5;
      for (int idx = 0; idx < ne; ++idx) {
        const dd xa = (ef[idx] - mu) / T;
        nf0[idx] = (xa.hi < 700.0) ? dd_make((double)(4 * (idx % Lp) + 2)) / (dd_exp(xa) + 1.0) : dd_make(0.0);
        nf1[idx] = nf0[idx];
      }
    } else {
      for (int idx = 0; idx < ne; ++idx) {
        nf0[idx] = dd_make(bt.nfix[((size_t)b * bt.L + idx % Lp) * bt.nh + idx / Lp]);
        nf1[idx] = nf0[idx];
      }
    }
    int cnt = 0;
    for (int idx = 0; idx < ne; ++idx)
      if (nf0[idx].hi != 0.0 && cnt < DD_MAXORB) {
        bt.orb_lk[(size_t)b * DD_MAXORB + cnt] = ((idx % Lp) << 8) | (idx / Lp);
        bt.orb_n0[(size_t)b * DD_MAXORB + cnt] = nf0[idx];
        bt.orb_n1[(size_t)b * DD_MAXORB + cnt] = nf0[idx];
        ++cnt;
      }
    bt.norb[b] = cnt;
    bt.degen[b] = 0;
    bt.redo[b] = 0;
    for (int l = 0; l < Lp; ++l) bt.kdone[b * bt.L + l] = nhp;
    if (bt.eig_fail[b]) { bt.status[b] = AKS_ENAN; bt.rec64[b].status = AKS_ENAN; }
    return;
  }
  dd left = dd_make(bt.N[b]);
  int i = 0, degen = 0, err = 0;
  const dd tol = dd_make(bt.degen_tol);
  dd e_last = dd_make(-1e300);       // head of the last block the fill examined
  int top_ex[8];                     // highest level per channel the fill examined (L <= 8 channels tracked)
  for (int l = 0; l < 8; ++l) top_ex[l] = -1;
  bool touched_missing = false;
  while (left > dd_make(0.0) && i < ne) {
    int blk[8];
    int len = 1;
    blk[0] = order[i];
    int j = i + 1;
    while (j < ne && dd_abs(ef[order[j]] - ef[blk[0]]) < tol && len < bt.max_degen && len < 8) { blk[len++] = order[j]; ++j; }
    i += len;
    e_last = ef[blk[0]];
    dd tot = dd_make(0.0);
    for (int q = 0; q < len; ++q) {
      tot = tot + (double)(4 * (blk[q] % Lp) + 2);
      if (ef[blk[q]].hi > 0.5 * DD_EPS_MISSING) touched_missing = true;
      const int lq = blk[q] % Lp, kq = blk[q] / Lp;
      if (lq < 8 && kq > top_ex[lq]) top_ex[lq] = kq;
    }
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
  // ---- certificate of the eigenpair window: the levels of a channel are ordered, so if the highest solved level
  // of every truncated channel lies above (last block head + tol), no unsolved level could have been occupied or
  // have joined a degenerate block -- the fill is the one of the all-nh solve.  Otherwise: solve the rest (pass 1).
  if (pass == 0 && bt.eig_window) {
    bool unsafe = touched_missing;
    for (int l = 0; l < Lp; ++l) {
      const int ka = bt.kact[b * bt.L + l];
      bt.kdone[b * bt.L + l] = ka;
      if (ka < nhp) {
        const dd top = bt.eps[((size_t)b * bt.L + l) * bt.nh + ka - 1];
        if (!(top > e_last + tol)) unsafe = true;
      }
    }
    if (unsafe) {
      bt.redo[b] = 1;
      for (int l = 0; l < Lp; ++l) bt.kact[b * bt.L + l] = nhp;
      return;
    }
    bt.redo[b] = 0;
  } else {
    for (int l = 0; l < Lp; ++l) bt.kdone[b * bt.L + l] = nhp;
    bt.redo[b] = 0;
  }
  if (bt.eig_window)
    // window of the next step: every level of the channel up to (last block head + tol), one level above it (the
    // certificate needs it) and one spare; shrinks by at most one level per step
    for (int l = 0; l < Lp; ++l) {
      const int solved = bt.kdone[b * bt.L + l];
      int cnt = 0;
      for (int k = 0; k < solved; ++k)
        if (!(bt.eps[((size_t)b * bt.L + l) * bt.nh + k] > e_last + tol)) cnt = k + 1;
      if (l < 8 && top_ex[l] + 1 > cnt) cnt = top_ex[l] + 1;
      bt.kact[b * bt.L + l] = min(nhp, max(cnt + 2, solved - 1));
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
// every level is current / the next step solves every level (reset, after the completion pass of aks_get_state)
__global__ void ddk_window_set(DDBatchDev bt, int set_kact, int set_kdone) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bt.B * bt.L) return;
  const int nhp = bt.nhp[i / bt.L];
  if (set_kact) bt.kact[i] = nhp;
  if (set_kdone) bt.kdone[i] = set_kdone > 0 ? nhp : 0;
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
  dd_solve<1>(dv.Afac, S, bo, wo, dv.l2g, C, n);
}

// ------------------------------------------------------------------ occupations, energies, ODA parameter
// Optim.jl's Brent() as called from line_search.jl:59-67, in dd; one thread drives the
// state machine, the CTA evaluates the objective between steps
struct DDBrent {
  dd lo, hi, x, fx, ox, ofx, oox, oofx, step, old_step, abs_tol, rel_tol, next;
  int it, maxit;
};
__device__ inline dd dd_golden() { return (dd_make(3.0) - dd_sqrt(dd_make(5.0))) * 0.5; }
__device__ inline void ddb_init(DDBrent& s, dd lo, dd hi, dd abs_tol, dd rel_tol, int maxit) {
  s.lo = lo; s.hi = hi; s.abs_tol = abs_tol; s.rel_tol = rel_tol; s.maxit = maxit; s.it = 0;
  s.step = dd_make(0.0); s.old_step = dd_make(0.0);
  s.next = lo + dd_golden() * (hi - lo);
}
__device__ inline void ddb_first(DDBrent& s, dd f) {
  s.x = s.ox = s.oox = s.next;
  s.fx = s.ofx = s.oofx = f;
}
__device__ inline bool ddb_propose(DDBrent& s) {
  if (s.it >= s.maxit) return false;
  const dd zero = dd_make(0.0);
  dd p = zero, q = zero;
  const dd x_tol = s.rel_tol * dd_abs(s.x) + s.abs_tol;
  const dd mid = (s.hi + s.lo) * 0.5;
  if (dd_abs(s.x - mid) <= x_tol * 2.0 - (s.hi - s.lo) * 0.5) return false;
  s.it += 1;
  if (dd_abs(s.old_step) > x_tol) {
    const dd r = (s.x - s.ox) * (s.fx - s.oofx);
    q = (s.x - s.oox) * (s.fx - s.ofx);
    p = (s.x - s.oox) * q - (s.x - s.ox) * r;
    q = (q - r) * 2.0;
    if (q > zero) p = -p; else q = -q;
  }
  if (dd_abs(p) < dd_abs(q * s.old_step * 0.5) && p < q * (s.hi - s.x) && p < q * (s.x - s.lo)) {
    s.old_step = s.step;
    s.step = p / q;
    const dd xt = s.x + s.step;
    if ((xt - s.lo) < x_tol * 2.0 || (s.hi - xt) < x_tol * 2.0) s.step = (s.x < mid) ? x_tol : -x_tol;
  } else {
    s.old_step = (s.x < mid) ? s.hi - s.x : s.lo - s.x;
    s.step = dd_golden() * s.old_step;
  }
  if (dd_abs(s.step) >= x_tol) s.next = s.x + s.step;
  else s.next = s.x + ((s.step > zero) ? x_tol : -x_tol);
  return true;
}
__device__ inline void ddb_update(DDBrent& s, dd nf) {
  const dd nx = s.next;
  if (nf < s.fx) {
    if (nx < s.x) s.hi = s.x; else s.lo = s.x;
    s.oox = s.ox; s.oofx = s.ofx;
    s.ox = s.x; s.ofx = s.fx;
    s.x = nx; s.fx = nf;
  } else {
    if (nx < s.x) s.lo = nx; else s.hi = nx;
    if (nf <= s.ofx || s.ox == s.x) {
      s.oox = s.ox; s.oofx = s.ofx;
      s.ox = nx; s.ofx = nf;
    } else if (nf <= s.oofx || s.oox == s.x || s.oox == s.ox) {
      s.oox = nx; s.oofx = nf;
    }
  }
}

#define DD_DECIDE_NT 256   // threads per CTA of ddk_decide
#define DD_DECIDE_CL 4     // CTAs (one thread-block cluster) per problem: the grid evaluations of Exc are split
                           // over the cluster and summed through distributed shared memory
struct DDDecideShared {
  dd part;          // this CTA's share of the last grid sum (read by the whole cluster)
  dd g[DD_MAXINV * DD_MAXINV];   // Gram matrix b_a . w_b of the involved orbitals
  dd hp[DD_MAXINV];               // b_a . W_prev
  dd nfin[DD_MAXINV];
  dd red[DD_DECIDE_NT];
  DDBrent br;
  dd t, f;
  int go;
};

// Exc on the global grid for rho = a new0 + bq new1 + cq prev (compute_exc_energy, energies.jl:217-233).  The
// points are dealt out to the CTAs of the cluster; every CTA adds the partial sums of all ranks in the same fixed
// order, so the whole cluster holds the same value (the Brent state machines of the CTAs stay in lockstep).
__device__ dd dd_exc_grid(const DDDiscDev& dv, const DDBatchDev& bt, int b, dd a, dd bq, dd cq, bool use1,
                          DDDecideShared& S) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int R = (int)cl.num_blocks(), r = (int)cl.block_rank();
  const int G = dv.G, ex = bt.ex[b], ec = bt.ec[b];
  const dd* n0 = bt.rho_y_new + ((size_t)b * 2 + 0) * G;
  const dd* n1 = bt.rho_y_new + ((size_t)b * 2 + 1) * G;
  const dd* pr = bt.rho_y + (size_t)b * G;
  dd acc = dd_make(0.0);
  for (int q = r * (int)blockDim.x + (int)threadIdx.x; q < G; q += R * (int)blockDim.x) {
    dd rr = a * n0[q] + cq * pr[q];
    if (use1) rr = rr + bq * n1[q];
    const dd val = rr * dd_xc_zk(dd_max(rr, dd_make(0.0)), ex, ec);
    acc = acc + dv.wy[q] * (val * (dv.y[q] * dv.y[q]));
  }
  __syncthreads();
  S.red[threadIdx.x] = acc;
  __syncthreads();
  for (int h = DD_DECIDE_NT / 2; h > 0; h >>= 1) {   // fixed tree: bit reproducible
    if ((int)threadIdx.x < h) S.red[threadIdx.x] = S.red[threadIdx.x] + S.red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) S.part = S.red[0];
  cl.sync();
  dd sum = dd_make(0.0);
  for (int rr = 0; rr < R; ++rr) sum = sum + *cl.map_shared_rank(&S.part, rr);
  cl.sync();   // nobody overwrites its share before every rank has read it
  return sum * DD_FOURPI;
}

// minimise a t^2 + b t + c + F(t) on [0, 1] (line_search_energy, line_search.jl:45-80).  F(t) = Exc[t rho_A +
// (1-t) rho_B] with (A, B) = (filling 0, filling 1) [mode 0] or (trial, previous) [mode 1]; without a
// functional the quadratic is minimised in closed form.  Called by every thread of the CTA.
__device__ void dd_line_search(const DDDiscDev& dv, const DDBatchDev& bt, int b, dd qa, dd qb, dd qc, bool has_xc,
                               int mode, dd td, bool degen, double abstol, double reltol, int maxit,
                               DDDecideShared& S, dd& t_out, dd& f_out) {
  const dd zero = dd_make(0.0), one = dd_make(1.0);
  if (!has_xc) {
    dd t;
    if (qa > zero) t = dd_min(dd_max(-qb / (qa * 2.0), zero), one);
    else t = (qc <= qc + qb + qa) ? zero : one;
    t_out = t;
    f_out = qc + qb * t + qa * t * t;
    return;
  }
  if (threadIdx.x == 0) {
    ddb_init(S.br, zero, one, dd_make(abstol), dd_make(reltol), maxit);
    S.t = S.br.next;
    S.go = 1;
  }
  __syncthreads();
  bool first = true;
  while (true) {
    const dd t = S.t;
    dd F;
    if (mode == 0) F = dd_exc_grid(dv, bt, b, t, one - t, zero, true, S);
    else F = dd_exc_grid(dv, bt, b, t * td, t * (one - td), one - t, degen, S);
    if (threadIdx.x == 0) {
      const dd fv = qa * (t * t) + qb * t + qc + F;
      if (first) ddb_first(S.br, fv); else ddb_update(S.br, fv);
      S.go = ddb_propose(S.br) ? 1 : 0;
      S.t = S.br.next;
      if (!S.go) { S.t = S.br.x; S.f = S.br.fx; }
    }
    first = false;
    __syncthreads();
    const int go = S.go;
    __syncthreads();
    if (!go) break;
  }
  t_out = S.t;
  f_out = S.f;
  __syncthreads();
}

// The scalar logic is executed redundantly by every thread (identical results), so that the CTA-wide
// evaluations of Exc sit in uniform control flow; thread 0 writes.
__global__ void __cluster_dims__(DD_DECIDE_CL, 1, 1) __launch_bounds__(DD_DECIDE_NT) ddk_decide(DDDiscDev dv, DDBatchDev bt) {
  const int b = blockIdx.y, Nh = dv.Nh, tid = threadIdx.x;
  const bool writer = (blockIdx.x == 0);   // every CTA of the cluster computes the same numbers, rank 0 stores them
  if (bt.status[b] != 0) return;
  __shared__ DDDecideShared S;
  const int no = bt.norb[b];
  const dd* borb = bt.borb + (size_t)b * DD_MAXORB * Nh;
  const dd* worb = bt.worb + (size_t)b * DD_MAXORB * Nh;
  const dd* Wp = bt.Wv + (size_t)b * Nh;
  const bool har_on = bt.hartree[b] != 0.0;
  const bool has_xc = bt.ex[b] != 0 || bt.ec[b] != 0;
  const bool degen = bt.degen[b] != 0;
  const dd zero = dd_make(0.0), one = dd_make(1.0);
  if (no > DD_MAXINV) { if (tid == 0) { bt.status[b] = AKS_EINVAL; bt.rec64[b].status = AKS_EINVAL; } return; }
  for (int pq = tid; pq < no * no + no; pq += blockDim.x) {
    dd acc = zero;
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
    if (pq < no * no) S.g[pq] = acc; else S.hp[pq - no * no] = acc;
  }
  __syncthreads();
  const dd* n0 = bt.orb_n0 + (size_t)b * DD_MAXORB;
  const dd* n1 = bt.orb_n1 + (size_t)b * DD_MAXORB;
  const int* lk = bt.orb_lk + (size_t)b * DD_MAXORB;
  const dd Nb = dd_make(bt.N[b]);
  const dd self = har_on ? (Nb * Nb) / dv.Rmax : zero;   // N^2 / Rmax
  dd kin[2], cou[2], har[2], seps[2], corr[2];
  for (int s = 0; s < 2; ++s) {
    const dd* ns = s ? n1 : n0;
    kin[s] = cou[s] = seps[s] = corr[s] = zero;
    dd q = zero;
    for (int a = 0; a < no; ++a) {
      const size_t pair = ((size_t)b * bt.L + (lk[a] >> 8)) * bt.nh + (lk[a] & 0xff);
      kin[s] = kin[s] + ns[a] * bt.ekin_orb[pair];
      cou[s] = cou[s] + ns[a] * bt.ecou_orb[pair];
      seps[s] = seps[s] + ns[a] * bt.eps[pair];
      dd r = zero;
      for (int c2 = 0; c2 < no; ++c2) r = r + S.g[a * no + c2] * ns[c2];
      q = q + r * ns[a];
    }
    har[s] = har_on ? (q + self) * 0.5 : zero;
    if (has_xc)
      for (int c = 0; c < dv.C; ++c) corr[s] = corr[s] + bt.corr_part[((size_t)b * 2 + s) * dv.C + c];
  }
  dd e[5];   // Etot, Ekin, Ecou, Ehar, Eexc of the trial density
  dd td = one;
  if (degen) {
    // resolve_twofold_degeneracy!: line search between the two extreme fillings
    dd q01 = zero;
    for (int a = 0; a < no; ++a) {
      dd r = zero;
      for (int c2 = 0; c2 < no; ++c2) r = r + S.g[a * no + c2] * n0[c2];   // B[D2] . W[D1]
      q01 = q01 + r * n1[a];
    }
    const dd har01 = har_on ? (q01 + self) * 0.5 : zero;
    const dd A0 = kin[0] + cou[0], A1 = kin[1] + cou[1];
    const dd qa = har[0] + har[1] - har01 * 2.0, qb = (A0 - A1) + (har01 - har[1]) * 2.0, qc = A1 + har[1];
    dd emin;
    // the degenerate search runs with the default tolerances of line_search_energy, 1e4 eps(Double64)
    dd_line_search(dv, bt, b, qa, qb, qc, has_xc, 0, one, true, 4.930380657631324e-28, 4.930380657631324e-28, 100, S,
                   td, emin);
    const dd um = one - td;
    if (tid < no) S.nfin[tid] = td * n0[tid] + um * n1[tid];
    e[0] = emin;
    e[1] = td * kin[0] + um * kin[1];
    e[2] = td * cou[0] + um * cou[1];
    e[3] = td * td * har[0] + um * um * har[1] + td * um * har01 * 2.0;
    e[4] = e[0] - e[1] - e[2] - e[3];
  } else {
    if (tid < no) S.nfin[tid] = n0[tid];
    e[1] = kin[0]; e[2] = cou[0]; e[3] = har[0];
    e[4] = has_xc ? dd_exc_grid(dv, bt, b, one, zero, zero, false, S) : zero;
    // compute_total_energy (energies.jl:7-39): sum n eps - Ehar [+ Exc - <Vxc[D_prev], D>]
    e[0] = has_xc ? (seps[0] - har[0] + e[4] - corr[0]) : (seps[0] - har[0]);
  }
  __syncthreads();
  // ODA step (oda/loop.jl:79-116)
  dd t = bt.t[b], tmix = one;
  dd en[5];
  for (int q = 0; q < 5; ++q) en[q] = e[q];
  if (bt.niter[b] > 0) {
    const dd* ep = bt.E + (size_t)b * 5;
    dd q01 = zero;
    for (int a = 0; a < no; ++a) q01 = q01 + S.hp[a] * S.nfin[a];
    const dd har01 = har_on ? (q01 + self) * 0.5 : zero;
    dd etot = zero;
    if (bt.frozen_t) {
      const dd um = one - t;
      en[1] = t * e[1] + um * ep[1];
      en[2] = t * e[2] + um * ep[2];
      en[3] = t * t * e[3] + um * um * ep[3] + t * um * har01 * 2.0;
      en[4] = has_xc ? dd_exc_grid(dv, bt, b, t * td, t * (one - td), um, degen, S) : zero;
      en[0] = en[1] + en[2] + en[3] + en[4];
    } else {
      const dd A0 = e[1] + e[2], A1 = ep[1] + ep[2];
      const dd qa = e[3] + ep[3] - har01 * 2.0, qb = (A0 - A1) + (har01 - ep[3]) * 2.0, qc = A1 + ep[3];
      dd_line_search(dv, bt, b, qa, qb, qc, has_xc, 1, td, degen, bt.abstol_ls, bt.reltol_ls, bt.maxiter_ls, S, t, etot);
      const dd um = one - t;
      en[1] = t * e[1] + um * ep[1];
      en[2] = t * e[2] + um * ep[2];
      en[3] = t * t * e[3] + um * um * ep[3] + t * um * har01 * 2.0;
      en[0] = etot;
      en[4] = etot - en[1] - en[2] - en[3];
    }
    tmix = t;
  }
  if (tid == 0 && writer) {
    for (int q = 0; q < 5; ++q) bt.Enew[(size_t)b * 5 + q] = en[q];
    bt.tnew[b] = t;
    bt.tmix[b] = tmix;
    bt.td[b] = td;
    // occupations of the trial density, (l, k) layout
    for (int q = 0; q < bt.L * bt.nh; ++q) bt.occ[(size_t)b * bt.L * bt.nh + q] = zero;
    for (int a = 0; a < no; ++a) {
      bt.orb_n[(size_t)b * DD_MAXORB + a] = S.nfin[a];
      bt.occ[((size_t)b * bt.L + (lk[a] >> 8)) * bt.nh + (lk[a] & 0xff)] = S.nfin[a];
    }
  }
  // B, W of the trial density are linear in the occupations
  for (int i = blockIdx.x * blockDim.x + tid; i < Nh; i += gridDim.x * blockDim.x) {
    dd bs = zero, ws = zero;
    for (int a = 0; a < no; ++a) {
      bs = bs + borb[(size_t)a * Nh + i] * S.nfin[a];
      ws = ws + worb[(size_t)a * Nh + i] * S.nfin[a];
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
  if (bt.ex[b] != 0 || bt.ec[b] != 0) {
    // the linear images of D the next step reads: rho at the Gauss nodes and on the energy grid
    const dd td = bt.td[b], ud = dd_make(1.0) - td;
    const bool degen = bt.degen[b] != 0;
    const size_t nq = (size_t)dv.C * dv.Q, ny = dv.G;
    for (size_t i = tid; i < nq + ny; i += blockDim.x) {
      dd *dst;
      const dd *a0, *a1;
      if (i < nq) {
        dst = bt.rho_q + (size_t)b * nq + i;
        a0 = bt.rho_q_new + ((size_t)b * 2 + 0) * nq + i; a1 = bt.rho_q_new + ((size_t)b * 2 + 1) * nq + i;
      } else {
        dst = bt.rho_y + (size_t)b * ny + (i - nq);
        a0 = bt.rho_y_new + ((size_t)b * 2 + 0) * ny + (i - nq); a1 = bt.rho_y_new + ((size_t)b * 2 + 1) * ny + (i - nq);
      }
      const dd trial = degen ? (td * (*a0) + ud * (*a1)) : *a0;
      *dst = t * trial + um * (*dst);
    }
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
