// k_reduce + k_eig: lowest-nh generalized eigenpairs of every (problem, l, sigma) channel.
//
// Replaces find_orbital! (src/algorithms/oda/loop.jl:65-77): the reference forms the dense
// S*H*S with S = sqrt(inv(M0)) and calls LAPACK syevr, O(Nh^3) per channel.  Here the pencil
// stays UNASSEMBLED, (H_c, M_c) per cell c, hats (2) + bubbles (nb = p-1):
//
//   k_reduce  (once per channel and SCF step, one warp per cell): the bubble block pencil is
//     reduced to standard tridiagonal form, exactly the classical two-stage reduction of the
//     north star but per cell:  M_bb = L L^T (L^-1 prefactored per cell at setup),
//     C = L^-1 H_bb L^-T,  C = Q T Q^T by Householder,  P = Q^T L^-1.  With x_b = P^T x~:
//         P (H_bb - s M_bb) P^T = T - s I,     P (H_bh - s M_bh) = h~ - s m~ .
//   k_eig  (one CTA per eigenpair): for any shift s the statically condensed operator costs
//     O(nb) per cell instead of O(nb^3): one THREAD per cell runs the tridiagonal LDL^T of T - sI,
//     W = (T - sI)^-1 (h~ - s m~) and the 2x2 Schur complement; the hats of all cells form a
//     (C-1) tridiagonal system.  By Sylvester/Haynsworth #{eig < s} = sum_c neg(T_c - sI) +
//     neg(hat Schur complement): every factorisation is also a Sturm count.
//       warm path: Rayleigh-quotient iteration from the previous SCF eigenvector; the counts of
//         the visited shifts are kept as witnesses and the converged pair is CERTIFIED as
//         eigenvalue number k (count == k just below, == k+1 just above);
//       cold path / failed certificate: bracket, bisect until isolated and narrow, same RQI.
//     The whole iteration runs in the reduced coordinates (hats, x~): M~ = [M_hh m~^T; m~ I] and
//     H~ = [H_hh h~^T; h~ T] act in O(nb) per cell too, so a factorisation, a solve and a mass
//     product are all thread-per-cell sweeps over shared memory.  Dense (nb x nb) products with P
//     appear only twice per eigenpair: x~ = P M_bb x_b for the warm start, x_b = P^T x~ at the end.
//     The returned eigenvalue is the Rayleigh quotient with the ORIGINAL blocks (second-order
//     accurate in the vector), so the O(eps*cond(M_bb)) error of the reduction does not reach it;
//     eigenvectors carry that error divided by the relative gap (~1e-13).
//
// Eigenvalues are accurate to a few ulp (the dense LAPACK route loses ~eps*||M0^-1 H||, 4e-9
// absolute at Nh = 799).
#pragma once
#include <float.h>

#include "aks_device.cuh"
#include "aks_types.cuh"

// per-cell reduced data of a channel, field-major ([field][cell]: consecutive cells are consecutive addresses)
//   td[nb] te[nb] ht0[nb] ht1[nb] mt0[nb] mt1[nb] Hhh[3] pad[1]      (te[nb-1] unused)
__host__ __device__ inline int red_stride(int nb) { return 6 * nb + 4; }

// ------------------------------------------------------------------------------------------
// k_reduce: one warp per (problem, channel, cell); lane i owns ROW i of the symmetric block C and
// COLUMN i of P, both in registers (NBM = nb rounded up, zero padded; every register index is a
// compile-time constant).  Shared memory holds only what all lanes read at the same address
// (H_bb, L^-1, the Householder vectors v and w: broadcast 128-bit loads) and the final P.
#define RED_NT 128
#define RED_NW (RED_NT / 32)
__host__ __device__ inline int reduce_nbm(int nb) { return nb <= 8 ? 8 : nb <= 12 ? 12 : nb <= 20 ? 20 : 28; }
__host__ __device__ inline size_t reduce_smem_bytes(int nb) {
  const int nbm = reduce_nbm(nb);
  return sizeof(double) * (size_t)RED_NW * (3 * nbm * nbm + nbm + 2 * 32);
}

template <int NBM>
__global__ void __launch_bounds__(RED_NT, 4)
k_reduce(DiscDev d, BatchDev b) {
  const int p = blockIdx.z, ch = blockIdx.y;
  if (b.status[p] != 0) return;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * RED_NW + warp;
  const int nb = d.nb, npk = d.npk, C = d.C;
  if (c >= C) return;
  constexpr int LDP = NBM | 1;
  extern __shared__ __align__(16) double sm_red[];
  double* Hs = sm_red + (size_t)warp * (3 * NBM * NBM + NBM + 64);   // [NBM][NBM]  H_bb (zero padded)
  double* Ls = Hs + NBM * NBM;                                       // [NBM][NBM]  L^-1 (lower, zero padded)
  double* Ps = Ls + NBM * NBM;                                       // [NBM][LDP]  final P
  double* vs = Hs + 3 * NBM * NBM + NBM;                             // [32], 16-byte aligned (NBM is even)
  double* ws = vs + 32;
  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  const double* Hc = b.Hpk + (chan * C + c) * npk;
  const double* Mc = d.M0pk + (size_t)c * npk;
  const double* Li = d.Linv + (size_t)c * nb * nb;
  const int off = 3 + 2 * nb;
  const int i = lane;
  // 1. stage H_bb (full symmetric) and L^-1
  for (int idx = lane; idx < NBM * NBM; idx += 32) {
    const int r = idx / NBM, cc = idx - r * NBM;
    const int hi = r >= cc ? r : cc, lo = r >= cc ? cc : r;
    const bool in = (r < nb) && (cc < nb);
    Hs[idx] = in ? Hc[off + hi * (hi + 1) / 2 + lo] : 0.0;
    Ls[idx] = (in && cc <= r) ? Li[r * nb + cc] : 0.0;
  }
  __syncwarp();
  const int ir = (i < NBM) ? i : 0;   // padded lanes compute on row 0 and are masked at the stores
  const bool act = i < nb;
  // 2. A = L^-1 H : row i in registers
  double row[NBM];
#pragma unroll
  for (int j = 0; j < NBM; ++j) row[j] = 0.0;
  {
    double lrow[NBM];
#pragma unroll
    for (int k = 0; k < NBM; ++k) lrow[k] = act ? Ls[ir * NBM + k] : 0.0;
#pragma unroll
    for (int k = 0; k < NBM; ++k) {
      const double a = lrow[k];
      const double2* h2 = reinterpret_cast<const double2*>(Hs + k * NBM);
#pragma unroll
      for (int j = 0; j < NBM / 2; ++j) {
        const double2 h = h2[j];
        row[2 * j] = fma(a, h.x, row[2 * j]);
        row[2 * j + 1] = fma(a, h.y, row[2 * j + 1]);
      }
    }
  }
  // 3. C = A L^-T : C[i][j] = sum_k A[i][k] Linv[j][k]  (row i stays in registers)
  {
    double crow[NBM];
#pragma unroll
    for (int j = 0; j < NBM; ++j) {
      const double2* l2 = reinterpret_cast<const double2*>(Ls + j * NBM);
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int k = 0; k < (j + 2) / 2; ++k) {   // Linv[j][k] = 0 for k > j
        const double2 lv = l2[k];
        a0 = fma(row[2 * k], lv.x, a0);
        a1 = fma(row[2 * k + 1], lv.y, a1);
      }
      crow[j] = a0 + a1;
    }
#pragma unroll
    for (int j = 0; j < NBM; ++j) row[j] = crow[j];
  }
  // symmetrise exactly: C[i][j] for j > i is taken from lane j's C[j][i]
#pragma unroll
  for (int j = 0; j < NBM; ++j) {
    // value C[j][i] lives in lane j as row[i] (runtime index there): exchange through shared memory
    if (i < NBM) Ps[i * LDP + j] = row[j];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NBM; ++j)
    if (j > i) row[j] = Ps[j * LDP + ir];
  __syncwarp();
  // 4. P = L^-1 : column i in registers
  double pcol[NBM];
#pragma unroll
  for (int m = 0; m < NBM; ++m) pcol[m] = act ? Ls[m * NBM + ir] : 0.0;
  double* red = b.red + chan * C * red_stride(nb) + c;   // field-major: entry f of cell c at red[f * C]
  // 5. Householder tridiagonalisation C <- H_k C H_k, P <- H_k P
#pragma unroll
  for (int k = 0; k + 2 < NBM; ++k) {
    if (k + 2 < nb) {
      const double x = (i > k && act) ? row[k] : 0.0;
      const double sig2 = warp_sum((i > k + 1) ? x * x : 0.0);
      const double xk1 = __shfl_sync(0xffffffffu, x, k + 1);
      const double dk = __shfl_sync(0xffffffffu, row[k], k);
      if (lane == 0) red[(k) * C] = dk;
      if (sig2 == 0.0) {  // already tridiagonal in this column (warp-uniform)
        if (lane == 0) red[(nb + k) * C] = xk1;
      } else {
        const double alpha = -copysign(sqrt(xk1 * xk1 + sig2), xk1);
        const double vk1 = xk1 - alpha;
        const double v = (i == k + 1) ? vk1 : ((i > k + 1) ? x : 0.0);
        const double tau = 2.0 / (sig2 + vk1 * vk1);
        if (lane == 0) red[(nb + k) * C] = alpha;
        vs[lane] = v;
        __syncwarp();
        const int j0 = (k + 1) / 2;   // v[j] = 0 for j <= k: start at the even index below k+1
        const double2* v2 = reinterpret_cast<const double2*>(vs);
        double pa = 0.0, pb = 0.0;
#pragma unroll
        for (int j = 0; j < NBM / 2; ++j) {
          if (j < j0) continue;   // folds after unrolling
          const double2 vv = v2[j];
          pa = fma(row[2 * j], vv.x, pa);
          pb = fma(row[2 * j + 1], vv.y, pb);
        }
        const double pacc = (i > k && act) ? (pa + pb) * tau : 0.0;
        const double K = 0.5 * tau * warp_sum(pacc * v);
        const double w = pacc - K * v;
        ws[lane] = w;
        __syncwarp();
        const double2* w2 = reinterpret_cast<const double2*>(ws);
#pragma unroll
        for (int j = 0; j < NBM / 2; ++j) {
          if (j < j0) continue;
          const double2 vv = v2[j], ww = w2[j];
          row[2 * j] -= v * ww.x + w * vv.x;
          row[2 * j + 1] -= v * ww.y + w * vv.y;
        }
        // P <- (I - tau v v^T) P, lane = column
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int m = 0; m < NBM / 2; ++m) {
          if (m < j0) continue;
          const double2 vv = v2[m];
          ta = fma(vv.x, pcol[2 * m], ta);
          tb = fma(vv.y, pcol[2 * m + 1], tb);
        }
        const double t = (ta + tb) * tau;
#pragma unroll
        for (int m = 0; m < NBM / 2; ++m) {
          if (m < j0) continue;
          const double2 vv = v2[m];
          pcol[2 * m] -= vv.x * t;
          pcol[2 * m + 1] -= vv.y * t;
        }
        __syncwarp();
      }
    }
  }
  // trailing 2x2 of T, the hat-hat block
  {
    // row[nb-2], row[nb-1] have runtime indices: go through shared memory once
#pragma unroll
    for (int j = 0; j < NBM; ++j)
      if (i < NBM) Ps[i * LDP + j] = row[j];
    __syncwarp();
    if (lane == 0) {
      if (nb >= 2) {
        red[(nb - 2) * C] = Ps[(nb - 2) * LDP + nb - 2];
        red[(nb + nb - 2) * C] = Ps[(nb - 1) * LDP + nb - 2];
      }
      red[(nb - 1) * C] = Ps[(nb - 1) * LDP + nb - 1];
      red[(nb + nb - 1) * C] = 0.0;
      red[(6 * nb + 0) * C] = Hc[0];
      red[(6 * nb + 1) * C] = Hc[1];
      red[(6 * nb + 2) * C] = Hc[2];
      red[(6 * nb + 3) * C] = 0.0;
    }
    __syncwarp();
  }
  // P to shared memory: Ps[m][i] = pcol[m]
#pragma unroll
  for (int m = 0; m < NBM; ++m)
    if (i < NBM) Ps[m * LDP + i] = pcol[m];
  __syncwarp();
  // 6. h~ = P H_bh, m~ = P M_bh
  {
    const double h0 = act ? Hc[3 + i] : 0.0, h1 = act ? Hc[3 + nb + i] : 0.0;
    const double m0 = act ? Mc[3 + i] : 0.0, m1 = act ? Mc[3 + nb + i] : 0.0;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int j = 0; j < nb; ++j) {
      const double pj = Ps[ir * LDP + j];
      a0 = fma(pj, __shfl_sync(0xffffffffu, h0, j), a0);
      a1 = fma(pj, __shfl_sync(0xffffffffu, h1, j), a1);
      a2 = fma(pj, __shfl_sync(0xffffffffu, m0, j), a2);
      a3 = fma(pj, __shfl_sync(0xffffffffu, m1, j), a3);
    }
    if (act) { red[(2 * nb + i) * C] = a0; red[(3 * nb + i) * C] = a1; red[(4 * nb + i) * C] = a2; red[(5 * nb + i) * C] = a3; }
  }
  // 7. P (row-major) and P^T
  double* Pm = b.Pm + (chan * C + c) * (size_t)nb * nb;
  double* Pt = b.Pt + (chan * C + c) * (size_t)nb * nb;
  for (int idx = lane; idx < nb * nb; idx += 32) {
    const int r = idx / nb, cc = idx - r * nb;
    Pm[idx] = Ps[r * LDP + cc];
    Pt[idx] = Ps[cc * LDP + r];
  }
}

// ------------------------------------------------------------------------------------------
// Vectors in reduced coordinates: hats h[C] (h[j] = hat node j, j < C-1) and bubbles t[nb][C]
// (field-major: t[j*C + c] = j-th transformed bubble coefficient of cell c).
struct RedVec { double* h; double* t; };

struct EigShared {
  const double* __restrict__ red;   // [red_stride][C]  reduced cell data of the channel (global memory, L1-cached)
  double* mhh;      // [3][C]           M hat-hat blocks
  double* fac;      // [4 nb][C]        dinv[nb] | l[nb] | W0[nb] | W1[nb]
  RedVec x, r, y;   // iterate, r = M~ x, solve output
  double* hatpart;  // [2C]
  double* sc;       // [EIG_NG][3C]  Schur blocks of the current factorisation(s)
  double* tri_di;   // [C]   reciprocal pivots of the hat system
  double* tri_l;    // [C]
  double* xh;       // [C]
  double* red8;     // [NW + 8]
  int* negc;        // [EIG_NG][C + 2]
  int* cntg;        // [EIG_NG + 2] inertia counts of the shifts of the last eig_factor call
  int* hatg;        // [C]   global index of hat j (right hat of cell j)
  int* bub0;        // [C]   global index of the first bubble of cell c (bubbles are contiguous)
};

__host__ __device__ inline size_t eig_smem_bytes(int C, int nb, int Nh) {
  size_t nd = 3 * C + (size_t)4 * nb * C + 3 * ((size_t)nb * C + C) +
              2 * C + 6 * 3 * C + 3 * C + AKS_NW + 8;
  return nd * sizeof(double) + (size_t)(6 * (C + 2) + 2 * C + 16) * sizeof(int);
}
// EIG_RED_SMEM = 1: k_eig stages the channel's reduced cell data (red_stride x C doubles, 37.8 KB at order 20) in
// shared memory with one bulk copy.  Measured (profiles/r02_k_eig_experiments.md): a factorisation drops from 20.4 k
// to 14.2 k cycles, but only two CTAs fit per SM instead of four and the kernel as a whole is slower (0.261 vs
// 0.235 ms): the sweeps are bound by the latency of dependent FP64 operations, which only more resident
// eigenpairs per SM (or shorter dependency chains) hide.  Off by default.
#ifndef EIG_PIGGYBACK
#define EIG_PIGGYBACK 0
#endif
#ifndef EIG_RED_SMEM
#define EIG_RED_SMEM 0
#endif
__host__ __device__ inline size_t eig_red_smem_bytes(int C, int nb) {
  return EIG_RED_SMEM ? sizeof(double) * (size_t)(6 * nb + 4) * C : 0;
}

#define EIG_NG 6   // shifts evaluated concurrently in the count-only states (groups of C threads)

// Reciprocal of a pivot on the sequential critical path: hardware seed + two Newton steps (1 ulp), about half
// the dependent FP64 operations of the correctly rounded division.  Pivots are clamped away from zero before.
__device__ __forceinline__ double pivot_rcp(double dval) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(dval));
  double err = fma(-dval, x, 1.0);
  x = fma(x, err, x);
  err = fma(-dval, x, 1.0);
  x = fma(x, err, x);
  return x;
}

// Inertia count of the hat system alone (count-only shifts): one thread runs both sweeps of the twisted
// factorisation interleaved (two independent dependency chains).  Not inlined: it is called from three places of
// eig_factor and would otherwise push k_eig over its register budget.
__device__ __noinline__ int eig_hat_count(const double* sc, const int* ngc, int C) {
  const int m = C - 1, tw = m / 2;

    // count only: one thread per shift runs both sweeps interleaved (two independent dependency chains)
    int neg = 0;
    for (int cc = 0; cc < C; ++cc) neg += ngc[cc];
    double dinvp = 0.0, offprev = 0.0, dinvn = 0.0, offnext = 0.0, scale = 0.0;
    const int nup = m - 1 - tw;  // rows of the bottom sweep
#pragma unroll 2
    for (int q = 0; q < max(tw, nup); ++q) {
      if (q < tw) {
        const int j = q;
        const double diag = sc[3 * j] + sc[3 * (j + 1) + 2];
        scale = fmax(scale, fabs(diag));
        double dj = diag;
        if (j > 0) dj = diag - (offprev * dinvp) * offprev;
        const double pivmin = fmax(scale * 4.9e-32, DBL_MIN * 1e8);
        if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
        neg += (dj < 0.0);
        dinvp = pivot_rcp(dj);
        offprev = sc[3 * (j + 1) + 1];
      }
      if (q < nup) {
        const int j = m - 1 - q;
        const double diag = sc[3 * j] + sc[3 * (j + 1) + 2];
        scale = fmax(scale, fabs(diag));
        double dj = diag;
        if (j < m - 1) dj = diag - (offnext * dinvn) * offnext;
        const double pivmin = fmax(scale * 4.9e-32, DBL_MIN * 1e8);
        if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
        neg += (dj < 0.0);
        dinvn = pivot_rcp(dj);
        offnext = sc[3 * j + 1];
      }
    }
    const double diag = sc[3 * tw] + sc[3 * (tw + 1) + 2];
    double gam = diag;
    if (tw > 0) gam -= offprev * dinvp * offprev;
    if (tw < m - 1) gam -= offnext * dinvn * offnext;
    return neg + (gam < 0.0);
    }

// ---- factor T(sigma) for up to EIG_NG shifts at once: thread t = g*C + c handles cell c of shift g.
// store != 0: keep the factors of shift 0 for eig_solve; further shifts (g >= 1) ride along count-only (the
// certificate's shifts next to a Rayleigh-quotient iteration, on threads that would idle).  Inertia counts -> S.cntg[g].
#ifdef AKS_EIG_TIMING
#define EIG_FT_DECL , long long* ft
#define EIG_FT_ARG , ft_
#define EIG_FT_MARK(i) do { if (ft) { const long long now_ = clock64(); ft[i] += now_ - ft[3]; ft[3] = now_; } } while (0)
#else
#define EIG_FT_DECL
#define EIG_FT_ARG
#define EIG_FT_MARK(i)
#endif
__device__ __forceinline__ void eig_factor(const EigShared& S, int C, int nb, int ng, double sigma, int store EIG_FT_DECL) {
  const int g = threadIdx.x / C, c = threadIdx.x - g * C;
#ifdef AKS_EIG_TIMING
  if (ft) ft[3] = clock64();
#endif
  if (g < ng) {
    const double* R = S.red;
    double* F = S.fac;
    double* sc = S.sc + (size_t)g * 3 * C;
    const double pivmin0 = DBL_MIN * 1e8;
    double scale = 0.0;
    int neg = 0;
    double dprev_inv = 0.0, y0p = 0.0, y1p = 0.0, s00 = 0.0, s10 = 0.0, s11 = 0.0;
    const int nbC = nb * C;
    const double* Rj = R + c;      // field td, advanced by C per step
    double* Fj = F + c;
    // software pipelined: the six table values of step j + 1 are requested before the pivot recurrence of
    // step j (a division and four dependent FMAs) is evaluated
    double n_td = Rj[0], n_te = 0.0, n_h0 = Rj[2 * nbC], n_h1 = Rj[3 * nbC], n_m0 = Rj[4 * nbC], n_m1 = Rj[5 * nbC];
    for (int j = 0; j < nb; ++j, Rj += C, Fj += C) {
      const double c_td = n_td, c_te = n_te, c_h0 = n_h0, c_h1 = n_h1, c_m0 = n_m0, c_m1 = n_m1;
      if (j + 1 < nb) {
        n_td = Rj[C]; n_te = Rj[nbC]; n_h0 = Rj[2 * nbC + C]; n_h1 = Rj[3 * nbC + C];
        n_m0 = Rj[4 * nbC + C]; n_m1 = Rj[5 * nbC + C];
      }
      const double tdj = c_td - sigma;
      scale = fmax(scale, fabs(tdj));
      double lj = 0.0, dj = tdj;
      if (j > 0) {
        const double te = c_te;
        lj = te * dprev_inv;
        dj = tdj - lj * te;
      }
      const double pivmin = fmax(scale * 4.9e-32, pivmin0);
      if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
      neg += (dj < 0.0);
      const double dinv = pivot_rcp(dj);
      const double g0 = c_h0 - sigma * c_m0;
      const double g1 = c_h1 - sigma * c_m1;
      const double y0 = g0 - lj * y0p, y1 = g1 - lj * y1p;
      s00 = fma(y0 * dinv, y0, s00);
      s10 = fma(y1 * dinv, y0, s10);
      s11 = fma(y1 * dinv, y1, s11);
      if (store && g == 0) {
        Fj[0] = dinv;
        Fj[nbC] = lj;
        Fj[2 * nbC] = y0 * dinv;
        Fj[3 * nbC] = y1 * dinv;
      }
      dprev_inv = dinv; y0p = y0; y1p = y1;
    }
    if (store && g == 0) {  // W = L^-T D^-1 y
      double w0n = 0.0, w1n = 0.0;
#pragma unroll 4
      for (int j = nb - 1; j >= 0; --j) {
        double w0 = F[(2 * nb + j) * C + c], w1 = F[(3 * nb + j) * C + c];
        if (j < nb - 1) {
          const double ln = F[(nb + j + 1) * C + c];
          w0 -= ln * w0n; w1 -= ln * w1n;
        }
        F[(2 * nb + j) * C + c] = w0;
        F[(3 * nb + j) * C + c] = w1;
        w0n = w0; w1n = w1;
      }
    }
    sc[3 * c] = (R[(6 * nb) * C + c] - sigma * S.mhh[c]) - s00;
    sc[3 * c + 1] = (R[(6 * nb + 1) * C + c] - sigma * S.mhh[C + c]) - s10;
    sc[3 * c + 2] = (R[(6 * nb + 2) * C + c] - sigma * S.mhh[2 * C + c]) - s11;
    S.negc[g * (C + 2) + c] = neg;
  }
  __syncthreads();
  EIG_FT_MARK(0);   // cell sweeps (+ barrier)
  // hat node j sits between cell j (its right hat, g=0) and cell j+1 (its left hat, g=1).
  // Twisted factorisation of the (C-1) tridiagonal: a sweep down from the top and a sweep up from the
  // bottom meet at the twist index tw; inertia = negatives of both sweeps + sign of the twist pivot.
  const int m = C - 1, tw = m / 2;
  auto hat_count = [&](int gg) { S.cntg[gg] = eig_hat_count(S.sc + (size_t)gg * 3 * C, S.negc + gg * (C + 2), C); };
  if (store) {
    // shift 0, factors kept: thread 0 runs the top sweep, thread 32 the bottom sweep
    const double* sc = S.sc;
    // shifts riding along count-only: one thread each, in warps of their own (64: shift 1, 96: shift 2)
    if (ng >= 2 && threadIdx.x == 64) hat_count(1);
    if (ng >= 3 && threadIdx.x == 96) hat_count(2);
    if (threadIdx.x == 0) {
      int neg = 0;
      for (int cc = 0; cc < C; ++cc) neg += S.negc[cc];
      double dinvp = 0.0, offprev = 0.0, scale = 0.0;
#pragma unroll 4
      for (int j = 0; j < tw; ++j) {
        const double diag = sc[3 * j] + sc[3 * (j + 1) + 2];
        scale = fmax(scale, fabs(diag));
        double l = 0.0, dj = diag;
        if (j > 0) { l = offprev * dinvp; dj = diag - l * offprev; }
        const double pivmin = fmax(scale * 4.9e-32, DBL_MIN * 1e8);
        if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
        neg += (dj < 0.0);
        dinvp = pivot_rcp(dj);
        S.tri_di[j] = dinvp;
        if (j > 0) S.tri_l[j - 1] = l;
        offprev = sc[3 * (j + 1) + 1];
      }
      S.xh[0] = (tw > 0) ? offprev * dinvp * offprev : 0.0;
      if (tw > 0) S.tri_l[tw - 1] = offprev * dinvp;
      S.negc[C] = neg;
    } else if (threadIdx.x == 32) {
      int neg = 0;
      double dinvn = 0.0, offnext = 0.0, scale = 0.0;
#pragma unroll 4
      for (int j = m - 1; j > tw; --j) {
        const double diag = sc[3 * j] + sc[3 * (j + 1) + 2];
        scale = fmax(scale, fabs(diag));
        double u = 0.0, dj = diag;
        if (j < m - 1) { u = offnext * dinvn; dj = diag - u * offnext; }
        const double pivmin = fmax(scale * 4.9e-32, DBL_MIN * 1e8);
        if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
        neg += (dj < 0.0);
        dinvn = pivot_rcp(dj);
        S.tri_di[j] = dinvn;
        if (j < m - 1) S.tri_l[j] = u;
        offnext = sc[3 * j + 1];
      }
      S.xh[1] = (tw < m - 1) ? offnext * dinvn * offnext : 0.0;
      if (tw < m - 1) S.tri_l[tw] = offnext * dinvn;
      S.negc[C + 1] = neg;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const double diag = sc[3 * tw] + sc[3 * (tw + 1) + 2];
      double gam = diag - S.xh[0] - S.xh[1];
      const double pivmin = fmax(fabs(diag) * 4.9e-32, DBL_MIN * 1e8);
      if (fabs(gam) < pivmin) gam = (gam < 0.0) ? -pivmin : pivmin;
      S.tri_di[tw] = pivot_rcp(gam);
      S.cntg[0] = S.negc[C] + S.negc[C + 1] + (gam < 0.0);
    }
  } else if (c == 0 && g < ng) {
    hat_count(g);
  }
  __syncthreads();
  EIG_FT_MARK(1);   // hat system (+ barriers)
}

// ---- y = T~(sigma)^{-1} r in reduced coordinates with the stored factors
__device__ __forceinline__ void eig_solve(const EigShared& S, int C, int nb) {
  // B. thread per cell: z = (T - sI)^-1 r_b ; hat partials W^T r_b
  {
    const int c = threadIdx.x;
    if (c < C) {
      const double* F = S.fac;
      double t0 = 0.0, t1 = 0.0, zp = 0.0;
#pragma unroll 4
      for (int j = 0; j < nb; ++j) {
        const double rj = S.r.t[j * C + c];
        t0 = fma(F[(2 * nb + j) * C + c], rj, t0);
        t1 = fma(F[(3 * nb + j) * C + c], rj, t1);
        const double z = rj - F[(nb + j) * C + c] * zp;
        S.y.t[j * C + c] = z;
        zp = z;
      }
      double zn = 0.0;
#pragma unroll 4
      for (int j = nb - 1; j >= 0; --j) {
        double z = S.y.t[j * C + c] * F[j * C + c];
        if (j < nb - 1) z -= F[(nb + j + 1) * C + c] * zn;
        S.y.t[j * C + c] = z;
        zn = z;
      }
      S.hatpart[2 * c] = t0;
      S.hatpart[2 * c + 1] = t1;
    }
  }
  __syncthreads();
  // C. hat system, twisted substitution (thread 0 from the top, thread 32 from the bottom)
  {
    const int m = C - 1, tw = m / 2;
    if (threadIdx.x == 0) {
      double zprev = 0.0;
      for (int j = 0; j <= tw; ++j) {
        double rh = S.r.h[j] - S.hatpart[2 * j] - S.hatpart[2 * (j + 1) + 1];
        if (j > 0) rh -= S.tri_l[j - 1] * zprev;
        if (j < tw) { S.xh[j] = rh; zprev = rh; } else S.red8[AKS_NW] = rh;
      }
    } else if (threadIdx.x == 32) {
      double znext = 0.0;
      for (int j = m - 1; j > tw; --j) {
        double rh = S.r.h[j] - S.hatpart[2 * j] - S.hatpart[2 * (j + 1) + 1];
        if (j < m - 1) rh -= S.tri_l[j] * znext;
        S.xh[j] = rh;
        znext = rh;
      }
      S.red8[AKS_NW + 1] = (tw < m - 1) ? S.tri_l[tw] * znext : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0 || threadIdx.x == 32) {
      const double xt = (S.red8[AKS_NW] - S.red8[AKS_NW + 1]) * S.tri_di[tw];
      if (threadIdx.x == 0) {
        S.y.h[tw] = xt;
        double xn = xt;
        for (int j = tw - 1; j >= 0; --j) {
          const double v = S.xh[j] * S.tri_di[j] - S.tri_l[j] * xn;
          xn = v;
          S.y.h[j] = v;
        }
      } else {
        double xp = xt;
        for (int j = tw + 1; j < m; ++j) {
          const double v = S.xh[j] * S.tri_di[j] - S.tri_l[j - 1] * xp;
          xp = v;
          S.y.h[j] = v;
        }
      }
    }
    __syncthreads();
  }
  // D. y_b = z - W y_h
  for (int idx = threadIdx.x; idx < nb * C; idx += blockDim.x) {
    const int j = idx / C, c = idx - j * C;
    const double xh0 = (c < C - 1) ? S.y.h[c] : 0.0;
    const double xh1 = (c > 0) ? S.y.h[c - 1] : 0.0;
    S.y.t[idx] -= S.fac[(2 * nb + j) * C + c] * xh0 + S.fac[(3 * nb + j) * C + c] * xh1;
  }
  __syncthreads();
}

#include "k_eig_pf.cuh"

// EIG_PF = 1: product-form factorisation / no-W solve of k_eig_pf.cuh (the ratio form above stays the fallback for
// degenerate shifts).  It shortens the dependent FP64 chain of a pivot step from 92 to 22 cycles, yet it is SLOWER
// in the kernel (tools/micro/eig_factor_bench.cu on a B200: cell sweeps 26.9 k vs 12.0 k cycles per factorisation,
// solve 19.1 k vs 10.9 k): with one or two active warps per scheduler the sweeps are bound by the number of
// instructions issued (~5 cycles each, little ILP), not by the FP64 latency of the recurrence, and the product form
// issues ~50 % more of them.  Off by default; kept as the measured experiment (profiles/r02_k_eig_experiments.md).
#ifndef EIG_PF
#define EIG_PF 0
#endif
__device__ __forceinline__ bool eig_factor_any(const EigShared& S, int C, int nb, int ng, double sigma, int store EIG_FT_DECL) {
#if EIG_PF
  if (eig_factor_pf(S, C, nb, ng, sigma, store
#ifdef AKS_EIG_TIMING
                    , ft
#endif
                    )) return true;
#endif
  eig_factor(S, C, nb, ng, sigma, store
#ifdef AKS_EIG_TIMING
             , ft
#endif
             );
  return false;
}
__device__ __forceinline__ void eig_solve_any(const EigShared& S, int C, int nb, bool pf_layout) {
  if (pf_layout) eig_solve_pf(S, C, nb);
  else eig_solve(S, C, nb);
}

// ---- out = M~ in  (which = 0)  or  out = H~ in  (which = 1), reduced coordinates, O(nb) per cell
__device__ __forceinline__ void eig_apply_red(const EigShared& S, int C, int nb, int which, const RedVec& in,
                                              const RedVec& out) {
  const int c = threadIdx.x;
  if (c < C) {
    const double* R = S.red;
    const double xR = (c < C - 1) ? in.h[c] : 0.0, xL = (c > 0) ? in.h[c - 1] : 0.0;
    const int o0 = which ? 2 * nb : 4 * nb, o1 = which ? 3 * nb : 5 * nb;
    double p0 = 0.0, p1 = 0.0, xprev = 0.0, xj = in.t[c];
#pragma unroll 4
    for (int j = 0; j < nb; ++j) {
      const double xnext = (j + 1 < nb) ? in.t[(j + 1) * C + c] : 0.0;
      const double c0 = R[(o0 + j) * C + c], c1 = R[(o1 + j) * C + c];
      double v;
      if (which) {
        v = R[j * C + c] * xj;
        if (j > 0) v = fma(R[(nb + j - 1) * C + c], xprev, v);
        if (j + 1 < nb) v = fma(R[(nb + j) * C + c], xnext, v);
      } else v = xj;
      out.t[j * C + c] = v + c0 * xR + c1 * xL;
      p0 = fma(c0, xj, p0);
      p1 = fma(c1, xj, p1);
      xprev = xj; xj = xnext;
    }
    double h00, h10, h11;
    if (which) { h00 = R[(6 * nb) * C + c]; h10 = R[(6 * nb + 1) * C + c]; h11 = R[(6 * nb + 2) * C + c]; }
    else { h00 = S.mhh[c]; h10 = S.mhh[C + c]; h11 = S.mhh[2 * C + c]; }
    S.hatpart[2 * c] = p0 + h00 * xR + h10 * xL;
    S.hatpart[2 * c + 1] = p1 + h10 * xR + h11 * xL;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C - 1; j += blockDim.x) out.h[j] = S.hatpart[2 * j] + S.hatpart[2 * (j + 1) + 1];
  __syncthreads();
}

__device__ __forceinline__ double eig_dot_red(const EigShared& S, int C, int nb, const RedVec& a, const RedVec& bv) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb * C; i += blockDim.x) acc = fma(a.t[i], bv.t[i], acc);
  for (int i = threadIdx.x; i < C - 1; i += blockDim.x) acc = fma(a.h[i], bv.h[i], acc);
  return block_sum(acc, S.red8);
}

#define EIG_NBMAX 32
// sum_j base[j*ld + lane] * x_j with x_j held by lane j: all loads are issued before the first use
// (they are L2 hits of ~500 cycles each; a dependent load-FMA loop would expose every one of them)
__device__ __forceinline__ double warp_colmatvec(const double* __restrict__ base, int rows, int ld, int lane,
                                                 bool act, double xl) {
  double pv[EIG_NBMAX];
#pragma unroll
  for (int j = 0; j < EIG_NBMAX; ++j) pv[j] = (j < rows && act) ? base[j * ld + lane] : 0.0;
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < EIG_NBMAX; ++j)
    if (j < rows) acc = fma(pv[j], __shfl_sync(0xffffffffu, xl, j), acc);
  return acc;
}

__device__ inline void eig_carve(const DiscDev& d, double* sm, EigShared& S) {
  const int C = d.C, n = d.n, nb = d.nb;
  double* ptr = sm;
  S.mhh = ptr; ptr += 3 * C;
  S.fac = ptr; ptr += (size_t)4 * nb * C;
  S.x.t = ptr; ptr += (size_t)nb * C; S.x.h = ptr; ptr += C;
  S.r.t = ptr; ptr += (size_t)nb * C; S.r.h = ptr; ptr += C;
  S.y.t = ptr; ptr += (size_t)nb * C; S.y.h = ptr; ptr += C;
  S.hatpart = ptr; ptr += 2 * C;
  S.sc = ptr; ptr += EIG_NG * 3 * C;
  S.tri_di = ptr; ptr += C;
  S.tri_l = ptr; ptr += C;
  S.xh = ptr; ptr += C;
  S.red8 = ptr; ptr += AKS_NW + 8;
  S.negc = (int*)ptr;
  S.cntg = S.negc + EIG_NG * (C + 2);
  S.hatg = S.cntg + EIG_NG + 2;
  S.bub0 = S.hatg + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    S.hatg[c] = (c < C - 1) ? d.l2g[c * n + 0] : 0;
    S.bub0[c] = d.l2g[c * n + 2];
    S.mhh[c] = d.M0pk[(size_t)c * d.npk + 0];
    S.mhh[C + c] = d.M0pk[(size_t)c * d.npk + 1];
    S.mhh[2 * C + c] = d.M0pk[(size_t)c * d.npk + 2];
  }
}

// levels [klo, khi) of a channel handled by a pass (see k_eig)
__device__ __forceinline__ bool eig_window(const BatchDev& b, int p, size_t chan, int mode, int& klo, int& khi) {
  if (mode != 2 && b.status[p] != 0) return false;
  const int nhp = b.nhp[p];
  if (mode == 0) { klo = 0; khi = min(nhp, b.kact[chan]); }
  else {
    if (b.redo[p] == 0) return false;
    klo = b.kdone[chan];
    khi = (mode == 1) ? min(nhp, b.kreq[chan]) : nhp;   // refill: exactly the levels the certificate missed
  }
  return klo < khi;
}

enum { EST_BRACKET_LO = 0, EST_BRACKET_HI, EST_BISECT, EST_RQI, EST_CERT_LO, EST_CERT_HI, EST_FINAL, EST_DONE };

// Eigenpair window.  Only levels that can influence the aufbau have to be current in every SCF step:
// mode 0 solves the first kact[channel] levels of each channel (occupied levels of the previous step
// + margin; the others get the sentinel AKS_EPS_MISSING and sort last); k_eig_cert then CERTIFIES by an
// inertia count that every unsolved level lies above everything the aufbau examined, else raises redo[p], for which
// mode 1 solves the remaining levels (cold) before the aufbau is repeated.  Mode 2 is the same fill
// pass on request of the host (aks_get_state / end of a run: the reference returns all nh pairs).
#ifndef EIG_NT
#define EIG_NT 128   // threads per eigenpair CTA: the sweeps are thread-per-cell, occupancy (4 CTAs/SM) matters more
#endif
__global__ void __launch_bounds__(EIG_NT, EIG_RED_SMEM ? 2 : ((EIG_NT == 128) ? 4 : 2))
k_eig(DiscDev d, BatchDev b, int mode) {
  const int k = blockIdx.x, ch = blockIdx.y, p = blockIdx.z;
  if (mode != 2 && b.status[p] != 0) return;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p] || k >= b.nhp[p]) return;
  {
    const size_t chan0 = ((size_t)p * b.s + e) * b.Lmax + l;
    if (mode == 0) {
      if (k >= b.kact[chan0]) {
        if (threadIdx.x == 0) b.eps_new[chan0 * b.nhmax + k] = AKS_EPS_MISSING;
        return;
      }
    } else if (b.redo[p] == 0 || k < b.kdone[chan0] || (mode == 1 && k >= b.kreq[chan0])) return;
  }
  extern __shared__ __align__(16) double sm[];
  const int C = d.C, nb = d.nb;
#if EIG_RED_SMEM
  // one bulk copy (TMA engine) of the channel's reduced data, overlapped with the set-up below
  __shared__ __align__(8) uint64_t red_bar;
  double* red_s = sm;
  {
    const size_t chan_ = ((size_t)p * b.s + e) * b.Lmax + l;
    const uint32_t bytes = (uint32_t)(sizeof(double) * (size_t)red_stride(nb) * C);
    if (threadIdx.x == 0) {
      mbar_init(&red_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&red_bar, bytes);
      bulk_g2s(red_s, b.red + chan_ * C * red_stride(nb), bytes, &red_bar);
    }
  }
  double* sm_rest = sm + (size_t)red_stride(nb) * C;
#else
  double* sm_rest = sm;
#endif
#ifdef AKS_EIG_TIMING
  long long t_set = 0, t_fac = 0, t_sol = 0, t_app = 0, t_fin = 0, t_all = clock64(), t0_ = clock64();
  long long ftv_[4] = {0, 0, 0, 0};
  long long* ft_ = ftv_;
#define TSTART() t0_ = clock64()
#define TADD(v) v += clock64() - t0_
#else
#define TSTART()
#define TADD(v)
#endif
  EigShared S;
  eig_carve(d, sm_rest, S);

  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  const double* eps_ch = b.eps + chan * b.nhmax;
  const int nhp = b.nhp[p];
  // eps / U of a level outside the last windows are older than one step: still a start, never trusted
  // (every result is certified by inertia counts)
  const bool have_prev = b.warm[p] != 0;
#if EIG_RED_SMEM
  S.red = red_s;
  __syncthreads();            // red_bar initialised by thread 0 before anyone waits on it
  mbar_wait(&red_bar, 0);     // the bulk copy has landed (visible to every thread that observed the phase)
#else
  S.red = b.red + chan * C * red_stride(nb);
  __syncthreads();
#endif

  // bracket == witnesses: lo = largest shift known with count <= k (clo), hi = smallest with
  // count >= k+1 (chi); -1 = not known yet
  double lo = -DBL_MAX, hi = DBL_MAX, step = 1.0, sigma = 0.0, rho = 0.0, prev_upd = DBL_MAX;
  int clo = -1, chi = -1;
  int state, nfac = 0, its = 0, its_cap = 12, ncount = 0;
  int bis_left = 0, rebisect = 0;   // cold path: extra bisection rounds between inverse iterations (close pairs)
  bool from_warm = false, need_start = false, certified = false, converged = false, refined = false, cluster = false, loose = false;
  // distance to the neighbouring levels of the previous step: RQI converges cubically, so an
  // eigenvalue update below 1e-10 of it means the current vector is already converged
  // Only eigenvalues of the LAST step qualify: levels outside the window keep older values, and an
  // overestimated gap would end the iteration early (eigenvector error ~ update / gap).  So the estimate is
  // used in the window pass only, from neighbours that are inside the window themselves.
  double gap_est = 0.0;
  if (have_prev && mode == 0) {
    const double ep = eps_ch[k];
    const int kwin = min(nhp, b.kact[chan]);
    gap_est = DBL_MAX;
    if (k > 0) gap_est = fmin(gap_est, fabs(ep - eps_ch[k - 1]));
    if (k + 1 < kwin) gap_est = fmin(gap_est, fabs(eps_ch[k + 1] - ep));
    if (!(gap_est < 1e300)) gap_est = 0.0;
  }

  auto cold_init = [&]() {
    from_warm = false;
    its = 0; its_cap = 12; prev_upd = DBL_MAX; converged = false; loose = false;
    bis_left = 0; rebisect = 0;
    need_start = true;
    if (clo >= 0 && chi >= 0 && lo < hi) { state = EST_BISECT; sigma = 0.5 * (lo + hi); return; }
    double blo, bhi;
    if (have_prev) {
      const double ep = eps_ch[k];
      double gap = DBL_MAX;
      if (k > 0) gap = fmin(gap, ep - eps_ch[k - 1]);
      if (k + 1 < nhp) gap = fmin(gap, eps_ch[k + 1] - ep);
      if (!(gap < 1e300) || !(gap > 0.0)) gap = fmax(fabs(ep), 1.0);
      const double dlt = fmax(0.5 * gap, 1e-12 * fmax(fabs(ep), 1e-3));
      blo = ep - dlt; bhi = ep + dlt;
    } else if (b.niter[p] == 0) {
      // first iteration from D = 0: the Hamiltonian is bare Coulomb, level k of channel l sits at the
      // hydrogenic -Z^2 / (2 n^2), n = k + l + 1 (box states aside); only a bracket to start from
      const double Z = b.Z[p], nq = (double)(k + l + 1);
      const double eh = -0.5 * Z * Z / (nq * nq), eh1 = -0.5 * Z * Z / ((nq + 1.0) * (nq + 1.0));
      const double dlt = 0.25 * (eh1 - eh);
      blo = eh - dlt; bhi = eh + dlt;
    } else {
      const double Z = b.Z[p];
      blo = -(Z * Z + 10.0); bhi = 1.0;
    }
    if (clo >= 0) blo = lo;
    if (chi >= 0) bhi = hi;
    if (!(blo < bhi)) { if (clo >= 0) bhi = blo + fmax(fabs(blo), 1.0); else blo = bhi - fmax(fabs(bhi), 1.0); }
    step = fmax(bhi - blo, 1e-3 * fmax(fabs(blo), fabs(bhi)));
    if (clo < 0) { lo = blo; hi = (chi >= 0) ? hi : bhi; state = EST_BRACKET_LO; sigma = blo; }
    else { hi = bhi; state = EST_BRACKET_HI; sigma = bhi; }
  };
  // M~-normalise x, r = M~ x
  auto normalise = [&]() {
    eig_apply_red(S, C, nb, 0, S.x, S.r);
    const double xmx = eig_dot_red(S, C, nb, S.x, S.r);
    const double sc = 1.0 / sqrt(xmx);
    for (int i = threadIdx.x; i < nb * C; i += blockDim.x) { S.x.t[i] *= sc; S.r.t[i] *= sc; }
    for (int i = threadIdx.x; i < C - 1; i += blockDim.x) { S.x.h[i] *= sc; S.r.h[i] *= sc; }
    __syncthreads();
  };

  double* Xt = b.Xt + (chan * b.nhmax + k) * (size_t)(nb * C + C);
  // refill pass after a failed certificate: the missing levels lie between the last solved level of the
  // channel and e_cert, and the certificate's count is a witness for the upper end -- start from that bracket
  bool bracketed = false;
  if (mode == 1 && b.full_refill[p] == 0) {
    const int kd = b.kdone[chan];
    if (kd >= 1 && k >= kd) {
      const double el = b.eps_new[chan * b.nhmax + kd - 1];
      lo = el - fmax(1e-8 * fabs(el), 1e-12); clo = kd - 1;
      hi = b.e_cert[p]; chi = b.cert_cnt[chan];
      bracketed = (lo < hi) && (chi >= k + 1);
      if (!bracketed) { lo = -DBL_MAX; hi = DBL_MAX; clo = chi = -1; }
    }
  }
  if (have_prev && b.stop[p] < 0.3 && !bracketed) {
    // ---- warm start: previous eigenvector, already in reduced coordinates (k_eig_init);
    // its Rayleigh quotient with the new H~ is the first shift
    for (int i = threadIdx.x; i < nb * C; i += blockDim.x) S.x.t[i] = Xt[i];
    for (int i = threadIdx.x; i < C - 1; i += blockDim.x) S.x.h[i] = Xt[nb * C + i];
    __syncthreads();
    normalise();
    eig_apply_red(S, C, nb, 1, S.x, S.y);
    sigma = eig_dot_red(S, C, nb, S.x, S.y);
    state = EST_RQI; from_warm = true; its_cap = 6;
    if (!isfinite(sigma)) cold_init();
  } else {
    cold_init();
  }
  TADD(t_set);

  auto narrow = [&]() { return clo == k && chi == k + 1 && (hi - lo) <= 0.05 * fmax(fmax(fabs(lo), fabs(hi)), 1e-12); };
  auto fail = [&]() {
    if (from_warm) cold_init();
    else { state = EST_DONE; certified = false; }
  };
  auto witness = [&](double s, int cnt) {
    if (cnt <= k) { if (s > lo) { lo = s; clo = cnt; } }
    else if (s < hi) { hi = s; chi = cnt; }
  };

  while (state != EST_DONE && nfac < 600) {
    if (state == EST_RQI && need_start) {
      // fixed start, a different pattern for every level: two members of a cluster of one channel (same shift,
      // same bracket) then converge to different vectors of its invariant subspace (k_eig_scale orthogonalises them)
      const int pk = 2 * k + 5;
      for (int i = threadIdx.x; i < nb * C; i += blockDim.x) S.x.t[i] = 0.3 + (double)((i * (k + 2)) % pk) / (double)pk - 0.07 * (double)((i * 7 + 3 * k) % 11);
      for (int i = threadIdx.x; i < C - 1; i += blockDim.x) S.x.h[i] = 0.6 - (double)((i * (k + 3)) % pk) / (double)pk + 0.03 * (double)((i * 5 + k) % 7);
      __syncthreads();
      normalise();
      need_start = false;
    }
    if (state == EST_FINAL) break;
    const int store = (state == EST_RQI) ? 1 : 0;
    // shifts of this round: one (RQI) or a ladder / partition / certificate pair evaluated concurrently
    const int ngmax = min(EIG_NG, (int)blockDim.x / C);
    int ng = 1;
    const int gme = threadIdx.x / C;
    double sg_mine = sigma;
    // EIG_PIGGYBACK = 1: in Rayleigh-quotient rounds the certificate's two shifts (current estimate -/+ dl) ride along
    // count-only on the threads that would idle (3 x C <= 128), so that a converged pair is usually certified by these
    // witnesses without a separate certificate round.  Measured on the 86-atom sweep: count-only rounds per pair drop
    // from 1.7 to 1.4, but every RQI round pays for two more hat sweeps (11.8 k vs 9.0 k cycles): k_eig 0.281 vs
    // 0.237 ms.  Off.
    const double pref = (its == 0) ? sigma : rho;
    const double dlp = fmax(1e-8 * fabs(pref), 1e-13);
    auto shift_of = [&](int g) -> double {
      if (state == EST_RQI) return (g == 0) ? sigma : ((g == 1) ? pref - dlp : pref + dlp);
      if (state == EST_BRACKET_LO) return sigma - step * (double)((1 << g) - 1);
      if (state == EST_BRACKET_HI) return sigma + step * (double)((1 << g) - 1);
      if (state == EST_BISECT) return lo + (hi - lo) * (double)(g + 1) / (double)(ng + 1);
      if (state == EST_CERT_LO) return (g == 0) ? sigma : rho + fmax(1e-8 * fabs(rho), 1e-13);
      return sigma;
    };
    if (state == EST_BRACKET_LO || state == EST_BRACKET_HI || state == EST_BISECT) ng = ngmax;
    else if (state == EST_CERT_LO) ng = min(2, ngmax);
#if EIG_PIGGYBACK
    else if (state == EST_RQI && ngmax >= 3 && isfinite(pref)) ng = 3;
#endif
    sg_mine = shift_of(min(gme, ng - 1));
    TSTART();
    const bool pf_layout = eig_factor_any(S, C, nb, ng, sg_mine, store EIG_FT_ARG);
    TADD(t_fac);
    nfac += 1;
    if (!store) ncount += 1;   // count-only rounds (each evaluates up to EIG_NG shifts)
    const int cnt = S.cntg[0];
    if (state == EST_BRACKET_LO) {
      // ladder going down: first rung with count <= k closes the bracket from below
      int gfound = -1;
      for (int g = 0; g < ng; ++g) {
        const int cg = S.cntg[g];
        const double sgv = shift_of(g);
        if (cg <= k) { gfound = g; lo = sgv; clo = cg; break; }
        hi = sgv; chi = cg;
      }
      if (gfound < 0) { sigma = shift_of(ng - 1) - step * (double)(1 << (ng - 1)); step *= (double)(1 << ng); }
      else if (chi >= 0) { state = narrow() ? EST_RQI : EST_BISECT; sigma = 0.5 * (lo + hi); }
      else { state = EST_BRACKET_HI; sigma = hi; }
    } else if (state == EST_BRACKET_HI) {
      int gfound = -1;
      for (int g = 0; g < ng; ++g) {
        const int cg = S.cntg[g];
        const double sgv = shift_of(g);
        if (cg >= k + 1) { gfound = g; hi = sgv; chi = cg; break; }
        lo = sgv; clo = cg;
      }
      if (gfound < 0) { sigma = shift_of(ng - 1) + step * (double)(1 << (ng - 1)); step *= (double)(1 << ng); }
      else { state = narrow() ? EST_RQI : EST_BISECT; sigma = 0.5 * (lo + hi); }
    } else if (state == EST_BISECT) {
      double nlo = lo, nhi = hi;
      int nclo = clo, nchi = chi;
      for (int g = 0; g < ng; ++g) {
        const int cg = S.cntg[g];
        const double sgv = shift_of(g);
        if (cg <= k) { if (sgv > nlo) { nlo = sgv; nclo = cg; } }
        else if (sgv < nhi) { nhi = sgv; nchi = cg; }
      }
      lo = nlo; hi = nhi; clo = nclo; chi = nchi;
      const double mid = 0.5 * (lo + hi);
      if (bis_left > 0) { if (--bis_left == 0 || !(mid > lo && mid < hi)) { bis_left = 0; state = EST_RQI; } }
      else if (narrow() || !(mid > lo && mid < hi)) state = EST_RQI;
      sigma = mid;
    } else if (state == EST_RQI) {
      witness(sigma, cnt);
      if (ng >= 3) { witness(shift_of(1), S.cntg[1]); witness(shift_of(2), S.cntg[2]); }
      TSTART();
      eig_solve_any(S, C, nb, pf_layout);                         // y = (H~ - sigma M~)^{-1} M~ x
      TADD(t_sol);
      TSTART();
      const double yr = eig_dot_red(S, C, nb, S.y, S.r);          // y^T M~ x
      eig_apply_red(S, C, nb, 0, S.y, S.r);                       // r <- M~ y
      const double ymy = eig_dot_red(S, C, nb, S.y, S.r);
      rho = sigma + yr / ymy;
      const double sc = 1.0 / sqrt(ymy);
      for (int i = threadIdx.x; i < nb * C; i += blockDim.x) { S.x.t[i] = S.y.t[i] * sc; S.r.t[i] *= sc; }
      for (int i = threadIdx.x; i < C - 1; i += blockDim.x) { S.x.h[i] = S.y.h[i] * sc; S.r.h[i] *= sc; }
      __syncthreads();
      TADD(t_app);
      ++its;
      const double upd = fabs(rho - sigma);
      if (!isfinite(rho)) { fail(); continue; }
      // third alternative: the iteration stagnates with updates below the certificate's resolution -- two levels of the
      // channel closer than that (1e-9 relative: a resonance crossing a box state) make the Rayleigh quotient hop
      // between them; the iterate lies in their invariant subspace and the certificate below classifies it (CLUSTER)
      const bool stagnant = its >= 3 && upd >= 0.25 * prev_upd;
      const bool tight = upd <= 1e-13 * fabs(rho) || upd <= 1e-10 * gap_est || (stagnant && upd <= 1e-9 * fmax(fabs(rho), 1e-6));
      if (tight || (stagnant && upd <= 0.5e-8 * fmax(fabs(rho), 1e-6))) {
        converged = true;
        loose = !tight;   // acceptable only as a member of a cluster (decided by the certificate)
        const double dl = fmax(1e-8 * fabs(rho), 1e-13);
        if (loose || !(clo == k && lo < rho) || !(chi == k + 1 && hi > rho)) { state = EST_CERT_LO; sigma = rho - dl; }
        else { state = EST_FINAL; certified = true; }
        continue;
      }
      prev_upd = upd;
      if (its >= its_cap) { fail(); continue; }
      if (rho > lo && rho < hi) sigma = rho;
      else if (from_warm) { fail(); continue; }
      else if (clo == k && chi == k + 1 && rebisect < 8) {
        // Level k is isolated in (lo, hi) but the Rayleigh quotient left the bracket: a neighbour as close as the
        // bracket is wide (a box state crossing a resonance, 1e-10 apart) still weighs in the iterate.  Two more
        // bisection rounds narrow the bracket 16x (EIG_NG = 3 shifts a round), then the inverse iteration goes on
        // FROM THE CURRENT VECTOR with the shift at the new midpoint: its contraction improves 16x per cycle.
        state = EST_BISECT; bis_left = 2; rebisect += 1; its_cap += 3;
        sigma = 0.5 * (lo + hi);
      } else sigma = 0.5 * (lo + hi);
    } else if (state == EST_CERT_LO) {
      // sigma = rho - dl (shift 0); with two groups the upper certificate rho + dl rides along (shift 1)
      bool ok = true;
      const int c_lo = cnt, c_hi = (ng >= 2) ? S.cntg[1] : -1;
      if (!(clo == k && lo < rho)) {
        if (cnt == k) { lo = sigma; clo = k; } else { witness(sigma, cnt); ok = false; }
      }
      if (ok && !(chi == k + 1 && hi > rho)) {
        if (ng >= 2) {
          const int c2 = S.cntg[1];
          const double s2 = shift_of(1);
          if (c2 == k + 1) { hi = s2; chi = c2; } else { witness(s2, c2); ok = false; }
        } else { state = EST_CERT_HI; sigma = rho + fmax(1e-8 * fabs(rho), 1e-13); continue; }
      }
      if (ok && !loose) { state = EST_FINAL; certified = true; }
      else if (ok) { converged = false; loose = false; fail(); }   // an isolated level has to converge properly
      else if (converged && c_hi >= k + 1 && c_lo <= k && c_hi - c_lo <= 3) {
        // CLUSTER: the converged pair is an eigenpair (Rayleigh-quotient iteration converged on it) and level k lies
        // in (rho - dl, rho + dl) together with one or two neighbours of the same channel (an unbound resonance crossing
        // a box state: 3e-12 apart for H-).  Its energy is right to the width of the cluster, which inertia counts
        // cannot resolve further; k_eig_scale checks that the members of a cluster did not converge on the SAME
        // vector and makes them M0-orthogonal.
        state = EST_FINAL; certified = true; cluster = true;
      } else fail();
    } else if (state == EST_CERT_HI) {
      if (cnt == k + 1) { state = EST_FINAL; certified = true; }
      else { witness(sigma, cnt); fail(); }
    }
  }
  // ---- hand the converged reduced vector to k_eig_final (back-transform, exact M0 normalisation,
  // Rayleigh quotient with the ORIGINAL blocks)
  TSTART();
  refined = (state == EST_FINAL);
  for (int i = threadIdx.x; i < nb * C; i += blockDim.x) Xt[i] = S.x.t[i];
  for (int i = threadIdx.x; i < C - 1; i += blockDim.x) Xt[nb * C + i] = S.x.h[i];
  if (threadIdx.x == 0) {
    b.eps_new[chan * b.nhmax + k] = rho;
    b.eig_info[chan * b.nhmax + k] = (ncount << 8) | ((nfac - ncount) & 0x3f) | (cluster ? 0x40 : 0) | ((certified && converged && refined) ? 0 : 0x80);
#ifdef AKS_EIG_TIMING
    TADD(t_fin);
    long long* dbg = b.eig_dbg + (chan * b.nhmax + k) * 8;
    dbg[0] = t_set; dbg[1] = t_fac; dbg[2] = t_sol; dbg[3] = t_app; dbg[4] = t_fin; dbg[5] = clock64() - t_all;
    dbg[6] = ftv_[0]; dbg[7] = ftv_[1];   // split of the factorisations: cell sweeps, hat system
#endif
  }
}

// ------------------------------------------------------------------------------------------
// k_eig_cert: certificate of the eigenpair window.  For every channel of which only the first kdone levels
// were solved: #{eigenvalues < e_cert} (ONE inertia count, the cost of a single RQI factorisation) must equal
// the number of solved levels below e_cert; otherwise an unsolved level could have been occupied and the
// problem is marked for the refill pass.  One CTA per (channel, problem).
__global__ void __launch_bounds__(EIG_NT, (EIG_NT == 128) ? 4 : 2)
k_eig_cert(DiscDev d, BatchDev b) {
  const int ch = blockIdx.x, p = blockIdx.y;
  if (b.status[p] != 0) return;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p]) return;
  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  if (b.full_refill[p]) {   // the fill ran out of solved levels (k_aufbau): everything that is left
    if (threadIdx.x == 0) b.kreq[chan] = b.nhp[p];
    return;
  }
  const int nhp = b.nhp[p], kd = min(nhp, b.kdone[chan]);
  if (threadIdx.x == 0) b.kreq[chan] = kd;
  if (kd >= nhp || b.aufbau_kind != 0 || b.redo[p] != 0) return;   // nothing truncated / occupations do not depend on eps
  extern __shared__ double sm[];
  const int C = d.C, nb = d.nb;
  EigShared S;
  eig_carve(d, sm, S);
  S.red = b.red + chan * C * red_stride(nb);
  __syncthreads();
  const double sigma = b.e_cert[p];
#ifdef AKS_EIG_TIMING
  long long* ft_ = nullptr;
#endif
  eig_factor_any(S, C, nb, 1, sigma, 0 EIG_FT_ARG);
  if (threadIdx.x == 0) {
    int below = 0;
    for (int k = 0; k < kd; ++k) below += (b.eps_new[chan * b.nhmax + k] < sigma);
    b.cert_cnt[chan] = S.cntg[0];
    if (S.cntg[0] > below) {
      // exactly cnt - below unsolved levels lie below e_cert.  Once they are solved the aufbau is final: more
      // low candidates can only lower the last level it examines, so e_cert still bounds the unsolved rest.
      b.kreq[chan] = min(nhp, kd + (S.cntg[0] - below));
      if (atomicExch(&b.redo[p], 1) == 0) b.redo_cnt[p] += 1;
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_eig_init: previous eigenvectors of a channel into the reduced coordinates of the NEW reduction,
//   x~_b = P (M_bb x_b)   (exact inverse of x_b = P^T x~_b because P M_bb P^T = I),  hats unchanged.
// One CTA per (cell chunk, channel): the cell's M_bb and P are staged once in shared memory and
// reused by all nh vectors (one warp per vector).
#ifndef EIGC_CHUNK
#define EIGC_CHUNK 4
#endif
__host__ __device__ inline size_t eig_chan_smem_bytes(int n, int nb) {
  return sizeof(double) * ((size_t)nb * nb + 4 * (size_t)n * n + 8);
}

#ifndef EIGI_NT
#define EIGI_NT 128   // threads per CTA of k_eig_init (one warp per vector of the window: rarely more than four)
#endif
__global__ void __launch_bounds__(EIGI_NT)
k_eig_init(DiscDev d, BatchDev b, int mode) {
  const int p = blockIdx.z, ch = blockIdx.y;
  if (b.warm[p] == 0 || !(b.stop[p] < 0.3)) return;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p]) return;
  extern __shared__ double sm[];
  const int C = d.C, nb = d.nb, n = d.n, Nh = d.Nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Ps = sm;              // [nb][nb]  P^T (Ps[j*nb+i] = P[i][j])
  double* Ms = Ps + nb * nb;    // [nb][nb]  M_bb
  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  int klo, nhp;                                  // only the levels of this pass are warm-started
  if (!eig_window(b, p, chan, mode, klo, nhp)) return;
  const size_t xs = (size_t)(nb * C + C);
  for (int c = blockIdx.x * EIGC_CHUNK; c < min(C, (blockIdx.x + 1) * EIGC_CHUNK); ++c) {
    __syncthreads();
    const double* Ptc = b.Pt + (chan * C + c) * (size_t)nb * nb;
    const double* Mc = d.M0_loc + (size_t)c * n * n;
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
      Ps[idx] = Ptc[idx];
      const int r = idx / nb, cc = idx - r * nb;
      Ms[idx] = Mc[(2 + r) * n + 2 + cc];
    }
    __syncthreads();
    const int g0 = d.l2g[c * n + 2];
    for (int k = klo + warp; k < nhp; k += EIGI_NT / 32) {
      const double* u = b.U + (chan * b.nhmax + k) * Nh;
      const double xb = (lane < nb) ? u[g0 + lane] : 0.0;
      double mb = 0.0;
      for (int j = 0; j < nb; ++j) {
        const double xj = __shfl_sync(0xffffffffu, xb, j);
        if (lane < nb) mb = fma(Ms[j * nb + lane], xj, mb);
      }
      double xt = 0.0;
      for (int j = 0; j < nb; ++j) {
        const double mj = __shfl_sync(0xffffffffu, mb, j);
        if (lane < nb) xt = fma(Ps[j * nb + lane], mj, xt);
      }
      if (lane < nb) b.Xt[(chan * b.nhmax + k) * xs + (size_t)lane * C + c] = xt;
    }
  }
  if (blockIdx.x == 0)
    for (int idx = threadIdx.x; idx < (nhp - klo) * (C - 1); idx += blockDim.x) {
      const int k = klo + idx / (C - 1), j = idx - (k - klo) * (C - 1);
      b.Xt[(chan * b.nhmax + k) * xs + (size_t)nb * C + j] = b.U[(chan * b.nhmax + k) * Nh + d.l2g[j * n]];
    }
}

// ------------------------------------------------------------------------------------------
// k_eig_final: one warp per (eigenvector, cell): x_b = P^T x~_b written to U (unnormalised), and the
// cell's share of x^T M0 x, x^T H x (ORIGINAL blocks, unassembled sums) and of the sign sum.
// k_eig_scale then adds the shares in a fixed order: U = x / sqrt(x^T M0 x) with a fixed sign,
// eps = x^T H x / x^T M0 x.  The kinetic and Coulomb blocks ride along (same staging, two more FMAs per
// entry), which gives the per-orbital energies of energies.jl:47-87 without another pass over U.
#ifndef EIGF_NT
#define EIGF_NT 128   // threads per CTA of k_eig_final: nv x EIGC_CHUNK warp items per CTA -- with 4 warps no warp idles for any nv
#endif
__global__ void __launch_bounds__(EIGF_NT)
k_eig_final(DiscDev d, BatchDev b, int mode) {
  const int p = blockIdx.z, ch = blockIdx.y, chunk = blockIdx.x;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p]) return;
  const int C = d.C, nb = d.nb, n = d.n, Nh = d.Nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  int klo, nhp;
  if (!eig_window(b, p, chan, mode, klo, nhp)) return;
  const double ll = (double)(l * (l + 1));
  const size_t xs = (size_t)(nb * C + C);
  const int c0 = chunk * EIGC_CHUNK, nc = min(C, c0 + EIGC_CHUNK) - c0, nv = nhp - klo;
  // (two vectors per warp item -- every loaded block entry feeding both -- was measured SLOWER, 0.142 vs 0.127 ms: the kernel
  // is bound by the latency of its loads, and half as many items means half as many warps hiding it)
  // one warp per (vector, cell) item; no staging and no CTA barrier: the blocks are read straight
  // through L1 (rows are coalesced, and the warps of a CTA share the cells of its chunk)
  for (int item = warp; item < nv * nc; item += EIGF_NT / 32) {
    const int k = klo + item / nc, c = c0 + item % nc;
    const double* __restrict__ Pc = b.Pm + (chan * C + c) * (size_t)nb * nb;
    const double* __restrict__ Mc = d.M0_loc + (size_t)c * n * n;
    const double* __restrict__ Hc = b.Hfull + (chan * C + c) * (size_t)n * n;
    const double* __restrict__ Ac = d.A_loc + (size_t)c * n * n;
    const double* __restrict__ M2c = d.Mm2_loc + (size_t)c * n * n;
    const double* __restrict__ M1c = d.Mm1_loc + (size_t)c * n * n;
    const double* X = b.Xt + (chan * b.nhmax + k) * xs;
    double* u = b.U + (chan * b.nhmax + k) * Nh;
    const int g0 = d.l2g[c * n + 2];
    const int gR = (c < C - 1) ? d.l2g[c * n] : -1;
    const double xt = (lane < nb) ? X[(size_t)lane * C + c] : 0.0;
    const bool inb = lane < nb, inn = lane < n;
    double xb = 0.0;   // x_b[lane] = sum_i P[i][lane] x~_i
#pragma unroll 4
    for (int i = 0; i < nb; ++i) {
      const double pv = inb ? Pc[i * nb + lane] : 0.0;
      xb = fma(pv, __shfl_sync(0xffffffffu, xt, i), xb);
    }
    // local vector in generator order: lane 0 right hat, lane 1 left hat, lanes 2.. bubbles
    double xl = __shfl_up_sync(0xffffffffu, xb, 2);
    if (lane == 0) xl = (c < C - 1) ? X[(size_t)nb * C + c] : 0.0;
    if (lane == 1) xl = (c > 0) ? X[(size_t)nb * C + c - 1] : 0.0;
    if (!inn) xl = 0.0;
    double mx = 0.0, hx = 0.0, kx = 0.0, cx = 0.0;
#pragma unroll 3
    for (int a = 0; a < n; ++a) {
      const int o = a * n + lane;
      const double mv = inn ? Mc[o] : 0.0, hv = inn ? Hc[o] : 0.0, cv = inn ? M1c[o] : 0.0;
      const double av = inn ? Ac[o] : 0.0, m2 = (inn && l > 0) ? M2c[o] : 0.0;
      const double xa = __shfl_sync(0xffffffffu, xl, a);
      mx = fma(mv, xa, mx); hx = fma(hv, xa, hx); cx = fma(cv, xa, cx);
      kx = fma((av + ll * m2) / 2.0, xa, kx);
    }
    if (lane >= 2 && inn) u[g0 + lane - 2] = xl;
    if (lane == 0 && gR >= 0) u[gR] = xl;
    const double own = ((lane >= 2 && inn) || (lane == 0 && gR >= 0)) ? xl : 0.0;   // entries this cell owns
    const double m = warp_sum(mx * xl), h = warp_sum(hx * xl), sg = warp_sum(own);
    const double kk = warp_sum(kx * xl), cc = warp_sum(cx * xl);
    if (lane == 0) {
      double* part = b.eig_part + ((chan * b.nhmax + k) * C + c) * 6;
      part[0] = m; part[1] = h; part[2] = sg; part[3] = kk; part[4] = cc;
    }
  }
}

__global__ void __launch_bounds__(AKS_NT)
k_eig_scale(DiscDev d, BatchDev b, int nchunk, int mode) {
  const int p = blockIdx.y, ch = blockIdx.x;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p]) return;
  const int Nh = d.Nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  int klo, nhp;
  if (!eig_window(b, p, chan, mode, klo, nhp)) return;
  __syncthreads();   // every thread has read kdone before it is advanced
  if (threadIdx.x == 0) b.kdone[chan] = nhp;
  for (int k = klo + warp; k < nhp; k += AKS_NW) {
    const double* part = b.eig_part + ((chan * b.nhmax + k) * nchunk) * 6;
    double m = 0.0, h = 0.0, sg = 0.0, kk = 0.0, cc = 0.0;
    for (int q = 0; q < nchunk; ++q) {
      m += part[6 * q]; h += part[6 * q + 1]; sg += part[6 * q + 2]; kk += part[6 * q + 3]; cc += part[6 * q + 4];
    }
    double* u = b.U + (chan * b.nhmax + k) * Nh;
    const double sc = ((sg < 0.0) ? -1.0 : 1.0) / sqrt(m);
    for (int i = lane; i < Nh; i += 32) u[i] *= sc;
    if (lane == 0) {
      const double rq = h / m;
      const int info = b.eig_info[chan * b.nhmax + k];
      if (isfinite(rq) && !(info & 0x80)) b.eps_new[chan * b.nhmax + k] = rq;
      else if (mode != 2) b.eig_fail[p] = 1;   // the aufbau would run on an uncertified level: k_finalize raises AKS_EEIG
      if (mode == 2) b.eps[chan * b.nhmax + k] = b.eps_new[chan * b.nhmax + k];   // no aufbau follows
      // u^T Kin_l u and u^T Coulomb u of the normalised orbital (compute_kinetic_energy /
      // compute_coulomb_energy, src/discretization/energies.jl:47-87)
      b.ekin_orb[chan * b.nhmax + k] = kk / m;
      b.ecou_orb[chan * b.nhmax + k] = -b.Z[p] * (cc / m);
    }
  }
  // ---- clusters (k_eig: info bit 0x40): neighbouring levels of this channel whose certificates overlap.  Each was
  // converged independently; make the pair M0-orthonormal (modified Gram-Schmidt, lower level first) and refuse a
  // pair that converged on the same vector (the other member of the cluster would be missing from the density).
  __shared__ double cl_red[AKS_NW];
  __shared__ int cl_any;
  if (threadIdx.x == 0) {
    int any = 0;
    for (int k = klo; k < nhp; ++k) any |= (b.eig_info[chan * b.nhmax + k] & 0x40);
    cl_any = any;
  }
  __syncthreads();
  if (!cl_any) return;
  const int n = d.n, C = d.C;
  for (int k = max(klo, 1); k < nhp; ++k) {
    __syncthreads();   // u and the info words as the previous pair left them; every branch below is CTA-uniform
    const int i0 = b.eig_info[chan * b.nhmax + k - 1], i1 = b.eig_info[chan * b.nhmax + k];
    if (!((i0 | i1) & 0x40)) continue;
    const double e0 = b.eps_new[chan * b.nhmax + k - 1], e1 = b.eps_new[chan * b.nhmax + k];
    if (!(fabs(e1 - e0) <= 4e-8 * fmax(fmax(fabs(e0), fabs(e1)), 1e-5))) continue;
    double* u0 = b.U + (chan * b.nhmax + k - 1) * Nh;
    double* u1 = b.U + (chan * b.nhmax + k) * Nh;
    double part = 0.0;
    for (int c = warp; c < C; c += AKS_NW) {
      const double* Mc = d.M0_loc + (size_t)c * n * n;
      const int gi = (lane < n) ? d.l2g[c * n + lane] : -1;
      const double a = (gi >= 0) ? u0[gi] : 0.0, bb = (gi >= 0) ? u1[gi] : 0.0;
      double t = 0.0;
      for (int j = 0; j < n; ++j) {
        const double bj = __shfl_sync(0xffffffffu, bb, j);
        if (lane < n) t = fma(Mc[lane * n + j], bj, t);
      }
      part += warp_sum(a * t);
    }
    if (lane == 0) cl_red[warp] = part;
    __syncthreads();
    double ov = 0.0;
    for (int w = 0; w < AKS_NW; ++w) ov += cl_red[w];
    if (fabs(ov) > 0.98) {   // the same eigenvector twice: a genuine failure
      if (threadIdx.x == 0 && mode != 2) b.eig_fail[p] = 1;
      if (threadIdx.x == 0) b.eig_info[chan * b.nhmax + k] |= 0x80;
      continue;
    }
    const double sc2 = 1.0 / sqrt(1.0 - ov * ov);
    for (int i = threadIdx.x; i < Nh; i += blockDim.x) u1[i] = (u1[i] - ov * u0[i]) * sc2;
  }
}
