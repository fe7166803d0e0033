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

// per-cell reduced data, field-major in shared memory ([field][cell]); in global [cell][field]
//   td[nb] te[nb] ht0[nb] ht1[nb] mt0[nb] mt1[nb] Hhh[3] pad[1]      (te[nb-1] unused)
__host__ __device__ inline int red_stride(int nb) { return 6 * nb + 4; }

// ------------------------------------------------------------------------------------------
// k_reduce: one warp per (problem, channel, cell)
__host__ __device__ inline size_t reduce_smem_bytes(int nb) {
  const int ld = nb | 1;
  return sizeof(double) * (size_t)AKS_NW * (2 * nb * ld + 8);
}

__global__ void __launch_bounds__(AKS_NT)
k_reduce(DiscDev d, BatchDev b) {
  const int p = blockIdx.z, ch = blockIdx.y;
  if (b.status[p] != 0) return;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * AKS_NW + warp;
  const int nb = d.nb, npk = d.npk, C = d.C, ld = nb | 1;
  if (c >= C) return;
  extern __shared__ double sm[];
  double* Cs = sm + (size_t)warp * (2 * nb * ld + 8);
  double* Ps = Cs + nb * ld;
  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  const double* Hc = b.Hpk + (chan * C + c) * npk;
  const double* Mc = d.M0pk + (size_t)c * npk;
  const double* Li = d.Linv + (size_t)c * nb * nb;
  const int off = 3 + 2 * nb;
  const int i = lane;
  // 1. H_bb, full symmetric
  for (int idx = lane; idx < nb * nb; idx += 32) {
    const int r = idx / nb, cc = idx - r * nb;
    const int hi = r >= cc ? r : cc, lo = r >= cc ? cc : r;
    Cs[r * ld + cc] = Hc[off + hi * (hi + 1) / 2 + lo];
  }
  __syncwarp();
  // 2. A = Linv * H  (row i per lane) -> Ps
  if (i < nb)
    for (int j = 0; j < nb; ++j) {
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc = fma(Li[i * nb + k], Cs[k * ld + j], acc);
      Ps[i * ld + j] = acc;
    }
  __syncwarp();
  // 3. C = A * Linv^T (lower triangle computed, mirrored)
  if (i < nb)
    for (int j = 0; j <= i; ++j) {
      double acc = 0.0;
      for (int k = 0; k <= j; ++k) acc = fma(Ps[i * ld + k], Li[j * nb + k], acc);
      Cs[i * ld + j] = acc;
    }
  __syncwarp();
  if (i < nb)
    for (int j = 0; j < i; ++j) Cs[j * ld + i] = Cs[i * ld + j];
  // 4. P = Linv
  if (i < nb)
    for (int j = 0; j < nb; ++j) Ps[i * ld + j] = (j <= i) ? Li[i * nb + j] : 0.0;
  __syncwarp();
  double* red = b.red + (chan * C + c) * red_stride(nb);
  // 5. Householder tridiagonalisation, P <- H_k P
  for (int k = 0; k + 2 < nb; ++k) {
    const double x = (i > k && i < nb) ? Cs[i * ld + k] : 0.0;
    const double sig2 = warp_sum((i > k + 1) ? x * x : 0.0);
    const double xk1 = __shfl_sync(0xffffffffu, x, k + 1);
    if (lane == 0) red[k] = Cs[k * ld + k];
    if (sig2 == 0.0) {  // already tridiagonal in this column (warp-uniform)
      if (lane == 0) red[nb + k] = xk1;
      continue;
    }
    const double alpha = -copysign(sqrt(xk1 * xk1 + sig2), xk1);
    const double v = (i == k + 1) ? xk1 - alpha : ((i > k + 1) ? x : 0.0);
    const double vk1 = xk1 - alpha;
    const double tau = 2.0 / (sig2 + vk1 * vk1);
    if (lane == 0) red[nb + k] = alpha;
    double pacc = 0.0;
    for (int j = k + 1; j < nb; ++j) {
      const double vj = __shfl_sync(0xffffffffu, v, j);
      if (i > k && i < nb) pacc = fma(Cs[i * ld + j], vj, pacc);
    }
    pacc *= tau;
    const double K = 0.5 * tau * warp_sum(pacc * v);
    const double w = pacc - K * v;
    for (int j = k + 1; j < nb; ++j) {
      const double vj = __shfl_sync(0xffffffffu, v, j), wj = __shfl_sync(0xffffffffu, w, j);
      if (i > k && i < nb) Cs[i * ld + j] -= v * wj + w * vj;
    }
    // P <- (I - tau v v^T) P, lane = column
    double t = 0.0;
    for (int m = k + 1; m < nb; ++m) {
      const double vm = __shfl_sync(0xffffffffu, v, m);
      if (i < nb) t = fma(vm, Ps[m * ld + i], t);
    }
    t *= tau;
    for (int m = k + 1; m < nb; ++m) {
      const double vm = __shfl_sync(0xffffffffu, v, m);
      if (i < nb) Ps[m * ld + i] -= vm * t;
    }
    __syncwarp();
  }
  __syncwarp();
  if (lane == 0) {
    if (nb >= 2) {
      red[nb - 2] = Cs[(nb - 2) * ld + nb - 2];
      red[nb + nb - 2] = Cs[(nb - 1) * ld + nb - 2];
    }
    red[nb - 1] = Cs[(nb - 1) * ld + nb - 1];
    red[nb + nb - 1] = 0.0;
    red[6 * nb + 0] = Hc[0];
    red[6 * nb + 1] = Hc[1];
    red[6 * nb + 2] = Hc[2];
    red[6 * nb + 3] = 0.0;
  }
  // 6. h~ = P H_bh, m~ = P M_bh
  {
    const double h0 = (i < nb) ? Hc[3 + i] : 0.0, h1 = (i < nb) ? Hc[3 + nb + i] : 0.0;
    const double m0 = (i < nb) ? Mc[3 + i] : 0.0, m1 = (i < nb) ? Mc[3 + nb + i] : 0.0;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int j = 0; j < nb; ++j) {
      const double pj = (i < nb) ? Ps[i * ld + j] : 0.0;
      a0 = fma(pj, __shfl_sync(0xffffffffu, h0, j), a0);
      a1 = fma(pj, __shfl_sync(0xffffffffu, h1, j), a1);
      a2 = fma(pj, __shfl_sync(0xffffffffu, m0, j), a2);
      a3 = fma(pj, __shfl_sync(0xffffffffu, m1, j), a3);
    }
    if (i < nb) { red[2 * nb + i] = a0; red[3 * nb + i] = a1; red[4 * nb + i] = a2; red[5 * nb + i] = a3; }
  }
  // 7. P (row-major) and P^T
  double* Pm = b.Pm + (chan * C + c) * (size_t)nb * nb;
  double* Pt = b.Pt + (chan * C + c) * (size_t)nb * nb;
  for (int idx = lane; idx < nb * nb; idx += 32) {
    const int r = idx / nb, cc = idx - r * nb;
    Pm[idx] = Ps[r * ld + cc];
    Pt[idx] = Ps[cc * ld + r];
  }
}

// ------------------------------------------------------------------------------------------
struct EigShared {
  double* red;      // [red_stride][C]  reduced cell data (field-major)
  double* mhh;      // [3][C]           M hat-hat blocks
  double* fac;      // [4 nb][C]        dinv[nb] | l[nb] | W0[nb] | W1[nb]
  double* zt;       // [nb][C]          transformed bubble vectors
  double* vx;       // [Nh]
  double* vr;       // [Nh]  r = M x
  double* vy;       // [Nh]
  double* hatpart;  // [2C]
  double* sc;       // [3C]  Schur blocks of the current factorisation
  double* tri_di;   // [C]   reciprocal pivots of the hat system
  double* tri_l;    // [C]
  double* xh;       // [C]
  double* red8;     // [NW + 8]
  int* negc;        // [C + 2]
  int* hatg;        // [C]   global index of hat j (right hat of cell j)
  int* bub0;        // [C]   global index of the first bubble of cell c (bubbles are contiguous)
};

__host__ __device__ inline size_t eig_smem_bytes(int C, int nb, int Nh) {
  size_t nd = (size_t)red_stride(nb) * C + 3 * C + (size_t)4 * nb * C + (size_t)nb * C + 3 * (size_t)Nh +
              2 * C + 3 * C + 3 * C + AKS_NW + 8;
  return nd * sizeof(double) + (size_t)(3 * C + 8) * sizeof(int);
}

// ---- factor T(sigma): thread c handles cell c; returns the inertia count
__device__ __forceinline__ int eig_factor(const EigShared& S, int C, int nb, double sigma, int store) {
  const int c = threadIdx.x;
  if (c < C) {
    const double* R = S.red;
    double* F = S.fac;
    const double pivmin0 = DBL_MIN * 1e8;
    double scale = 0.0;
    int neg = 0;
    double dprev_inv = 0.0, y0p = 0.0, y1p = 0.0, s00 = 0.0, s10 = 0.0, s11 = 0.0;
    for (int j = 0; j < nb; ++j) {
      const double tdj = R[j * C + c] - sigma;
      scale = fmax(scale, fabs(tdj));
      double lj = 0.0, dj = tdj;
      if (j > 0) {
        const double te = R[(nb + j - 1) * C + c];
        lj = te * dprev_inv;
        dj = tdj - lj * te;
      }
      const double pivmin = fmax(scale * 4.9e-32, pivmin0);
      if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
      neg += (dj < 0.0);
      const double dinv = 1.0 / dj;
      const double g0 = R[(2 * nb + j) * C + c] - sigma * R[(4 * nb + j) * C + c];
      const double g1 = R[(3 * nb + j) * C + c] - sigma * R[(5 * nb + j) * C + c];
      const double y0 = g0 - lj * y0p, y1 = g1 - lj * y1p;
      s00 = fma(y0 * dinv, y0, s00);
      s10 = fma(y1 * dinv, y0, s10);
      s11 = fma(y1 * dinv, y1, s11);
      if (store) {
        F[j * C + c] = dinv;
        F[(nb + j) * C + c] = lj;
        F[(2 * nb + j) * C + c] = y0 * dinv;
        F[(3 * nb + j) * C + c] = y1 * dinv;
      }
      dprev_inv = dinv; y0p = y0; y1p = y1;
    }
    if (store) {  // W = L^-T D^-1 y
      double w0n = 0.0, w1n = 0.0;
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
    S.sc[3 * c] = (R[(6 * nb) * C + c] - sigma * S.mhh[c]) - s00;
    S.sc[3 * c + 1] = (R[(6 * nb + 1) * C + c] - sigma * S.mhh[C + c]) - s10;
    S.sc[3 * c + 2] = (R[(6 * nb + 2) * C + c] - sigma * S.mhh[2 * C + c]) - s11;
    S.negc[c] = neg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int neg = 0;
    for (int cc = 0; cc < C; ++cc) neg += S.negc[cc];
    // hat node j sits between cell j (its right hat, g=0) and cell j+1 (its left hat, g=1)
    double dinvp = 0.0, offprev = 0.0, scale = 0.0;
    for (int j = 0; j < C - 1; ++j) {
      const double diag = S.sc[3 * j] + S.sc[3 * (j + 1) + 2];
      scale = fmax(scale, fabs(diag));
      double l = 0.0, dj = diag;
      if (j > 0) { l = offprev * dinvp; dj = diag - l * offprev; }
      const double pivmin = fmax(scale * 4.9e-32, DBL_MIN * 1e8);
      if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
      neg += (dj < 0.0);
      dinvp = 1.0 / dj;
      S.tri_di[j] = dinvp;
      if (j > 0) S.tri_l[j - 1] = l;
      offprev = (j + 1 < C - 1) ? S.sc[3 * (j + 1) + 1] : 0.0;
    }
    S.negc[C] = neg;
  }
  __syncthreads();
  return S.negc[C];
}

// ---- y = T(sigma)^{-1} r with the stored factors (r in S.vr, y to S.vy)
__device__ __forceinline__ void eig_solve(const EigShared& S, int C, int nb, const double* __restrict__ Pm,
                                          const double* __restrict__ Pt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // A. r~ = P r_b  (warp per cell; P^T storage makes lane i's reads contiguous)
  for (int c = warp; c < C; c += AKS_NW) {
    const double rb = (lane < nb) ? S.vr[S.bub0[c] + lane] : 0.0;
    const double* Ptc = Pt + (size_t)c * nb * nb;
    double acc = 0.0;
    for (int j = 0; j < nb; ++j) {
      const double rj = __shfl_sync(0xffffffffu, rb, j);
      if (lane < nb) acc = fma(Ptc[j * nb + lane], rj, acc);
    }
    if (lane < nb) S.zt[lane * C + c] = acc;
  }
  __syncthreads();
  // B. thread per cell: z = (T - sI)^-1 r~ ; hat partials W^T r~
  {
    const int c = threadIdx.x;
    if (c < C) {
      const double* F = S.fac;
      double t0 = 0.0, t1 = 0.0, zp = 0.0;
      for (int j = 0; j < nb; ++j) {
        const double rj = S.zt[j * C + c];
        t0 = fma(F[(2 * nb + j) * C + c], rj, t0);
        t1 = fma(F[(3 * nb + j) * C + c], rj, t1);
        const double z = rj - F[(nb + j) * C + c] * zp;
        S.zt[j * C + c] = z;
        zp = z;
      }
      double zn = 0.0;
      for (int j = nb - 1; j >= 0; --j) {
        double z = S.zt[j * C + c] * F[j * C + c];
        if (j < nb - 1) z -= F[(nb + j + 1) * C + c] * zn;
        S.zt[j * C + c] = z;
        zn = z;
      }
      S.hatpart[2 * c] = t0;
      S.hatpart[2 * c + 1] = t1;
    }
  }
  __syncthreads();
  // C. hat system
  if (threadIdx.x == 0) {
    const int m = C - 1;
    double zprev = 0.0;
    for (int j = 0; j < m; ++j) {
      double rh = S.vr[S.hatg[j]] - S.hatpart[2 * j] - S.hatpart[2 * (j + 1) + 1];
      if (j > 0) rh -= S.tri_l[j - 1] * zprev;
      S.xh[j] = rh;
      zprev = rh;
    }
    double xnext = 0.0;
    for (int j = m - 1; j >= 0; --j) {
      double v = S.xh[j] * S.tri_di[j];
      if (j < m - 1) v -= S.tri_l[j] * xnext;
      S.xh[j] = v;
      xnext = v;
      S.vy[S.hatg[j]] = v;
    }
  }
  __syncthreads();
  // D+E. y~ = z - W y_h ;  y_b = P^T y~   (warp per cell, lane = column of P)
  for (int c = warp; c < C; c += AKS_NW) {
    const double xh0 = (c < C - 1) ? S.xh[c] : 0.0;
    const double xh1 = (c > 0) ? S.xh[c - 1] : 0.0;
    double yt = 0.0;
    if (lane < nb)
      yt = S.zt[lane * C + c] - S.fac[(2 * nb + lane) * C + c] * xh0 - S.fac[(3 * nb + lane) * C + c] * xh1;
    const double* Pc = Pm + (size_t)c * nb * nb;
    double acc = 0.0;
    for (int i = 0; i < nb; ++i) {
      const double yi = __shfl_sync(0xffffffffu, yt, i);
      if (lane < nb) acc = fma(Pc[i * nb + lane], yi, acc);
    }
    if (lane < nb) S.vy[S.bub0[c] + lane] = acc;
  }
  __syncthreads();
}

// ---- out = (unassembled SYMMETRIC operator [C][n][n]) * in   (shared vectors, global numbering)
__device__ __noinline__ void eig_apply(int n, int C, const int* hatg, const int* bub0, double* hatpart,
                                       const double* __restrict__ mat, const double* in, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += AKS_NW) {
    int gi = -1;
    if (lane >= 2 && lane < n) gi = bub0[c] + lane - 2;
    else if (lane == 0 && c < C - 1) gi = hatg[c];
    else if (lane == 1 && c > 0) gi = hatg[c - 1];
    const double xl = (gi >= 0) ? in[gi] : 0.0;
    const double* Mc = mat + (size_t)c * n * n;
    double r = 0.0;
    for (int a = 0; a < n; ++a) {
      const double xa = __shfl_sync(0xffffffffu, xl, a);
      if (lane < n) r = fma(Mc[a * n + lane], xa, r);   // symmetric: column access is coalesced
    }
    if (lane >= 2 && gi >= 0) out[gi] = r;
    if (lane < 2) hatpart[2 * c + lane] = r;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C - 1; j += blockDim.x)
    out[hatg[j]] = hatpart[2 * j] + hatpart[2 * (j + 1) + 1];
  __syncthreads();
}

// ---- out = H * in with the packed per-channel blocks
__device__ __noinline__ void eig_apply_pk(int n, int C, int nb, int npk, const int* hatg, const int* bub0,
                                          double* hatpart, const double* __restrict__ Hpk, const double* in,
                                          double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += AKS_NW) {
    int gi = -1;
    if (lane >= 2 && lane < n) gi = bub0[c] + lane - 2;
    else if (lane == 0 && c < C - 1) gi = hatg[c];
    else if (lane == 1 && c > 0) gi = hatg[c - 1];
    const double xl = (gi >= 0) ? in[gi] : 0.0;
    const double* Hc = Hpk + (size_t)c * npk;
    double r = 0.0;
    for (int a = 0; a < n; ++a) {
      const double xa = __shfl_sync(0xffffffffu, xl, a);
      if (lane < n) {
        const int hi = (lane >= a) ? lane : a, lo = (lane >= a) ? a : lane;
        r = fma(Hc[pk_index(hi, lo, nb)], xa, r);
      }
    }
    if (lane >= 2 && gi >= 0) out[gi] = r;
    if (lane < 2) hatpart[2 * c + lane] = r;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C - 1; j += blockDim.x)
    out[hatg[j]] = hatpart[2 * j] + hatpart[2 * (j + 1) + 1];
  __syncthreads();
}

__device__ __noinline__ double eig_dot(double* red, const double* a, const double* bvec, int Nh) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) acc = fma(a[i], bvec[i], acc);
  return block_sum(acc, red);
}

__device__ inline void eig_carve(const DiscDev& d, double* sm, EigShared& S) {
  const int C = d.C, Nh = d.Nh, n = d.n, nb = d.nb;
  double* ptr = sm;
  S.red = ptr; ptr += (size_t)red_stride(nb) * C;
  S.mhh = ptr; ptr += 3 * C;
  S.fac = ptr; ptr += (size_t)4 * nb * C;
  S.zt = ptr; ptr += (size_t)nb * C;
  S.vx = ptr; ptr += Nh;
  S.vr = ptr; ptr += Nh;
  S.vy = ptr; ptr += Nh;
  S.hatpart = ptr; ptr += 2 * C;
  S.sc = ptr; ptr += 3 * C;
  S.tri_di = ptr; ptr += C;
  S.tri_l = ptr; ptr += C;
  S.xh = ptr; ptr += C;
  S.red8 = ptr; ptr += AKS_NW + 8;
  S.negc = (int*)ptr;
  S.hatg = S.negc + C + 2;
  S.bub0 = S.hatg + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    S.hatg[c] = (c < C - 1) ? d.l2g[c * n + 0] : 0;
    S.bub0[c] = d.l2g[c * n + 2];
    S.mhh[c] = d.M0pk[(size_t)c * d.npk + 0];
    S.mhh[C + c] = d.M0pk[(size_t)c * d.npk + 1];
    S.mhh[2 * C + c] = d.M0pk[(size_t)c * d.npk + 2];
  }
}

enum { EST_BRACKET_LO = 0, EST_BRACKET_HI, EST_BISECT, EST_RQI, EST_CERT_LO, EST_CERT_HI, EST_REFINE, EST_DONE };

__global__ void __launch_bounds__(AKS_NT)
k_eig(DiscDev d, BatchDev b) {
  const int k = blockIdx.x, ch = blockIdx.y, p = blockIdx.z;
  if (b.status[p] != 0) return;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p] || k >= b.nhp[p]) return;
  extern __shared__ double sm[];
  const int C = d.C, npk = d.npk, Nh = d.Nh, nb = d.nb, n = d.n;
  EigShared S;
  eig_carve(d, sm, S);

  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  const double* Hpk = b.Hpk + chan * C * npk;
  const double* Pm = b.Pm + chan * C * (size_t)nb * nb;
  const double* Pt = b.Pt + chan * C * (size_t)nb * nb;
  const double* eps_ch = b.eps + chan * b.nhmax;
  double* Uout = b.U + (chan * b.nhmax + k) * Nh;
  const int nhp = b.nhp[p];
  const bool have_prev = b.warm[p] != 0;
  {  // reduced data of this channel, transposed to field-major
    const int rs = red_stride(nb);
    const double* rg = b.red + chan * C * rs;
    for (int idx = threadIdx.x; idx < C * rs; idx += blockDim.x) {
      const int c = idx / rs, f = idx - c * rs;
      S.red[f * C + c] = rg[idx];
    }
  }
  __syncthreads();

  // bracket == witnesses: lo = largest shift known with count <= k (clo), hi = smallest with
  // count >= k+1 (chi); -1 = not known yet
  double lo = -DBL_MAX, hi = DBL_MAX, step = 1.0, sigma = 0.0, rho = 0.0, prev_upd = DBL_MAX;
  int clo = -1, chi = -1;
  int state, nfac = 0, its = 0, its_cap = 12, ncount = 0;
  bool from_warm = false, need_start = false, certified = false, converged = false, refined = false;

  auto cold_init = [&]() {
    from_warm = false;
    its = 0; its_cap = 12; prev_upd = DBL_MAX; converged = false;
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

  if (have_prev && b.stop[p] < 0.3) {
    // ---- warm start: previous eigenvector, Rayleigh quotient with the new H as first shift
    for (int i = threadIdx.x; i < Nh; i += blockDim.x) S.vx[i] = Uout[i];
    __syncthreads();
    eig_apply(n, C, S.hatg, S.bub0, S.hatpart, d.M0_loc, S.vx, S.vr);
    const double xmx = eig_dot(S.red8, S.vx, S.vr, Nh);
    const double sc = 1.0 / sqrt(xmx);
    for (int i = threadIdx.x; i < Nh; i += blockDim.x) { S.vx[i] *= sc; S.vr[i] *= sc; }
    __syncthreads();
    eig_apply_pk(n, C, nb, npk, S.hatg, S.bub0, S.hatpart, Hpk, S.vx, S.vy);
    sigma = eig_dot(S.red8, S.vx, S.vy, Nh);
    state = EST_RQI; from_warm = true; its_cap = 6;
    if (!isfinite(sigma)) cold_init();
  } else {
    cold_init();
  }

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
      for (int i = threadIdx.x; i < Nh; i += blockDim.x) S.vx[i] = 1.0 + 0.37 * (double)((i * 7 + 3 * k) % 11) / 11.0;
      __syncthreads();
      eig_apply(n, C, S.hatg, S.bub0, S.hatpart, d.M0_loc, S.vx, S.vr);
      const double xmx = eig_dot(S.red8, S.vx, S.vr, Nh);
      const double sc = 1.0 / sqrt(xmx);
      for (int i = threadIdx.x; i < Nh; i += blockDim.x) { S.vx[i] *= sc; S.vr[i] *= sc; }
      __syncthreads();
      need_start = false;
    }
    if (state == EST_REFINE) {
      // eigenvalue = Rayleigh quotient with the ORIGINAL blocks (x is M-normalised): second-order
      // accurate in the eigenvector, so the reduction's O(eps cond(M_bb)) error does not reach it
      eig_apply_pk(n, C, nb, npk, S.hatg, S.bub0, S.hatpart, Hpk, S.vx, S.vy);
      const double rq = eig_dot(S.red8, S.vx, S.vy, Nh);
      if (isfinite(rq)) rho = rq;
      refined = true;
      state = EST_DONE;
      break;
    }
    const int store = (state == EST_RQI) ? 1 : 0;
    const int cnt = eig_factor(S, C, nb, sigma, store);
    ++nfac;
    if (!store) ++ncount;
    if (state == EST_BRACKET_LO) {
      if (cnt > k) { hi = sigma; chi = cnt; lo = sigma - step; step *= 2.0; sigma = lo; }
      else {
        lo = sigma; clo = cnt;
        if (chi >= 0) { state = narrow() ? EST_RQI : EST_BISECT; sigma = 0.5 * (lo + hi); }
        else { state = EST_BRACKET_HI; sigma = hi; }
      }
    } else if (state == EST_BRACKET_HI) {
      if (cnt < k + 1) { lo = sigma; clo = cnt; hi = sigma + step; step *= 2.0; sigma = hi; }
      else { hi = sigma; chi = cnt; state = narrow() ? EST_RQI : EST_BISECT; sigma = 0.5 * (lo + hi); }
    } else if (state == EST_BISECT) {
      if (cnt <= k) { lo = sigma; clo = cnt; } else { hi = sigma; chi = cnt; }
      const double mid = 0.5 * (lo + hi);
      if (narrow() || !(mid > lo && mid < hi)) state = EST_RQI;
      sigma = mid;
    } else if (state == EST_RQI) {
      witness(sigma, cnt);
      eig_solve(S, C, nb, Pm, Pt);                                          // y = (H - sigma M)^{-1} M x
      const double yr = eig_dot(S.red8, S.vy, S.vr, Nh);                    // y^T M x
      eig_apply(n, C, S.hatg, S.bub0, S.hatpart, d.M0_loc, S.vy, S.vr);     // r <- M y
      const double ymy = eig_dot(S.red8, S.vy, S.vr, Nh);
      rho = sigma + yr / ymy;
      const double sc = 1.0 / sqrt(ymy);
      for (int i = threadIdx.x; i < Nh; i += blockDim.x) { S.vx[i] = S.vy[i] * sc; S.vr[i] *= sc; }
      __syncthreads();
      ++its;
      const double upd = fabs(rho - sigma);
      if (!isfinite(rho)) { fail(); continue; }
      if (upd <= 1e-13 * fabs(rho) || (its >= 3 && upd >= 0.25 * prev_upd && upd <= 1e-9 * fmax(fabs(rho), 1e-6))) {
        converged = true;
        const double dl = fmax(1e-8 * fabs(rho), 1e-13);
        if (!(clo == k && lo < rho)) { state = EST_CERT_LO; sigma = rho - dl; }
        else if (!(chi == k + 1 && hi > rho)) { state = EST_CERT_HI; sigma = rho + dl; }
        else { state = EST_REFINE; certified = true; }
        continue;
      }
      prev_upd = upd;
      if (its >= its_cap) { fail(); continue; }
      if (rho > lo && rho < hi) sigma = rho;
      else if (from_warm) { fail(); continue; }
      else sigma = 0.5 * (lo + hi);
    } else if (state == EST_CERT_LO) {
      if (cnt == k) {
        lo = sigma; clo = k;
        if (!(chi == k + 1 && hi > rho)) { state = EST_CERT_HI; sigma = rho + fmax(1e-8 * fabs(rho), 1e-13); }
        else { state = EST_REFINE; certified = true; }
      } else { witness(sigma, cnt); fail(); }
    } else if (state == EST_CERT_HI) {
      if (cnt == k + 1) { state = EST_REFINE; certified = true; }
      else { witness(sigma, cnt); fail(); }
    }
  }
  // ---- write back: eigenvalue, vector with a fixed sign convention
  double big = 0.0;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) big += S.vx[i];
  big = block_sum(big, S.red8);
  const double sgn = (big < 0.0) ? -1.0 : 1.0;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) Uout[i] = sgn * S.vx[i];
  if (threadIdx.x == 0) {
    b.eps_new[chan * b.nhmax + k] = rho;
    b.eig_info[chan * b.nhmax + k] = (ncount << 8) | ((nfac - ncount) & 0x7f) | ((certified && converged && refined) ? 0 : 0x80);
  }
}
