// k_eig: lowest-nh generalized eigenpairs of every (problem, l, sigma) channel.
//
// Replaces find_orbital! (src/algorithms/oda/loop.jl:65-77): the reference forms the dense
// S*H*S with S = sqrt(inv(M0)) and calls LAPACK syevr, O(Nh^3) per channel.  Here one CTA owns
// one eigenpair (problem, channel, k) and works on the UNASSEMBLED pencil (H_c, M_c), c = cells:
//
//   * static condensation: for a shift s, T_c = H_c - s M_c is split into hats (2) and bubbles
//     (nb = p-1); the bubble block is factored T_bb = L D L^T by one warp per cell (lane = row,
//     packed lower storage in shared memory), X = T_bb^{-1} T_bh and the 2x2 Schur complement
//     follow, and the hats of all cells form a (C-1) tridiagonal system.  Cost ~ C nb^3/3 instead
//     of Nh b^2 for a band factorisation, and the dependency chain is nb + C long instead of Nh.
//   * inertia: by Sylvester/Haynsworth, #{eig < s} = sum_c neg(D_c) + neg(tridiagonal LDL^T).
//     Used to bracket and bisect eigenvalue k until it is isolated (count(lo) = k, count(hi) = k+1).
//   * Rayleigh-quotient iteration inside the isolating interval (cubic convergence; a quotient
//     leaving the interval triggers more bisection), warm-started from the previous SCF iterate.
//
// Every returned pair is therefore certified by two inertia counts; eigenvalues are accurate to
// a few ulp (the dense LAPACK route loses ~eps*||M0^-1 H||, 4e-9 absolute at Nh = 799).
#pragma once
#include <float.h>

#include "aks_device.cuh"
#include "aks_types.cuh"

struct EigShared {
  double* fac;      // [C][npk] factor storage (shared or global)
  double* vx;       // [Nh]
  double* vr;       // [Nh]  r = M x
  double* vy;       // [Nh]
  double* hatpart;  // [2C]
  double* tri_d;    // [C]
  double* tri_l;    // [C]
  double* xh;       // [C]
  double* red;      // [NW + 8]
  int* negc;        // [C + 2]
};

__host__ __device__ inline size_t eig_smem_bytes(int C, int npk, int Nh, bool fac_in_smem) {
  size_t nd = (size_t)3 * Nh + 2 * C + 3 * C + AKS_NW + 8;
  if (fac_in_smem) nd += (size_t)C * npk;
  return nd * sizeof(double) + (size_t)(C + 4) * sizeof(int);
}

// ---- factor T(sigma) = H - sigma M cell by cell; returns #negative pivots (inertia)
__device__ int eig_factor(const DiscDev& d, const EigShared& S, const double* __restrict__ Hpk,
                          double sigma) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = d.nb, npk = d.npk, C = d.C;
  for (int c = warp; c < C; c += AKS_NW) {
    double* f = S.fac + (size_t)c * npk;
    const double* Hc = Hpk + (size_t)c * npk;
    const double* Mc = d.M0pk + (size_t)c * npk;
    for (int idx = lane; idx < npk; idx += 32) f[idx] = Hc[idx] - sigma * Mc[idx];
    __syncwarp();
    double* bb = f + 3 + 2 * nb;
    const int i = lane;
    const int rowi = i * (i + 1) / 2;
    // pivot floor relative to the block's diagonal scale
    double dmax = (i < nb) ? fabs(bb[rowi + i]) : 0.0;
    dmax = warp_max(dmax);
    const double pivmin = fmax(dmax * 4.9e-32, DBL_MIN * 1e8);
    int neg = 0;
    for (int j = 0; j < nb; ++j) {
      const int rowj = j * (j + 1) / 2;
      double dj = bb[rowj + j];
      if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
      neg += (dj < 0.0);
      const double inv = 1.0 / dj;
      double lij = 0.0;
      const bool act = (i > j) && (i < nb);
      if (act) {
        lij = bb[rowi + j] * inv;
        for (int k = j + 1; k <= i; ++k) bb[rowi + k] = fma(-lij, bb[k * (k + 1) / 2 + j], bb[rowi + k]);
      }
      __syncwarp();
      if (act) bb[rowi + j] = lij;
      if (lane == 0) bb[rowj + j] = dj;
      __syncwarp();
    }
    // X = T_bb^{-1} T_bh for both hats, Schur complement of the hats
    double h0 = 0.0, h1 = 0.0, di = 1.0;
    if (i < nb) { h0 = f[3 + i]; h1 = f[3 + nb + i]; di = bb[rowi + i]; }
    double v0 = h0, v1 = h1;
    for (int j = 0; j < nb; ++j) {
      const double a0 = __shfl_sync(0xffffffffu, v0, j), a1 = __shfl_sync(0xffffffffu, v1, j);
      if (i > j && i < nb) { const double l = bb[rowi + j]; v0 = fma(-l, a0, v0); v1 = fma(-l, a1, v1); }
    }
    v0 /= di; v1 /= di;
    for (int j = nb - 1; j >= 0; --j) {
      const double a0 = __shfl_sync(0xffffffffu, v0, j), a1 = __shfl_sync(0xffffffffu, v1, j);
      if (i < j) { const double l = bb[j * (j + 1) / 2 + i]; v0 = fma(-l, a0, v0); v1 = fma(-l, a1, v1); }
    }
    const double s00 = warp_sum(h0 * v0), s10 = warp_sum(h1 * v0), s11 = warp_sum(h1 * v1);
    __syncwarp();
    if (i < nb) { f[3 + i] = v0; f[3 + nb + i] = v1; }
    if (lane == 0) { f[0] -= s00; f[1] -= s10; f[2] -= s11; S.negc[c] = neg; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int neg = 0;
    for (int c = 0; c < C; ++c) neg += S.negc[c];
    // hat node j sits between cell j (its right hat, g=0) and cell j+1 (its left hat, g=1)
    double dprev = 0.0, offprev = 0.0, scale = 0.0;
    for (int j = 0; j < C - 1; ++j) {
      const double diag = S.fac[(size_t)j * npk + 0] + S.fac[(size_t)(j + 1) * npk + 2];
      scale = fmax(scale, fabs(diag));
      double l = 0.0, dj = diag;
      if (j > 0) { l = offprev / dprev; dj = diag - l * offprev; }
      const double pivmin = fmax(scale * 4.9e-32, DBL_MIN * 1e8);
      if (fabs(dj) < pivmin) dj = (dj < 0.0) ? -pivmin : pivmin;
      neg += (dj < 0.0);
      S.tri_d[j] = dj;
      if (j > 0) S.tri_l[j - 1] = l;
      dprev = dj;
      offprev = (j + 1 < C - 1) ? S.fac[(size_t)(j + 1) * npk + 1] : 0.0;
    }
    S.negc[C] = neg;
  }
  __syncthreads();
  return S.negc[C];
}

// ---- y = T(sigma)^{-1} r with the stored factors (r in S.vr, y to S.vy)
__device__ void eig_solve(const DiscDev& d, const EigShared& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = d.nb, npk = d.npk, C = d.C, n = d.n;
  const int i = lane;
  for (int c = warp; c < C; c += AKS_NW) {
    const double* f = S.fac + (size_t)c * npk;
    double rb = 0.0, x0 = 0.0, x1 = 0.0;
    if (i < nb) { rb = S.vr[d.l2g[c * n + 2 + i]]; x0 = f[3 + i]; x1 = f[3 + nb + i]; }
    const double t0 = warp_sum(x0 * rb), t1 = warp_sum(x1 * rb);
    if (lane == 0) { S.hatpart[2 * c] = t0; S.hatpart[2 * c + 1] = t1; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int m = C - 1;
    // forward
    double zprev = 0.0;
    for (int j = 0; j < m; ++j) {
      double rh = S.vr[d.l2g[j * n + 0]] - S.hatpart[2 * j] - S.hatpart[2 * (j + 1) + 1];
      if (j > 0) rh -= S.tri_l[j - 1] * zprev;
      S.xh[j] = rh;
      zprev = rh;
    }
    // diagonal + backward
    double xnext = 0.0;
    for (int j = m - 1; j >= 0; --j) {
      double v = S.xh[j] / S.tri_d[j];
      if (j < m - 1) v -= S.tri_l[j] * xnext;
      S.xh[j] = v;
      xnext = v;
      S.vy[d.l2g[j * n + 0]] = v;
    }
  }
  __syncthreads();
  for (int c = warp; c < C; c += AKS_NW) {
    const double* f = S.fac + (size_t)c * npk;
    const double* bb = f + 3 + 2 * nb;
    const int rowi = i * (i + 1) / 2;
    double v = 0.0, x0 = 0.0, x1 = 0.0, di = 1.0;
    int gi = -1;
    if (i < nb) { gi = d.l2g[c * n + 2 + i]; v = S.vr[gi]; x0 = f[3 + i]; x1 = f[3 + nb + i]; di = bb[rowi + i]; }
    for (int j = 0; j < nb; ++j) {
      const double a = __shfl_sync(0xffffffffu, v, j);
      if (i > j && i < nb) v = fma(-bb[rowi + j], a, v);
    }
    v /= di;
    for (int j = nb - 1; j >= 0; --j) {
      const double a = __shfl_sync(0xffffffffu, v, j);
      if (i < j) v = fma(-bb[j * (j + 1) / 2 + i], a, v);
    }
    const double xh0 = (c < C - 1) ? S.xh[c] : 0.0;
    const double xh1 = (c > 0) ? S.xh[c - 1] : 0.0;
    if (i < nb) S.vy[gi] = v - x0 * xh0 - x1 * xh1;
  }
  __syncthreads();
}

// ---- out = Mloc-assembled matrix times in  (in/out: shared vectors, global numbering).
// `mat` is an unassembled operator [C][n][n]; optional second operator with weight.
__device__ void eig_apply(const DiscDev& d, const EigShared& S, const double* __restrict__ mat,
                          const double* in, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = d.n, C = d.C;
  for (int c = warp; c < C; c += AKS_NW) {
    const int gi = (lane < n) ? d.l2g[c * n + lane] : -1;
    const double xl = (gi >= 0) ? in[gi] : 0.0;
    const double* Mc = mat + (size_t)c * n * n + (size_t)lane * n;
    double r = 0.0;
    for (int a = 0; a < n; ++a) {
      const double xa = __shfl_sync(0xffffffffu, xl, a);
      if (lane < n) r = fma(Mc[a], xa, r);
    }
    if (lane >= 2 && gi >= 0) out[gi] = r;
    if (lane < 2) S.hatpart[2 * c + lane] = r;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C - 1; j += blockDim.x)
    out[d.l2g[j * n + 0]] = S.hatpart[2 * j] + S.hatpart[2 * (j + 1) + 1];
  __syncthreads();
}

// x^T H x with the packed per-channel blocks (for the warm-start Rayleigh quotient)
__device__ double eig_quadform_pk(const DiscDev& d, const EigShared& S, const double* __restrict__ Hpk,
                                  const double* x) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = d.n, C = d.C, nb = d.nb, npk = d.npk;
  double acc = 0.0;
  for (int c = warp; c < C; c += AKS_NW) {
    const int gi = (lane < n) ? d.l2g[c * n + lane] : -1;
    const double xl = (gi >= 0) ? x[gi] : 0.0;
    const double* Hc = Hpk + (size_t)c * npk;
    double r = 0.0;
    for (int a = 0; a < n; ++a) {
      const double xa = __shfl_sync(0xffffffffu, xl, a);
      if (lane < n) {
        const int hi = (lane >= a) ? lane : a, lo = (lane >= a) ? a : lane;
        r = fma(Hc[pk_index(hi, lo, nb)], xa, r);
      }
    }
    acc += r * xl;
  }
  return block_sum(acc, S.red);
}

__device__ double eig_dot(const EigShared& S, const double* a, const double* bvec, int Nh) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) acc = fma(a[i], bvec[i], acc);
  return block_sum(acc, S.red);
}

__global__ void __launch_bounds__(AKS_NT)
k_eig(DiscDev d, BatchDev b, double* __restrict__ fac_global, int fac_in_smem) {
  const int k = blockIdx.x, ch = blockIdx.y, p = blockIdx.z;
  if (b.status[p] != 0) return;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p] || k >= b.nhp[p]) return;
  extern __shared__ double sm[];
  const int C = d.C, npk = d.npk, Nh = d.Nh;
  EigShared S;
  double* ptr = sm;
  if (fac_in_smem) { S.fac = ptr; ptr += (size_t)C * npk; }
  else S.fac = fac_global + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * C * npk;
  S.vx = ptr; ptr += Nh;
  S.vr = ptr; ptr += Nh;
  S.vy = ptr; ptr += Nh;
  S.hatpart = ptr; ptr += 2 * C;
  S.tri_d = ptr; ptr += C;
  S.tri_l = ptr; ptr += C;
  S.xh = ptr; ptr += C;
  S.red = ptr; ptr += AKS_NW + 8;
  S.negc = (int*)ptr;

  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  const double* Hpk = b.Hpk + chan * C * npk;
  double* eps_ch = b.eps + chan * b.nhmax;
  double* Uout = b.U + (chan * b.nhmax + k) * Nh;
  const bool warm = b.warm[p] != 0;
  const int nhp = b.nhp[p];

  // ---- initial bracket
  double lo, hi;
  double eprev = 0.0;
  if (warm) {
    eprev = eps_ch[k];
    double gap = 1e300;
    if (k > 0) gap = fmin(gap, eprev - eps_ch[k - 1]);
    if (k + 1 < nhp) gap = fmin(gap, eps_ch[k + 1] - eprev);
    if (!(gap < 1e299)) gap = fmax(fabs(eprev), 1.0);
    const double delta = fmax(0.25 * gap, 1e-13 * fmax(fabs(eprev), 1e-3));
    lo = eprev - delta;
    hi = eprev + delta;
  } else {
    const double Z = b.Z[p];
    lo = -(Z * Z + 10.0);
    hi = 1.0;
  }
  int ncount = 0;
  int clo = eig_factor(d, S, Hpk, lo); ++ncount;
  {
    double step = fmax(hi - lo, 1e-3 * fabs(lo));
    while (clo > k && ncount < 400) { hi = lo; lo -= step; step *= 2.0; clo = eig_factor(d, S, Hpk, lo); ++ncount; }
  }
  int chi = eig_factor(d, S, Hpk, hi); ++ncount;
  {
    double step = fmax(hi - lo, 1e-3 * fabs(hi));
    while (chi < k + 1 && ncount < 400) { hi += step; step *= 2.0; chi = eig_factor(d, S, Hpk, hi); ++ncount; }
  }
  // ---- bisection until eigenvalue k is isolated in a narrow interval
  while (!(clo == k && chi == k + 1 && (hi - lo) <= 0.05 * fmax(fmax(fabs(lo), fabs(hi)), 1e-12)) &&
         ncount < 400) {
    const double mid = 0.5 * (lo + hi);
    if (!(mid > lo && mid < hi)) break;
    const int cm = eig_factor(d, S, Hpk, mid); ++ncount;
    if (cm <= k) { lo = mid; clo = cm; } else { hi = mid; chi = cm; }
  }

  // ---- start vector, M-normalised; r = M x
  if (warm) {
    for (int i = threadIdx.x; i < Nh; i += blockDim.x) S.vx[i] = Uout[i];
  } else {
    for (int i = threadIdx.x; i < Nh; i += blockDim.x)
      S.vx[i] = 1.0 + 0.37 * (double)((i * 7 + 3 * k) % 11) / 11.0;
  }
  __syncthreads();
  eig_apply(d, S, d.M0_loc, S.vx, S.vr);
  double xmx = eig_dot(S, S.vx, S.vr, Nh);
  {
    const double sc = 1.0 / sqrt(xmx);
    for (int i = threadIdx.x; i < Nh; i += blockDim.x) { S.vx[i] *= sc; S.vr[i] *= sc; }
    __syncthreads();
  }
  double sigma = 0.5 * (lo + hi);
  if (warm) {
    const double rq = eig_quadform_pk(d, S, Hpk, S.vx);
    if (rq > lo && rq < hi) sigma = rq;
  }
  // ---- Rayleigh quotient iteration, safeguarded by the isolating interval
  double rho = sigma;
  int its = 0;
  bool converged = false;
  for (; its < 14; ++its) {
    eig_factor(d, S, Hpk, sigma);
    eig_solve(d, S);                                  // y = (H - sigma M)^{-1} M x
    const double yr = eig_dot(S, S.vy, S.vr, Nh);     // y^T M x
    eig_apply(d, S, d.M0_loc, S.vy, S.vr);            // r <- M y
    const double ymy = eig_dot(S, S.vy, S.vr, Nh);
    rho = sigma + yr / ymy;
    const double sc = 1.0 / sqrt(ymy);
    for (int i = threadIdx.x; i < Nh; i += blockDim.x) { S.vx[i] = S.vy[i] * sc; S.vr[i] *= sc; }
    __syncthreads();
    if (!(rho > lo && rho < hi)) {
      // left the isolating interval (or NaN): shrink it and restart from its midpoint
      for (int r3 = 0; r3 < 3; ++r3) {
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo && mid < hi)) break;
        const int cm = eig_factor(d, S, Hpk, mid); ++ncount;
        if (cm <= k) lo = mid; else hi = mid;
      }
      sigma = 0.5 * (lo + hi);
      rho = sigma;
      continue;
    }
    if (fabs(rho - sigma) <= 1e-14 * fabs(rho)) { converged = true; ++its; break; }
    sigma = rho;
  }
  // ---- write back: eigenvalue, vector with a fixed sign convention
  double big = 0.0;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) big += S.vx[i];
  big = block_sum(big, S.red);
  const double sgn = (big < 0.0) ? -1.0 : 1.0;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) Uout[i] = sgn * S.vx[i];
  if (threadIdx.x == 0) {
    b.eps_new[chan * b.nhmax + k] = rho;
    b.eig_info[chan * b.nhmax + k] = (ncount << 8) | (its & 0x7f) | (converged ? 0 : 0x80);
  }
}
