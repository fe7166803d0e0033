// Auxiliary kernels: state from a dense density matrix (restart / "identical input D"
// parity hooks), eigen-residuals, Exc of the current state, FP64 peak probe.
#pragma once
#include "aks_device.cuh"
#include "aks_types.cuh"
#include "k_eig.cuh"
#include "k_scf.cuh"

// rho at the Gauss nodes of cell c and the cell's share of B = F:D from a DENSE D:
//   rho_sigma(r_q) = (g_q^T D_loc g_q) / r_q^2 / (4 pi)     (optimized_eval_density!,
//   src/discretization/assemble.jl:91-125, in factored form)
template <int NMAX>
__global__ void __launch_bounds__(AKS_NT) k_cells_from_dense(DiscDev d, BatchDev b, int p) {
  const int c = blockIdx.x;
  extern __shared__ double sm[];
  const int n = d.n, Q = d.Q, C = d.C, s = b.s, Nh = d.Nh;
  double* Dl = sm;  // [s][n*n]
  for (int idx = threadIdx.x; idx < s * n * n; idx += AKS_NT) {
    const int e = idx / (n * n), r = idx - e * n * n, i = r / n, j = r - i * n;
    const int gi = d.l2g[c * n + i], gj = d.l2g[c * n + j];
    Dl[idx] = (gi >= 0 && gj >= 0) ? b.Dd[((size_t)p * s + e) * Nh * Nh + (size_t)gi * Nh + gj] : 0.0;
  }
  __syncthreads();
  const double inv1 = d.inv1[c], inv2 = d.inv2[c];
  for (int q = threadIdx.x; q < Q; q += AKS_NT) {
    double g[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) g[i] = (i < n) ? d.Pg[(size_t)i * d.Qs + q] : 0.0;
    const double X = inv1 * d.x[q] + inv2;
    for (int e = 0; e < s; ++e) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) {
        double col = 0.0;
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
          if (i < n) col = fma(Dl[e * n * n + i * n + j], g[i], col);
        // g[j] with a runtime index: read it back from global (L1 hit)
        acc = fma(col, d.Pg[(size_t)j * d.Qs + q], acc);
      }
      b.rho_q[((size_t)(p * s + e) * C + c) * Q + q] = (acc / (X * X)) * c_xc.inv4pi;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* Fc = d.F_loc + (size_t)c * n * n * n;
  for (int m = warp; m < n; m += AKS_NW) {
    double acc = 0.0;
    for (int idx = lane; idx < n * n; idx += 32) {
      double dsum = Dl[idx];
      if (s == 2) dsum += Dl[n * n + idx];
      acc = fma(dsum, Fc[(size_t)m * n * n + idx], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) b.Bloc_new[(((size_t)p * 2 + 0) * C + c) * n + m] = acc;
  }
}

// rho on the global energy grid from a DENSE D (eval_density!, density.jl:97-126)
__global__ void __launch_bounds__(AKS_NT) k_grid_from_dense(DiscDev d, BatchDev b, int p) {
  const int q = blockIdx.x * AKS_NT + threadIdx.x;
  if (q >= d.G) return;
  const int n = d.n, c = d.ycell[q], s = b.s, Nh = d.Nh;
  const double yq = d.y[q];
  for (int e = 0; e < s; ++e) {
    const double* D = b.Dd + ((size_t)p * s + e) * Nh * Nh;
    double acc = 0.0;
    for (int a = 0; a < n; ++a) {
      const int ga = d.l2g[c * n + a];
      if (ga < 0) continue;
      const double ea = d.Pgy[(size_t)a * d.G + q];
      for (int bb = 0; bb < n; ++bb) {
        const int gb = d.l2g[c * n + bb];
        if (gb < 0) continue;
        acc += ea * D[(size_t)ga * Nh + gb] * d.Pgy[(size_t)bb * d.G + q];
      }
    }
    b.rho_y[((size_t)p * s + e) * d.G + q] = acc / (c_xc.fourpi * (yq * yq));
  }
}

// density!(D, U, n) (src/discretization/density.jl:15-41) of problem p as a dense matrix: the trial density
// the SCF loop itself never materialises (k_mix_dense folds it into the mixing).  out [s][Nh][Nh].
__global__ void __launch_bounds__(AKS_NT) k_density_dense(DiscDev d, BatchDev b, int p, double* __restrict__ out) {
  const int Nh = d.Nh, e = blockIdx.z;
  const int i = blockIdx.y * 16 + (threadIdx.x >> 4), j = blockIdx.x * 16 + (threadIdx.x & 15);
  if (i >= Nh || j >= Nh) return;
  double acc = 0.0;
  for (int l = 0; l < b.Lp[p]; ++l)
    for (int k = 0; k < b.nhp[p]; ++k) {
      const double nk = b.occ[occ_off(b, p, 0, e, l, k)];
      if (nk == 0.0) continue;
      const double* u = b.U + orb_off(b, p, e, l, k) * Nh;
      acc = fma(nk * u[i], u[j], acc);
    }
  out[((size_t)e * Nh + i) * Nh + j] = acc;
}

// Exc of the current (mixed) state of problem p -> out[0]
__global__ void __launch_bounds__(AKS_NT) k_exc_state(DiscDev d, BatchDev b, int p, double* out) {
  __shared__ double red[AKS_NW + 2];
  const bool has_xc = (b.ex[p] != 0) || (b.ec[p] != 0);
  const double v = has_xc ? exc_grid(d, b, p, 0.0, 0.0, 1.0, false, red) : 0.0;
  if (threadIdx.x == 0) out[0] = v;
}

// out = (unassembled SYMMETRIC operator [C][n][n]) * in   (vectors in the original numbering)
__device__ __noinline__ void dense_apply(int n, int C, const int* hatg, const int* bub0, double* hatpart,
                                         const double* __restrict__ mat, const double* in, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += AKS_NW) {
    int gi = -1;
    if (lane >= 2 && lane < n) gi = bub0[c] + lane - 2;
    else if (lane == 0 && c < C - 1) gi = hatg[c];
    else if (lane == 1 && c > 0) gi = hatg[c - 1];
    const double xl = (gi >= 0) ? in[gi] : 0.0;
    const double r = warp_colmatvec(mat + (size_t)c * n * n, n, n, lane, lane < n, xl);
    if (lane >= 2 && gi >= 0) out[gi] = r;
    if (lane < 2) hatpart[2 * c + lane] = r;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C - 1; j += blockDim.x) out[hatg[j]] = hatpart[2 * j] + hatpart[2 * (j + 1) + 1];
  __syncthreads();
}

// ||(H - eps M0) u||_2 for every eigenpair of problem p (diagnostic of aks_eig_lowest), original blocks
__global__ void __launch_bounds__(AKS_NT) k_eig_residual(DiscDev d, BatchDev b, int p, double* out) {
  const int k = blockIdx.x, ch = blockIdx.y;
  const int e = ch / b.Lmax, l = ch - e * b.Lmax;
  if (l >= b.Lp[p] || k >= b.nhp[p]) return;
  extern __shared__ double sm[];
  const int C = d.C, Nh = d.Nh, n = d.n;
  double* vu = sm;            // [Nh]
  double* vm = vu + Nh;       // [Nh]
  double* vh = vm + Nh;       // [Nh]
  double* hatpart = vh + Nh;  // [2C]
  double* red8 = hatpart + 2 * C;
  int* hatg = (int*)(red8 + AKS_NW + 8);
  int* bub0 = hatg + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { hatg[c] = (c < C - 1) ? d.l2g[c * n] : 0; bub0[c] = d.l2g[c * n + 2]; }
  const size_t chan = ((size_t)p * b.s + e) * b.Lmax + l;
  const double* Hfull = b.Hfull + chan * C * (size_t)n * n;
  const double* u = b.U + (chan * b.nhmax + k) * Nh;
  const double lam = b.eps_new[chan * b.nhmax + k];
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) vu[i] = u[i];
  __syncthreads();
  dense_apply(n, C, hatg, bub0, hatpart, d.M0_loc, vu, vm);
  dense_apply(n, C, hatg, bub0, hatpart, Hfull, vu, vh);
  double acc = 0.0;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) {
    const double r = vh[i] - lam * vm[i];
    acc = fma(r, r, acc);
  }
  acc = block_sum(acc, red8);
  if (threadIdx.x == 0) out[(size_t)ch * b.nhmax + k] = sqrt(acc);
}

// eigenpair window (k_eig): [p0, p0+np) back to "solve every level", nothing pending
__global__ void k_window_full(BatchDev b, int p0, int np) {
  const int per = b.s * b.Lmax;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np * per; i += gridDim.x * blockDim.x) {
    b.kact[(size_t)p0 * per + i] = b.nhmax;
    b.kdone[(size_t)p0 * per + i] = b.nhmax;
    for (int h = 0; h < AKS_KHIST; ++h) b.khist[((size_t)p0 * per + i) * AKS_KHIST + h] = 0;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) { b.redo[p0 + i] = 0; b.full_refill[p0 + i] = 0; }
}
__global__ void k_set_redo(BatchDev b, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < b.B; i += gridDim.x * blockDim.x) b.redo[i] = v;
}

// FP64 FMA throughput probe: 8 independent register chains per thread
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = 1.1, a2 = 1.2, a3 = 1.3, a4 = 1.4, a5 = 1.5, a6 = 1.6, a7 = 1.7;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// FP64 tensor-core throughput probe: mma.sync m8n8k4 f64, 8 independent accumulator tiles per warp
__global__ void k_dmma_peak(double* out, int iters) {
  double a = 1.0 + (threadIdx.x & 31) * 1e-9, b = 0.999999;
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) { c[t][0] = 1e-7 * t; c[t][1] = 2e-7 * t; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------
// k_eval_grid: the solution of one problem on arbitrary radial points (post-processing on the device),
// straight from the device state D, U, W -- nothing is copied to the host but the values asked for.
//   kind 0  u_{k,l,sigma}(x) / x            eval_orbital             src/solution/postprocess.jl:34-62
//   kind 5  u_{k,l,sigma}(x)                the H^1_0 function itself (h10 = true)
//   kind 1  rho_sigma(x) (c = -1: spin sum) eval_density             postprocess.jl:81-182:  g^T D_loc g / (4 pi x^2)
//   kind 2  W(x) / x + N / Rmax             eval_hartree             postprocess.jl:207-222
//   kind 3  v_xc,sigma[rho](x)              eval_vxc                 postprocess.jl:328-352
//   kind 4  -Z/x + l(l+1)/(2x^2) + hartree V_H + v_xc   eval_effective_potential   postprocess.jl:378-387
// One thread per point: cell by bisection on the cell boundaries (findindex, src/fem/mesh.jl:64-75: right-owning at interior
// nodes, the last node belongs to the last cell, 0 outside), generators by the Legendre three-term recurrence
// (src/fem/generators.jl:79-138: 1 + xi, 1 - xi, (P_{k+1} - P_{k-1}) / (2k + 1)).  Only the lower triangle of D is kept
// current during the iterations (k_mix_dense): D is read symmetrically.
#define EVAL_NMAX 32
__global__ void __launch_bounds__(128) k_eval_grid(DiscDev d, BatchDev b, int p, int kind, int qa, int qb, int qc,
                                                   long long npts, const double* __restrict__ xs, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  const double x = xs[t];
  const int n = d.n, C = d.C, Nh = d.Nh, s = b.s;
  // cell: last c with lower boundary <= x
  int c = -1;
  if (x >= d.inv2[0] - d.inv1[0] && x <= d.Rmax) {
    int lo = 0, hi = C - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (d.inv2[mid] - d.inv1[mid] <= x) lo = mid; else hi = mid - 1;
    }
    c = lo;
  }
  double g[EVAL_NMAX];
  int gi[EVAL_NMAX];
  if (c >= 0) {
    const double xi = (x - d.inv2[c]) / d.inv1[c];
    g[0] = 1.0 + xi;
    g[1] = 1.0 - xi;
    double pm = 1.0, pk = xi;   // P_{k-1}, P_k
    for (int k = 1; k + 1 < n; ++k) {
      const double pn = ((2.0 * k + 1.0) * xi * pk - (double)k * pm) / (double)(k + 1);
      g[k + 1] = (pn - pm) / (2.0 * k + 1.0);
      pm = pk; pk = pn;
    }
    for (int i = 0; i < n; ++i) gi[i] = d.l2g[c * n + i];
  }
  auto fem = [&](const double* coef) {   // sum_i coef[global(i)] g_i(x)
    double v = 0.0;
    if (c >= 0)
      for (int i = 0; i < n; ++i)
        if (gi[i] >= 0) v = fma(coef[gi[i]], g[i], v);
    return v;
  };
  auto dens = [&](int e) {   // rho_e(x)
    double acc = 0.0;
    if (c >= 0) {
      const double* D = b.Dd + ((size_t)p * s + e) * Nh * Nh;
      for (int i = 0; i < n; ++i) {
        if (gi[i] < 0) continue;
        double row = 0.0;
        for (int j = 0; j < n; ++j) {
          if (gi[j] < 0) continue;
          const int r = gi[i] >= gi[j] ? gi[i] : gi[j], cc = gi[i] >= gi[j] ? gi[j] : gi[i];
          row = fma(D[(size_t)r * Nh + cc], g[j], row);
        }
        acc = fma(g[i], row, acc);
      }
    }
    return acc / (c_xc.fourpi * x * x);
  };
  auto hartree = [&]() {
    const double bd = b.N[p] / d.Rmax;
    return b.hartree[p] == 0.0 ? bd : fem(b.Wv + (size_t)p * Nh) / x + bd;
  };
  auto vxc = [&](int e) {
    const int ex = b.ex[p], ec = b.ec[p];
    if (ex == 0 && ec == 0) return 0.0;
    if (s == 1) return xc_vrho_unpol(fmax(dens(0), 0.0), ex, ec);
    double vu, vd;
    xc_vrho_pol(fmax(dens(0), 0.0), fmax(dens(1), 0.0), ex, ec, vu, vd);
    return e == 0 ? vu : vd;
  };
  double v = 0.0;
  switch (kind) {
    case 0: v = fem(b.U + orb_off(b, p, qc, qa, qb) * Nh) / x; break;
    case 5: v = fem(b.U + orb_off(b, p, qc, qa, qb) * Nh); break;
    case 1: v = qc < 0 ? (s == 2 ? dens(0) + dens(1) : dens(0)) : dens(qc); break;
    case 2: v = hartree(); break;
    case 3: v = vxc(qc); break;
    default: {
      const double l = (double)qa;
      v = -b.Z[p] / x + l * (l + 1.0) / (2.0 * x * x) + b.hartree[p] * hartree() + vxc(qc);
    }
  }
  out[t] = v;
}
