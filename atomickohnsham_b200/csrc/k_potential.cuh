// k_potential: density-dependent potential blocks of one (problem, cell).
//
// Replaces, for all spins at once,
//   assemble_exc!  (src/discretization/assemble.jl:277-319) with its per-cell body
//                  fill_local_matrix!(K, ::GaussLegendre, ::FunWeight, ...)
//                  (src/fem/local matrix.jl:129-155) and evaluate_vrho!
//                  (src/physics/models.jl:97-105),
//   the F.W contraction + N/Rmax M0 shift of assemble_hartree! (assemble.jl:78-81),
//   assemble_hamiltonian! (assemble.jl:334-358): H_l = Hfix_l + Vxc + Hartree.
//
// The XC block is the factored contraction K = B^T diag(w v) B with B = generator values at
// the Gauss nodes (north star), not the reference's Qgenx mat-vec: rho at the nodes is part of
// the device state (it mixes linearly with D), so only the second GEMM-like half remains:
//   K[i][j] = inv1 * sum_q g_i(q) (w_q v(rho_q)) g_j(q)          2 n^2 Q flop per cell and spin.
// One warp owns a q-slice; each lane owns 3x3 register tiles of the symmetric K (upper tile
// triangle only); generator values are staged per 64-node chunk in shared memory (TMA, 4 stages).
#pragma once

#include "aks_device.cuh"
#include "aks_types.cuh"

// Node chunk staged per TMA transfer and its padded row stride in shared memory: 66 doubles keeps every
// row 16-byte aligned (bulk copies need it) and maps the 7 rows a warp touches (3*ti*66 mod 16 doubles)
// onto distinct bank groups.  POT_ST chunks are in flight / in use at any time.
#define POT_QC 64
#define POT_QCS 66
#define POT_ST 4

// A CTA owns one cell and NU "units" = (problem, spin) pairs: the generator chunks, the mass-tensor
// block F_c and the fixed blocks of the cell are fetched once for all of them, and the 9 products
// g_i g_j of a lane's tile are formed once per node and reused by every unit (9 DMUL + NU x 9 DFMA).
// Orders p <= 20 need one 3x3 tile per lane (NTS = 1, NU = 4); larger orders two (NTS = 2, NU = 2).
__host__ __device__ inline int potential_nts(int n) { const int T = (n + 2) / 3; return (T * (T + 1) / 2 <= 32) ? 1 : 2; }
__host__ __device__ inline int potential_nu(int n) { return potential_nts(n) == 1 ? 4 : 2; }
__host__ __device__ inline int potential_np(int n, int s) { const int np = potential_nu(n) / s; return np < 1 ? 1 : np; }

__host__ __device__ inline size_t potential_smem_bytes(int n, int Q, int s) {
  const int T = (n + 2) / 3, n3 = 3 * T;
  const int Qpad = ((Q + POT_QC - 1) / POT_QC) * POT_QC;
  const int nu = potential_nu(n) < s ? s : potential_nu(n);
  size_t stage = (size_t)POT_ST * n3 * POT_QCS, kp = (size_t)AKS_NW * n3 * n3;
  if (kp > stage) stage = kp;
  return sizeof(double) * ((size_t)nu * Qpad + stage + 2 * (size_t)nu * n * n + (size_t)nu * n + 8);
}

// (mbarrier / bulk-copy primitives: aks_device.cuh)

// k_xc: pointwise XC potential at the Gauss nodes of every cell, f_sigma[q] = w_q v_sigma(rho(q))
// (evaluate_vrho!, src/physics/models.jl:97-105; clamp at 0, BuiltinFunctional.jl:112).  One thread per node:
// long dependent FP64 chains (rcbrt, sqrt, log, one division), so it runs at full occupancy on its own
// instead of inside the register-heavy contraction kernel.  k_dens_cells reads f again for <Vxc[D_prev], D_new>.
__global__ void __launch_bounds__(256) k_xc(DiscDev d, BatchDev b, int p0, int np, int force) {
  const int p = p0 + blockIdx.z;
  if (p >= p0 + np || (!force && b.status[p] != 0)) return;
  const int ex = b.ex[p], ec = b.ec[p];
  if (ex == 0 && ec == 0) return;
  const int c = blockIdx.y, q = blockIdx.x * 256 + threadIdx.x;
  const int Q = d.Q, C = d.C, s = b.s;
  if (q >= Q) return;
  const double wq = d.w[q];
  const size_t o0 = ((size_t)(p * s + 0) * C + c) * Q + q;
  if (s == 1) {
    b.fxc[o0] = xc_vrho_unpol(fmax(b.rho_q[o0], 0.0), ex, ec) * wq;
  } else {
    const size_t o1 = ((size_t)(p * s + 1) * C + c) * Q + q;
    double vu, vd;
    xc_vrho_pol(fmax(b.rho_q[o0], 0.0), fmax(b.rho_q[o1], 0.0), ex, ec, vu, vd);
    b.fxc[o0] = vu * wq;
    b.fxc[o1] = vd * wq;
  }
}

template <int NTS, int NU>
__global__ void __launch_bounds__(AKS_NT, 2)
k_potential(DiscDev d, BatchDev b, double* __restrict__ vxc_loc_out, double* __restrict__ har_loc_out,
            int p0, int np, int force) {
  const int c = blockIdx.x;
  const int n = d.n, Q = d.Q, s = b.s, C = d.C;
  constexpr int NUs = NU;                    // units of this CTA
  const int NP = (NUs / s) < 1 ? 1 : NUs / s;  // problems of this CTA (NU >= s is guaranteed by the host)
  const int pbase = p0 + blockIdx.y * NP;
  // unit u -> problem pbase + u / s, spin u % s
  bool uval[NU];
  bool any = false;
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int pu = pbase + u / s;
    uval[u] = (u / s < NP) && (pu < p0 + np) && (force || b.status[pu] == 0);
    any = any || uval[u];
  }
  if (!any) return;
  extern __shared__ __align__(16) double sm_pot[];
  double* sm = sm_pot;
  const int T = (n + 2) / 3, n3 = 3 * T, ntile = T * (T + 1) / 2;
  const int Qpad = ((Q + POT_QC - 1) / POT_QC) * POT_QC;
  size_t stage = (size_t)POT_ST * n3 * POT_QCS;
  if ((size_t)AKS_NW * n3 * n3 > stage) stage = (size_t)AKS_NW * n3 * n3;
  double* fs = sm;                           // [NU][Qpad]   w_q v_u(rho_q)
  double* Gs = fs + (size_t)NU * Qpad;       // [POT_ST][n3][POT_QCS] node chunks (16 B aligned)
  double* Kp = Gs;                           // [NW][n3*n3]  cross-warp reduction, after the chunks are done
  double* Ks = Gs + stage;                   // [NU][n*n]    XC block per unit
  double* Hs = Ks + (size_t)NU * n * n;      // [NP][n*n]    Hartree block per problem
  double* wl = Hs + (size_t)NU * n * n;      // [NP][n]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double inv1 = d.inv1[c];

  // ---- 1. f_u[q] = w_q v_u(rho(q)) of this CTA's units, evaluated by k_xc (pointwise, full occupancy)
  bool any_xc = false;
  for (int pp = 0; pp < NP; ++pp) {
    const int p = pbase + pp;
    const bool val = uval[pp * s];
    const bool has_xc = val && ((b.ex[p] != 0) || (b.ec[p] != 0));
    any_xc = any_xc || has_xc;
    for (int e = 0; e < s; ++e) {
      double* fd = fs + (size_t)(pp * s + e) * Qpad;
      const double* fg = b.fxc + ((size_t)(p * s + e) * C + c) * Q;
      for (int q = tid; q < Qpad; q += AKS_NT) fd[q] = (has_xc && q < Q) ? fg[q] : 0.0;
    }
  }
  // units beyond the problems of this CTA (NU > NP * s cannot happen; kept for safety)
  for (int u = NP * s; u < NU; ++u)
    for (int q = tid; q < Qpad; q += AKS_NT) fs[(size_t)u * Qpad + q] = 0.0;

  // ---- 2. K_u = sum_q f_u[q] g g^T, 3x3 register tiles
  int ti[NTS], tj[NTS];
#pragma unroll
  for (int ts = 0; ts < NTS; ++ts) {
    int tile = lane + 32 * ts;
    ti[ts] = -1;
    tj[ts] = -1;
    if (tile < ntile) {
      int a = 0, rem = tile;
      while (rem >= T - a) { rem -= (T - a); ++a; }  // row a holds tiles (a, a..T-1)
      ti[ts] = a;
      tj[ts] = a + rem;
    }
  }
  double acc[NTS][NU][9];
#pragma unroll
  for (int a = 0; a < NTS; ++a)
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[a][u][k] = 0.0;

  if (any_xc) {
    // Generator values at the nodes are streamed through shared memory in 64-node chunks by the TMA
    // engine (cp.async.bulk, one 512 B row per transfer), POT_ST stages deep, full/empty mbarrier pair per
    // stage: chunk ch + POT_ST is requested as soon as every warp has released chunk ch.
    __shared__ __align__(8) uint64_t full_bar[POT_ST], empty_bar[POT_ST];
    const int nch = Qpad / POT_QC;
    const bool tma_ok = ((Q % 2) == 0) && ((d.Qs % 2) == 0);   // 16-byte granularity of every row piece
    if (tid == 0) {
#pragma unroll
      for (int st = 0; st < POT_ST; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], AKS_NW); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // rows >= n and the tail of the last chunk are never written by the copies: zero the buffers once
    for (int idx = tid; idx < POT_ST * n3 * POT_QCS; idx += AKS_NT) Gs[idx] = 0.0;
    __syncthreads();
    if (tid == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // zeros before the first bulk writes
    auto issue = [&](int ch) {   // thread 0 only
      const int st = ch % POT_ST;
      double* dst = Gs + (size_t)st * n3 * POT_QCS;
      const int q0 = ch * POT_QC, len = min(POT_QC, Q - q0);
      mbar_expect_tx(&full_bar[st], (uint32_t)(n * len * 8));
      for (int g = 0; g < n; ++g) bulk_g2s(dst + g * POT_QCS, d.Pg + (size_t)g * d.Qs + q0, (uint32_t)(len * 8), &full_bar[st]);
    };
    if (tma_ok && tid == 0)
      for (int ch = 0; ch < POT_ST && ch < nch; ++ch) issue(ch);
    for (int ch = 0; ch < nch; ++ch) {
      const int q0 = ch * POT_QC, st = ch % POT_ST;
      const uint32_t parity = (uint32_t)((ch / POT_ST) & 1);
      double* Gb = Gs + (size_t)st * n3 * POT_QCS;
      if (tma_ok) {
        mbar_wait(&full_bar[st], parity);   // the chunk has landed; no CTA-wide barrier: warps drift up to POT_ST chunks
      } else {
        __syncthreads();
        for (int idx = tid; idx < n * POT_QC; idx += AKS_NT) {
          const int g = idx / POT_QC, j = idx - g * POT_QC;
          Gb[g * POT_QCS + j] = (q0 + j < Q) ? d.Pg[(size_t)g * d.Qs + q0 + j] : 0.0;
        }
        __syncthreads();
      }
      // hot loop: two nodes per iteration -- every shared-memory access is a 128-bit load (the rows of a chunk and the
      // weights are 16-byte aligned, a warp's slice starts at an even node): 6 + NU LDS.128 feed 18 DMUL + NU x 18 DFMA
      const int jq0 = warp * (POT_QC / AKS_NW);
#pragma unroll
      for (int ts = 0; ts < NTS; ++ts) {
        if (ti[ts] < 0) continue;
        const double2* gi = reinterpret_cast<const double2*>(Gb + (3 * ti[ts]) * POT_QCS + jq0);
        const double2* gj = reinterpret_cast<const double2*>(Gb + (3 * tj[ts]) * POT_QCS + jq0);
        const double2* f0 = reinterpret_cast<const double2*>(fs + q0 + jq0);
#pragma unroll 1
        for (int jq = 0; jq < POT_QC / AKS_NW / 2; ++jq) {
          const double2 a0 = gi[jq], a1 = gi[POT_QCS / 2 + jq], a2 = gi[POT_QCS + jq];
          const double2 b0 = gj[jq], b1 = gj[POT_QCS / 2 + jq], b2 = gj[POT_QCS + jq];
          double2 fq[NU];
#pragma unroll
          for (int u = 0; u < NU; ++u) fq[u] = f0[(size_t)u * (Qpad / 2) + jq];
          {
            const double p00 = a0.x * b0.x, p01 = a0.x * b1.x, p02 = a0.x * b2.x;
            const double p10 = a1.x * b0.x, p11 = a1.x * b1.x, p12 = a1.x * b2.x;
            const double p20 = a2.x * b0.x, p21 = a2.x * b1.x, p22 = a2.x * b2.x;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
              const double fv = fq[u].x;
              double* A = acc[ts][u];
              A[0] = fma(p00, fv, A[0]); A[1] = fma(p01, fv, A[1]); A[2] = fma(p02, fv, A[2]);
              A[3] = fma(p10, fv, A[3]); A[4] = fma(p11, fv, A[4]); A[5] = fma(p12, fv, A[5]);
              A[6] = fma(p20, fv, A[6]); A[7] = fma(p21, fv, A[7]); A[8] = fma(p22, fv, A[8]);
            }
          }
          {
            const double p00 = a0.y * b0.y, p01 = a0.y * b1.y, p02 = a0.y * b2.y;
            const double p10 = a1.y * b0.y, p11 = a1.y * b1.y, p12 = a1.y * b2.y;
            const double p20 = a2.y * b0.y, p21 = a2.y * b1.y, p22 = a2.y * b2.y;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
              const double fv = fq[u].y;
              double* A = acc[ts][u];
              A[0] = fma(p00, fv, A[0]); A[1] = fma(p01, fv, A[1]); A[2] = fma(p02, fv, A[2]);
              A[3] = fma(p10, fv, A[3]); A[4] = fma(p11, fv, A[4]); A[5] = fma(p12, fv, A[5]);
              A[6] = fma(p20, fv, A[6]); A[7] = fma(p21, fv, A[7]); A[8] = fma(p22, fv, A[8]);
            }
          }
        }
      }
      if (tma_ok) {
        // this warp is done with the stage; thread 0 refills it with chunk ch + POT_ST once all warps are
        // (a shorter last chunk leaves older, finite values in the tail columns; f is zero there)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
        if (tid == 0 && ch + POT_ST < nch) {
          mbar_wait(&empty_bar[st], parity);
          issue(ch + POT_ST);
        }
        __syncwarp();
      }
    }
  }
  // cross-warp reduction, fixed order; symmetrise diagonal tiles; scale by inv1
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    __syncthreads();
    if (any_xc) {
#pragma unroll
      for (int ts = 0; ts < NTS; ++ts) {
        if (ti[ts] < 0) continue;
        double* K = Kp + (size_t)warp * n3 * n3;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb)
            K[(3 * ti[ts] + a) * n3 + 3 * tj[ts] + bb] = acc[ts][u][3 * a + bb];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += AKS_NT) {
      const int i = idx / n, j = idx - i * n;
      double v = 0.0;
      if (any_xc) {
        const int ii = (i / 3 <= j / 3) ? i : j, jj = (i / 3 <= j / 3) ? j : i;
        double s1 = 0.0, s2 = 0.0;
        for (int wv = 0; wv < AKS_NW; ++wv) {
          s1 += Kp[(size_t)wv * n3 * n3 + ii * n3 + jj];
          if (i / 3 == j / 3) s2 += Kp[(size_t)wv * n3 * n3 + jj * n3 + ii];
        }
        v = (i / 3 == j / 3) ? 0.5 * (s1 + s2) : s1;
        v *= inv1;
      }
      Ks[(size_t)u * n * n + idx] = v;
    }
  }

  // ---- 3. Hartree block per problem: coeff * (sum_m W_m F[ij][m] + N/Rmax M0); F_c is read once
  for (int idx = tid; idx < NP * n; idx += AKS_NT) {
    const int pp = idx / n, m = idx - pp * n;
    const int p = pbase + pp;
    const int gi = d.l2g[c * n + m];
    const bool val = uval[pp * s];
    wl[idx] = (val && gi >= 0 && b.hartree[p] != 0.0) ? b.Wv[(size_t)p * d.Nh + gi] : 0.0;
  }
  __syncthreads();
  {
    const double* Fc = d.F_loc + (size_t)c * n * n * n;
    const double* M0c = d.M0_loc + (size_t)c * n * n;
    bool any_h = false;
    for (int pp = 0; pp < NP; ++pp) any_h = any_h || (uval[pp * s] && b.hartree[pbase + pp] != 0.0);
    for (int idx = tid; idx < n * n; idx += AKS_NT) {
      double fw[NU];
#pragma unroll
      for (int pp = 0; pp < NU; ++pp) fw[pp] = 0.0;
      if (any_h)
#pragma unroll 7
        for (int m = 0; m < n; ++m) {
          const double fv = Fc[(size_t)m * n * n + idx];
#pragma unroll
          for (int pp = 0; pp < NU; ++pp)
            if (pp < NP) fw[pp] = fma(wl[pp * n + m], fv, fw[pp]);
        }
      const double m0 = M0c[idx];
#pragma unroll
      for (int pp = 0; pp < NU; ++pp)
        if (pp < NP) {
          const int p = pbase + pp;
          double v = 0.0;
          if (uval[pp * s]) {
            const double coeff = b.hartree[p];
            if (coeff != 0.0) v = (fw[pp] + (b.N[p] / d.Rmax) * m0) * coeff;
          }
          Hs[(size_t)pp * n * n + idx] = v;
        }
    }
  }
  __syncthreads();

  // optional local outputs for the step-level hooks (problem p0 only: launched with np = 1)
  if (vxc_loc_out && blockIdx.y == 0)
    for (int e = 0; e < s; ++e)
      for (int idx = tid; idx < n * n; idx += AKS_NT)
        vxc_loc_out[((size_t)e * C + c) * n * n + idx] = Ks[(size_t)e * n * n + idx];
  if (har_loc_out && blockIdx.y == 0)
    for (int idx = tid; idx < n * n; idx += AKS_NT) har_loc_out[(size_t)c * n * n + idx] = Hs[idx];

  // ---- 4. H_l = (Kin_l + Coulomb) + Vxc + Hartree, packed per (problem, sigma, l, cell).  The fixed blocks
  // of the cell and the symmetric-twin / packing tables are staged once (the chunk buffers are free now):
  // the (problem, sigma, l) loop then reads shared memory only.
  double* As = Gs;                 // [n*n]  A / 2
  double* M2s = As + n * n;        // [n*n]  M_-2 / 2
  double* M1s = M2s + n * n;       // [n*n]  M_-1
  int* los = reinterpret_cast<int*>(M1s + n * n);   // [n*n]
  int* pks = los + n * n;                           // [n*n]
  {
    const double* Ac = d.A_loc + (size_t)c * n * n;
    const double* M1c = d.Mm1_loc + (size_t)c * n * n;
    const double* M2c = d.Mm2_loc + (size_t)c * n * n;
    const bool has_m2 = d.Mm2_loc != nullptr;
    for (int idx = tid; idx < n * n; idx += AKS_NT) {
      As[idx] = Ac[idx] * 0.5;
      M2s[idx] = has_m2 ? M2c[idx] * 0.5 : 0.0;
      M1s[idx] = M1c[idx];
      los[idx] = d.lo_tab[idx];   // index of the lower-triangle twin: exactly symmetric output
      pks[idx] = d.pk_tab[idx];   // packed position, or -1 above the diagonal
    }
  }
  __syncthreads();
  for (int pp = 0; pp < NP; ++pp) {
    if (!uval[pp * s]) continue;
    const int p = pbase + pp;
    const double Zp = b.Z[p];
    const int Lp = b.Lp[p];
    for (int e = 0; e < s; ++e) {
      const double* Ku = Ks + (size_t)(pp * s + e) * n * n;
      const double* Hp = Hs + (size_t)pp * n * n;
      for (int l = 0; l < Lp; ++l) {
        double* out = b.Hpk + ((((size_t)p * s + e) * b.Lmax + l) * C + c) * d.npk;
        double* outf = b.Hfull + ((((size_t)p * s + e) * b.Lmax + l) * C + c) * (size_t)n * n;
        const double ll = (double)(l * (l + 1));
        for (int idx = tid; idx < n * n; idx += AKS_NT) {
          const int lo_idx = los[idx], pk = pks[idx];
          const double kin = As[lo_idx] + ll * M2s[lo_idx];
          const double cou = -Zp * M1s[lo_idx];
          const double h = ((kin + cou) + Ku[lo_idx]) + Hp[lo_idx];
          outf[idx] = h;
          if (pk >= 0) out[pk] = h;
        }
      }
    }
  }
}
