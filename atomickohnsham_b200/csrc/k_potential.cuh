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

// -DAKS_POT_TIMING: cycle stamps of thread 0 at the phase boundaries of the potential kernels, 8 per CTA, in eig_dbg
// (tools/gpu_pot_cycles.py)
#ifdef AKS_POT_TIMING
#define POT_STAMP_DECL() long long* stamp_ = b.eig_dbg + ((size_t)blockIdx.y * d.C + blockIdx.x) * 8; int nstamp_ = 0
#define POT_STAMP() do { if (threadIdx.x == 0 && nstamp_ < 8) stamp_[nstamp_++] = clock64(); } while (0)
#else
#define POT_STAMP_DECL() do { } while (0)
#define POT_STAMP() do { } while (0)
#endif

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
  POT_STAMP_DECL();
  POT_STAMP();

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

  POT_STAMP();
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
  POT_STAMP();
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

  POT_STAMP();
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

  POT_STAMP();
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
  __syncthreads();
  POT_STAMP();
}


// ---------------------------------------------------------------------------------------------------------------------
// k_potential_tc + k_hassemble: the same blocks with the contraction on the FP64 tensor cores (mma.sync m8n8k4 ->
// DMMA.8x8x4), for orders p <= 23 (n <= 24 generators: three 8-row tiles) and Q a multiple of 4.
//
//   K_u = G^T diag(f_u) G:  per 4 nodes, A = (G diag(f_u))^T (8 generators x 4 nodes) and B = G (4 nodes x 8 generators)
//   have the SAME fragment layout (row / column lane / 4, node lane % 4): one shared-memory load per 8-row tile serves as
//   B fragment and, scaled by f_u of the lane's node, as A fragment.
//
// Work split of k_potential_tc: a CTA owns one cell and 4 units; warp w owns unit w / 2 and three of its six upper 8x8
// tiles ((0,0) (0,1) (0,2) for even w, (1,1) (1,2) (2,2) for odd w) over ALL nodes -- the accumulators of a tile live in
// one warp from the first node to the last, so there is no cross-warp reduction, and a lane carries 12 accumulator
// doubles (even / odd 4-node steps kept apart: six independent DMMA chains) instead of 36.  f_u[q] is streamed with the
// generator chunks (TMA, POT_TC_ST stages, full / empty mbarrier pair per stage) instead of being staged whole.  The kernel
// is nothing but the contraction (its CTAs would otherwise run their latency-bound prologue and epilogue in lockstep, wave
// by wave, with the FP64 pipe idle): the XC blocks go to `Kxc`, and the Hartree block + assembly of H_l -- streaming
// work -- are k_hassemble's.
#define POT_TC_QCS 68       // row stride of a node chunk: 4 mod 16 doubles -> the 8 x 4 fragment hits 16 distinct banks per half warp
#define POT_TC_ROWS 24
#define POT_TC_ST 4
#define POT_TC_NU 4
#define POT_TC_NT (AKS_NT + 32)   // eight computing warps + the producer warp
__host__ __device__ inline bool potential_tc_ok(int n, int Q, int Qs, int s) {
  return n <= POT_TC_ROWS && (Q % 4) == 0 && (Qs % 2) == 0 && s <= POT_TC_NU;
}
__host__ __device__ inline size_t potential_tc_smem_bytes() {
  size_t stage = (size_t)POT_TC_ST * (POT_TC_ROWS * POT_TC_QCS + POT_TC_NU * POT_QC);
  const size_t tiles = (size_t)POT_TC_NU * POT_TC_ROWS * POT_TC_ROWS;
  if (tiles > stage) stage = tiles;
  return sizeof(double) * (stage + 8);
}
__host__ __device__ inline size_t hassemble_smem_bytes(int n) {
  return sizeof(double) * (4 * (size_t)n * n + 2 * (size_t)POT_TC_NU * n * n + (size_t)POT_TC_NU * n + 8);
}

__global__ void __launch_bounds__(POT_TC_NT, 3)
k_potential_tc(DiscDev d, BatchDev b, int p0, int np, int force) {
  constexpr int NU = POT_TC_NU, QCS = POT_TC_QCS, ROWS = POT_TC_ROWS, ST = POT_TC_ST;
  const int c = blockIdx.x;
  const int n = d.n, Q = d.Q, s = b.s, C = d.C;
  const int NP = NU / s;
  const int pbase = p0 + blockIdx.y * NP;
  unsigned vmask = 0, xmask = 0;   // units that run / that have a functional
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int pu = pbase + u / s;
    const bool val = (u / s < NP) && (pu < p0 + np) && (force || b.status[pu] == 0);
    if (val) vmask |= 1u << u;
    if (val && ((b.ex[pu] != 0) || (b.ec[pu] != 0))) xmask |= 1u << u;
  }
  if (!vmask) return;
  auto uxc = [&](int u) { return ((xmask >> u) & 1u) != 0; };
  extern __shared__ __align__(16) double sm_pot[];
  constexpr int STAGE = ROWS * QCS + NU * POT_QC;     // doubles per stage: generator rows, then f of the 4 units
  double* Gs = sm_pot;                       // [ST][STAGE]; later the tiles [NU][24][24]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double inv1 = d.inv1[c];
  POT_STAMP_DECL();
  POT_STAMP();

  __shared__ __align__(8) uint64_t full_bar[ST], empty_bar[ST];
  const int nch = (Q + POT_QC - 1) / POT_QC;
  // One lane of the producer warp (warp 8) issues a chunk: its image of the generator table (DiscDev::Pgc: one 13 KB copy
  // instead of n row copies of 512 B -- small bulk copies retire at ~1 per 60-100 cycles per SM), then the f rows of the units
  // with a functional.  Bulk copies are warp-uniform instructions (UBLKCP): a computing warp that issued them would be the
  // straggler of its CTA, so the ninth warp does nothing else.
  auto issue = [&](int ch) {
    const int st = ch % ST;
    double* dst = Gs + (size_t)st * STAGE;
    const int q0 = ch * POT_QC, len = min(POT_QC, Q - q0);
    mbar_expect_tx(&full_bar[st], (uint32_t)(ROWS * QCS * 8 + __popc(xmask) * len * 8));
    bulk_g2s(dst, d.Pgc + (size_t)ch * POT_IMG_ROWS * QCS, (uint32_t)(ROWS * QCS * 8), &full_bar[st]);   // the chunk image: one copy
#pragma unroll
    for (int u = 0; u < NU; ++u)
      if (uxc(u)) {
        const int pu = pbase + u / s, e = u % s;
        bulk_g2s(dst + ROWS * QCS + u * POT_QC, b.fxc + ((size_t)(pu * s + e) * C + c) * Q + q0, (uint32_t)(len * 8), &full_bar[st]);
      }
  };
  const int wu = warp >> 1, wh = warp & 1;           // unit and tile group of this warp
  double ca[3][2], cb[3][2];                        // even / odd 4-node steps
#pragma unroll
  for (int a = 0; a < 3; ++a) ca[a][0] = ca[a][1] = cb[a][0] = cb[a][1] = 0.0;
  if (xmask) {
    if (tid == 0) {
#pragma unroll
      for (int st = 0; st < ST; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], AKS_NW); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    POT_STAMP();
    if (warp == AKS_NW) {   // producer
      if (lane == 0) {
        for (int ch = 0; ch < nch; ++ch) {
          if (ch >= ST) mbar_wait(&empty_bar[ch % ST], (uint32_t)(((ch - ST) / ST) & 1));   // all eight warps released it
          issue(ch);
        }
      }
      __syncwarp();
    }
    const bool mine = uxc(wu);
    for (int ch = 0; warp < AKS_NW && ch < nch; ++ch) {
      const int st = ch % ST;
      const uint32_t parity = (uint32_t)((ch / ST) & 1);
      // every warp waits for the chunk, also one whose unit has no functional: its arrival on the empty barrier below
      // must not run a phase ahead of the others
      mbar_wait(&full_bar[st], parity);
      if (mine) {
        const double* Gb = Gs + (size_t)st * STAGE;
        const int nk = min(POT_QC, Q - ch * POT_QC) / 4;   // a shorter last chunk leaves older values in the tail columns
        const double* gb = Gb + (lane >> 2) * QCS + (lane & 3);
        const double* fb = Gb + ROWS * QCS + wu * POT_QC + (lane & 3);
        if (wh == 0) {
#pragma unroll 2
          for (int k4 = 0; k4 < nk; k4 += 2) {
            {
              const double g0 = gb[4 * k4], g1 = gb[8 * QCS + 4 * k4], g2 = gb[16 * QCS + 4 * k4];
              const double a0 = g0 * fb[4 * k4];
              dmma_m8n8k4(ca[0], a0, g0); dmma_m8n8k4(ca[1], a0, g1); dmma_m8n8k4(ca[2], a0, g2);
            }
            if (k4 + 1 < nk) {
              const double g0 = gb[4 * k4 + 4], g1 = gb[8 * QCS + 4 * k4 + 4], g2 = gb[16 * QCS + 4 * k4 + 4];
              const double a0 = g0 * fb[4 * k4 + 4];
              dmma_m8n8k4(cb[0], a0, g0); dmma_m8n8k4(cb[1], a0, g1); dmma_m8n8k4(cb[2], a0, g2);
            }
          }
        } else {
#pragma unroll 2
          for (int k4 = 0; k4 < nk; k4 += 2) {
            {
              const double g1 = gb[8 * QCS + 4 * k4], g2 = gb[16 * QCS + 4 * k4], fv = fb[4 * k4];
              const double a1 = g1 * fv, a2 = g2 * fv;
              dmma_m8n8k4(ca[0], a1, g1); dmma_m8n8k4(ca[1], a1, g2); dmma_m8n8k4(ca[2], a2, g2);
            }
            if (k4 + 1 < nk) {
              const double g1 = gb[8 * QCS + 4 * k4 + 4], g2 = gb[16 * QCS + 4 * k4 + 4], fv = fb[4 * k4 + 4];
              const double a1 = g1 * fv, a2 = g2 * fv;
              dmma_m8n8k4(cb[0], a1, g1); dmma_m8n8k4(cb[1], a1, g2); dmma_m8n8k4(cb[2], a2, g2);
            }
          }
        }
      }
      // this warp is done with the stage (or never reads it); the producer refills it once all eight warps are
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
  }
  POT_STAMP();
  // tiles -> shared (the chunk buffers are free once every warp is past the loop), then K_u: mirror the upper tiles,
  // average the two halves of the diagonal tiles (exactly symmetric output), scale by inv1
  __syncthreads();
  if (warp < AKS_NW && uxc(wu)) {
    double* Kt = Gs + (size_t)wu * ROWS * ROWS + (lane >> 2) * ROWS + 2 * (lane & 3);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ti = wh == 0 ? 0 : (a == 2 ? 2 : 1), tj = wh == 0 ? a : (a == 0 ? 1 : 2);
      Kt[(8 * ti) * ROWS + 8 * tj] = ca[a][0] + cb[a][0];
      Kt[(8 * ti) * ROWS + 8 * tj + 1] = ca[a][1] + cb[a][1];
    }
  }
  __syncthreads();
  const int nn = n * n;
  for (int idx = tid; idx < NU * nn; idx += POT_TC_NT) {
    const int u = idx / nn, r = idx - u * nn;
    if (!((vmask >> u) & 1u)) continue;
    const int i = r / n, j = r - i * n;
    double v = 0.0;
    if (uxc(u)) {
      const double* Kt = Gs + (size_t)u * ROWS * ROWS;
      const int ii = (i / 8 <= j / 8) ? i : j, jj = (i / 8 <= j / 8) ? j : i;
      const double s1 = Kt[ii * ROWS + jj];
      v = (i / 8 == j / 8) ? 0.5 * (s1 + Kt[jj * ROWS + ii]) : s1;
      v *= inv1;
    }
    const int pu = pbase + u / s, e = u % s;
    b.Kxc[((size_t)(pu * s + e) * C + c) * nn + r] = v;
  }
  POT_STAMP();
}

// k_hassemble: Hartree block of the cell, coeff * (sum_m W_m F[ij][m] + N/Rmax M0)  (assemble_hartree!, assemble.jl:78-81),
// and H_l = (Kin_l + Coulomb) + Vxc + Hartree (assemble_hamiltonian!, assemble.jl:334-358) packed per (problem, sigma, l,
// cell).  One CTA per cell and POT_TC_NU / s problems: F_c (L2 resident) is read once for all of them.  Streaming work:
// full occupancy, every load of a round in flight at once.
__global__ void __launch_bounds__(AKS_NT)
k_hassemble(DiscDev d, BatchDev b, double* __restrict__ vxc_loc_out, double* __restrict__ har_loc_out,
            int p0, int np, int force) {
  constexpr int NU = POT_TC_NU;
  const int c = blockIdx.x;
  const int n = d.n, s = b.s, C = d.C, nn = n * n;
  const int NP = NU / s;
  const int pbase = p0 + blockIdx.y * NP;
  unsigned vmask = 0;
#pragma unroll
  for (int pp = 0; pp < NU; ++pp) {
    const int p = pbase + pp;
    if (pp < NP && p < p0 + np && (force || b.status[p] == 0)) vmask |= 1u << pp;
  }
  if (!vmask) return;
  auto pval = [&](int pp) { return ((vmask >> pp) & 1u) != 0; };
  extern __shared__ __align__(16) double sm_pot[];
  double* As = sm_pot;             // [n*n]  A / 2
  double* M2s = As + nn;           // [n*n]  M_-2 / 2
  double* M1s = M2s + nn;          // [n*n]  M_-1
  int* los = reinterpret_cast<int*>(M1s + nn);   // [n*n]  index of the lower-triangle twin: exactly symmetric output
  int* pks = los + nn;                           // [n*n]  packed position, or -1 above the diagonal
  double* Hs = M1s + 2 * nn;       // [NP][n*n]  Hartree block per problem
  double* Ks = Hs + (size_t)NU * nn;   // [NU][n*n]  XC block per unit (from k_potential_tc)
  double* wl = Ks + (size_t)NU * nn;   // [NP][n]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NU * nn; idx += AKS_NT) {
    const int u = idx / nn, r = idx - u * nn;
    Ks[idx] = pval(u / s) ? b.Kxc[((size_t)((pbase + u / s) * s + u % s) * C + c) * nn + r] : 0.0;
  }
  for (int idx = tid; idx < NP * n; idx += AKS_NT) {
    const int pp = idx / n, m = idx - pp * n;
    const int p = pbase + pp;
    const int gi = d.l2g[c * n + m];
    wl[idx] = (pval(pp) && gi >= 0 && b.hartree[p] != 0.0) ? b.Wv[(size_t)p * d.Nh + gi] : 0.0;
  }
  {
    const double* Ac = d.A_loc + (size_t)c * nn;
    const double* M1c = d.Mm1_loc + (size_t)c * nn;
    const double* M2c = d.Mm2_loc + (size_t)c * nn;
    const bool has_m2 = d.Mm2_loc != nullptr;
    for (int idx = tid; idx < nn; idx += AKS_NT) {
      As[idx] = Ac[idx] * 0.5;
      M2s[idx] = has_m2 ? M2c[idx] * 0.5 : 0.0;
      M1s[idx] = M1c[idx];
      los[idx] = d.lo_tab[idx];
      pks[idx] = d.pk_tab[idx];
    }
  }
  __syncthreads();
  {
    const double* Fc = d.F_loc + (size_t)c * nn * n;
    const double* M0c = d.M0_loc + (size_t)c * nn;
    bool any_h = false;
    for (int pp = 0; pp < NP; ++pp) any_h = any_h || (pval(pp) && b.hartree[pbase + pp] != 0.0);
    for (int idx = tid; idx < nn; idx += AKS_NT) {
      double fw[NU];
#pragma unroll
      for (int pp = 0; pp < NU; ++pp) fw[pp] = 0.0;
      if (any_h) {
        int m = 0;
        for (; m + 8 <= n; m += 8) {   // 8 loads of F_c in flight per thread
          double fv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) fv[k] = Fc[(size_t)(m + k) * nn + idx];
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int pp = 0; pp < NU; ++pp)
              if (pp < NP) fw[pp] = fma(wl[pp * n + m + k], fv[k], fw[pp]);
        }
        for (; m < n; ++m) {
          const double fv = Fc[(size_t)m * nn + idx];
#pragma unroll
          for (int pp = 0; pp < NU; ++pp)
            if (pp < NP) fw[pp] = fma(wl[pp * n + m], fv, fw[pp]);
        }
      }
      const double m0 = M0c[idx];
#pragma unroll
      for (int pp = 0; pp < NU; ++pp)
        if (pp < NP) {
          const int p = pbase + pp;
          double v = 0.0;
          if (pval(pp)) {
            const double coeff = b.hartree[p];
            if (coeff != 0.0) v = (fw[pp] + (b.N[p] / d.Rmax) * m0) * coeff;
          }
          Hs[(size_t)pp * nn + idx] = v;
        }
    }
  }
  __syncthreads();
  // optional local outputs for the step-level hooks (problem p0 only: launched with np = 1)
  if (vxc_loc_out && blockIdx.y == 0)
    for (int e = 0; e < s; ++e)
      for (int idx = tid; idx < nn; idx += AKS_NT)
        vxc_loc_out[((size_t)e * C + c) * nn + idx] = Ks[(size_t)e * nn + idx];
  if (har_loc_out && blockIdx.y == 0)
    for (int idx = tid; idx < nn; idx += AKS_NT) har_loc_out[(size_t)c * nn + idx] = Hs[idx];
  // H_l of every (problem, sigma, l): shared memory in, two coalesced streams out
  for (int pp = 0; pp < NP; ++pp) {
    if (!pval(pp)) continue;
    const int p = pbase + pp;
    const double Zp = b.Z[p];
    const int Lp = b.Lp[p];
    for (int e = 0; e < s; ++e) {
      const double* Ku = Ks + (size_t)(pp * s + e) * nn;
      const double* Hp = Hs + (size_t)pp * nn;
      for (int idx = tid; idx < nn; idx += AKS_NT) {
        const int lo_idx = los[idx], pk = pks[idx];
        const double a = As[lo_idx], m2 = M2s[lo_idx];
        const double cou = -Zp * M1s[lo_idx], ku = Ku[lo_idx], hp = Hp[lo_idx];
        for (int l = 0; l < Lp; ++l) {
          const size_t blk = (((size_t)p * s + e) * b.Lmax + l) * C + c;
          const double kin = a + (double)(l * (l + 1)) * m2;
          const double h = ((kin + cou) + ku) + hp;
          b.Hfull[blk * (size_t)nn + idx] = h;
          if (pk >= 0) b.Hpk[blk * d.npk + pk] = h;
        }
      }
    }
  }
}
