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
// triangle only), generator values are staged per 256-node chunk in shared memory with a row
// stride = 1 (mod 16) doubles so that the 7 distinct rows a warp touches hit distinct banks.
#pragma once
#include <cuda/barrier>

#include "aks_device.cuh"
#include "aks_types.cuh"

// node chunk staged per TMA transfer and its padded row stride in shared memory: 130 doubles keeps every
// row 16-byte aligned (bulk copies need it) and maps the 7 rows a warp touches (3*ti*130 mod 16 doubles)
// onto distinct bank groups
#define POT_QC 128
#define POT_QCS 130

__host__ __device__ inline size_t potential_smem_bytes(int n, int Q, int s) {
  const int T = (n + 2) / 3, n3 = 3 * T;
  const int Qpad = ((Q + POT_QC - 1) / POT_QC) * POT_QC;
  return sizeof(double) * ((size_t)s * Qpad + 2 * (size_t)n3 * POT_QCS + (size_t)AKS_NW * n3 * n3 +
                           (size_t)(s + 1) * n * n + n + 8);
}

__global__ void __launch_bounds__(AKS_NT)
k_potential(DiscDev d, BatchDev b, double* __restrict__ vxc_loc_out, double* __restrict__ har_loc_out,
            int p0, int force) {
  const int c = blockIdx.x, p = blockIdx.y + p0;
  if (!force && b.status[p] != 0) return;
  extern __shared__ __align__(16) double sm_pot[];
  double* sm = sm_pot;
  const int n = d.n, Q = d.Q, s = b.s, C = d.C;
  const int T = (n + 2) / 3, n3 = 3 * T, ntile = T * (T + 1) / 2;
  const int Qpad = ((Q + POT_QC - 1) / POT_QC) * POT_QC;
  double* fs = sm;                           // [s][Qpad]
  double* Gs = fs + (size_t)s * Qpad;        // [2][n3][POT_QCS]  double-buffered node chunks (16 B aligned)
  double* Kp = Gs + 2 * (size_t)n3 * POT_QCS;  // [NW][n3*n3]
  double* Ks = Kp + (size_t)AKS_NW * n3 * n3;  // [s][n*n]   XC block per spin
  double* Hs = Ks + (size_t)s * n * n;       // [n*n]      Hartree block
  double* wl = Hs + n * n;                   // [n]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ex = b.ex[p], ec = b.ec[p];
  const bool has_xc = (ex != 0) || (ec != 0);
  const double inv1 = d.inv1[c];

  // ---- 1. pointwise XC potential at the Gauss nodes: f_sigma[q] = w_q v_sigma(rho(q))
  if (has_xc) {
    const double* r0 = b.rho_q + ((size_t)(p * s + 0) * C + c) * Q;
    const double* r1 = (s == 2) ? b.rho_q + ((size_t)(p * s + 1) * C + c) * Q : nullptr;
    double* fo0 = b.fxc + ((size_t)(p * s + 0) * C + c) * Q;
    double* fo1 = (s == 2) ? b.fxc + ((size_t)(p * s + 1) * C + c) * Q : nullptr;
    for (int q = tid; q < Qpad; q += AKS_NT) {
      double f0 = 0.0, f1 = 0.0;
      if (q < Q) {
        const double wq = d.w[q];
        if (s == 1) {
          f0 = xc_vrho_unpol(fmax(r0[q], 0.0), ex, ec) * wq;
          fo0[q] = f0;
        } else {
          double vu, vd;
          xc_vrho_pol(fmax(r0[q], 0.0), fmax(r1[q], 0.0), ex, ec, vu, vd);
          f0 = vu * wq;
          f1 = vd * wq;
          fo0[q] = f0;
          fo1[q] = f1;
        }
      }
      fs[q] = f0;
      if (s == 2) fs[Qpad + q] = f1;
    }
  }

  // ---- 2. K_sigma = sum_q f_sigma[q] g g^T, 3x3 register tiles
  int ti[2], tj[2];
  for (int ts = 0; ts < 2; ++ts) {
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
  double acc[2][2][9];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[a][e][k] = 0.0;

  if (has_xc) {
    // Generator values at the nodes are streamed through shared memory in 128-node chunks by the TMA
    // engine (cp.async.bulk, one 1 KB row per transfer, completion on an mbarrier), double buffered:
    // chunk c+2 is in flight while chunk c is consumed.
    using barrier_t = cuda::barrier<cuda::thread_scope_block>;
    namespace cde = cuda::device::experimental;
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar[2];
    const int nch = Qpad / POT_QC;
    const bool tma_ok = ((Q % 2) == 0) && ((d.Qs % 2) == 0);   // 16-byte granularity of every row piece
    if (tid == 0) {
      init(&bar[0], AKS_NT);
      init(&bar[1], AKS_NT);
      cde::fence_proxy_async_shared_cta();
    }
    // rows >= n and the tail of the last chunk are never written by the copies: zero both buffers once
    for (int idx = tid; idx < 2 * n3 * POT_QCS; idx += AKS_NT) Gs[idx] = 0.0;
    __syncthreads();
    barrier_t::arrival_token tok[2];
    auto issue = [&](int ch) {   // thread 0 only
      double* dst = Gs + (size_t)(ch & 1) * n3 * POT_QCS;
      const int q0 = ch * POT_QC, len = min(POT_QC, Q - q0);
      for (int g = 0; g < n; ++g)
        cde::cp_async_bulk_global_to_shared(dst + g * POT_QCS, d.Pg + (size_t)g * d.Qs + q0, (uint32_t)(len * 8), bar[ch & 1]);
      tok[ch & 1] = cuda::device::barrier_arrive_tx(bar[ch & 1], 1, (size_t)n * len * 8);
    };
    if (tma_ok && tid == 0) {
      issue(0);
      if (nch > 1) issue(1);
    }
    for (int ch = 0; ch < nch; ++ch) {
      const int q0 = ch * POT_QC;
      double* Gb = Gs + (size_t)(ch & 1) * n3 * POT_QCS;
      if (tma_ok) {
        if (tid == 0) bar[ch & 1].wait(std::move(tok[ch & 1]));
        else bar[ch & 1].arrive_and_wait();
      } else {
        __syncthreads();
        for (int idx = tid; idx < n * POT_QC; idx += AKS_NT) {
          const int g = idx / POT_QC, j = idx - g * POT_QC;
          Gb[g * POT_QCS + j] = (q0 + j < Q) ? d.Pg[(size_t)g * d.Qs + q0 + j] : 0.0;
        }
        __syncthreads();
      }
      // hot loop: per node 6 generator values + f, 3 DMUL + 9 DFMA per spin for this lane's 3x3 tile.
      // Slot 0 covers the first 32 tiles (all of them for p <= 20); slot 1 only exists for p > 21.
      const int jq0 = warp * (POT_QC / AKS_NW);
#pragma unroll
      for (int ts = 0; ts < 2; ++ts) {
        if (ti[ts] < 0) continue;
        if (ts == 1 && ntile <= 32) break;
        const double* gi = Gb + (3 * ti[ts]) * POT_QCS + jq0;
        const double* gj = Gb + (3 * tj[ts]) * POT_QCS + jq0;
        const double* f0 = fs + q0 + jq0;
        double* A0 = acc[ts][0];
        double* A1 = acc[ts][1];
        if (s == 1) {
#pragma unroll 4
          for (int jq = 0; jq < POT_QC / AKS_NW; ++jq) {
            const double a0 = gi[jq], a1 = gi[POT_QCS + jq], a2 = gi[2 * POT_QCS + jq];
            const double fq = f0[jq];
            const double t0 = fq * gj[jq], t1 = fq * gj[POT_QCS + jq], t2 = fq * gj[2 * POT_QCS + jq];
            A0[0] = fma(a0, t0, A0[0]); A0[1] = fma(a0, t1, A0[1]); A0[2] = fma(a0, t2, A0[2]);
            A0[3] = fma(a1, t0, A0[3]); A0[4] = fma(a1, t1, A0[4]); A0[5] = fma(a1, t2, A0[5]);
            A0[6] = fma(a2, t0, A0[6]); A0[7] = fma(a2, t1, A0[7]); A0[8] = fma(a2, t2, A0[8]);
          }
        } else {
#pragma unroll 2
          for (int jq = 0; jq < POT_QC / AKS_NW; ++jq) {
            const double a0 = gi[jq], a1 = gi[POT_QCS + jq], a2 = gi[2 * POT_QCS + jq];
            const double b0 = gj[jq], b1 = gj[POT_QCS + jq], b2 = gj[2 * POT_QCS + jq];
            const double fu = f0[jq], fd = f0[Qpad + jq];
            double t0 = fu * b0, t1 = fu * b1, t2 = fu * b2;
            A0[0] = fma(a0, t0, A0[0]); A0[1] = fma(a0, t1, A0[1]); A0[2] = fma(a0, t2, A0[2]);
            A0[3] = fma(a1, t0, A0[3]); A0[4] = fma(a1, t1, A0[4]); A0[5] = fma(a1, t2, A0[5]);
            A0[6] = fma(a2, t0, A0[6]); A0[7] = fma(a2, t1, A0[7]); A0[8] = fma(a2, t2, A0[8]);
            t0 = fd * b0; t1 = fd * b1; t2 = fd * b2;
            A1[0] = fma(a0, t0, A1[0]); A1[1] = fma(a0, t1, A1[1]); A1[2] = fma(a0, t2, A1[2]);
            A1[3] = fma(a1, t0, A1[3]); A1[4] = fma(a1, t1, A1[4]); A1[5] = fma(a1, t2, A1[5]);
            A1[6] = fma(a2, t0, A1[6]); A1[7] = fma(a2, t1, A1[7]); A1[8] = fma(a2, t2, A1[8]);
          }
        }
      }
      if (tma_ok) {
        __syncthreads();   // everyone is done with this buffer: refill it with chunk ch+2
        if (tid == 0 && ch + 2 < nch) {
          // (a shorter last chunk leaves older, finite values in the tail columns; f is zero there)
          issue(ch + 2);
        }
      }
    }
  }
  // cross-warp reduction, fixed order; symmetrise diagonal tiles; scale by inv1
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if (e >= s) break;
    __syncthreads();
    if (has_xc) {
#pragma unroll
      for (int ts = 0; ts < 2; ++ts) {
        if (ti[ts] < 0) continue;
        double* K = Kp + (size_t)warp * n3 * n3;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb)
            K[(3 * ti[ts] + a) * n3 + 3 * tj[ts] + bb] = acc[ts][e][3 * a + bb];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += AKS_NT) {
      const int i = idx / n, j = idx - i * n;
      double v = 0.0;
      if (has_xc) {
        const int ii = (i / 3 <= j / 3) ? i : j, jj = (i / 3 <= j / 3) ? j : i;
        double s1 = 0.0, s2 = 0.0;
        for (int wv = 0; wv < AKS_NW; ++wv) {
          s1 += Kp[(size_t)wv * n3 * n3 + ii * n3 + jj];
          if (i / 3 == j / 3) s2 += Kp[(size_t)wv * n3 * n3 + jj * n3 + ii];
        }
        v = (i / 3 == j / 3) ? 0.5 * (s1 + s2) : s1;
        v *= inv1;
      }
      Ks[e * n * n + idx] = v;
    }
  }

  // ---- 3. Hartree block: coeff * (sum_m W_m F[ij][m] + N/Rmax M0)
  const double coeff = b.hartree[p];
  if (tid < n) {
    const int gi = d.l2g[c * n + tid];
    wl[tid] = (gi >= 0 && coeff != 0.0) ? b.Wv[(size_t)p * d.Nh + gi] : 0.0;
  }
  __syncthreads();
  {
    const double shift = b.N[p] / d.Rmax;
    const double* Fc = d.F_loc + (size_t)c * n * n * n;
    const double* M0c = d.M0_loc + (size_t)c * n * n;
    for (int idx = tid; idx < n * n; idx += AKS_NT) {
      double v = 0.0;
      if (coeff != 0.0) {
        double fw = 0.0;
        for (int m = 0; m < n; ++m) fw = fma(wl[m], Fc[(size_t)m * n * n + idx], fw);
        v = (fw + shift * M0c[idx]) * coeff;
      }
      Hs[idx] = v;
    }
  }
  __syncthreads();

  // optional local outputs for the step-level hooks
  if (vxc_loc_out)
    for (int e = 0; e < s; ++e)
      for (int idx = tid; idx < n * n; idx += AKS_NT)
        vxc_loc_out[((size_t)e * C + c) * n * n + idx] = Ks[e * n * n + idx];
  if (har_loc_out)
    for (int idx = tid; idx < n * n; idx += AKS_NT) har_loc_out[(size_t)c * n * n + idx] = Hs[idx];

  // ---- 4. H_l = (Kin_l + Coulomb) + Vxc + Hartree, packed per (sigma, l, cell)
  const double Zp = b.Z[p];
  const int Lp = b.Lp[p];
  const double* Ac = d.A_loc + (size_t)c * n * n;
  const double* M1c = d.Mm1_loc + (size_t)c * n * n;
  const double* M2c = d.Mm2_loc + (size_t)c * n * n;
  for (int e = 0; e < s; ++e) {
    for (int l = 0; l < Lp; ++l) {
      double* out = b.Hpk + ((((size_t)p * s + e) * b.Lmax + l) * C + c) * d.npk;
      double* outf = b.Hfull + ((((size_t)p * s + e) * b.Lmax + l) * C + c) * (size_t)n * n;
      const double ll = (double)(l * (l + 1));
      for (int idx = tid; idx < n * n; idx += AKS_NT) {
        const int lo_idx = d.lo_tab[idx];   // index of the lower-triangle twin: exactly symmetric output
        const int pk = d.pk_tab[idx];       // packed position, or -1 above the diagonal
        const double kin = (Ac[lo_idx] + ll * M2c[lo_idx]) * 0.5;
        const double cou = -Zp * M1c[lo_idx];
        const double h = ((kin + cou) + Ks[e * n * n + lo_idx]) + Hs[lo_idx];
        outf[idx] = h;
        if (pk >= 0) out[pk] = h;
      }
    }
  }
}
