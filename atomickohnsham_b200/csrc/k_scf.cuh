// Remaining kernels of one SCF step: aufbau, orbital energies, trial density in every
// representation the path needs, radial Poisson solve, energies + ODA line search, mixing
// and the stopping rule.  Reference functions replaced are cited per kernel.
#pragma once
#include <type_traits>
#include <float.h>

#include "aks_device.cuh"
#include "aks_types.cuh"

__device__ __forceinline__ size_t orb_off(const BatchDev& b, int p, int e, int l, int k) {
  return (((size_t)p * b.s + e) * b.Lmax + l) * b.nhmax + k;
}
__device__ __forceinline__ size_t occ_off(const BatchDev& b, int p, int slot, int e, int l, int k) {
  return ((((size_t)p * 2 + slot) * b.s + e) * b.Lmax + l) * b.nhmax + k;
}

// compact list of the occupied orbitals of (problem, slot), density! order (sigma, k, l); written by
// one thread of k_aufbau (both slots) and again by k_linesearch (slot 0, final occupations)
__device__ inline void write_occ_list(const BatchDev& b, int p, int slot) {
  int cnt = 0;
  int* elk = b.occ_elk + ((size_t)p * 2 + slot) * AKS_MAXORB;
  double* nn = b.occ_n + ((size_t)p * 2 + slot) * AKS_MAXORB;
  for (int e = 0; e < b.s; ++e)
    for (int k = 0; k < b.nhp[p]; ++k)
      for (int l = 0; l < b.Lp[p]; ++l) {
        const double v = b.occ[occ_off(b, p, slot, e, l, k)];
        if (v != 0.0 && cnt < AKS_MAXORB) { elk[cnt] = (e << 16) | (l << 8) | k; nn[cnt] = v; ++cnt; }
      }
  b.occ_cnt[2 * p + slot] = cnt;
}

// ------------------------------------------------------------------------------------------
// k_aufbau: aufbau!(n, eps, ..., ::OptimizedAufbauCache, niter)  (src/algorithms/aufbau/
// optimized.jl:99-222) -- stable sort of all eps, greedy filling by blocks of <= max_degen
// levels within `tol` of the block head; a partially filled two-fold block only records the
// two extreme fillings here (slot 0 / slot 1), the energy minimisation between them
// (optimized.jl:228-306) happens in k_linesearch.  One thread per problem.
// mode 0: all running problems, on the eigenpair window (k_eig); the result is kept only if it is
// certified: every level outside the window lies above everything the fill examined.  mode 1: the
// problems whose certificate failed, after their remaining levels were solved.
__global__ void __launch_bounds__(32) k_aufbau(DiscDev d, BatchDev b, int mode) {
  const int p = blockIdx.x;
  if (p >= b.B || b.status[p] != 0) return;
  if (mode == 1 && b.redo[p] == 0) {
    // the window pass was certified (k_eig_cert): commit what it had to defer
    if (threadIdx.x == 0 && b.degen3[p]) { b.status[p] = AKS_EDEGEN3; b.degen3[p] = 0; }
    return;
  }
  const int Lp = b.Lp[p], nhp = b.nhp[p], s = b.s;
  const int norb = Lp * nhp * s;
  // stage everything the serial logic touches in shared memory (one warp per problem): the
  // sort / fill below then never waits on global memory
  __shared__ double ev[AKS_MAXORB];
  __shared__ double occ_s[2][AKS_MAXORB];
  __shared__ int ord[AKS_MAXORB];
  const int lane = threadIdx.x;
  // flat index f = l + Lp*k + Lp*nhp*e  (convert_index, discretization.jl:379-392)
  for (int f = lane; f < norb; f += 32) {
    const int e = f / (Lp * nhp), r = f - e * Lp * nhp, k = r / Lp, l = r - k * Lp;
    const double v = b.eps_new[orb_off(b, p, e, l, k)];
    if (v < 0.5 * AKS_EPS_MISSING) b.eps[orb_off(b, p, e, l, k)] = v;   // unsolved levels keep their last value
    ev[f] = v;
    ord[f] = f;
    occ_s[0][f] = 0.0;
    occ_s[1][f] = 0.0;
  }
  __syncwarp();
  int degen = 0;
  if (b.aufbau_kind != 0) {
    // ---- SmearedAufbau (aufbau/smeared.jl:68-101): n_i = g_i / (1 + exp((eps_i - mu)/T)), mu by 100
    // bisection steps on [min eps - 50 T, max eps + 50 T];  FrozenAufbau (aufbau/frozen.jl:84-88): n = nfix
    if (b.aufbau_kind == 1) {
      const double T = b.temp, Np = b.N[p];
      double emin = DBL_MAX, emax = -DBL_MAX;
      for (int f = lane; f < norb; f += 32) { emin = fmin(emin, ev[f]); emax = fmax(emax, ev[f]); }
      emin = -warp_max(-emin); emax = warp_max(emax);
      double lo = emin - 50.0 * T, hi = emax + 50.0 * T;
      for (int itb = 0; itb < 100; ++itb) {
        const double mid = (lo + hi) / 2.0;
        double part = 0.0;
        for (int f = lane; f < norb; f += 32) {
          const int l = (f % (Lp * nhp)) % Lp;
          const double g = (s == 1) ? (double)(4 * l + 2) : (double)(2 * l + 1);
          part += g / (1.0 + exp((ev[f] - mid) / T));
        }
        const double occ = warp_sum(part);
        if (occ < Np) lo = mid; else hi = mid;
      }
      const double mu = (lo + hi) / 2.0;
      for (int f = lane; f < norb; f += 32) {
        const int l = (f % (Lp * nhp)) % Lp;
        const double g = (s == 1) ? (double)(4 * l + 2) : (double)(2 * l + 1);
        occ_s[0][f] = g / (1.0 + exp((ev[f] - mu) / T));
      }
    } else {
      for (int f = lane; f < norb; f += 32) {
        const int e = f / (Lp * nhp), r = f - e * Lp * nhp, k = r / Lp, l = r - k * Lp;
        occ_s[0][f] = b.nfix[orb_off(b, p, e, l, k)];
      }
    }
    __syncwarp();
    if (lane == 0) {
      b.redo[p] = 0;
      b.full_refill[p] = 0;
      b.degen[p] = 0;
      int cnt = 0;
      int* elk = b.occ_elk + ((size_t)p * 2) * AKS_MAXORB;
      double* nn = b.occ_n + ((size_t)p * 2) * AKS_MAXORB;
      for (int e = 0; e < s; ++e)
        for (int k = 0; k < nhp; ++k)
          for (int l = 0; l < Lp; ++l) {
            const double v = occ_s[0][l + Lp * k + Lp * nhp * e];
            if (v != 0.0) { elk[cnt] = (e << 16) | (l << 8) | k; nn[cnt] = v; ++cnt; }
          }
      b.occ_cnt[2 * p] = cnt;
      b.occ_cnt[2 * p + 1] = 0;
      // eigenpair window: smeared occupations need every level; frozen ones the prescribed levels
      for (int e = 0; e < s; ++e)
        for (int l = 0; l < Lp; ++l) {
          int top = 0;
          for (int k = 0; k < nhp; ++k)
            if (occ_s[0][l + Lp * k + Lp * nhp * e] != 0.0) top = k + 1;
          b.kact[((size_t)p * s + e) * b.Lmax + l] =
              (b.aufbau_kind == 1 || b.eig_margin < 0) ? b.nhmax : max(1, top);
        }
    }
  } else if (lane == 0) {
    for (int i = 1; i < norb; ++i) {  // stable insertion sort (sortperm!)
      const int oi = ord[i];
      const double vi = ev[oi];
      int j = i - 1;
      while (j >= 0 && ev[ord[j]] > vi) { ord[j + 1] = ord[j]; --j; }
      ord[j + 1] = oi;
    }
    double Nleft = b.N[p];
    const int niter = b.niter[p];
    int i = 0;
    int blk[8];
    const int maxd = b.max_degen < 8 ? b.max_degen : 8;
    double e_last = -1e300;   // level of the last block the fill looked at
    bool degen3 = false;
    while (Nleft > 0.0 && i < norb) {
      blk[0] = ord[i];
      int nbd = 1;
      const double e0 = ev[blk[0]];
      e_last = e0;
      int j = i + 1;
      while (j < norb && fabs(ev[ord[j]] - e0) < b.degen_tol && nbd < maxd) { blk[nbd++] = ord[j]; ++j; }
      i += nbd;
      int degs[8], tot = 0, ee[8], ll[8], kk[8];
      for (int m = 0; m < nbd; ++m) {
        const int f = blk[m];
        ee[m] = f / (Lp * nhp);
        const int r = f - ee[m] * Lp * nhp;
        kk[m] = r / Lp;
        ll[m] = r - kk[m] * Lp;
        degs[m] = (s == 1) ? 4 * ll[m] + 2 : 2 * ll[m] + 1;
        tot += degs[m];
      }
      if (Nleft >= (double)tot) {
        for (int m = 0; m < nbd; ++m) { occ_s[0][blk[m]] = (double)degs[m]; occ_s[1][blk[m]] = (double)degs[m]; }
        Nleft -= (double)tot;
      } else if (nbd == 1) {
        occ_s[0][blk[0]] = Nleft;
        break;
      } else if (nbd == 2) {
        if (b.handle_spin && niter == 0 && ll[0] == ll[1] && kk[0] == kk[1] && ee[0] != ee[1]) {
          occ_s[0][blk[0]] = Nleft / 2.0;
          occ_s[0][blk[1]] = Nleft / 2.0;
          break;
        }
        degen = 1;
        const double n10 = ((double)degs[0] >= Nleft) ? Nleft : (double)degs[0];
        const double n20 = Nleft - n10;
        const double n21 = ((double)degs[1] >= Nleft) ? Nleft : (double)degs[1];
        const double n11 = Nleft - n21;
        occ_s[0][blk[0]] = n10; occ_s[0][blk[1]] = n20;
        occ_s[1][blk[0]] = n11; occ_s[1][blk[1]] = n21;
        b.degen_idx[2 * p] = (ee[0] * b.Lmax + ll[0]) * b.nhmax + kk[0];
        b.degen_idx[2 * p + 1] = (ee[1] * b.Lmax + ll[1]) * b.nhmax + kk[1];
        b.degen_n[4 * p] = n10; b.degen_n[4 * p + 1] = n20;
        b.degen_n[4 * p + 2] = n11; b.degen_n[4 * p + 3] = n21;
        break;
      } else {
        degen3 = true;
        break;
      }
    }
    // Certificate of the window (mode 0): no unsolved level may lie below e_cert = (last level the fill
    // examined) + degen_tol.  k_eig_cert checks that with one inertia count per truncated channel; here only
    // the trivial failure (the fill ran out of solved levels) is decided.
    bool redo = false;
    if (mode == 0 && b.eig_margin >= 0) {
      redo = !(e_last < 0.5 * AKS_EPS_MISSING);
      b.full_refill[p] = redo ? 1 : 0;
      b.e_cert[p] = e_last + b.degen_tol;
    }
    b.redo[p] = redo ? 1 : 0;
    if (redo) b.redo_cnt[p] += 1;
    if (!redo) {
      if (mode == 0 && b.eig_margin >= 0) b.degen3[p] = degen3 ? 1 : 0;   // committed once certified
      else if (degen3) b.status[p] = AKS_EDEGEN3;
      // window of the next step: occupied levels (either slot) + margin
      for (int e = 0; e < s; ++e)
        for (int l = 0; l < Lp; ++l) {
          int top = 0;
          for (int k = 0; k < nhp; ++k) {
            const int f = l + Lp * k + Lp * nhp * e;
            if (occ_s[0][f] != 0.0 || occ_s[1][f] != 0.0) top = k + 1;
          }
          // highest need of the last AKS_KHIST steps: configurations that alternate between steps
          // (open d / f shells) would otherwise fail the certificate every other step
          const size_t chn = ((size_t)p * s + e) * b.Lmax + l;
          int* hist = b.khist + chn * AKS_KHIST;
          hist[niter % AKS_KHIST] = top;
          int need = top;
          for (int h = 0; h < AKS_KHIST; ++h) need = max(need, hist[h]);
          b.kact[chn] = (b.eig_margin < 0) ? b.nhmax : max(1, min(nhp, need + b.eig_margin));
        }
    }
    b.degen[p] = degen;
    // compact occupied lists, density! order (sigma, k, l)
    for (int slot = 0; slot < 2; ++slot) {
      int cnt = 0;
      int* elk = b.occ_elk + ((size_t)p * 2 + slot) * AKS_MAXORB;
      double* nn = b.occ_n + ((size_t)p * 2 + slot) * AKS_MAXORB;
      if (slot == 0 || degen)
        for (int e = 0; e < s; ++e)
          for (int k = 0; k < nhp; ++k)
            for (int l = 0; l < Lp; ++l) {
              const double v = occ_s[slot][l + Lp * k + Lp * nhp * e];
              if (v != 0.0) { elk[cnt] = (e << 16) | (l << 8) | k; nn[cnt] = v; ++cnt; }
            }
      b.occ_cnt[2 * p + slot] = cnt;
    }
  }
  __syncwarp();
  // occupations back to the [slot][sigma][Lmax][nhmax] layout (zero where a problem has no orbital)
  for (int idx = lane; idx < 2 * s * b.Lmax * b.nhmax; idx += 32) {
    const int k = idx % b.nhmax, l = (idx / b.nhmax) % b.Lmax, e = (idx / (b.nhmax * b.Lmax)) % s, slot = idx / (b.nhmax * b.Lmax * s);
    double v = 0.0;
    if (l < Lp && k < nhp) v = occ_s[slot][l + Lp * k + Lp * nhp * e];
    b.occ[occ_off(b, p, slot, e, l, k)] = v;
  }
}

// occupied-orbital list of (problem, slot) in the reference's density! order (sigma, k, l)
struct OccList {
  int count;
  int e[AKS_MAXORB], l[AKS_MAXORB], k[AKS_MAXORB];
  double n[AKS_MAXORB];
};
__device__ inline void build_occ_list(const BatchDev& b, int p, int slot, OccList* L) {
  const int cnt = b.occ_cnt[2 * p + slot];
  const int* elk = b.occ_elk + ((size_t)p * 2 + slot) * AKS_MAXORB;
  const double* nn = b.occ_n + ((size_t)p * 2 + slot) * AKS_MAXORB;
  for (int o = threadIdx.x; o < cnt; o += blockDim.x) {
    const int v = elk[o];
    L->e[o] = v >> 16; L->l[o] = (v >> 8) & 0xff; L->k[o] = v & 0xff; L->n[o] = nn[o];
  }
  if (threadIdx.x == 0) L->count = cnt;
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// k_dens_cells: trial density of one (cell, problem, slot) straight from the orbitals:
//   rho_sigma(r_q) = sum_k n_k (sum_i u_ik g_i(q))^2 / r_q^2 / (4 pi)
// (what optimized_eval_density!, assemble.jl:91-125, yields for D = sum n u u^T), the energy
// correction <Vxc[D_prev], D_new> (energies.jl:15-24) as sum_q fxc(q) q_new(q), and the cell's
// share of the Hartree rhs B = F:D (tensor_matrix_dict!, src/utils/free allocation.jl:17-36).
#define DENS_NT 128
template <int NMAX>
__global__ void __launch_bounds__(DENS_NT, 3) k_dens_cells(DiscDev d, BatchDev b) {
  const int c = blockIdx.x, p = blockIdx.y, slot = blockIdx.z;
  if (b.status[p] != 0) return;
  if (slot == 1 && !b.degen[p]) return;
  __shared__ OccList L;
  extern __shared__ __align__(16) double sm_dc[];
  double* sm = sm_dc;
  const int n = d.n, Q = d.Q, C = d.C, s = b.s;
  double* us = sm;                              // [count][NMAX]
  double* Dl = us + (size_t)AKS_MAXORB * NMAX;  // [n*n]
  double* red = Dl + n * n;                     // [NW]
  build_occ_list(b, p, slot, &L);
  const int cnt = L.count;
  for (int idx = threadIdx.x; idx < cnt * NMAX; idx += DENS_NT) {
    const int o = idx / NMAX, i = idx - o * NMAX;
    double v = 0.0;
    if (i < n) {
      const int gi = d.l2g[c * n + i];
      if (gi >= 0) v = b.U[orb_off(b, p, L.e[o], L.l[o], L.k[o]) * d.Nh + gi];
    }
    us[idx] = v * sqrt(L.n[o]);   // occupation folded in: rho = sum_o (sqrt(n_o) psi_o)^2, D_loc = us^T us
  }
  __syncthreads();
  // the list is ordered by spin (density! order: sigma, k, l): orbitals [0, cnt_up) have sigma = 0
  int cnt_up = cnt;
  for (int o = 0; o < cnt; ++o)
    if (L.e[o] == 1) { cnt_up = o; break; }
  const double inv1 = d.inv1[c], inv2 = d.inv2[c];
  const bool has_xc = (b.ex[p] != 0) || (b.ec[p] != 0);
  double corr = 0.0;
  // two nodes per thread: every (128-bit, broadcast) shared load of orbital coefficients feeds four FMAs --
  // with one node the kernel is bound by the shared-memory issue rate, not by the FP64 pipe
  for (int qa = threadIdx.x; qa < Q; qa += 2 * DENS_NT) {
    const int qb = qa + DENS_NT;
    const bool hb = qb < Q;
    double ga[NMAX], gb[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      ga[i] = (i < n) ? d.Pg[(size_t)i * d.Qs + qa] : 0.0;
      gb[i] = (i < n && hb) ? d.Pg[(size_t)i * d.Qs + qb] : 0.0;
    }
    double ra0 = 0.0, ra1 = 0.0, rb0 = 0.0, rb1 = 0.0;
    auto psi2 = [&](int o, double& ra, double& rb) {
      const double2* uo = reinterpret_cast<const double2*>(us + (size_t)o * NMAX);   // broadcast LDS.128
      double pa0 = 0.0, pa1 = 0.0, pb0 = 0.0, pb1 = 0.0;
#pragma unroll
      for (int i = 0; i < NMAX / 2; ++i) {
        const double2 u2 = uo[i];
        pa0 = fma(u2.x, ga[2 * i], pa0);
        pa1 = fma(u2.y, ga[2 * i + 1], pa1);
        pb0 = fma(u2.x, gb[2 * i], pb0);
        pb1 = fma(u2.y, gb[2 * i + 1], pb1);
      }
      const double psa = pa0 + pa1, psb = pb0 + pb1;
      ra = fma(psa, psa, ra);
      rb = fma(psb, psb, rb);
    };
    for (int o = 0; o < cnt_up; ++o) psi2(o, ra0, rb0);
    for (int o = cnt_up; o < cnt; ++o) psi2(o, ra1, rb1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = h ? qb : qa;
      if (h && !hb) break;
      const double r0 = h ? rb0 : ra0, r1 = h ? rb1 : ra1;
      const double X = inv1 * d.x[q] + inv2;
      double* out0 = b.rho_q_new + ((((size_t)p * 2 + slot) * s + 0) * C + c) * Q;
      out0[q] = (r0 / (X * X)) * c_xc.inv4pi;
      if (has_xc) corr = fma(b.fxc[((size_t)(p * s + 0) * C + c) * Q + q], r0, corr);
      if (s == 2) {
        double* out1 = b.rho_q_new + ((((size_t)p * 2 + slot) * s + 1) * C + c) * Q;
        out1[q] = (r1 / (X * X)) * c_xc.inv4pi;
        if (has_xc) corr = fma(b.fxc[((size_t)(p * s + 1) * C + c) * Q + q], r1, corr);
      }
    }
  }
  corr = block_sum(corr, red);
  if (threadIdx.x == 0) b.corr_part[((size_t)p * 2 + slot) * C + c] = corr * inv1;
  // local density block (both spins summed) and its contraction with the mass tensor
  if (b.hartree[p] != 0.0) {
    for (int idx = threadIdx.x; idx < n * n; idx += DENS_NT) {
      const int i = idx / n, j = idx - i * n;
      double v = 0.0;
      for (int o = 0; o < cnt; ++o) v = fma(us[(size_t)o * NMAX + i], us[(size_t)o * NMAX + j], v);
      Dl[idx] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Fc = d.F_loc + (size_t)c * n * n * n;
    // a warp owns the rows m = warp, warp + 4, ...: all of them advance together so that up to 8 loads of
    // the (L2-resident) tensor block are in flight per lane instead of one
    constexpr int NWD = DENS_NT / 32, MPW = (32 + NWD - 1) / NWD;
    double acc[MPW];
#pragma unroll
    for (int j = 0; j < MPW; ++j) acc[j] = 0.0;
    for (int idx = lane; idx < n * n; idx += 32) {
      const double dv = Dl[idx];
#pragma unroll
      for (int j = 0; j < MPW; ++j) {
        const int m = warp + NWD * j;
        if (m < n) acc[j] = fma(dv, Fc[(size_t)m * n * n + idx], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < MPW; ++j) {
      const int m = warp + NWD * j;
      if (m < n) {   // warp-uniform
        const double v = warp_sum(acc[j]);
        if (lane == 0) b.Bloc_new[(((size_t)p * 2 + slot) * C + c) * n + m] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_dens_cells_tc + k_fd: the same three outputs with the orbital values at the nodes on the FP64 tensor cores, for
// orders p <= 23 (n <= 24 generators) and Q a multiple of 8.
//
//   psi[q][o] = sum_i G[i][q] us[o][i]  as  C(8 nodes x 8 orbitals) += A(8 nodes x 4 generators) B(4 generators x 8
//   orbitals), mma.sync m8n8k4: six steps over the generators per 8-node group and 8-orbital tile.
//
// A CTA owns four cells of one (problem, slot): computing warp w owns cell c0 + w, a fifth warp streams the generator
// table through shared memory in 64-node chunks (TMA row copies, DENS_TC_ST stages, full / empty mbarrier pair per stage;
// the table is the same for every cell and problem, so one stream feeds the four cells).  The B fragments -- orbital
// coefficients of the warp's cell, occupation folded in, 16 orbitals of one spin per sweep over the nodes (one sweep for
// every atom of the periodic table without spin) -- are gathered from U straight into registers; the A fragments are
// conflict-free shared loads (row stride 68).  Two node groups advance together (up to four independent DMMA chains).  A
// lane ends a chain with psi[node lane / 4][orbitals 2 (lane % 4) + {0, 1}]: squares, two shuffles over lane % 4, then
// lane % 4 = 0 / 1 finish the node of the first / second group: division by r^2, store, <Vxc[D_prev], D_new>.
// Against k_dens_cells: no 48 generator values per thread (168 registers, 3 CTAs per SM), no reload of the table per cell.
// The mass-tensor contraction B = F:D moves to k_fd, which reads F_c once for four problems.
#define DENS_TC_ROWS 24
#define DENS_TC_QC 64
#define DENS_TC_QCS 68     // row stride of a node chunk: 4 mod 16 doubles
#define DENS_TC_ST 3
#define DENS_TC_NC 4       // cells (= computing warps) per CTA
#define DENS_TC_NT (32 * DENS_TC_NC + 32)
__host__ __device__ inline bool dens_tc_ok(int n, int Q, int Qs) { return n <= DENS_TC_ROWS && (Q % 8) == 0 && (Qs % 2) == 0; }
#define DENS_TC_STAGE (POT_IMG_ROWS * DENS_TC_QCS + DENS_TC_NC * DENS_TC_QC)   // chunk image (table + nodes), then f of the 4 cells
__host__ __device__ inline size_t dens_tc_smem_bytes() {
  return sizeof(double) * ((size_t)DENS_TC_ST * DENS_TC_STAGE + 8);
}
template <int SP>   // nspin
__global__ void __launch_bounds__(DENS_TC_NT, 3) k_dens_cells_tc(DiscDev d, BatchDev b) {
  const int c0 = blockIdx.x * DENS_TC_NC, p = blockIdx.y;
  if (b.status[p] != 0) return;
  const bool degen = b.degen[p] != 0;
  __shared__ OccList L;                    // union of the orbitals of the two slots; L.n = occupations of slot 0
  __shared__ double w1s[AKS_MAXORB];       // occupations of slot 1
  __shared__ __align__(8) uint64_t full_bar[DENS_TC_ST], empty_bar[DENS_TC_ST];
  extern __shared__ __align__(16) double sm_dc[];
  constexpr int NR = DENS_TC_ROWS, QCS = DENS_TC_QCS, ST = DENS_TC_ST, QC = DENS_TC_QC;
  const int n = d.n, Q = d.Q, C = d.C;
  double* Gs = sm_dc;                                   // [ST][25 x 68 image | 4 x 64 f]
  constexpr int STAGE = DENS_TC_STAGE;
#ifdef AKS_DENS_TIMING   // cycle stamps of thread 0 (tools/gpu_dens_cycles.py); problems 0..63
  long long* stamp_ = b.eig_dbg + ((size_t)p * gridDim.x + blockIdx.x) * 8; int nstamp_ = 0;
#define DENS_STAMP() do { if (threadIdx.x == 0 && p < 64 && nstamp_ < 8) stamp_[nstamp_++] = clock64(); } while (0)
#else
#define DENS_STAMP() do { } while (0)
#endif
  DENS_STAMP();
  // Both trial densities of a two-fold degenerate problem (slots 0 and 1: the two fillings of the degenerate pair,
  // optimized.jl:205-262) come out of ONE pass: psi_o(q) is computed once, rho_slot = sum_o n_o^slot psi_o^2.
  build_occ_list(b, p, 0, &L);
  if (degen) {
    if (threadIdx.x == 0) {
      int cnt = L.count;
      for (int o = 0; o < cnt; ++o) w1s[o] = 0.0;
      const int cnt1 = b.occ_cnt[2 * p + 1];
      const int* elk1 = b.occ_elk + ((size_t)p * 2 + 1) * AKS_MAXORB;
      const double* n1 = b.occ_n + ((size_t)p * 2 + 1) * AKS_MAXORB;
      bool appended = false;
      for (int j = 0; j < cnt1; ++j) {
        const int v = elk1[j];
        int o = 0;
        while (o < cnt && ((L.e[o] << 16) | (L.l[o] << 8) | L.k[o]) != v) ++o;
        if (o == cnt && cnt < AKS_MAXORB) {
          L.e[o] = v >> 16; L.l[o] = (v >> 8) & 0xff; L.k[o] = v & 0xff; L.n[o] = 0.0;
          ++cnt; appended = true;
        }
        if (o < AKS_MAXORB) w1s[o] = n1[j];
      }
      if (SP == 2 && appended)   // keep the list ordered by spin (stable insertion sort on sigma)
        for (int i = 1; i < cnt; ++i)
          for (int j = i; j > 0 && L.e[j - 1] > L.e[j]; --j) {
            int t;
            double tn;
            t = L.e[j]; L.e[j] = L.e[j - 1]; L.e[j - 1] = t;
            t = L.l[j]; L.l[j] = L.l[j - 1]; L.l[j - 1] = t;
            t = L.k[j]; L.k[j] = L.k[j - 1]; L.k[j - 1] = t;
            tn = L.n[j]; L.n[j] = L.n[j - 1]; L.n[j - 1] = tn;
            tn = w1s[j]; w1s[j] = w1s[j - 1]; w1s[j - 1] = tn;
          }
      L.count = cnt;
    }
    __syncthreads();
  }
  const int cnt = L.count;
  // the list is ordered by spin (density! order: sigma, k, l): orbitals [0, cnt_up) have sigma = 0
  int cnt_up = cnt;
  if (SP == 2)
    for (int o = 0; o < cnt; ++o)
      if (L.e[o] == 1) { cnt_up = o; break; }
  // sweeps over the nodes: per spin, 16 orbitals each (at least one: a spin without electrons still writes rho = 0)
  int nsw[2];
  nsw[0] = max(1, (cnt_up + 15) >> 4);
  nsw[1] = SP == 2 ? max(1, (cnt - cnt_up + 15) >> 4) : 0;
  const int nch = (Q + QC - 1) / QC, nvalid = min(DENS_TC_NC, C - c0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int st = 0; st < ST; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], nvalid); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  DENS_STAMP();
  const bool has_xc = (b.ex[p] != 0) || (b.ec[p] != 0);
  if (warp == DENS_TC_NC) {   // producer: (nsw[0] + nsw[1]) passes over the table: one bulk copy per chunk image, and in the
    if (lane == 0) {          // last sweep of a spin the f rows of the four cells (<Vxc[D_prev], D_new>)
      int it = 0;
      for (int e = 0; e < SP; ++e)
        for (int sw = 0; sw < nsw[e]; ++sw)
          for (int ch = 0; ch < nch; ++ch, ++it) {
            const int st = it % ST;
            if (it >= ST) mbar_wait(&empty_bar[st], (uint32_t)(((it - ST) / ST) & 1));
            const bool wf = has_xc && sw == nsw[e] - 1;
            const int q0 = ch * QC, len = min(QC, Q - q0);
            double* dst = Gs + (size_t)st * STAGE;
            mbar_expect_tx(&full_bar[st], (uint32_t)(POT_IMG_ROWS * QCS * 8 + (wf ? nvalid * len * 8 : 0)));
            bulk_g2s(dst, d.Pgc + (size_t)ch * POT_IMG_ROWS * QCS, (uint32_t)(POT_IMG_ROWS * QCS * 8), &full_bar[st]);
            if (wf)
              for (int u = 0; u < nvalid; ++u)
                bulk_g2s(dst + POT_IMG_ROWS * QCS + u * QC, b.fxc + ((size_t)(p * SP + e) * C + c0 + u) * Q + q0, (uint32_t)(len * 8),
                         &full_bar[st]);
          }
    }
    return;
  }
  const int c = c0 + warp;
  if (c >= C) return;
  const int lr = lane >> 2, lk = lane & 3;
  const double inv1 = d.inv1[c], inv2 = d.inv2[c];
  int gi[6];   // global index of the generators 4 ks + lane % 4 of this cell (-1: none)
#pragma unroll
  for (int ks = 0; ks < 6; ++ks) gi[ks] = (4 * ks + lk < n) ? d.l2g[c * n + 4 * ks + lk] : -1;
  double corr0 = 0.0, corr1 = 0.0;
  int it = 0;   // chunk counter of the stream
  for (int e = 0; e < SP; ++e) {
    const int o0 = e == 0 ? 0 : cnt_up, o1 = e == 0 ? cnt_up : cnt;
    double* out0 = b.rho_q_new + ((((size_t)p * 2 + 0) * SP + e) * C + c) * Q;
    double* out1 = b.rho_q_new + ((((size_t)p * 2 + 1) * SP + e) * C + c) * Q;
    for (int sw = 0; sw < nsw[e]; ++sw) {
      const bool first = sw == 0, last = sw == nsw[e] - 1;
      // B fragments: U_o[i], orbital o0 + 16 sw + 8 t + lane / 4, generator 4 ks + lane % 4
      double ba[6], bb[6];
      const int oa = o0 + 16 * sw + lr, ob = oa + 8;
      const bool two = o0 + 16 * sw + 8 < o1;
      {
        const double* Ua = oa < o1 ? b.U + orb_off(b, p, L.e[oa], L.l[oa], L.k[oa]) * d.Nh : nullptr;
        const double* Ub = ob < o1 ? b.U + orb_off(b, p, L.e[ob], L.l[ob], L.k[ob]) * d.Nh : nullptr;
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
          ba[ks] = (Ua && gi[ks] >= 0) ? Ua[gi[ks]] : 0.0;
          bb[ks] = (Ub && gi[ks] >= 0) ? Ub[gi[ks]] : 0.0;
        }
      }
      // occupations of the orbitals whose values this lane ends up with (columns 2 (lane % 4) + {0, 1} of the two tiles)
      double w0[2][2], w1[2][2];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int o = o0 + 16 * sw + 8 * t + 2 * lk + j;
          w0[t][j] = o < o1 ? L.n[o] : 0.0;
          w1[t][j] = (degen && o < o1) ? w1s[o] : 0.0;
        }
      for (int ch = 0; ch < nch; ++ch, ++it) {
        const int st = it % ST;
        mbar_wait(&full_bar[st], (uint32_t)((it / ST) & 1));
        const double* Gb = Gs + (size_t)st * STAGE + lk * QCS + lr;
        const double* xs = Gs + (size_t)st * STAGE + NR * QCS;                       // nodes of the chunk (image row 24)
        const double* fs = Gs + (size_t)st * STAGE + POT_IMG_ROWS * QCS + warp * QC;  // f of this warp's cell
        const int q0 = ch * QC, ngr = min(QC, Q - q0) >> 3;
        // Eight node groups at a time for one orbital tile, four for two (sixteen accumulator chains per warp; the even and
        // the odd steps over the generators of a (group, tile) accumulate apart and are added at the end).  With two groups
        // per iteration the time per group did not depend on the number of DMMAs at all: it was the per-iteration work
        // around them -- loop and address arithmetic, the two shuffles per group, a finalisation that kept 8 or 16 of 32 lanes
        // busy -- that bound the kernel.  A batch of 8 / 4 groups shares that work and finishes one node per lane.  (A
        // dependent DMMA costs 26 cycles and a warp issues one per 16-17.5, tools/micro/latency.cu: latency was never it.)
        auto run = [&](auto ng_c, auto nt_c) {
          constexpr int NG = decltype(ng_c)::value, NTL = decltype(nt_c)::value, NH = (NG + 3) / 4;
          for (int gp = 0; gp < ngr; gp += NG) {
            double acc[NG][NTL][2][2];
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
              for (int t = 0; t < NTL; ++t) acc[g][t][0][0] = acc[g][t][0][1] = acc[g][t][1][0] = acc[g][t][1][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 6; ++ks)
#pragma unroll
              for (int g = 0; g < NG; ++g) {
                const double av = gp + g < ngr ? Gb[(4 * ks) * QCS + 8 * (gp + g)] : 0.0;
                dmma_m8n8k4(acc[g][0][ks & 1], av, ba[ks]);
                if (NTL == 2) dmma_m8n8k4(acc[g][NTL - 1][ks & 1], av, bb[ks]);
              }
            // lane % 4 = k finishes node lane / 4 of the groups gp + k (+ 4)
            double my0[NH], my1[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) my0[h] = my1[h] = 0.0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              double r0 = 0.0, r1 = 0.0;
#pragma unroll
              for (int t = 0; t < NTL; ++t) {
                const double p0 = acc[g][t][0][0] + acc[g][t][1][0], p1 = acc[g][t][0][1] + acc[g][t][1][1];
                const double s0 = p0 * p0, s1 = p1 * p1;
                r0 = fma(w0[t][0], s0, fma(w0[t][1], s1, r0));
                r1 = fma(w1[t][0], s0, fma(w1[t][1], s1, r1));
              }
              r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
              r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
              if (lk == (g & 3)) my0[g >> 2] = r0;
              if (degen) {
                r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
                r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
                if (lk == (g & 3)) my1[g >> 2] = r1;
              }
            }
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const int gq = gp + lk + 4 * h, j = 8 * gq + lr, q = q0 + j;
              if (lk + 4 * h < NG && gq < ngr) {
                const double X = inv1 * xs[j] + inv2;
                const double sc = c_xc.inv4pi / (X * X);
                const double fxv = (has_xc && last) ? fs[j] : 0.0;
                double r = my0[h] + (first ? 0.0 : out0[q]);
                out0[q] = last ? r * sc : r;          // not last: raw sum of the sweeps so far
                corr0 = fma(fxv, r, corr0);
                if (degen) {
                  r = my1[h] + (first ? 0.0 : out1[q]);
                  out1[q] = last ? r * sc : r;
                  corr1 = fma(fxv, r, corr1);
                }
              }
            }
          }
        };
        if (two) run(std::integral_constant<int, 4>{}, std::integral_constant<int, 2>{});
        else run(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{});
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
      }
    }
  }
  DENS_STAMP();
  corr0 = warp_sum(corr0);
  corr1 = warp_sum(corr1);
  if (lane == 0) {
    b.corr_part[((size_t)p * 2 + 0) * C + c] = corr0 * inv1;
    if (degen) b.corr_part[((size_t)p * 2 + 1) * C + c] = corr1 * inv1;
  }
}

// k_fd: the cell's share of the Hartree rhs, B_m = sum_ij F_c[m][ij] D_loc[ij] with D_loc = sum_o n_o u_o u_o^T on the cell
// (tensor_matrix_dict!, src/utils/free allocation.jl:17-36).  One CTA per cell, slot and FD_NP problems: the (L2-resident)
// tensor block F_c is read once for all of them.
#define FD_NP 4
__host__ __device__ inline size_t fd_smem_bytes(int n) {
  return sizeof(double) * ((size_t)AKS_MAXORB * DENS_TC_ROWS + (size_t)FD_NP * n * n + 8);
}
__global__ void __launch_bounds__(AKS_NT) k_fd(DiscDev d, BatchDev b) {
  const int c = blockIdx.x, pbase = blockIdx.y * FD_NP, slot = blockIdx.z;
  __shared__ OccList L;
  extern __shared__ __align__(16) double sm_dc[];
  constexpr int NR = DENS_TC_ROWS;
  const int n = d.n, C = d.C, nn = n * n;
  double* us = sm_dc;                           // [cnt][24]
  double* Dl = us + (size_t)AKS_MAXORB * NR;    // [FD_NP][n*n]
  unsigned mask = 0;
  for (int pp = 0; pp < FD_NP; ++pp) {
    const int p = pbase + pp;
    if (p < b.B && b.status[p] == 0 && (slot == 0 || b.degen[p]) && b.hartree[p] != 0.0) mask |= 1u << pp;
  }
  if (!mask) return;
  for (int pp = 0; pp < FD_NP; ++pp) {
    if (!((mask >> pp) & 1u)) {
      for (int idx = threadIdx.x; idx < nn; idx += AKS_NT) Dl[pp * nn + idx] = 0.0;
      continue;
    }
    const int p = pbase + pp;
    __syncthreads();   // the previous problem's `us` and list are no longer read
    build_occ_list(b, p, slot, &L);
    const int cnt = L.count;
    for (int idx = threadIdx.x; idx < cnt * NR; idx += AKS_NT) {
      const int o = idx / NR, i = idx - o * NR;
      double v = 0.0;
      if (i < n) {
        const int gi = d.l2g[c * n + i];
        if (gi >= 0) v = b.U[orb_off(b, p, L.e[o], L.l[o], L.k[o]) * d.Nh + gi];
      }
      us[idx] = v * sqrt(L.n[o]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nn; idx += AKS_NT) {
      const int i = idx / n, j = idx - i * n;
      double v = 0.0;
      for (int o = 0; o < cnt; ++o) v = fma(us[(size_t)o * NR + i], us[(size_t)o * NR + j], v);
      Dl[pp * nn + idx] = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* Fc = d.F_loc + (size_t)c * nn * n;
  // a warp owns the rows m = warp, warp + 8, warp + 16: they advance together (three loads of F_c in flight per lane
  // and step, every value used for the four problems)
  constexpr int MPW = (DENS_TC_ROWS + AKS_NW - 1) / AKS_NW;
  double acc[MPW][FD_NP];
#pragma unroll
  for (int j = 0; j < MPW; ++j)
#pragma unroll
    for (int pp = 0; pp < FD_NP; ++pp) acc[j][pp] = 0.0;
#pragma unroll 2
  for (int idx = lane; idx < nn; idx += 32) {
    double dv[FD_NP];
#pragma unroll
    for (int pp = 0; pp < FD_NP; ++pp) dv[pp] = Dl[pp * nn + idx];
#pragma unroll
    for (int j = 0; j < MPW; ++j) {
      const int m = warp + AKS_NW * j;
      if (m < n) {
        const double fv = Fc[(size_t)m * nn + idx];
#pragma unroll
        for (int pp = 0; pp < FD_NP; ++pp) acc[j][pp] = fma(dv[pp], fv, acc[j][pp]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MPW; ++j) {
    const int m = warp + AKS_NW * j;
    if (m < n) {   // warp-uniform
#pragma unroll
      for (int pp = 0; pp < FD_NP; ++pp) {
        const double v = warp_sum(acc[j][pp]);
        if (lane == 0 && ((mask >> pp) & 1u)) b.Bloc_new[(((size_t)(pbase + pp) * 2 + slot) * C + c) * n + m] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_dens_grid: trial density on the global energy grid (eval_density!, density.jl:97-126)
template <int NMAX>
__global__ void __launch_bounds__(AKS_NT) k_dens_grid(DiscDev d, BatchDev b) {
  const int p = blockIdx.y, slot = blockIdx.z;
  if (b.status[p] != 0) return;
  if (slot == 1 && !b.degen[p]) return;
  __shared__ OccList L;
  build_occ_list(b, p, slot, &L);
  const int q = blockIdx.x * AKS_NT + threadIdx.x;
  if (q >= d.G) return;
  const int n = d.n, c = d.ycell[q], s = b.s;
  double g[NMAX];
  int gi[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    g[i] = (i < n) ? d.Pgy[(size_t)i * d.G + q] : 0.0;
    gi[i] = (i < n) ? d.l2g[c * n + i] : -1;
  }
  double r0 = 0.0, r1 = 0.0;
  for (int o = 0; o < L.count; ++o) {
    const double* u = b.U + orb_off(b, p, L.e[o], L.l[o], L.k[o]) * d.Nh;
    double psi = 0.0;
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
      if (gi[i] >= 0) psi = fma(u[gi[i]], g[i], psi);
    const double t = L.n[o] * psi * psi;
    if (L.e[o] == 0) r0 += t; else r1 += t;
  }
  const double yq = d.y[q];
  const double den = c_xc.fourpi * (yq * yq);
  b.rho_y_new[(((size_t)p * 2 + slot) * s + 0) * d.G + q] = r0 / den;
  if (s == 2) b.rho_y_new[(((size_t)p * 2 + slot) * s + 1) * d.G + q] = r1 / den;
}

// ------------------------------------------------------------------------------------------
// k_poisson: assemble B = F:D_new from the cell shares and solve A W = B with the prefactored
// statically condensed stiffness operator (the reference calls CHOLMOD on the sparse A every
// time: assemble.jl:71,76; energies.jl:105,111,128,137).  Also B.W and B_prev.W.
__global__ void __launch_bounds__(128) k_poisson(DiscDev d, BatchDev b, int p0, int force) {
  const int p = blockIdx.x + p0, slot = blockIdx.y;
  if (!force && b.status[p] != 0) return;
  if (slot == 1 && !b.degen[p]) return;
  extern __shared__ double sm[];
  const int C = d.C, n = d.n, nb = d.nb, Nh = d.Nh;
  double* Bs = sm;          // [Nh]
  double* Ws = Bs + Nh;     // [Nh]
  double* hp = Ws + Nh;     // [2C]
  double* xh = hp + 2 * C;  // [C]
  double* red = xh + C;     // [8]
  double* Bout = b.B_new + ((size_t)p * 2 + slot) * Nh;
  double* Wout = b.W_new + ((size_t)p * 2 + slot) * Nh;
  double* sc = b.scal + ((size_t)p * 2 + slot) * 4;
  if (b.hartree[p] == 0.0) {
    for (int i = threadIdx.x; i < Nh; i += blockDim.x) { Bout[i] = 0.0; Wout[i] = 0.0; }
    if (threadIdx.x == 0) { sc[0] = 0.0; sc[1] = 0.0; }
    return;
  }
  const double* Bl = b.Bloc_new + ((size_t)p * 2 + slot) * C * n;
  for (int idx = threadIdx.x; idx < C * n; idx += blockDim.x) {
    const int c = idx / n, m = idx - c * n;
    const int gi = d.l2g[idx];
    if (gi < 0) continue;
    if (m >= 2) Bs[gi] = Bl[idx];
    else if (m == 0) Bs[gi] = Bl[idx] + Bl[(c + 1) * n + 1];  // right hat of c == left hat of c+1
  }
  __syncthreads();
  // hat rhs: rh_j = B[hat j] - XA_j[0].B_b(j) - XA_{j+1}[1].B_b(j+1)
  for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
    const int c = idx >> 1, g = idx & 1;
    const double* X = d.XA + ((size_t)c * 2 + g) * nb;
    double acc = 0.0;
    for (int i = 0; i < nb; ++i) acc = fma(X[i], Bs[d.l2g[c * n + 2 + i]], acc);
    hp[idx] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int m = C - 1;
    double zprev = 0.0;
    for (int j = 0; j < m; ++j) {
      double rh = Bs[d.l2g[j * n]] - hp[2 * j] - hp[2 * (j + 1) + 1];
      if (j > 0) rh -= d.triA_l[j - 1] * zprev;
      xh[j] = rh;
      zprev = rh;
    }
    double xnext = 0.0;
    for (int j = m - 1; j >= 0; --j) {
      double v = xh[j] / d.triA_d[j];
      if (j < m - 1) v -= d.triA_l[j] * xnext;
      xh[j] = v;
      xnext = v;
      Ws[d.l2g[j * n]] = v;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * nb; idx += blockDim.x) {
    const int c = idx / nb, i = idx - c * nb;
    const double* Ai = d.Abb_inv + ((size_t)c * nb + i) * nb;
    double acc = 0.0;
    for (int j = 0; j < nb; ++j) acc = fma(Ai[j], Bs[d.l2g[c * n + 2 + j]], acc);
    const double xh0 = (c < C - 1) ? xh[c] : 0.0, xh1 = (c > 0) ? xh[c - 1] : 0.0;
    acc -= d.XA[((size_t)c * 2 + 0) * nb + i] * xh0 + d.XA[((size_t)c * 2 + 1) * nb + i] * xh1;
    Ws[d.l2g[c * n + 2 + i]] = acc;
  }
  __syncthreads();
  double bw = 0.0, bpw = 0.0;
  const double* Bp = b.Bv + (size_t)p * Nh;
  for (int i = threadIdx.x; i < Nh; i += blockDim.x) {
    bw = fma(Bs[i], Ws[i], bw);
    bpw = fma(Bp[i], Ws[i], bpw);
    Bout[i] = Bs[i];
    Wout[i] = Ws[i];
  }
  bw = block_sum(bw, red);
  bpw = block_sum(bpw, red);
  if (threadIdx.x == 0) { sc[0] = bw; sc[1] = bpw; }
}

// ------------------------------------------------------------------------------------------
// Exc on the global grid for rho = a*new0 + bq*new1 + cq*prev (compute_exc_energy,
// energies.jl:217-233 with evaluate_zk!, models.jl:107-116).  rho is linear in D, so the
// reference's F(t) = Exc[t D + (1-t) D_prev] (oda/types.jl:96-100) costs O(G) per evaluation.
__device__ double exc_grid(const DiscDev& d, const BatchDev& b, int p, double a, double bq, double cq,
                           bool use1, double* red) {
  const int G = d.G, s = b.s;
  const int ex = b.ex[p], ec = b.ec[p];
  const double* n0 = b.rho_y_new + (((size_t)p * 2 + 0) * s) * G;
  const double* n1 = b.rho_y_new + (((size_t)p * 2 + 1) * s) * G;
  const double* pr = b.rho_y + ((size_t)p * s) * G;
  double acc = 0.0;
  for (int q = threadIdx.x; q < G; q += blockDim.x) {
    double ru = a * n0[q] + cq * pr[q];
    if (use1) ru += bq * n1[q];
    double val;
    if (s == 1) {
      val = ru * xc_zk_unpol(fmax(ru, 0.0), ex, ec);
    } else {
      double rd = a * n0[G + q] + cq * pr[G + q];
      if (use1) rd += bq * n1[G + q];
      val = (ru + rd) * xc_zk_pol(fmax(ru, 0.0), fmax(rd, 0.0), ex, ec);
    }
    const double yq = d.y[q];
    acc = fma(d.wy[q], val * (yq * yq), acc);
  }
  return c_xc.fourpi * block_sum(acc, red);
}

// ------------------------------------------------------------------------------------------
// k_linesearch: energies of the trial density (compute_total_energy & co, energies.jl:7-141),
// two-fold degeneracy resolution (optimized.jl:228-306) and the ODA line search
// (line_search_density!, oda/loop.jl:79-116).  One CTA per problem.
//
// The whole Brent loop (Optim.jl's Brent(), ~30-60 SEQUENTIAL evaluations of Exc on the grid) runs inside the
// kernel, so the time of one evaluation is what matters.  Every thread keeps the grid densities of its points
// (trial slot 0 / slot 1 / previous, both spins) and the quadrature weights in REGISTERS for the whole kernel: an
// evaluation is then pure arithmetic (rho(t) is linear in t) + one warp shuffle reduction + ONE pass through shared
// memory with two barriers -- no global loads, no per-evaluation block_sum of 32 partials.
#define LS_NT 512
#define LS_RHO_MIN 1e-30   // densities at or below this are rounding noise of the orbital tails (see ls_exc)
#define LS_MAXP 4   // grid points per thread held in registers: G <= LS_NT * LS_MAXP = 2048
// MAXP points per thread (2 for G <= 1024, else 4), NS spins: compile-time so that the arrays live in registers
template <int MAXP, int NS>
struct LsGrid {
  double n0u[MAXP], n1u[MAXP], pru[MAXP], wy[MAXP], y2[MAXP];
  double n0d[NS == 2 ? MAXP : 1], n1d[NS == 2 ? MAXP : 1], prd[NS == 2 ? MAXP : 1];
};
struct LsShared { double t; double f; int go; };

// Exc[a*new0 + bq*new1 + cq*prev] from the register-resident grid (same per-point arithmetic as exc_grid)
template <int MAXP, int NS>
__device__ __forceinline__ double ls_exc(const LsGrid<MAXP, NS>& g, int G, int ex, int ec, double a, double bq, double cq,
                                         bool use1, double* red /*[LS_NT/32]*/) {
  constexpr int s = NS;
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < MAXP; ++j) {
    const int q = threadIdx.x + j * LS_NT;
    // The energy grid is ONE Gauss rule on [0, Rmax] (quadrature.jl / integration methods.jl): 70-80 % of its nodes lie where
    // the density of an atom has decayed to rounding noise.  A node whose three densities are all below LS_RHO_MIN adds
    // less than 1e-33 to a sum of order 0.1 ... 1e3 -- nothing, in double precision -- and is skipped: the functional is
    // evaluated where it matters (per evaluation ~250 nodes instead of 1000; whole warps of far-out nodes fall through).
    bool live = q < G && fmax(fmax(fabs(g.n0u[j]), fabs(g.pru[j])), use1 ? fabs(g.n1u[j]) : 0.0) > LS_RHO_MIN;
    if (s == 2) {
      constexpr int jd2 = (NS == 2) ? 1 : 0;
      live = live || (q < G && fmax(fmax(fabs(g.n0d[j * jd2]), fabs(g.prd[j * jd2])), use1 ? fabs(g.n1d[j * jd2]) : 0.0) > LS_RHO_MIN);
    }
    if (live) {
      double ru = a * g.n0u[j] + cq * g.pru[j];
      if (use1) ru += bq * g.n1u[j];
      double val;
      if (s == 1) {
        val = ru * xc_zk_unpol(fmax(ru, 0.0), ex, ec);
      } else {
        constexpr int jd = (NS == 2) ? 1 : 0;   // (index kept in range for the one-element arrays of NS = 1)
        double rd = a * g.n0d[j * jd] + cq * g.prd[j * jd];
        if (use1) rd += bq * g.n1d[j * jd];
        val = (ru + rd) * xc_zk_pol(fmax(ru, 0.0), fmax(rd, 0.0), ex, ec);
      }
      acc = fma(g.wy[j], val * g.y2[j], acc);
    }
  }
  acc = warp_sum(acc);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = acc;
  __syncthreads();
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < LS_NT / 32; ++i) sum += red[i];   // fixed order: every thread forms the same value
  return c_xc.fourpi * sum;
}

// minimise a t^2 + b t + c + F(t) on [0,1] (line_search_energy, line_search.jl:45-80);
// F(t) = Exc[t*rhoA + (1-t)*rhoB] where (A,B) = (new0,new1) [degenerate aufbau] or
// (trial, prev) [ODA].  All threads of the CTA call this; returns (tmin, fmin) to all.
// Thread 0 drives Brent's state machine; per evaluation: ls_exc (one barrier) + publish (one barrier).
template <int MAXP, int NS>
__device__ void line_search(const LsGrid<MAXP, NS>& g, int G, int ex, int ec, double qa, double qb, double qc,
                            bool has_xc, int mode /*0: new0 vs new1, 1: trial(td) vs prev*/, double td,
                            bool degen, double abstol, double reltol, int maxit, LsShared* ls,
                            double* red, double& t_out, double& f_out) {
  if (!has_xc) {
    double t;
    if (qa > 0.0) t = fmin(fmax(-qb / (2.0 * qa), 0.0), 1.0);
    else t = (qc <= qc + qb + qa) ? 0.0 : 1.0;
    t_out = t;
    f_out = qc + qb * t + qa * t * t;
    return;
  }
  BrentState st;
  brent_init(st, 0.0, 1.0, abstol, reltol, maxit);   // every thread: the first abscissa needs no broadcast
  double t = st.next;
  bool first = true;
  while (true) {
    double F;
    if (mode == 0) F = ls_exc(g, G, ex, ec, t, 1.0 - t, 0.0, true, red);
    else F = ls_exc(g, G, ex, ec, t * td, t * (1.0 - td), 1.0 - t, degen, red);
    if (threadIdx.x == 0) {
      const double fv = qa * (t * t) + qb * t + qc + F;
      if (first) brent_first(st, fv); else brent_update(st, fv);
      const int go = brent_propose(st) ? 1 : 0;
      ls->go = go;
      ls->t = go ? st.next : st.x;
      ls->f = st.fx;
    }
    first = false;
    __syncthreads();   // publishes ls; also orders this evaluation's reads of `red` before the next writes
    t = ls->t;
    if (!ls->go) break;
  }
  t_out = t;
  f_out = ls->f;
  __syncthreads();     // ls is rewritten by the next search
}

template <int MAXP, int NS>
__global__ void __launch_bounds__(LS_NT) k_linesearch(DiscDev d, BatchDev b) {
  const int p = blockIdx.x;
  if (b.status[p] != 0) return;
#ifdef AKS_LS_TIMING   // cycle stamps of thread 0 (tools/gpu_ls_cycles.py)
  long long* stamp_ = b.eig_dbg + (size_t)p * 8; int nstamp_ = 0;
#define LS_STAMP() do { if (threadIdx.x == 0 && nstamp_ < 8) stamp_[nstamp_++] = clock64(); } while (0)
#else
#define LS_STAMP() do { } while (0)
#endif
  LS_STAMP();
  __shared__ double red[LS_NT / 32 + 2];
  __shared__ LsShared ls;
  __shared__ double sh[16];
  __shared__ double s_occ[2][AKS_MAXORB], s_eps[AKS_MAXORB], s_ek[AKS_MAXORB], s_ec[AKS_MAXORB];
  constexpr int s = NS;
  const int C = d.C, Nh = d.Nh, G = d.G;
  const int ex = b.ex[p], ec = b.ec[p];
  const bool has_xc = (ex != 0) || (ec != 0);
  const bool har_on = b.hartree[p] != 0.0;
  const bool degen = b.degen[p] != 0;
  const double Np = b.N[p], Rmax = d.Rmax;
  const double n2r = Np * Np / Rmax;
  const double eps64 = 2.220446049250313e-16;
  const int Lp = b.Lp[p], nhp = b.nhp[p];
  // ---- grid densities of this thread's points -> registers (read once)
  LsGrid<MAXP, NS> g;
  const bool in_regs = (G <= LS_NT * MAXP) && (b.s == NS);
  if (has_xc) {
    const double* n0 = b.rho_y_new + (((size_t)p * 2 + 0) * s) * G;
    const double* n1 = b.rho_y_new + (((size_t)p * 2 + 1) * s) * G;
    const double* pr = b.rho_y + ((size_t)p * s) * G;
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      const int q = threadIdx.x + j * LS_NT;
      const bool in = q < G;
      g.n0u[j] = in ? n0[q] : 0.0;
      g.pru[j] = in ? pr[q] : 0.0;
      g.n1u[j] = (in && degen) ? n1[q] : 0.0;
      if (NS == 2) {
        g.n0d[j * (NS - 1)] = in ? n0[G + q] : 0.0;
        g.prd[j * (NS - 1)] = in ? pr[G + q] : 0.0;
        g.n1d[j * (NS - 1)] = (in && degen) ? n1[G + q] : 0.0;
      }
      g.wy[j] = in ? d.wy[q] : 0.0;
      const double yq = in ? d.y[q] : 0.0;
      g.y2[j] = yq * yq;
    }
  }
  LS_STAMP();
  if (!in_regs) {   // never with the supported shapes (npoints <= 2048); keep the failure loud
    if (threadIdx.x == 0) b.status[p] = AKS_EINVAL;
    return;
  }
  // ---- orbital sums for both slots: loads by all threads, sums by thread 0 in the reference's order
  // (sigma, l, k as in energies.jl:52-61)
  for (int f = threadIdx.x; f < s * Lp * nhp; f += LS_NT) {
    const int e = f / (Lp * nhp), r = f - e * Lp * nhp, l = r / nhp, k = r - l * nhp;
    s_occ[0][f] = b.occ[occ_off(b, p, 0, e, l, k)];
    s_occ[1][f] = degen ? b.occ[occ_off(b, p, 1, e, l, k)] : 0.0;
    s_eps[f] = b.eps[orb_off(b, p, e, l, k)];
    s_ek[f] = b.ekin_orb[orb_off(b, p, e, l, k)];
    s_ec[f] = b.ecou_orb[orb_off(b, p, e, l, k)];
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int slot = threadIdx.x;
    double sne = 0.0, ek = 0.0, ecs = 0.0;
    if (slot == 0 || degen)
      for (int f = 0; f < s * Lp * nhp; ++f) {
        const double nn = s_occ[slot][f];
        if (nn != 0.0) { sne += nn * s_eps[f]; ek += nn * s_ek[f]; ecs += nn * s_ec[f]; }
      }
    double corr = 0.0;
    if (has_xc && (slot == 0 || degen))
      for (int c = 0; c < C; ++c) corr += b.corr_part[((size_t)p * 2 + slot) * C + c];
    sh[4 * slot + 0] = sne; sh[4 * slot + 1] = ek; sh[4 * slot + 2] = ecs; sh[4 * slot + 3] = corr;
  }
  __syncthreads();
  LS_STAMP();
  const double* sc0 = b.scal + ((size_t)p * 2 + 0) * 4;
  const double* sc1 = b.scal + ((size_t)p * 2 + 1) * 4;
  double Ekin, Ecou, Ehar, Eexc, Etot, td = 1.0;
  if (degen) {
    // energies of the two extreme fillings and their Hartree cross term
    const double k0 = sh[1], c0 = sh[2], k1 = sh[5], c1 = sh[6];
    const double h0 = har_on ? (sc0[0] + n2r) / 2.0 : 0.0;
    const double h1 = har_on ? (sc1[0] + n2r) / 2.0 : 0.0;
    double h01 = 0.0;
    if (har_on) {
      const double* B1 = b.B_new + ((size_t)p * 2 + 1) * Nh;
      const double* W0 = b.W_new + ((size_t)p * 2 + 0) * Nh;
      double acc = 0.0;
      for (int i = threadIdx.x; i < Nh; i += blockDim.x) acc = fma(B1[i], W0[i], acc);
      h01 = (block_sum(acc, red) + n2r) / 2.0;
      __syncthreads();
    }
    const double A0 = k0 + c0, A1 = k1 + c1;
    const double qa = h0 + h1 - 2.0 * h01, qb = (A0 - A1) + 2.0 * (h01 - h1), qc = A1 + h1;
    double emin;
    line_search(g, G, ex, ec, qa, qb, qc, has_xc, 0, 1.0, true, 1e4 * eps64, 1e4 * eps64, 100, &ls, red, td, emin);
    Etot = emin;
    Ekin = td * k0 + (1.0 - td) * k1;
    Ecou = td * c0 + (1.0 - td) * c1;
    Ehar = td * td * h0 + (1.0 - td) * (1.0 - td) * h1 + 2.0 * td * (1.0 - td) * h01;
    Eexc = Etot - Ekin - Ecou - Ehar;
    if (threadIdx.x == 0) {
      const int o0 = b.degen_idx[2 * p], o1 = b.degen_idx[2 * p + 1];
      const double* dn = b.degen_n + 4 * p;
      const size_t base = ((size_t)p * 2 + 0) * s * b.Lmax * b.nhmax;
      b.occ[base + o0] = td * dn[0] + (1.0 - td) * dn[2];
      b.occ[base + o1] = td * dn[1] + (1.0 - td) * dn[3];
      write_occ_list(b, p, 0);
    }
  } else {
    Ekin = sh[1];
    Ecou = sh[2];
    Ehar = har_on ? (sc0[0] + n2r) / 2.0 : 0.0;
    Eexc = 0.0;
    if (has_xc) { Eexc = ls_exc(g, G, ex, ec, 1.0, 0.0, 0.0, false, red); __syncthreads(); }
    Etot = has_xc ? (sh[0] - Ehar + Eexc - sh[3]) : (sh[0] - Ehar);
  }
  LS_STAMP();
  // ---- ODA relaxation (niter > 0)
  const int niter = b.niter[p];
  const double* Ep = b.E + (size_t)p * 5;  // previous: Etot,Ekin,Ecou,Ehar,Eexc
  double t = b.t[p], tmix = 1.0;
  if (niter > 0) {
    double h01 = 0.0;
    if (har_on) {
      const double bpw = degen ? (td * sc0[1] + (1.0 - td) * sc1[1]) : sc0[1];
      h01 = (bpw + n2r) / 2.0;
    }
    double ekin_m, ecou_m, ehar_m, eexc_m, etot_m;
    if (b.frozen_p[p]) {
      ekin_m = t * Ekin + (1.0 - t) * Ep[1];
      ecou_m = t * Ecou + (1.0 - t) * Ep[2];
      ehar_m = (t * t) * Ehar + ((1.0 - t) * (1.0 - t)) * Ep[3] + 2.0 * t * (1.0 - t) * h01;
      eexc_m = 0.0;
      if (has_xc) { eexc_m = ls_exc(g, G, ex, ec, t * td, t * (1.0 - td), 1.0 - t, degen, red); __syncthreads(); }
      etot_m = ekin_m + ecou_m + ehar_m + eexc_m;
    } else {
      const double A0 = Ekin + Ecou, A1 = Ep[1] + Ep[2];
      const double qa = Ehar + Ep[3] - 2.0 * h01, qb = (A0 - A1) + 2.0 * (h01 - Ep[3]), qc = A1 + Ep[3];
      line_search(g, G, ex, ec, qa, qb, qc, has_xc, 1, td, degen, b.abstol_ls, b.reltol_ls, b.maxiter_ls, &ls,
                  red, t, etot_m);
      ekin_m = t * Ekin + (1.0 - t) * Ep[1];
      ecou_m = t * Ecou + (1.0 - t) * Ep[2];
      ehar_m = t * t * Ehar + (1.0 - t) * (1.0 - t) * Ep[3] + 2.0 * t * (1.0 - t) * h01;
      eexc_m = etot_m - ekin_m - ecou_m - ehar_m;
    }
    Ekin = ekin_m; Ecou = ecou_m; Ehar = ehar_m; Eexc = eexc_m; Etot = etot_m;
    tmix = t;
  }
  LS_STAMP();
  if (threadIdx.x == 0) {
    double* En = b.Enew + (size_t)p * 5;
    En[0] = Etot; En[1] = Ekin; En[2] = Ecou; En[3] = Ehar; En[4] = Eexc;
    b.t[p] = t;
    b.tmix[p] = tmix;
    b.td[p] = td;
  }
}

// ------------------------------------------------------------------------------------------
// k_mix_state: D <- t D_new + (1-t) D_prev applied to the linear images of D the next step
// reads (rho at the Gauss nodes, rho on the energy grid, Hartree rhs and solution).
__global__ void k_mix_state(DiscDev d, BatchDev b) {
  const int p = blockIdx.y;
  if (b.status[p] != 0) return;
  const double t = b.tmix[p], td = b.td[p];
  const bool degen = b.degen[p] != 0;
  const size_t nq = (size_t)b.s * d.C * d.Q, ny = (size_t)b.s * d.G, nv = d.Nh;
  const size_t total = nq + ny + 2 * nv;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    double *dst;
    const double *a0, *a1;
    if (idx < nq) {
      dst = b.rho_q + (size_t)p * nq + idx;
      a0 = b.rho_q_new + ((size_t)p * 2 + 0) * nq + idx; a1 = b.rho_q_new + ((size_t)p * 2 + 1) * nq + idx;
    } else if (idx < nq + ny) {
      const size_t i = idx - nq;
      dst = b.rho_y + (size_t)p * ny + i;
      a0 = b.rho_y_new + ((size_t)p * 2 + 0) * ny + i; a1 = b.rho_y_new + ((size_t)p * 2 + 1) * ny + i;
    } else if (idx < nq + ny + nv) {
      const size_t i = idx - nq - ny;
      dst = b.Bv + (size_t)p * nv + i;
      a0 = b.B_new + ((size_t)p * 2 + 0) * nv + i; a1 = b.B_new + ((size_t)p * 2 + 1) * nv + i;
    } else {
      const size_t i = idx - nq - ny - nv;
      dst = b.Wv + (size_t)p * nv + i;
      a0 = b.W_new + ((size_t)p * 2 + 0) * nv + i; a1 = b.W_new + ((size_t)p * 2 + 1) * nv + i;
    }
    double v = *a0;
    if (degen) v = td * v + (1.0 - td) * (*a1);
    *dst = t * v + (1.0 - t) * (*dst);
  }
}

// ------------------------------------------------------------------------------------------
// k_mix_dense: the dense density matrix D = sum n u u^T (density!, density.jl:15-41), mixed in
// place, and the Frobenius norm of the update (stopping rule, oda/loop.jl:43-45).  Pure HBM
// streaming: read D, write D; the trial D is never materialised.  One CTA per 64x64 tile; each thread
// owns 4 rows x 4 strided columns (8 shared loads per 16 FMAs of the rank-n_occ update).  The old values of the tile are
// PREFETCHED INTO L2 at kernel entry (one prefetch.global.L2 per 128-byte line) and only loaded into registers after
// the update has been accumulated: they travel from HBM while the CTA stages orbitals and computes, but hold no
// registers meanwhile -- 64 registers, 4 CTAs per SM instead of 126 and 2 (0.159 -> 0.139 ms).
#define MIX_TS 64
#define MIX_OCH 40   // orbitals staged at once: in practice every occupied orbital of one spin (<= 26 for Z <= 86)
__global__ void __launch_bounds__(AKS_NT, 4) k_mix_dense(DiscDev d, BatchDev b) {
  const int p = blockIdx.z, e = blockIdx.y;
  if (b.status[p] != 0) return;
  // 2 x 40 x 64 doubles = 40 KB of static shared memory
  __shared__ double Ui[MIX_OCH][MIX_TS];
  __shared__ double Uj[MIX_OCH][MIX_TS];
  __shared__ double red[AKS_NW];
  const int Nh = d.Nh;
  // D is symmetric: only the tiles on and below the diagonal are kept current during the iterations
  // (k_mirror_dense completes D when the host asks for it); blockIdx.x = tr (tr + 1) / 2 + tc
  int tr = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((tr + 1) * (tr + 2) / 2 <= (int)blockIdx.x) ++tr;
  while (tr * (tr + 1) / 2 > (int)blockIdx.x) --tr;
  const int tc = blockIdx.x - tr * (tr + 1) / 2;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 entries each
  const int r0 = tr * MIX_TS + 4 * ty, c0 = tc * MIX_TS + tx;   // columns c0 + 16 c: coalesced rows
  double* D = b.Dd + ((size_t)p * b.s + e) * Nh * Nh;
  // 1. start the tile's trip from HBM to L2 without holding registers for it: one prefetch per 128-byte line (a 64-entry
  // row piece covers 4 or 5 lines: the row stride Nh is odd), 320 lines per tile ...
  {
    const int rows = min(MIX_TS, Nh - tr * MIX_TS), cols = min(MIX_TS, Nh - tc * MIX_TS);
    for (int ln = threadIdx.x; ln < rows * 5; ln += AKS_NT) {
      const int r = ln / 5, sg = ln - r * 5;
      const double* row = D + (size_t)(tr * MIX_TS + r) * Nh + tc * MIX_TS;
      const char* line = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(row) & ~(uintptr_t)127) + 128 * sg;
      if (line < reinterpret_cast<const char*>(row + cols)) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
    }
  }
  // 2. ... and, while the lines are in flight, stage the two 64-entry slices of EVERY occupied orbital of this spin (L2):
  // the compact list of k_aufbau / k_linesearch is ordered by spin, so the entries of spin e are contiguous.  One
  // barrier for the whole tile (the chunked staging of round 1 had two per 16 orbitals plus three in a prologue
  // that rebuilt the list with shared-memory atomics; ncu showed the barrier as the top stall reason).
  const int tot = b.occ_cnt[2 * p];
  const int* elk = b.occ_elk + ((size_t)p * 2) * AKS_MAXORB;
  const double* nn = b.occ_n + ((size_t)p * 2) * AKS_MAXORB;
  const int sp = ((int)threadIdx.x < tot) ? (elk[threadIdx.x] >> 16) : 99;   // AKS_MAXORB <= blockDim.x
  const int first = __syncthreads_count(sp < e);
  const int cnt = __syncthreads_count(sp == e);
  const double* Ubase = b.U + (((size_t)p * b.s + e) * b.Lmax * b.nhmax) * Nh;
  const double t = b.tmix[p];
  double dn[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) dn[a][c] = 0.0;
  for (int o0 = 0; o0 < cnt; o0 += MIX_OCH) {   // one pass unless more than MIX_OCH orbitals of this spin are occupied
    const int oc = min(MIX_OCH, cnt - o0);
    if (o0 > 0) __syncthreads();
    for (int idx = threadIdx.x; idx < oc * MIX_TS; idx += AKS_NT) {
      const int o = idx / MIX_TS, r = idx - o * MIX_TS;
      const int v = elk[first + o0 + o];
      const double* u = Ubase + (size_t)((((v >> 8) & 0xff) * b.nhmax) + (v & 0xff)) * Nh;
      const int gi = tr * MIX_TS + r, gj = tc * MIX_TS + r;
      Ui[o][r] = (gi < Nh) ? nn[first + o0 + o] * u[gi] : 0.0;   // occupation folded into the row factor
      Uj[o][r] = (gj < Nh) ? u[gj] : 0.0;
    }
    __syncthreads();
#pragma unroll 2
    for (int o = 0; o < oc; ++o) {
      const double a0 = Ui[o][4 * ty], a1 = Ui[o][4 * ty + 1], a2 = Ui[o][4 * ty + 2], a3 = Ui[o][4 * ty + 3];
      const double b0 = Uj[o][tx], b1 = Uj[o][tx + 16], b2 = Uj[o][tx + 32], b3 = Uj[o][tx + 48];
      dn[0][0] = fma(a0, b0, dn[0][0]); dn[0][1] = fma(a0, b1, dn[0][1]); dn[0][2] = fma(a0, b2, dn[0][2]); dn[0][3] = fma(a0, b3, dn[0][3]);
      dn[1][0] = fma(a1, b0, dn[1][0]); dn[1][1] = fma(a1, b1, dn[1][1]); dn[1][2] = fma(a1, b2, dn[1][2]); dn[1][3] = fma(a1, b3, dn[1][3]);
      dn[2][0] = fma(a2, b0, dn[2][0]); dn[2][1] = fma(a2, b1, dn[2][1]); dn[2][2] = fma(a2, b2, dn[2][2]); dn[2][3] = fma(a2, b3, dn[2][3]);
      dn[3][0] = fma(a3, b0, dn[3][0]); dn[3][1] = fma(a3, b1, dn[3][1]); dn[3][2] = fma(a3, b2, dn[3][2]); dn[3][3] = fma(a3, b3, dn[3][3]);
    }
  }
  // 3. the old values now come from L2
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double old[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      old[c] = (r0 + a < Nh && c0 + 16 * c < Nh) ? D[(size_t)(r0 + a) * Nh + c0 + 16 * c] : 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (r0 + a < Nh && c0 + 16 * c < Nh) {
        const double nv = t * dn[a][c] + (1.0 - t) * old[c];
        D[(size_t)(r0 + a) * Nh + c0 + 16 * c] = nv;
        const double df = nv - old[c];
        acc = fma(df, df, acc);
      }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) b.norm_part[((size_t)p * b.s + e) * b.ntri + blockIdx.x] = (tr == tc) ? acc : 2.0 * acc;
}

// k_mix_dense_pipe: the same update by PERSISTENT CTAs (two per SM) that walk the (atom, spin, tile) list and keep the old
// values of the NEXT tile in flight while the current one is updated.  k_mix_dense issues its 16 loads per thread, then
// stages orbitals, computes, stores and reduces with nothing in flight: ncu shows 32 % of the DRAM throughput at
// 2 CTAs per SM (126 registers), a CTA living ~12 k cycles for 64 KB of traffic.  Here the old values go through shared
// memory by cp.async (8-byte copies: the rows of D have an odd stride, 16-byte alignment is not available): every thread
// copies exactly the 16 entries it will consume, into its own slots of a double buffer, so the buffer needs no barrier --
// only cp.async.wait_group -- and no registers while the copies fly.
// MEASURED SLOWER than k_mix_dense (0.177 vs 0.159 ms alone, 1.198 vs 1.178 ms per 8-lane step): the 8-byte cp.async copies
// cost an LSU instruction each like the loads they replace, the per-item barriers stay, and 124 registers still mean two
// CTAs per SM.  Off by default (AKS_MIX_PIPE=1 selects it); kept as the record of the experiment.
__host__ __device__ inline size_t mix_pipe_smem_bytes() {
  return sizeof(double) * (2 * (size_t)MIX_TS * MIX_TS + 2 * (size_t)MIX_OCH * MIX_TS + AKS_NW + 8);
}
__global__ void __launch_bounds__(AKS_NT, 2) k_mix_dense_pipe(DiscDev d, BatchDev b) {
  extern __shared__ __align__(16) double sm_mix[];
  double* obuf = sm_mix;                                   // [2][64][64] old values, thread-private slots
  double (*Ui)[MIX_TS] = reinterpret_cast<double (*)[MIX_TS]>(obuf + 2 * MIX_TS * MIX_TS);
  double (*Uj)[MIX_TS] = Ui + MIX_OCH;
  double* red = reinterpret_cast<double*>(Uj + MIX_OCH);
  const int Nh = d.Nh, s = b.s, ntri = b.ntri;
  const long long total = (long long)b.B * s * ntri;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 entries each
  // work item w -> (problem, spin, tile (tr, tc)); items of finished problems are skipped
  auto decode = [&](long long w, int& p, int& e, int& tile, int& tr, int& tc) {
    tile = (int)(w % ntri);
    const long long r = w / ntri;
    e = (int)(r % s);
    p = (int)(r / s);
    tr = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
    while ((tr + 1) * (tr + 2) / 2 <= tile) ++tr;
    while (tr * (tr + 1) / 2 > tile) --tr;
    tc = tile - tr * (tr + 1) / 2;
  };
  auto next_item = [&](long long w) {   // first item >= w of a running problem
    while (w < total && b.status[(int)(w / ((long long)s * ntri))] != 0) w += gridDim.x;
    return w;
  };
  auto prefetch = [&](long long w, int buf) {   // cp.async of this thread's 16 old values
    int p, e, tile, tr, tc;
    decode(w, p, e, tile, tr, tc);
    const double* D = b.Dd + ((size_t)p * s + e) * Nh * Nh;
    const int r0 = tr * MIX_TS + 4 * ty, c0 = tc * MIX_TS + tx;
    double* dst = obuf + (size_t)buf * MIX_TS * MIX_TS;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (r0 + a < Nh && c0 + 16 * c < Nh) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + (4 * ty + a) * MIX_TS + tx + 16 * c);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(D + (size_t)(r0 + a) * Nh + c0 + 16 * c) : "memory");
        }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  long long w = next_item(blockIdx.x);
  if (w >= total) return;
  prefetch(w, 0);
  for (int it = 0; w < total; ++it) {
    const int buf = it & 1;
    int p, e, tile, tr, tc;
    decode(w, p, e, tile, tr, tc);
    const long long wn = next_item(w + gridDim.x);
    const int r0 = tr * MIX_TS + 4 * ty, c0 = tc * MIX_TS + tx;
    double* D = b.Dd + ((size_t)p * s + e) * Nh * Nh;
    // occupied orbitals of this spin (the compact list is ordered by spin)
    const int tot = b.occ_cnt[2 * p];
    const int* elk = b.occ_elk + ((size_t)p * 2) * AKS_MAXORB;
    const double* nn = b.occ_n + ((size_t)p * 2) * AKS_MAXORB;
    const int sp = ((int)threadIdx.x < tot) ? (elk[threadIdx.x] >> 16) : 99;
    const int first = __syncthreads_count(sp < e);   // also: the previous item's Ui / Uj / red are no longer read
    const int cnt = __syncthreads_count(sp == e);
    const double* Ubase = b.U + (((size_t)p * s + e) * b.Lmax * b.nhmax) * Nh;
    const double t = b.tmix[p];
    double dn[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) dn[a][c] = 0.0;
    for (int o0 = 0; o0 < cnt || o0 == 0; o0 += MIX_OCH) {
      const int oc = min(MIX_OCH, cnt - o0);
      if (o0 > 0) __syncthreads();
      for (int idx = threadIdx.x; idx < oc * MIX_TS; idx += AKS_NT) {
        const int o = idx / MIX_TS, r = idx - o * MIX_TS;
        const int v = elk[first + o0 + o];
        const double* u = Ubase + (size_t)((((v >> 8) & 0xff) * b.nhmax) + (v & 0xff)) * Nh;
        const int gi = tr * MIX_TS + r, gj = tc * MIX_TS + r;
        Ui[o][r] = (gi < Nh) ? nn[first + o0 + o] * u[gi] : 0.0;   // occupation folded into the row factor
        Uj[o][r] = (gj < Nh) ? u[gj] : 0.0;
      }
      // the old values of the NEXT item start their trip while this one is updated
      if (o0 == 0 && wn < total) prefetch(wn, buf ^ 1);
      __syncthreads();
#pragma unroll 2
      for (int o = 0; o < oc; ++o) {
        const double a0 = Ui[o][4 * ty], a1 = Ui[o][4 * ty + 1], a2 = Ui[o][4 * ty + 2], a3 = Ui[o][4 * ty + 3];
        const double b0 = Uj[o][tx], b1 = Uj[o][tx + 16], b2 = Uj[o][tx + 32], b3 = Uj[o][tx + 48];
        dn[0][0] = fma(a0, b0, dn[0][0]); dn[0][1] = fma(a0, b1, dn[0][1]); dn[0][2] = fma(a0, b2, dn[0][2]); dn[0][3] = fma(a0, b3, dn[0][3]);
        dn[1][0] = fma(a1, b0, dn[1][0]); dn[1][1] = fma(a1, b1, dn[1][1]); dn[1][2] = fma(a1, b2, dn[1][2]); dn[1][3] = fma(a1, b3, dn[1][3]);
        dn[2][0] = fma(a2, b0, dn[2][0]); dn[2][1] = fma(a2, b1, dn[2][1]); dn[2][2] = fma(a2, b2, dn[2][2]); dn[2][3] = fma(a2, b3, dn[2][3]);
        dn[3][0] = fma(a3, b0, dn[3][0]); dn[3][1] = fma(a3, b1, dn[3][1]); dn[3][2] = fma(a3, b2, dn[3][2]); dn[3][3] = fma(a3, b3, dn[3][3]);
      }
    }
    // this item's old values have landed (the next item's group may still be in flight)
    if (wn < total) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    const double* ob = obuf + (size_t)buf * MIX_TS * MIX_TS;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (r0 + a < Nh && c0 + 16 * c < Nh) {
          const double old = ob[(4 * ty + a) * MIX_TS + tx + 16 * c];
          const double nv = t * dn[a][c] + (1.0 - t) * old;
          D[(size_t)(r0 + a) * Nh + c0 + 16 * c] = nv;
          const double df = nv - old;
          acc = fma(df, df, acc);
        }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) b.norm_part[((size_t)p * s + e) * ntri + tile] = (tr == tc) ? acc : 2.0 * acc;
    w = wn;
  }
}

// upper triangle of D from the lower one (see k_mix_dense); one CTA per 32 x 32 tile above the diagonal
__global__ void __launch_bounds__(AKS_NT) k_mirror_dense(DiscDev d, BatchDev b, int p) {
  __shared__ double tile[32][33];
  const int Nh = d.Nh, e = blockIdx.z;
  const int bi = blockIdx.y, bj = blockIdx.x;   // writes block (bi, bj), bj >= bi, from (bj, bi)
  if (bj < bi) return;
  double* D = b.Dd + ((size_t)p * b.s + e) * Nh * Nh;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += AKS_NW) {
    const int gi = bj * 32 + r, gj = bi * 32 + tx;
    tile[r][tx] = (gi < Nh && gj < Nh) ? D[(size_t)gi * Nh + gj] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += AKS_NW) {
    const int gi = bi * 32 + r, gj = bj * 32 + tx;
    if (gi < Nh && gj < Nh && gj > gi) D[(size_t)gi * Nh + gj] = tile[tx][r];
  }
}

// ------------------------------------------------------------------------------------------
// k_finalize: stopping criterion max(||D - D_prev||_F, |Etot - Etot_prev|) < scftol
// (oda/loop.jl:43-49), iteration counter, per-iteration record for LogBook / callbacks.
__global__ void k_finalize(DiscDev d, BatchDev b) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= b.B) return;
  if (b.status[p] != 0) {
    if (b.status[p] < 0) b.rec[p].status = b.status[p];   // an error raised inside this step (AKS_EDEGEN3)
    return;
  }
  const int npart = b.s * b.ntri;
  double ss = 0.0;
  for (int i = 0; i < npart; ++i) ss += b.norm_part[(size_t)p * npart + i];
  const double* En = b.Enew + (size_t)p * 5;
  double* E = b.E + (size_t)p * 5;
  const double stop = fmax(sqrt(ss), fabs(En[0] - E[0]));
  for (int i = 0; i < 5; ++i) E[i] = En[i];
  b.stop[p] = stop;
  b.niter[p] += 1;
  b.warm[p] = 1;
  const int efail = b.eig_fail[p];
  b.eig_fail[p] = 0;
  int st = 0;
  if (!(isfinite(En[0]) && isfinite(stop))) st = AKS_ENAN;
  else if (efail) st = AKS_EEIG;   // a level of the window was not converged + certified (k_eig): never report SUCCESS on it
  else if (stop < b.scftol_p[p]) st = AKS_SUCCESS;
  else if (b.maxiter_p[p] > 0 && b.niter[p] >= b.maxiter_p[p]) st = AKS_MAXITERS;   // groundstate(...; maxiter)
  b.status[p] = st;
  aks_iter_record r;
  r.niter = b.niter[p]; r.status = st; r.stop = stop;
  r.Etot = En[0]; r.Ekin = En[1]; r.Ecou = En[2]; r.Ehar = En[3]; r.Eexc = En[4];
  r.t = b.t[p];
  b.rec[p] = r;
  if (b.hist) {   // history of aks_scf_run, one slot per step of the run
    const int slot = r.niter - b.niter0[p] - 1;
    if (slot >= 0 && slot < b.hist_cap) b.hist[(size_t)slot * b.hist_stride + p] = r;
  }
}
