// Product-form factorisation of the shifted pencil for k_eig (included by k_eig.cuh).
//
// eig_factor / eig_solve (k_eig.cuh) run the textbook LDL^T recurrence d_j = a_j - e_{j-1}^2 / d_{j-1}: every step
// carries a reciprocal on its critical path (multiply, FMA, clamp, rcp.approx + two Newton steps: measured 92 cycles per
// step on sm_100a, tools/micro/latency.cu) and the 19 + 20 steps of a factorisation are strictly sequential, so a
// factorisation cost 14-20 k cycles (tools/micro/eig_factor_bench.cu, profiles/r02_k_eig_experiments.md).  The same
// pivots are ratios of the Sturm sequence
//        p_j = a_j p_{j-1} - e_{j-1}^2 p_{j-2},   d_j = p_j / p_{j-1},
// whose recurrence has ONE dependent FMA per step; the forward-substituted couplings y = L^-1 g follow the same
// pattern through z_j = y_j p_{j-1} = g_j p_{j-1} - e_{j-1} z_{j-1}.  All reciprocals (1/p_j) are off the critical path:
// in the cells they pipeline behind the recurrence, in the hat system they are computed by one thread per row after the
// sweep.  Inertia counts are the sign changes of the sequence (no division at all in the count-only rounds of the hat
// system).  Table values are prefetched EIG_PF_DEPTH steps ahead (the reduced cell data does not fit the L1 that is
// left beside four resident CTAs; each step used to wait for L2).
//
// A zero / non-finite / vanishing term of a sequence (the shift hits an eigenvalue of a leading block to 1e-32) makes
// the caller fall back to the safe ratio form (eig_factor), so the degenerate cases keep their old semantics.
//
// The factors are kept as (1/d_j, l_j, y0_j, y1_j) per cell: eig_solve_pf applies the couplings through the forward
// substituted vectors, t = y^T D^-1 L^-1 r, and folds the hat correction into the right-hand side of the backward
// substitution, so the back substitution W = L^-T D^-1 y of eig_factor (two more sequential sweeps) is not needed.
#pragma once

#define EIG_PF_DEPTH 4

// scale factor 2^(1023 - ex) for a value with biased exponent ex, and its inverse
__device__ __forceinline__ double pf_pow2(int biased) { return __hiloint2double(biased << 20, 0); }

// returns false when a term vanished (caller falls back to eig_factor)
__device__ __forceinline__ bool eig_factor_pf(const EigShared& S, int C, int nb, int ng, double sigma, int store EIG_FT_DECL) {
  const int g = threadIdx.x / C, c = threadIdx.x - g * C;
#ifdef AKS_EIG_TIMING
  if (ft) ft[3] = clock64();
#endif
  bool bad = false;
  if (g < ng) {
    const int nbC = nb * C;
    const double* R = S.red + c;
    double* F = S.fac + c;
    double* sc = S.sc + (size_t)g * 3 * C;
    double qtd[EIG_PF_DEPTH], qte[EIG_PF_DEPTH], qh0[EIG_PF_DEPTH], qh1[EIG_PF_DEPTH], qm0[EIG_PF_DEPTH], qm1[EIG_PF_DEPTH];
#pragma unroll
    for (int u = 0; u < EIG_PF_DEPTH; ++u)
      if (u < nb) {
        qtd[u] = R[u * C]; qte[u] = R[nbC + u * C]; qh0[u] = R[2 * nbC + u * C]; qh1[u] = R[3 * nbC + u * C];
        qm0[u] = R[4 * nbC + u * C]; qm1[u] = R[5 * nbC + u * C];
      }
    const double hh0 = R[6 * nbC], hh1 = R[6 * nbC + C], hh2 = R[6 * nbC + 2 * C];
    double pm = 1.0, pc = 1.0, z0 = 0.0, z1 = 0.0, rprev = 1.0, eprev = 0.0, dinv_prev = 0.0;
    double s00 = 0.0, s10 = 0.0, s11 = 0.0, scale = 0.0;
    int neg = 0;
    for (int j0 = 0; j0 < nb; j0 += EIG_PF_DEPTH) {
#pragma unroll
      for (int u = 0; u < EIG_PF_DEPTH; ++u) {
        const int j = j0 + u;
        if (j < nb) {
          const double td = qtd[u], te = qte[u], h0 = qh0[u], h1 = qh1[u], m0 = qm0[u], m1 = qm1[u];
          if (j + EIG_PF_DEPTH < nb) {
            const int jn = (j + EIG_PF_DEPTH) * C;
            qtd[u] = R[jn]; qte[u] = R[nbC + jn]; qh0[u] = R[2 * nbC + jn]; qh1[u] = R[3 * nbC + jn];
            qm0[u] = R[4 * nbC + jn]; qm1[u] = R[5 * nbC + jn];
          }
          const double a = td - sigma;
          scale = fmax(scale, fabs(a));
          const double pn = fma(a, pc, -((eprev * eprev) * pm));            // the only dependent operation per step
          const double z0n = fma(-eprev, z0, fma(-sigma, m0, h0) * pc);
          const double z1n = fma(-eprev, z1, fma(-sigma, m1, h1) * pc);
          neg += ((pn < 0.0) != (pc < 0.0));
          bad = bad || !(fabs(pn) >= fmax(scale * 4.9e-32, DBL_MIN * 1e8) * fabs(pc)) || !(fabs(pn) < 1e290);
          const double r = pivot_rcp(pn);                                    // off the critical path
          const double y0 = z0n * rprev, y1 = z1n * rprev;
          const double dinv = pc * r;
          s00 = fma(y0 * dinv, y0, s00);
          s10 = fma(y1 * dinv, y0, s10);
          s11 = fma(y1 * dinv, y1, s11);
          if (store) {
            F[j * C] = dinv;
            F[nbC + j * C] = eprev * dinv_prev;
            F[2 * nbC + j * C] = y0;
            F[3 * nbC + j * C] = y1;
          }
          pm = pc; pc = pn; z0 = z0n; z1 = z1n; rprev = r; eprev = te; dinv_prev = dinv;
          if ((j & 7) == 7) {   // keep the sequence in range: common power-of-two factor (exact)
            const int ex = (__double2hiint(pc) >> 20) & 0x7ff;
            const double dn = pf_pow2(2046 - ex), up = pf_pow2(ex);
            pc *= dn; pm *= dn; z0 *= dn; z1 *= dn; rprev *= up;
          }
        }
      }
    }
    sc[3 * c] = (hh0 - sigma * S.mhh[c]) - s00;
    sc[3 * c + 1] = (hh1 - sigma * S.mhh[C + c]) - s10;
    sc[3 * c + 2] = (hh2 - sigma * S.mhh[2 * C + c]) - s11;
    S.negc[g * (C + 2) + c] = neg;
  }
  bad = __syncthreads_or(bad ? 1 : 0) != 0;
  EIG_FT_MARK(0);   // cell sweeps (+ barrier)
  if (bad) return false;
  // ---- hat system: node j sits between cell j (right hat) and cell j+1 (left hat); tridiagonal of order m = C - 1,
  // diag_j = sc[3j] + sc[3(j+1)+2], off_j = sc[3(j+1)+1] couples j and j+1.  Twisted at tw: a sequence from the top
  // (P) and one from the bottom (Q) meet there.
  const int m = C - 1, tw = m / 2;
  double* pnum = S.hatpart;       // [C] scratch: numerator / denominator of 1/d_j = num_j / den_j, both in the scale the
  double* pden = S.hatpart + C;   // [C] sequence had at row j (the sequences are rescaled on the way)
  if (store) {
    const double* sv = S.sc;
    bool hbad = false;
    if (threadIdx.x == 0) {
      double pm = 1.0, pc = 1.0, offp = 0.0;
#pragma unroll 4
      for (int j = 0; j < tw; ++j) {
        const double diag = sv[3 * j] + sv[3 * (j + 1) + 2];
        const double pn = fma(diag, pc, -((offp * offp) * pm));
        hbad = hbad || !(fabs(pn) >= DBL_MIN * 1e8 * fabs(pc)) || !(fabs(pn) < 1e290);
        pnum[j] = pc; pden[j] = pn;
        pm = pc; pc = pn;
        offp = sv[3 * (j + 1) + 1];
        if ((j & 7) == 7) {
          const int ex = (__double2hiint(pc) >> 20) & 0x7ff;
          const double dn = pf_pow2(2046 - ex);
          pc *= dn; pm *= dn;
        }
      }
    } else if (threadIdx.x == 32) {
      double qn = 1.0, qc = 1.0, offn = 0.0;   // qc = Q_{j+1}, qn = Q_{j+2}
#pragma unroll 4
      for (int j = m - 1; j > tw; --j) {
        const double diag = sv[3 * j] + sv[3 * (j + 1) + 2];
        const double q = fma(diag, qc, -((offn * offn) * qn));
        hbad = hbad || !(fabs(q) >= DBL_MIN * 1e8 * fabs(qc)) || !(fabs(q) < 1e290);
        pnum[j] = qc; pden[j] = q;
        qn = qc; qc = q;
        offn = sv[3 * j + 1];   // couples j-1 and j
        if (((m - 1 - j) & 7) == 7) {
          const int ex = (__double2hiint(qc) >> 20) & 0x7ff;
          const double dn = pf_pow2(2046 - ex);
          qc *= dn; qn *= dn;
        }
      }
    }
    if (__syncthreads_or(hbad ? 1 : 0)) return false;
    // one thread per row: reciprocal pivot, multiplier, sign change
    int sgn = 0;
    const int j = threadIdx.x;
    if (j < tw) {
      const double pj = pden[j], pp = pnum[j];
      const double di = pp * pivot_rcp(pj);                 // 1 / d_j
      S.tri_di[j] = di;
      S.tri_l[j] = sv[3 * (j + 1) + 1] * di;                // multiplier of row j + 1 (tri_l[tw-1] feeds the twist row)
      sgn = (pj < 0.0) != (pp < 0.0);
    } else if (j > tw && j < m) {
      const double qj = pden[j], qq = pnum[j];
      const double di = qq * pivot_rcp(qj);
      S.tri_di[j] = di;
      S.tri_l[j - 1] = sv[3 * j + 1] * di;                  // multiplier of row j - 1
      sgn = (qj < 0.0) != (qq < 0.0);
    } else if (j == tw) {
      const double diag = sv[3 * tw] + sv[3 * (tw + 1) + 2];
      double gam = diag;
      if (tw > 0) { const double o = sv[3 * tw + 1]; gam -= o * (pnum[tw - 1] * pivot_rcp(pden[tw - 1])) * o; }
      if (tw < m - 1) { const double o = sv[3 * (tw + 1) + 1]; gam -= o * (pnum[tw + 1] * pivot_rcp(pden[tw + 1])) * o; }
      const double pivmin = fmax(fabs(diag) * 4.9e-32, DBL_MIN * 1e8);
      if (fabs(gam) < pivmin) gam = (gam < 0.0) ? -pivmin : pivmin;
      S.tri_di[tw] = pivot_rcp(gam);
      sgn = gam < 0.0;
    }
    // cells' negative pivots: thread c adds its own
    if (threadIdx.x < C) sgn += S.negc[threadIdx.x];
    // integer sum over the CTA (order independent)
    for (int o = 16; o > 0; o >>= 1) sgn += __shfl_xor_sync(0xffffffffu, sgn, o);
    if ((threadIdx.x & 31) == 0) S.negc[C + 2 + (threadIdx.x >> 5)] = sgn;   // scratch: group-1 slots are unused with ng == 1
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += S.negc[C + 2 + w];
      S.cntg[0] = tot;
    }
  } else {
    bool hbad = false;
    if (c == 0 && g < ng) {
      // count only: one thread per shift runs both sequences interleaved (two independent chains, no division)
      const double* sv = S.sc + (size_t)g * 3 * C;
      const int* ngc = S.negc + g * (C + 2);
      int neg = 0;
#pragma unroll 8
      for (int cc = 0; cc < C; ++cc) neg += ngc[cc];
      double pm = 1.0, pc = 1.0, offp = 0.0, qn = 1.0, qc = 1.0, offn = 0.0;
      const int nup = m - 1 - tw;
      const int nit = max(tw, nup);
#pragma unroll 2
      for (int q = 0; q < nit; ++q) {
        if (q < tw) {
          const int j = q;
          const double diag = sv[3 * j] + sv[3 * (j + 1) + 2];
          const double pn = fma(diag, pc, -((offp * offp) * pm));
          hbad = hbad || !(fabs(pn) >= DBL_MIN * 1e8 * fabs(pc)) || !(fabs(pn) < 1e290);
          neg += ((pn < 0.0) != (pc < 0.0));
          pm = pc; pc = pn;
          offp = sv[3 * (j + 1) + 1];
        }
        if (q < nup) {
          const int j = m - 1 - q;
          const double diag = sv[3 * j] + sv[3 * (j + 1) + 2];
          const double qv = fma(diag, qc, -((offn * offn) * qn));
          hbad = hbad || !(fabs(qv) >= DBL_MIN * 1e8 * fabs(qc)) || !(fabs(qv) < 1e290);
          neg += ((qv < 0.0) != (qc < 0.0));
          qn = qc; qc = qv;
          offn = sv[3 * j + 1];
        }
        if ((q & 7) == 7) {
          const int e1 = (__double2hiint(pc) >> 20) & 0x7ff, e2 = (__double2hiint(qc) >> 20) & 0x7ff;
          const double d1 = pf_pow2(2046 - e1), d2 = pf_pow2(2046 - e2);
          pc *= d1; pm *= d1; qc *= d2; qn *= d2;
        }
      }
      // twist pivot: gam = diag - off_{tw-1}^2 / d_{tw-1} - off_tw^2 / d_{tw+1}
      const double diag = sv[3 * tw] + sv[3 * (tw + 1) + 2];
      double gam = diag;
      if (tw > 0) gam -= offp * (pm / pc) * offp;
      if (tw < m - 1) gam -= offn * (qn / qc) * offn;
      S.cntg[g] = neg + (gam < 0.0);
    }
    if (__syncthreads_or(hbad ? 1 : 0)) return false;
  }
  __syncthreads();
  EIG_FT_MARK(1);   // hat system (+ barriers)
  return true;
}

// ---- y = T~(sigma)^{-1} r with the factors of eig_factor_pf: fac = 1/d | l | y0 | y1 per cell
__device__ __forceinline__ void eig_solve_pf(const EigShared& S, int C, int nb) {
  const int m = C - 1, tw = m / 2;
  const int nbC = nb * C;
  // A. thread per cell: z = L^-1 r_b (kept in y.t), hat partials t = y^T D^-1 z
  {
    const int c = threadIdx.x;
    if (c < C) {
      const double* F = S.fac + c;
      const double* rt = S.r.t + c;
      double* yt = S.y.t + c;
      double t0 = 0.0, t1 = 0.0, zp = 0.0;
      for (int j0 = 0; j0 < nb; j0 += 4) {
        double rj[4], lj[4], dj[4], a0[4], a1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < nb) {
            const int o = (j0 + u) * C;
            rj[u] = rt[o]; dj[u] = F[o]; lj[u] = F[nbC + o]; a0[u] = F[2 * nbC + o]; a1[u] = F[3 * nbC + o];
          }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < nb) {
            const double z = fma(-lj[u], zp, rj[u]);
            yt[(j0 + u) * C] = z;
            const double zd = z * dj[u];
            t0 = fma(a0[u], zd, t0);
            t1 = fma(a1[u], zd, t1);
            zp = z;
          }
      }
      S.hatpart[2 * c] = t0;
      S.hatpart[2 * c + 1] = t1;
    }
  }
  __syncthreads();
  // B. hat right-hand sides (one thread per row), then the twisted substitution
  for (int j = threadIdx.x; j < m; j += blockDim.x) S.xh[j] = S.r.h[j] - S.hatpart[2 * j] - S.hatpart[2 * (j + 1) + 1];
  __syncthreads();
  if (threadIdx.x == 0) {
    double zprev = 0.0;
    for (int j0 = 0; j0 <= tw; j0 += 4) {
      double rh[4], ll[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u <= tw) { rh[u] = S.xh[j0 + u]; ll[u] = (j0 + u > 0) ? S.tri_l[j0 + u - 1] : 0.0; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u <= tw) {
          const double v = fma(-ll[u], zprev, rh[u]);
          if (j0 + u < tw) { S.xh[j0 + u] = v; zprev = v; } else S.red8[AKS_NW] = v;
        }
    }
  } else if (threadIdx.x == 32) {
    double znext = 0.0;
    for (int j0 = m - 1; j0 > tw; j0 -= 4) {
      double rh[4], ll[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 - u > tw) { rh[u] = S.xh[j0 - u]; ll[u] = (j0 - u < m - 1) ? S.tri_l[j0 - u] : 0.0; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 - u > tw) {
          const double v = fma(-ll[u], znext, rh[u]);
          S.xh[j0 - u] = v;
          znext = v;
        }
    }
    S.red8[AKS_NW + 1] = (tw < m - 1) ? S.tri_l[tw] * znext : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0 || threadIdx.x == 32) {
    const double xt = (S.red8[AKS_NW] - S.red8[AKS_NW + 1]) * S.tri_di[tw];
    if (threadIdx.x == 0) {
      S.y.h[tw] = xt;
      double xn = xt;
      for (int j0 = tw - 1; j0 >= 0; j0 -= 4) {
        double xd[4], ll[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 - u >= 0) { xd[u] = S.xh[j0 - u] * S.tri_di[j0 - u]; ll[u] = S.tri_l[j0 - u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 - u >= 0) {
            const double v = fma(-ll[u], xn, xd[u]);
            xn = v;
            S.y.h[j0 - u] = v;
          }
      }
    } else {
      double xp = xt;
      for (int j0 = tw + 1; j0 < m; j0 += 4) {
        double xd[4], ll[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < m) { xd[u] = S.xh[j0 + u] * S.tri_di[j0 + u]; ll[u] = S.tri_l[j0 + u - 1]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < m) {
            const double v = fma(-ll[u], xp, xd[u]);
            xp = v;
            S.y.h[j0 + u] = v;
          }
      }
    }
  }
  __syncthreads();
  // C. thread per cell: x_b = L^-T D^-1 (z - y0 x_R - y1 x_L)
  {
    const int c = threadIdx.x;
    if (c < C) {
      const double* F = S.fac + c;
      double* yt = S.y.t + c;
      const double xh0 = (c < C - 1) ? S.y.h[c] : 0.0, xh1 = (c > 0) ? S.y.h[c - 1] : 0.0;
      double xn = 0.0;
      for (int j0 = nb - 1; j0 >= 0; j0 -= 4) {
        double vv[4], ln[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 - u >= 0) {
            const int j = j0 - u, o = j * C;
            const double zz = yt[o] - F[2 * nbC + o] * xh0 - F[3 * nbC + o] * xh1;
            vv[u] = zz * F[o];
            ln[u] = (j < nb - 1) ? F[nbC + o + C] : 0.0;
          }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 - u >= 0) {
            const double x = fma(-ln[u], xn, vv[u]);
            yt[(j0 - u) * C] = x;
            xn = x;
          }
      }
    }
  }
  __syncthreads();
}
