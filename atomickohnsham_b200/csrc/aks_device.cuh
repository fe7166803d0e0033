// Device helpers: deterministic reductions, LDA/LSDA exchange-correlation, Brent.
//
// XC formulas follow the reference's builtin functionals
//   src/physics/exchange correlation/SlaterXa.jl:11-30   (Slater exchange)
//   src/physics/exchange correlation/PerdewWang.jl:15-101 (PW92 correlation)
//   src/physics/exchange correlation/BuiltinFunctional.jl:74-142 (clamp at 0, zero at rho = 0)
// x^(1/3), x^(3/2), x^(4/3) are evaluated from ONE rcbrt(rho) per point (rho^(1/3) = rho * rcbrt^2,
// rs = (3/4pi)^(1/3) * rcbrt), x*sqrt(x), x*cbrt(x) (the reference calls pow), and the two quotients of
// G' share one division; the results differ by a few ulp, far inside the 1e-10 parity tolerance.
// Brent follows Optim.jl's `Brent()` as called from
//   src/algorithms/optimization/line_search.jl:59-67.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define AKS_NT 256          // threads per CTA for the cell / eigen kernels
#define AKS_NW (AKS_NT / 32)
#define POT_IMG_ROWS 25     // rows of a 64-node chunk image of the generator table in DiscDev::Pgc (row 24: the nodes)
#define AKS_MAXORB 96       // max orbitals (L*nh*nspin) per problem held in shared lists

struct XcConsts {
  double cx;      // (3/pi)^(1/3)
  double cx34;    // 3/4 (3/pi)^(1/3)
  double cx6_34;  // 3/4 (6/pi)^(1/3)
  double fourpi;  // 4 pi
  double inv4pi;  // 1/(4 pi)
  double f2p0;    // 4/(9(2^(1/3)-1))
  double two43m2; // 2^(4/3)-2
  double third;   // 1/3 as a double
  double crs;     // (3/(4 pi))^(1/3): rs = crs * rho^(-1/3)
  double icrs;    // 1/crs
  double four3;   // 4/3 as a double
};
__constant__ XcConsts c_xc;

// ---------------------------------------------------------------- reductions (fixed order)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Sum over the CTA; every thread receives the same value. `red` holds >= blockDim.x/32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+): a "full" barrier per stage completes on the
// bytes of the TMA transfers, an "empty" barrier per stage counts the warps that are done with the stage
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- PW92 pieces
struct PwPar { double A, a1, b1, b2, b3, b4; };
__device__ __forceinline__ PwPar pw_p0() { return {0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294}; }
__device__ __forceinline__ PwPar pw_p1() { return {0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517}; }
__device__ __forceinline__ PwPar pw_pa() { return {0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671}; }

// rho^(1/3) and rs = (3/(4 pi rho))^(1/3) from one reciprocal cube root
__device__ __forceinline__ void rho_roots(double rho, double& cr, double& rs) {
  const double ic = rcbrt(rho);
  cr = (rho * ic) * ic;
  rs = c_xc.crs * ic;
}
// G(rs) and dG/drs (PerdewWang.jl:15-27); sq = sqrt(rs), isq = 1/sqrt(rs)
__device__ __forceinline__ void pw_G(double rs, double sq, double isq, const PwPar& P, double& G, double& dG) {
  const double rs32 = rs * sq;
  const double Q0 = -2.0 * P.A * (1.0 + P.a1 * rs);
  const double dQ0 = -2.0 * P.A * P.a1;
  const double Q1 = 2.0 * P.A * (P.b1 * sq + P.b2 * rs + P.b3 * rs32 + P.b4 * (rs * rs));
  const double dQ1 = 2.0 * P.A * (0.5 * P.b1 * isq + P.b2 + 1.5 * P.b3 * sq + 2.0 * P.b4 * rs);
  const double iq = 1.0 / (Q1 * (Q1 + 1.0));        // 1/Q1 = (Q1 + 1) * iq
  const double lg = log(1.0 + (Q1 + 1.0) * iq);
  G = Q0 * lg;
  dG = dQ0 * lg - Q0 * dQ1 * iq;
}
// G(rs) alone
__device__ __forceinline__ double pw_G0(double rs, double sq, const PwPar& P) {
  const double Q0 = -2.0 * P.A * (1.0 + P.a1 * rs);
  const double Q1 = 2.0 * P.A * (P.b1 * sq + P.b2 * rs + P.b3 * (rs * sq) + P.b4 * (rs * rs));
  return Q0 * log(1.0 + 1.0 / Q1);
}

// ---------------------------------------------------------------- unpolarised
// vxc = vx + vc at density rho (already clamped >= 0)
__device__ __forceinline__ double xc_vrho_unpol(double rho, int ex, int ec) {
  double v = 0.0;
  if (rho <= 0.0) return 0.0;
  double cr, rs;
  rho_roots(rho, cr, rs);
  if (ec == 12) {
    const double sq = sqrt(rs);
    double G, dG;
    pw_G(rs, sq, sq * (cr * c_xc.icrs), pw_p0(), G, dG);   // 1/sqrt(rs) = sqrt(rs) / rs, 1/rs = rho^(1/3) / crs
    v += G - rs / 3.0 * dG;
  }
  if (ex == 1) v += -c_xc.cx * cr;
  return v;
}
// exc per particle = ex + ec
__device__ __forceinline__ double xc_zk_unpol(double rho, int ex, int ec) {
  double e = 0.0;
  if (rho <= 0.0) return 0.0;
  double cr, rs;
  rho_roots(rho, cr, rs);
  if (ec == 12) e += pw_G0(rs, sqrt(rs), pw_p0());
  if (ex == 1) e += -c_xc.cx34 * cr;
  return e;
}

// ---------------------------------------------------------------- polarised
__device__ __forceinline__ double pw_fz(double z) {
  return ((1.0 + z) * cbrt(1.0 + z) + (1.0 - z) * cbrt(1.0 - z) - 2.0) / c_xc.two43m2;
}
__device__ __forceinline__ double pw_dfz(double z) {
  return c_xc.four3 * (cbrt(1.0 + z) - cbrt(1.0 - z)) / c_xc.two43m2;
}

__device__ __forceinline__ void xc_vrho_pol(double ru, double rd, int ex, int ec, double& vu,
                                            double& vd) {
  vu = 0.0;
  vd = 0.0;
  const double rho = ru + rd;
  if (ec == 12 && rho > 0.0) {
    const double z = (ru - rd) / rho;
    double cr, rs;
    rho_roots(rho, cr, rs);
    const double sq = sqrt(rs), isq = sq * (cr * c_xc.icrs);
    double e0, d0, e1, d1, ea, da;
    pw_G(rs, sq, isq, pw_p0(), e0, d0);
    pw_G(rs, sq, isq, pw_p1(), e1, d1);
    pw_G(rs, sq, isq, pw_pa(), ea, da);
    const double ac = -ea, dac = -da;
    const double f = pw_fz(z), df = pw_dfz(z);
    const double z3 = z * z * z, z4 = z3 * z;
    const double ecz = e0 + ac * (f / c_xc.f2p0) * (1.0 - z4) + (e1 - e0) * f * z4;
    const double decdrs = d0 + dac * (f / c_xc.f2p0) * (1.0 - z4) + (d1 - d0) * f * z4;
    const double decdz = ac / c_xc.f2p0 * (df * (1.0 - z4) - 4.0 * z3 * f) +
                         (e1 - e0) * (df * z4 + 4.0 * z3 * f);
    const double common = ecz - rs / 3.0 * decdrs;
    vu += common + (1.0 - z) * decdz;
    vd += common - (1.0 + z) * decdz;
  }
  if (ex == 1) {
    vu += (ru > 0.0) ? -c_xc.cx * cbrt(2.0 * ru) : 0.0;
    vd += (rd > 0.0) ? -c_xc.cx * cbrt(2.0 * rd) : 0.0;
  }
}

__device__ __forceinline__ double xc_zk_pol(double ru, double rd, int ex, int ec) {
  const double rho = ru + rd;
  if (rho <= 0.0) return 0.0;
  double e = 0.0;
  if (ec == 12) {
    const double z = (ru - rd) / rho;
    double cr, rs;
    rho_roots(rho, cr, rs);
    const double sq = sqrt(rs);
    const double e0 = pw_G0(rs, sq, pw_p0()), e1 = pw_G0(rs, sq, pw_p1()), ea = pw_G0(rs, sq, pw_pa());
    const double f = pw_fz(z);
    const double z4 = z * z * z * z;
    e += e0 + (-ea) * (f / c_xc.f2p0) * (1.0 - z4) + (e1 - e0) * f * z4;
  }
  if (ex == 1) e += -c_xc.cx6_34 * (ru * cbrt(ru) + rd * cbrt(rd)) / rho;
  return e;
}

// ---------------------------------------------------------------- Brent (one thread drives)
// State machine form so that a whole CTA can evaluate f(t) between steps.
struct BrentState {
  double lo, hi, x, fx, ox, ofx, oox, oofx, step, old_step;
  double abs_tol, rel_tol;
  int it, maxit;
  double next;  // abscissa to evaluate next
};
#define AKS_GOLDEN 0.3819660112501051  // (3 - sqrt 5)/2

__device__ inline void brent_init(BrentState& s, double lo, double hi, double abs_tol,
                                  double rel_tol, int maxit) {
  s.lo = lo; s.hi = hi; s.abs_tol = abs_tol; s.rel_tol = rel_tol; s.maxit = maxit; s.it = 0;
  s.step = 0.0; s.old_step = 0.0;
  s.next = lo + AKS_GOLDEN * (hi - lo);
}
__device__ inline void brent_first(BrentState& s, double f) {
  s.x = s.ox = s.oox = s.next;
  s.fx = s.ofx = s.oofx = f;
}
// Propose the next abscissa; returns false when converged / out of iterations.
__device__ inline bool brent_propose(BrentState& s) {
  if (s.it >= s.maxit) return false;
  double p = 0.0, q = 0.0;
  const double x_tol = s.rel_tol * fabs(s.x) + s.abs_tol;
  const double mid = (s.hi + s.lo) / 2.0;
  if (fabs(s.x - mid) <= 2.0 * x_tol - (s.hi - s.lo) / 2.0) return false;
  s.it += 1;
  if (fabs(s.old_step) > x_tol) {
    const double r = (s.x - s.ox) * (s.fx - s.oofx);
    q = (s.x - s.oox) * (s.fx - s.ofx);
    p = (s.x - s.oox) * q - (s.x - s.ox) * r;
    q = 2.0 * (q - r);
    if (q > 0.0) p = -p; else q = -q;
  }
  if (fabs(p) < fabs(q * s.old_step / 2.0) && p < q * (s.hi - s.x) && p < q * (s.x - s.lo)) {
    s.old_step = s.step;
    s.step = p / q;
    const double xt = s.x + s.step;
    if ((xt - s.lo) < 2.0 * x_tol || (s.hi - xt) < 2.0 * x_tol)
      s.step = (s.x < mid) ? x_tol : -x_tol;
  } else {
    s.old_step = (s.x < mid) ? s.hi - s.x : s.lo - s.x;
    s.step = AKS_GOLDEN * s.old_step;
  }
  if (fabs(s.step) >= x_tol) s.next = s.x + s.step;
  else s.next = s.x + ((s.step > 0.0) ? x_tol : -x_tol);
  return true;
}
__device__ inline void brent_update(BrentState& s, double nf) {
  const double nx = s.next;
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

// FP64 tensor-core tile update C(8x8) += A(8x4) B(4x8), mma.sync m8n8k4: lane holds A[lane/4][lane%4], B[lane%4][lane/4]
// and C[lane/4][2 (lane%4) + {0,1}]
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
