// Double-double arithmetic (the reference's T = Double64, DoubleFloats.jl: an unevaluated sum hi + lo of
// two Float64, ~106 significant bits) for the Double64 instantiation of the hot path (SURVEY.md section 8
// row a20).  Error-free transformations: TwoSum (Knuth) and TwoProd through the hardware FMA, as the
// north star names them.  Every routine is __host__ __device__ so that the setup code of aks_disc_create
// (table splitting) can use the same arithmetic; there is no host path through the SCF loop.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

struct dd {
  double hi, lo;
};

#define DD_HD __host__ __device__ __forceinline__

DD_HD dd dd_make(double hi, double lo = 0.0) { dd r; r.hi = hi; r.lo = lo; return r; }
DD_HD dd dd_two_sum(double a, double b) {
  const double s = a + b, bb = s - a;
  return dd_make(s, (a - (s - bb)) + (b - bb));
}
DD_HD dd dd_quick_two_sum(double a, double b) {   // |a| >= |b|
  const double s = a + b;
  return dd_make(s, b - (s - a));
}
DD_HD dd dd_two_prod(double a, double b) {
  const double p = a * b;
  return dd_make(p, fma(a, b, -p));
}
// accurate ("IEEE") addition: both components through TwoSum
DD_HD dd operator+(dd a, dd b) {
  dd s = dd_two_sum(a.hi, b.hi);
  const dd t = dd_two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = dd_quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return dd_quick_two_sum(s.hi, s.lo);
}
DD_HD dd operator-(dd a) { return dd_make(-a.hi, -a.lo); }
DD_HD dd operator-(dd a, dd b) { return a + (-b); }
DD_HD dd operator+(dd a, double b) {
  dd s = dd_two_sum(a.hi, b);
  s.lo += a.lo;
  return dd_quick_two_sum(s.hi, s.lo);
}
DD_HD dd operator-(dd a, double b) { return a + (-b); }
DD_HD dd operator*(dd a, dd b) {
  dd p = dd_two_prod(a.hi, b.hi);
  p.lo += fma(a.hi, b.lo, a.lo * b.hi);
  return dd_quick_two_sum(p.hi, p.lo);
}
DD_HD dd operator*(dd a, double b) {
  dd p = dd_two_prod(a.hi, b);
  p.lo = fma(a.lo, b, p.lo);
  return dd_quick_two_sum(p.hi, p.lo);
}
DD_HD dd operator*(double b, dd a) { return a * b; }
// a + b*c
DD_HD dd dd_fma(dd b, dd c, dd a) { return a + b * c; }
DD_HD dd operator/(dd a, dd b) {
  // long division with three quotient terms (full double-double accuracy)
  const double q1 = a.hi / b.hi;
  dd r = a - b * q1;
  const double q2 = r.hi / b.hi;
  r = r - b * q2;
  const double q3 = r.hi / b.hi;
  dd q = dd_quick_two_sum(q1, q2);
  return q + q3;
}
DD_HD dd operator/(dd a, double b) { return a / dd_make(b); }
DD_HD dd dd_inv(dd b) { return dd_make(1.0) / b; }
DD_HD dd dd_sqrt(dd a) {
  if (a.hi <= 0.0) return dd_make(0.0);
  // Karp's trick: x = 1/sqrt(a.hi); sqrt(a) ~ a*x + (a - (a*x)^2) * x / 2
  const double x = 1.0 / sqrt(a.hi), ax = a.hi * x;
  const dd sq = dd_two_prod(ax, ax);
  const dd r = a - sq;
  dd s = dd_two_sum(ax, r.hi * (x * 0.5));
  // one more Newton step in full dd for the last bits
  const dd e = a - s * s;
  return s + dd_make(e.hi / (2.0 * s.hi));
}
DD_HD bool operator<(dd a, dd b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
DD_HD bool operator>(dd a, dd b) { return b < a; }
DD_HD bool operator<=(dd a, dd b) { return !(b < a); }
DD_HD bool operator>=(dd a, dd b) { return !(a < b); }
DD_HD bool operator==(dd a, dd b) { return a.hi == b.hi && a.lo == b.lo; }
DD_HD dd dd_abs(dd a) { return (a.hi < 0.0 || (a.hi == 0.0 && a.lo < 0.0)) ? -a : a; }
DD_HD dd dd_min(dd a, dd b) { return (a < b) ? a : b; }
DD_HD dd dd_max(dd a, dd b) { return (a < b) ? b : a; }
DD_HD bool dd_isfinite(dd a) { return isfinite(a.hi) && isfinite(a.lo); }

// 1/k for a small integer k to double-double accuracy: hi = fl(1/k), lo = (1 - k hi)/k with the residual exact (FMA)
DD_HD dd dd_inv_int(double k) {
  const double hi = 1.0 / k;
  return dd_make(hi, fma(-hi, k, 1.0) / k);
}
// ---- elementary functions (the reference's Double64 functionals call log, ^ and sqrt of DoubleFloats)
#define DD_LN2_HI 6.931471805599452862e-01
#define DD_LN2_LO 2.319046813846299558e-17
DD_HD dd dd_ldexp(dd a, int e) { return dd_make(ldexp(a.hi, e), ldexp(a.lo, e)); }
// exp: x = k ln2 + r, r / 256 by Taylor on expm1, eight doublings p <- 2p + p^2, 2^k
DD_HD dd dd_exp(dd x) {
  if (x.hi <= -709.0) return dd_make(0.0);
  if (x.hi >= 709.0) return dd_make(INFINITY);
  const double kf = floor(x.hi / DD_LN2_HI + 0.5);
  const dd ln2 = dd_make(DD_LN2_HI, DD_LN2_LO);
  dd r = dd_ldexp(x - ln2 * kf, -8);
  // expm1(r) = r + r^2/2! + ... (|r| < 1.4e-3: 11 terms reach 1e-36)
  dd term = r, p = r;
  for (int n = 2; n <= 11; ++n) {
    term = (term * r) * dd_inv_int((double)n);
    p = p + term;
  }
  for (int i = 0; i < 8; ++i) p = p * 2.0 + p * p;
  return dd_ldexp(p + 1.0, (int)kf);
}
// log by one Newton step on exp from the double logarithm: y <- y + x exp(-y) - 1
DD_HD dd dd_log(dd x) {
  if (x.hi <= 0.0) return dd_make(x.hi == 0.0 ? -INFINITY : NAN);
  dd y = dd_make(log(x.hi));
  y = y + (x * dd_exp(-y) - 1.0);
  return y;
}
// log(1 + e) without forming 1 + e when e is small: 2 atanh(e / (2 + e))
DD_HD dd dd_log1p(dd e) {
  if (fabs(e.hi) > 0.25) return dd_log(e + 1.0);
  const dd z = e / (e + 2.0), z2 = z * z;
  dd term = z, acc = z;
  for (int k = 3; k <= 61; k += 2) {
    term = term * z2;
    const dd add = term * dd_inv_int((double)k);
    acc = acc + add;
    if (fabs(add.hi) < 1e-36 * fabs(acc.hi)) break;
  }
  return acc * 2.0;
}
DD_HD dd dd_pow(dd x, dd e) {   // x > 0
  if (x.hi <= 0.0) return dd_make(0.0);
  return dd_exp(e * dd_log(x));
}
DD_HD dd dd_cbrt(dd x) {        // x > 0: two Newton steps from the double cube root
  if (x.hi <= 0.0) return dd_make(0.0);
  dd y = dd_make(cbrt(x.hi));
  for (int i = 0; i < 2; ++i) {
    const dd y2 = y * y;
    y = y - (y2 * y - x) / (y2 * 3.0);
  }
  return y;
}

// ---- a scalar trait so that the cell-local linear algebra (static condensation, LDL^T, Sturm counts) is
// written once and instantiated for double (bracketing) and dd (the results)
template <class R> struct Sc;
template <> struct Sc<double> {
  DD_HD static double from(dd v) { return v.hi + v.lo; }
  DD_HD static double fromd(double v) { return v; }
  DD_HD static double hi(double v) { return v; }
  DD_HD static double inv(double v) { return 1.0 / v; }
};
template <> struct Sc<dd> {
  DD_HD static dd from(dd v) { return v; }
  DD_HD static dd fromd(double v) { return dd_make(v); }
  DD_HD static double hi(dd v) { return v.hi; }
  DD_HD static dd inv(dd v) { return dd_inv(v); }
};
