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
