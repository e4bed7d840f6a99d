// Scalar arithmetic for the real and USE_COMPLEX builds (reference src/include/scalar.h:12-24).
//
// `cplx` restates what g++ emits for std::complex<double> on finite operands:
//   complex*complex  -> (ac - bd, ad + bc)               (libgcc __muldc3 fast path)
//   complex*double, complex/double, complex+-double      -> component-wise (libstdc++ <complex>)
//   complex/complex  -> Smith's scaled division           (agrees with libgcc __divdc3 to rounding)
// No FMA contraction: the translation units that include this for TABLE construction are compiled
// with -fmad=false so that each product/sum rounds once, like the reference's x86-64 build.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define PS_HD __host__ __device__ __forceinline__
#else
#define PS_HD inline   /* plain host compilers (libpspace.so) instantiate the same formulas */
#endif

struct cplx {
  double re, im;
};

PS_HD cplx mkc(double r, double i = 0.0) { cplx c; c.re = r; c.im = i; return c; }

// ---- a tiny "scalar trait" layer: the same formulas compile for double and cplx -------------------
PS_HD double s_make(double r, double) { return r; }
PS_HD double s_add(double a, double b) { return a + b; }
PS_HD double s_sub(double a, double b) { return a - b; }
PS_HD double s_mul(double a, double b) { return a * b; }
PS_HD double s_div(double a, double b) { return a / b; }
PS_HD double s_neg(double a) { return -a; }
PS_HD double s_muld(double a, double d) { return a * d; }   // scalar * double
PS_HD double s_divd(double a, double d) { return a / d; }   // scalar / double
PS_HD double s_addd(double a, double d) { return a + d; }   // scalar + double
PS_HD double s_subd(double a, double d) { return a - d; }   // scalar - double
PS_HD double s_dsub(double d, double a) { return d - a; }   // double - scalar
PS_HD double s_sqrt(double a) { return sqrt(a); }
PS_HD double s_from(double d, double) { return d; }

PS_HD cplx s_add(cplx a, cplx b) { return mkc(a.re + b.re, a.im + b.im); }
PS_HD cplx s_sub(cplx a, cplx b) { return mkc(a.re - b.re, a.im - b.im); }
PS_HD cplx s_mul(cplx a, cplx b) { return mkc(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
PS_HD cplx s_neg(cplx a) { return mkc(-a.re, -a.im); }
PS_HD cplx s_muld(cplx a, double d) { return mkc(a.re * d, a.im * d); }
PS_HD cplx s_divd(cplx a, double d) { return mkc(a.re / d, a.im / d); }
PS_HD cplx s_addd(cplx a, double d) { return mkc(a.re + d, a.im); }
PS_HD cplx s_subd(cplx a, double d) { return mkc(a.re - d, a.im); }
PS_HD cplx s_dsub(double d, cplx a) { return mkc(d - a.re, -a.im); }
PS_HD cplx s_div(cplx a, cplx b) {
  // Smith (1962): scale by the larger component of the divisor
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re, den = b.re + b.im * r;
    return mkc((a.re + a.im * r) / den, (a.im - a.re * r) / den);
  } else {
    double r = b.re / b.im, den = b.re * r + b.im;
    return mkc((a.re * r + a.im) / den, (a.im * r - a.re) / den);
  }
}
PS_HD cplx s_sqrt(cplx a) {
  // principal square root; exact (sqrt(x), +0) for a = x + 0i, x >= 0 — the only use in the tables
  if (a.im == 0.0 && a.re >= 0.0) return mkc(sqrt(a.re), 0.0);
  double m = hypot(a.re, a.im);
  double sr = sqrt(0.5 * (m + fabs(a.re)));
  double si = a.im / (2.0 * sr);
  if (a.re >= 0.0) return mkc(sr, si);
  return mkc(fabs(si), a.im < 0.0 ? -sr : sr);
}

template <typename S> struct scalar_of;
template <> struct scalar_of<double> {
  static PS_HD double from(double d) { return d; }
  static PS_HD double load(const double *p, long long i) { return p[i]; }
  static PS_HD void store(double *p, long long i, double v) { p[i] = v; }
  static constexpr int width = 1;
};
template <> struct scalar_of<cplx> {
  static PS_HD cplx from(double d) { return mkc(d, 0.0); }
  static PS_HD cplx load(const double *p, long long i) { return mkc(p[2 * i], p[2 * i + 1]); }
  static PS_HD void store(double *p, long long i, cplx v) { p[2 * i] = v.re; p[2 * i + 1] = v.im; }
  static constexpr int width = 2;
};
