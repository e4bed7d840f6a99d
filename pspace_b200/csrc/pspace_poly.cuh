// Orthonormal 1-D polynomial families and the affine maps of the 1-D Gauss rules, as host+device
// templates over the scalar type.  These restate the reference's *formula mix* (not textbook
// recurrences), because the parity target is the reference arithmetic (SURVEY.md §7 "hard parts"):
//   Hermite  (probabilists')   d<=4 explicit monomials, d>=5  h_d = z*h_{d-1} - (d-1)*h_{d-2}
//                              unit = h / sqrt(d!)                 OrthogonalPolynomials.cpp:28-44,148-180
//   Laguerre                   d<=4 explicit / d!, d>=5 ((2d-1 - z)*l_{d-1} - (d-1)*l_{d-2})/d
//                              unit = l                            OrthogonalPolynomials.cpp:52-68,188-224
//   shifted Legendre on [0,1]  d<=4 explicit, d>=5 (-1)^d sum_k C(d,k) C(d+k,k) (-z)^k
//                              unit = p * sqrt(2d+1)               OrthogonalPolynomials.cpp:76-102,232-261
// The reference's naive double recursion (recursive_hermite/laguerre) performs exactly the arithmetic
// of the iterative three-term recurrence started from (1, z) resp. (1, 1-z); the iteration below is
// that recurrence.  Must be compiled with -fmad=false where bit-exact tables are wanted.
#pragma once
#include "pspace_scalar.cuh"

#define PSPACE_KIND_NORMAL 0
#define PSPACE_KIND_UNIFORM 1
#define PSPACE_KIND_EXPONENTIAL 2

// d! as the reference builds it: literals to 10!, running product above (OrthogonalPolynomials.cpp:109-140)
PS_HD double ps_factorial(int n) {
  double f = 1.0;
  for (int i = 2; i <= n; i++) f *= (double)i;   // exact in double for n <= 22, equal to the literals for n <= 10
  return f;
}
// nCr = n!/(r!(n-r)!) with the reference's grouping nfact/(rfact*nrfact) (OrthogonalPolynomials.cpp:97-102)
PS_HD double ps_comb(int n, int r) { return ps_factorial(n) / (ps_factorial(r) * ps_factorial(n - r)); }

// x^k for integer k >= 0.
//  real:    the reference calls libm pow(double,double) (std::pow(double,int) promotes).  glibc's pow is
//           correctly rounded in all but pathological cases; here x^k is formed in double-double
//           (error-free products) and rounded once, which gives the same bits wherever glibc rounds correctly.
//  complex: libstdc++ std::pow(complex<double>,int) is binary powering (__complex_pow_unsigned) — restated.
PS_HD void ps_two_prod(double a, double b, double &p, double &e) {
  p = a * b;
  e = fma(a, b, -p);
}
PS_HD double ps_powi(double x, int k) {
  if (k == 0) return 1.0;
  double hi = x, lo = 0.0;       // running power as hi + lo
  for (int i = 1; i < k; i++) {
    double p, e;
    ps_two_prod(hi, x, p, e);
    e = fma(lo, x, e);            // fma here is deliberate (error term), unaffected by -fmad
    double s = p + e;
    lo = e - (s - p);
    hi = s;
  }
  return hi;
}
PS_HD cplx ps_powi(cplx x, int k) {
  unsigned n = (unsigned)k;
  cplx y = (n & 1u) ? x : mkc(1.0, 0.0);
  while (n >>= 1) {
    x = s_mul(x, x);
    if (n & 1u) y = s_mul(y, x);
  }
  return y;
}

template <typename S> PS_HD S ps_hermite(S z, int d) {
  typedef scalar_of<S> T;
  if (d == 0) return T::from(1.0);
  if (d == 1) return z;
  if (d == 2) return s_subd(s_mul(z, z), 1.0);                                                      // z*z - 1.0
  if (d == 3) return s_sub(s_mul(s_mul(z, z), z), s_muld(z, 3.0));                                  // z*z*z - 3.0*z
  if (d == 4) return s_addd(s_sub(s_mul(s_mul(s_mul(z, z), z), z), s_mul(s_muld(z, 6.0), z)), 3.0); // z*z*z*z - 6.0*z*z + 3.0
  S h0 = T::from(1.0), h1 = z;
  for (int n = 2; n <= d; n++) {
    S h2 = s_sub(s_mul(z, h1), s_mul(T::from((double)(n - 1)), h0));
    h0 = h1; h1 = h2;
  }
  return h1;
}
template <typename S> PS_HD S ps_unit_hermite(S z, int d) {
  typedef scalar_of<S> T;
  return s_div(ps_hermite(z, d), s_sqrt(T::from(ps_factorial(d))));
}

template <typename S> PS_HD S ps_laguerre(S z, int d) {
  typedef scalar_of<S> T;
  if (d == 0) return T::from(1.0);
  if (d == 1) return s_dsub(1.0, z);
  if (d == 2) {  // (z*z - 4.0*z + 2.0)/2!
    S l = s_addd(s_sub(s_mul(z, z), s_muld(z, 4.0)), 2.0);
    return s_div(l, T::from(2.0));
  }
  if (d == 3) {  // (-z*z*z + 9.0*z*z - 18.0*z + 6.0)/3!
    S l = s_mul(s_mul(s_neg(z), z), z);
    l = s_add(l, s_mul(s_muld(z, 9.0), z));
    l = s_sub(l, s_muld(z, 18.0));
    l = s_addd(l, 6.0);
    return s_div(l, T::from(6.0));
  }
  if (d == 4) {  // (z^4 - 16 z^3 + 72 z^2 - 96 z + 24)/4!
    S l = s_mul(s_mul(s_mul(z, z), z), z);
    l = s_sub(l, s_mul(s_mul(s_muld(z, 16.0), z), z));
    l = s_add(l, s_mul(s_muld(z, 72.0), z));
    l = s_sub(l, s_muld(z, 96.0));
    l = s_addd(l, 24.0);
    return s_div(l, T::from(24.0));
  }
  S l0 = T::from(1.0), l1 = s_dsub(1.0, z);
  for (int n = 2; n <= d; n++) {
    S a = s_sub(T::from((double)(2 * n - 1)), z);
    S l2 = s_sub(s_mul(a, l1), s_mul(T::from((double)(n - 1)), l0));
    l2 = s_div(l2, T::from((double)n));
    l0 = l1; l1 = l2;
  }
  return l1;
}
template <typename S> PS_HD S ps_unit_laguerre(S z, int d) { return ps_laguerre(z, d); }

template <typename S> PS_HD S ps_legendre(S z, int d) {
  typedef scalar_of<S> T;
  if (d == 0) return T::from(1.0);
  if (d == 1) return s_subd(s_muld(z, 2.0), 1.0);                                       // 2.0*z - 1.0
  if (d == 2) return s_addd(s_sub(s_mul(s_muld(z, 6.0), z), s_muld(z, 6.0)), 1.0);      // 6.0*z*z - 6.0*z + 1.0
  if (d == 3) {  // 20.0*z*z*z - 30.0*z*z + 12.0*z - 1.0
    S p = s_mul(s_mul(s_muld(z, 20.0), z), z);
    p = s_sub(p, s_mul(s_muld(z, 30.0), z));
    p = s_add(p, s_muld(z, 12.0));
    return s_subd(p, 1.0);
  }
  if (d == 4) {  // 70 z^4 - 140 z^3 + 90 z^2 - 20 z + 1
    S p = s_mul(s_mul(s_mul(s_muld(z, 70.0), z), z), z);
    p = s_sub(p, s_mul(s_mul(s_muld(z, 140.0), z), z));
    p = s_add(p, s_mul(s_muld(z, 90.0), z));
    p = s_sub(p, s_muld(z, 20.0));
    return s_addd(p, 1.0);
  }
  S p = T::from(0.0);
  S mz = s_neg(z);
  for (int k = 0; k <= d; k++) {
    S coef = s_mul(T::from(ps_comb(d, k)), T::from(ps_comb(d + k, k)));
    p = s_add(p, s_mul(coef, ps_powi(mz, k)));
  }
  return s_muld(p, (d & 1) ? -1.0 : 1.0);   // pow(-1.0, d)
}
template <typename S> PS_HD S ps_unit_legendre(S z, int d) {
  typedef scalar_of<S> T;
  return s_mul(ps_legendre(z, d), s_sqrt(T::from((double)(2 * d + 1))));
}

// {Normal,Uniform,Exponential}Parameter::basis (…Parameter.cpp:42-44)
template <typename S> PS_HD S ps_param_basis(int kind, S z, int d) {
  if (kind == PSPACE_KIND_NORMAL) return ps_unit_hermite(z, d);
  if (kind == PSPACE_KIND_UNIFORM) return ps_unit_legendre(z, d);
  return ps_unit_laguerre(z, d);
}

// Affine maps of the 1-D rules: raw node x / weight w  ->  (z, y, w)
//   Hermite  y = mu + sigma*sqrt(2.0)*x ; z = (y-mu)/sigma ; w /= sqrt(4.0*atan(1.0))   GaussianQuadrature.cpp:510-514
//   Legendre y = ((b-a)/2.0)*x + (b+a)/2.0 ; z = (y-a)/(b-a) ; w /= 2.0                 GaussianQuadrature.cpp:1026-1032
//   Laguerre y = mu + beta*x ; z = x ; w = w                                            GaussianQuadrature.cpp:1479-1482
// In the complex build x and w are themselves `scalar` (x + 0i), hence complex*complex below.
template <typename S> PS_HD void ps_map_rule(int kind, S p1, S p2, double x, double w, S &z, S &y, S &wo) {
  typedef scalar_of<S> T;
  S xs = T::from(x), ws = T::from(w);
  if (kind == PSPACE_KIND_NORMAL) {
    const double sqrt2 = 1.4142135623730951;      // sqrt(2.0) correctly rounded
    const double sqrtpi = 1.7724538509055159;     // sqrt(4.0*atan(1.0)) = sqrt(0x400921FB54442D18) correctly rounded
    y = s_add(p1, s_mul(s_muld(p2, sqrt2), xs));
    z = s_div(s_sub(y, p1), p2);
    wo = s_divd(ws, sqrtpi);
  } else if (kind == PSPACE_KIND_UNIFORM) {
    S shift = s_divd(s_add(p2, p1), 2.0), scale = s_divd(s_sub(p2, p1), 2.0);
    y = s_add(s_mul(scale, xs), shift);
    z = s_div(s_sub(y, p1), s_sub(p2, p1));
    wo = s_divd(ws, 2.0);
  } else {
    y = s_add(p1, s_mul(p2, xs));
    z = xs;
    wo = ws;
  }
}
