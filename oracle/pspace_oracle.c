/* TEST INFRASTRUCTURE — see pspace_oracle.h.  CPU restatement of the reference's hot path:
 *   src/cpp/OrthogonalPolynomials.cpp, GaussianQuadrature.cpp (maps), BasisHelper.cpp,
 *   QuadratureHelper.cpp, ParameterContainer.cpp, and the projection loops of
 *   examples/stacs/cpp/TACSStochasticElement.cpp:283-311, 380-414.
 * Built with -ffp-contract=off so no FMA contraction changes a rounding.
 */
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pspace_oracle.h"

#define PSPACE_GAUSS_TABLE(name) static const double name[]
#include "gauss_tables.inc"

int oracle_is_complex(void) { return 0; }

static inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
void oracle_fill_synthetic(double *out, long long n, unsigned long long seed, unsigned long long offset) {
  for (long long i = 0; i < n; i++) {
    unsigned long long x = splitmix64(seed ^ (offset + (unsigned long long)i));
    out[i] = (double)(x >> 11) * 0x1.0p-53 - 0.5;
  }
}

/* ---- real build ---- */
#define S double
#define SFX
#define MKS(x) (x)
#define LOADS(p, i) ((p)[i])
#define SQRT(x) sqrt(x)
#define POWI(x, k) pow((x), (double)(k)) /* std::pow(double,int) promotes to pow(double,double) */
#include "pspace_oracle_impl.inc"
#undef S
#undef SFX
#undef MKS
#undef LOADS
#undef SQRT
#undef POWI

/* ---- USE_COMPLEX build ---- */
/* libstdc++ std::pow(complex<double>, int) (bits of <complex>: __complex_pow_unsigned): binary powering
 * y = n odd ? x : 1 ; while (n >>= 1) { x *= x; if (n odd) y *= x; } */
static double _Complex cpowi(double _Complex x, int n) {
  unsigned un = (unsigned)n;
  double _Complex y = (un % 2) ? x : (double _Complex)1.0;
  while (un >>= 1) {
    x = x * x;
    if (un % 2) y = y * x;
  }
  return y;
}
#define S double _Complex
#define SFX _z
#define MKS(x) ((double _Complex)(x))
#define LOADS(p, i) (CMPLX((p)[2 * (i)], (p)[2 * (i) + 1]))
#define SQRT(x) csqrt(x)
#define POWI(x, k) cpowi((x), (k))
#include "pspace_oracle_impl.inc"
