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

/* ---- TACS-interleaved layouts (index loops of the reference caller; width = doubles per scalar) ---------------- */
/* TACSStochasticElement.cpp:316-322: for n, for d: res[n*nsvpn + i*ndvpn + d] += rtmpi[n*ndvpn + d]   (nsvpn = N*ndvpn) */
void oracle_scatter_vector_tacs(int N, int nnodes, int ndvpn, int width, const double *F, double *res) {
  const long long nsvpn = (long long)N * ndvpn, m = (long long)nnodes * ndvpn;
  for (int i = 0; i < N; i++)
    for (int n = 0; n < nnodes; n++) {
      const long long lptr = (long long)n * ndvpn, gptr = n * nsvpn + (long long)i * ndvpn;
      for (int d = 0; d < ndvpn; d++)
        for (int w = 0; w < width; w++) res[width * (gptr + d) + w] += F[width * (i * m + lptr + d) + w];
    }
}
/* TACSStochasticElement.cpp:417-432: mat[(giptr+di)*nsdof + gjptr+dj] += A[(liptr+di)*nddof + ljptr+dj] */
void oracle_scatter_matrix_tacs(int N, int nnodes, int ndvpn, int width, const double *C, double *mat) {
  const long long nsvpn = (long long)N * ndvpn, nddof = (long long)nnodes * ndvpn, nsdof = nddof * N;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      const double *A = C + width * (((long long)i * N + j) * nddof * nddof);
      for (int ni = 0; ni < nnodes; ni++) {
        const long long liptr = (long long)ni * ndvpn, giptr = ni * nsvpn + (long long)i * ndvpn;
        for (int di = 0; di < ndvpn; di++)
          for (int nj = 0; nj < nnodes; nj++) {
            const long long ljptr = (long long)nj * ndvpn, gjptr = nj * nsvpn + (long long)j * ndvpn;
            for (int dj = 0; dj < ndvpn; dj++)
              for (int w = 0; w < width; w++)
                mat[width * ((giptr + di) * nsdof + gjptr + dj) + w] += A[width * ((liptr + di) * nddof + ljptr + dj) + w];
          }
      }
    }
}
/* nonzero(): TACSStochasticElement.cpp:7-21 (BasisHelper::sparse, BasisHelper.cpp:76-88) on a degree table [N][nvars] */
void oracle_pair_mask(int N, int nvars, const int *deg, const int *dmapf, unsigned char *mask) {
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      int prod = 1;
      for (int d = 0; d < nvars; d++) {
        int diff = deg[i * nvars + d] - deg[j * nvars + d];
        if (diff < 0) diff = -diff;
        prod *= (diff <= dmapf[d]) ? 1 : 0;
      }
      mask[(long long)i * N + j] = (unsigned char)prod;
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
