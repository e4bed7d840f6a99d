/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the pspace stochastic Galerkin hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / parity / --impl reference legs (plus the
 * developer scripts under tools/ that generate fixtures or time the CPU side) may load this library, and only as the
 * checker / CPU baseline.  The product path (libpspace_cuda.so,
 * libpspace.so, pspace_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Every function is checked bit for bit against the compiled, unmodified
 * reference (oracle/_ref, built from /root/reference by oracle/Makefile) by tests/test_oracle_vs_ref.py
 * (run in the build container, where the reference exists) and against the golden vectors under
 * tests/golden/ that were generated from the compiled reference (tools/gen_golden.py).
 * For nvars > 5 the reference has no implementation (BasisHelper.cpp:63-65, QuadratureHelper.cpp:34-132
 * end at 5 variables); there the oracle extends the same enumeration / grid / product rules
 * ("extended oracle") and is pinned only through nvars <= 5.
 *
 * Scalars cross this ABI as double (real entry points) or interleaved (re,im) pairs (entry points
 * with the _z suffix = the reference's USE_COMPLEX build, src/include/scalar.h:18-24).
 * The API mirrors oracle/ref_shim.cpp one to one so a single Python harness drives both.
 */
#ifndef PSPACE_ORACLE_H
#define PSPACE_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_VARS 32

/* the caller's deterministic element for (f4): adds scale * g(psi_q, u_q, y_q) into dfdx[dvlen]; every scalar array is
 * double (real entry points) or interleaved (re,im) (_z entry points), scale points at one scalar */
typedef void (*oracle_adjres_element)(void *user, const double *scale, const double *psi_q, const double *u_q,
                                      const double *y_q, int m, int nvars, int dvlen, double *dfdx);

/* kinds: 0 normal(mu,sigma)  1 uniform(a,b)  2 exponential(mu,beta) */
#define ORACLE_DECL(SFX)                                                                              \
  void *oracle_create##SFX(int nvars, const int *kinds, const double *p1, const double *p2,           \
                           const int *dmax, int basis_type);                                          \
  void oracle_destroy##SFX(void *h);                                                                  \
  void oracle_initialize##SFX(void *h);                                                               \
  void oracle_initialize_basis##SFX(void *h, const int *pmax);                                        \
  void oracle_initialize_quadrature##SFX(void *h, const int *nqpts);                                  \
  int oracle_num_basis##SFX(void *h);                                                                 \
  int oracle_num_quad##SFX(void *h);                                                                  \
  int oracle_num_params##SFX(void *h);                                                                \
  void oracle_degrees##SFX(void *h, int *out /*[N][nvars]*/);                                         \
  void oracle_max_degrees##SFX(void *h, int *out /*[nvars]*/);                                        \
  void oracle_grid##SFX(void *h, double *Z /*[Q][nvars]*/, double *Y /*[Q][nvars]*/, double *W);      \
  void oracle_basis_at##SFX(void *h, int npts, const double *Z /*[npts][nvars]*/, double *out);       \
  void oracle_param_quadrature##SFX(void *h, int pid, int npoints, double *z, double *y, double *w);  \
  void oracle_param_basis##SFX(void *h, int pid, int npts, const double *z, int d, double *out);      \
  void oracle_project_vector##SFX(void *h, int m, const double *f, int k0, int k1, double *F);        \
  void oracle_project_matrix##SFX(void *h, int m2, const double *J, int i0, int i1, int j0, int j1,   \
                                  double *C, int nthreads);                                           \
  void oracle_moments##SFX(void *h, int m, const double *f, long long q0, long long q1, double *mean,   \
                           double *second, double *var);                                               \
  void oracle_evaluate_expansion##SFX(void *h, int m, const double *V, int k0, int k1, long long q0,    \
                                      long long q1, double *U);                                        \
  void oracle_adj_res_product##SFX(void *h, int nnodes, int ndvpn, int dvlen, const double *scale,      \
                                   const double *v, const double *adj, oracle_adjres_element elem,      \
                                   void *user, double *dfdx);                                           \
  void oracle_synthetic_adjres_element##SFX(void *user, const double *scale, const double *psi_q,       \
                                            const double *u_q, const double *y_q, int m, int nvars,      \
                                            int dvlen, double *dfdx);                                    \
  /* q-range restricted variants used by the multi-GPU (sharded) parity tests */                      \
  void oracle_project_matrix_qrange##SFX(void *h, int m2, const double *J, int i0, int i1, int j0,    \
                                         int j1, long long q0, long long q1, double *C, int nthreads);

ORACLE_DECL()
ORACLE_DECL(_z)

void oracle_scatter_vector_tacs(int N, int nnodes, int ndvpn, int width, const double *F, double *res);
void oracle_scatter_matrix_tacs(int N, int nnodes, int ndvpn, int width, const double *C, double *mat);
void oracle_pair_mask(int N, int nvars, const int *deg, const int *dmapf, unsigned char *mask);

int oracle_is_complex(void);   /* 0; the _z entry points are the complex build */

/* counter-based synthetic block generator shared (by restating, not linking) with the product bench:
 * x = splitmix64(seed ^ idx) ; value = (x >> 11) * 2^-53 - 0.5           */
void oracle_fill_synthetic(double *out, long long n, unsigned long long seed, unsigned long long offset);

#ifdef __cplusplus
}
#endif
#endif
