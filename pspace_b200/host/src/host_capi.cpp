// extern "C" handles over the C++ host classes, so that Python (ctypes, pspace_b200/PSPACE.py) can drive
// ParameterFactory / ParameterContainer / AbstractParameter exactly as the reference's Cython wrapper does
// (pspace/PSPACE.pyx:35-131) without a compiled extension module.  Scalars cross as double[PSPACE_SCALAR_WIDTH].
#include "ParameterContainer.h"
#include "ParameterFactory.h"

namespace {
inline scalar get(const double *p) {
#ifdef USE_COMPLEX
  return scalar(p[0], p[1]);
#else
  return p[0];
#endif
}
inline void put(double *p, const scalar &s) {
#ifdef USE_COMPLEX
  p[0] = s.real(); p[1] = s.imag();
#else
  p[0] = s;
#endif
}
}  // namespace

extern "C" {
int pspace_host_is_complex() { return PSPACE_SCALAR_WIDTH == 2; }

void *pspace_host_factory_new() { return new ParameterFactory(); }
void pspace_host_factory_delete(void *f) { delete (ParameterFactory *)f; }
void *pspace_host_create_normal(void *f, const double *mu, const double *sigma, int dmax) {
  return ((ParameterFactory *)f)->createNormalParameter(get(mu), get(sigma), dmax);
}
void *pspace_host_create_uniform(void *f, const double *a, const double *b, int dmax) {
  return ((ParameterFactory *)f)->createUniformParameter(get(a), get(b), dmax);
}
void *pspace_host_create_exponential(void *f, const double *mu, const double *beta, int dmax) {
  return ((ParameterFactory *)f)->createExponentialParameter(get(mu), get(beta), dmax);
}
void pspace_host_param_basis(void *p, const double *z, int d, double *out) { put(out, ((AbstractParameter *)p)->basis(get(z), d)); }
void pspace_host_param_quadrature(void *p, int n, double *z, double *y, double *w) {
  ((AbstractParameter *)p)->quadrature(n, reinterpret_cast<scalar *>(z), reinterpret_cast<scalar *>(y), reinterpret_cast<scalar *>(w));
}
int pspace_host_param_id(void *p) { return ((AbstractParameter *)p)->getParameterID(); }
int pspace_host_param_max_degree(void *p) { return ((AbstractParameter *)p)->getMaxDegree(); }

void *pspace_host_container_new(int basis_type, int quadrature_type) { return new ParameterContainer(basis_type, quadrature_type); }
void pspace_host_container_delete(void *c) { delete (ParameterContainer *)c; }
void pspace_host_add_parameter(void *c, void *p) { ((ParameterContainer *)c)->addParameter((AbstractParameter *)p); }
void pspace_host_initialize(void *c) { ((ParameterContainer *)c)->initialize(); }
void pspace_host_initialize_basis(void *c, const int *pmax) { ((ParameterContainer *)c)->initializeBasis(pmax); }
void pspace_host_initialize_quadrature(void *c, const int *nq) { ((ParameterContainer *)c)->initializeQuadrature(nq); }
int pspace_host_num_basis(void *c) { return ((ParameterContainer *)c)->getNumBasisTerms(); }
int pspace_host_num_params(void *c) { return ((ParameterContainer *)c)->getNumParameters(); }
int pspace_host_num_quad(void *c) { return ((ParameterContainer *)c)->getNumQuadraturePoints(); }
void pspace_host_quadrature(void *c, int q, double *zq, double *yq, double *wq) {
  put(wq, ((ParameterContainer *)c)->quadrature(q, reinterpret_cast<scalar *>(zq), reinterpret_cast<scalar *>(yq)));
}
void pspace_host_basis(void *c, int k, double *z, double *out) { put(out, ((ParameterContainer *)c)->basis(k, reinterpret_cast<scalar *>(z))); }
void pspace_host_basis_param_deg(void *c, int k, int *degs) { ((ParameterContainer *)c)->getBasisParamDeg(k, degs); }
void pspace_host_basis_param_max_deg(void *c, int *pmax) { ((ParameterContainer *)c)->getBasisParamMaxDeg(pmax); }
void pspace_host_eval_basis(void *c, int npts, const double *z, double *psi) {
  ((ParameterContainer *)c)->evalBasis(npts, reinterpret_cast<const scalar *>(z), reinterpret_cast<scalar *>(psi));
}
void pspace_host_evaluate_expansion(void *c, int m, const double *V, double *U) {
  ((ParameterContainer *)c)->evaluateExpansion(m, reinterpret_cast<const scalar *>(V), reinterpret_cast<scalar *>(U));
}
void pspace_host_project_vector(void *c, int m, const double *f, double *F) {
  ((ParameterContainer *)c)->projectVector(m, reinterpret_cast<const scalar *>(f), reinterpret_cast<scalar *>(F));
}
void pspace_host_project_matrix(void *c, int m2, const double *J, double *C) {
  ((ParameterContainer *)c)->projectMatrix(m2, reinterpret_cast<const scalar *>(J), reinterpret_cast<scalar *>(C));
}
void pspace_host_project_matrix_window(void *c, int m2, const double *J, int i0, int i1, int j0, int j1, double *C) {
  ((ParameterContainer *)c)->projectMatrixWindow(m2, reinterpret_cast<const scalar *>(J), i0, i1, j0, j1, reinterpret_cast<scalar *>(C));
}
void pspace_host_adjoint_states(void *c, int nnodes, int ndvpn, const double *v, const double *adj, double *UP) {
  ((ParameterContainer *)c)->adjointStates(nnodes, ndvpn, reinterpret_cast<const scalar *>(v), reinterpret_cast<const scalar *>(adj), reinterpret_cast<scalar *>(UP));
}
void pspace_host_adjoint_product(void *c, int dvlen, const double *scale, const double *G, double *dfdx) {
  ((ParameterContainer *)c)->adjointProduct(dvlen, *reinterpret_cast<const scalar *>(scale), reinterpret_cast<const scalar *>(G), reinterpret_cast<scalar *>(dfdx));
}
void *pspace_host_device_context(void *c) { return ((ParameterContainer *)c)->deviceContext(); }
}
