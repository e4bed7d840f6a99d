// ParameterContainer (host side): orchestration + host mirrors; all batched work goes through the C-ABI of
// include/pspace_cuda.h.  Reference: src/cpp/ParameterContainer.cpp:11-237.
//
// Error convention: the reference has none (printf and continue, no status codes, Cython methods are not
// `except +`).  Methods keep their void signatures; a failing device call prints the C-ABI message to stderr and
// abort()s — there is no CPU fallback to continue on.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ParameterContainer.h"
#include "pspace_cuda.h"

namespace {
void die(const char *what, int rc) {
  fprintf(stderr, "pspace: %s failed (%d): %s\n", what, rc, pspace_cuda_last_error());
  abort();
}
#define PS_MUST(call)                 \
  do {                                \
    int rc__ = (call);                \
    if (rc__ != PSPACE_OK) die(#call, rc__); \
  } while (0)
}  // namespace

ParameterContainer::ParameterContainer(int basis_type_, int quadrature_type_)
    : basis_type(basis_type_), quadrature_type(quadrature_type_), tnum_parameters(0), tnum_basis_terms(0),
      tnum_quadrature_points(0), ctx(0), device_(-1), d_ws(0), ws_bytes(0) {}

ParameterContainer::~ParameterContainer() { releaseDevice(); }

extern "C" {
// tiny CUDA-runtime shims exported by libpspace_cuda.so so this translation unit needs no CUDA headers
int pspace_cuda_malloc(void **p, size_t bytes);
int pspace_cuda_free(void *p);
int pspace_cuda_memcpy(void *dst, const void *src, size_t bytes, int to_device);
}

void ParameterContainer::releaseDevice() {
  if (ctx) {
    pspace_cuda_activate(ctx);
    if (d_ws) pspace_cuda_free(d_ws);       // workspace of the device-pointer forms (allocated through the shim)
    pspace_cuda_destroy(ctx);
    ctx = 0;
  }
  d_ws = 0; ws_bytes = 0;
}

void ParameterContainer::addParameter(AbstractParameter *param) {
  // the reference warns beyond 5 parameters (ParameterContainer.cpp:50-56) because its helpers stop there;
  // this implementation takes up to PSPACE_MAX_VARS
  if (tnum_parameters >= PSPACE_MAX_VARS) {
    fprintf(stderr, "pspace: parameter container is full (%d parameters)\n", PSPACE_MAX_VARS);
    abort();
  }
  pmap.insert(std::pair<int, AbstractParameter *>(param->getParameterID(), param));
  tnum_parameters++;
  releaseDevice();   // the device context describes a fixed parameter set
  tnum_basis_terms = 0; tnum_quadrature_points = 0;
}

int ParameterContainer::getNumBasisTerms() { return tnum_basis_terms; }
int ParameterContainer::getNumParameters() { return tnum_parameters; }
int ParameterContainer::getNumQuadraturePoints() { return (int)tnum_quadrature_points; }

void ParameterContainer::ensureContext() {
  // the shims (malloc / memcpy) act on the CURRENT device: select the context's device on every entry, so that a
  // process that switches devices between calls (several containers / torch) stays on the right GPU
  if (ctx) { PS_MUST(pspace_cuda_activate(ctx)); return; }
  const int nv = tnum_parameters;
  std::vector<int> kinds(nv), dmax(nv);
  std::vector<scalar> p1(nv), p2(nv);
  int expect = 0;
  for (std::map<int, AbstractParameter *>::iterator it = pmap.begin(); it != pmap.end(); ++it, ++expect) {
    if (it->first != expect) {
      fprintf(stderr, "pspace: parameter ids must be 0..%d (one ParameterFactory per container); found id %d\n", nv - 1, it->first);
      abort();
    }
    kinds[expect] = it->second->kind();
    dmax[expect] = it->second->getMaxDegree();
    it->second->distribution(&p1[expect], &p2[expect]);
  }
  int device = 0;
  const char *env = getenv("PSPACE_DEVICE");      // one process per GPU: LOCAL_RANK selects the device
  if (!env) env = getenv("LOCAL_RANK");
  if (env) device = atoi(env);
  if (device_ >= 0) device = device_;
  PS_MUST(pspace_cuda_create(&ctx, device, PSPACE_SCALAR_WIDTH == 2, nv, kinds.data(), (const double *)p1.data(),
                             (const double *)p2.data(), dmax.data(), basis_type));
}

void ParameterContainer::initializeBasis(const int *pmax) {
  ensureContext();
  PS_MUST(pspace_cuda_initialize_basis(ctx, pmax, 0));
  tnum_basis_terms = pspace_cuda_num_basis(ctx);
  dindex.resize((size_t)tnum_basis_terms * tnum_parameters);
  PS_MUST(pspace_cuda_get_degrees(ctx, dindex.data()));
}

void ParameterContainer::initializeQuadrature(const int *nqpts) {
  ensureContext();
  PS_MUST(pspace_cuda_initialize_quadrature(ctx, nqpts, 0));
  tnum_quadrature_points = pspace_cuda_num_quad(ctx);
  nqpts_.assign(nqpts, nqpts + tnum_parameters);
  mirrorQuadrature();
}

void ParameterContainer::mirrorQuadrature() {
  const int nv = tnum_parameters;
  zp.assign(nv, std::vector<scalar>());
  yp.assign(nv, std::vector<scalar>());
  for (int d = 0; d < nv; d++) {
    zp[d].resize(nqpts_[d]);
    yp[d].resize(nqpts_[d]);
    PS_MUST(pspace_cuda_get_rules(ctx, d, (double *)zp[d].data(), (double *)yp[d].data(), 0));
  }
  W.resize((size_t)tnum_quadrature_points);
  PS_MUST(pspace_cuda_get_weights(ctx, 0, tnum_quadrature_points, (double *)W.data()));
}

void ParameterContainer::initialize() {
  const int nv = tnum_parameters;
  std::vector<int> pmax(nv), nq(nv);
  for (std::map<int, AbstractParameter *>::iterator it = pmap.begin(); it != pmap.end(); ++it) {
    pmax[it->first] = it->second->getMaxDegree();
    nq[it->first] = it->second->getMaxDegree() + 1;
  }
  initializeBasis(pmax.data());
  initializeQuadrature(nq.data());
}

// weight of point q; zq/yq receive its coordinates.  Grid index: last parameter fastest
// (reference QuadratureHelper.cpp:105-124); values come from the device-built rules and weights.
scalar ParameterContainer::quadrature(int q, scalar *zq, scalar *yq) {
  long long r = q;
  for (int d = tnum_parameters - 1; d >= 0; d--) {
    const int i = (int)(r % nqpts_[d]);
    r /= nqpts_[d];
    zq[d] = zp[d][i];
    yq[d] = yp[d][i];
  }
  return W[(size_t)q];
}
void ParameterContainer::quadrature(int q, scalar *zq, scalar *yq, scalar *wq) { *wq = quadrature(q, zq, yq); }

// psi_k(z) = ((1*b_0)*b_1)*...  in ascending parameter id (reference ParameterContainer.cpp:184-192)
scalar ParameterContainer::basis(int k, scalar *z) {
  scalar psi = 1.0;
  const int nv = tnum_parameters;
  for (std::map<int, AbstractParameter *>::iterator it = pmap.begin(); it != pmap.end(); ++it)
    psi *= it->second->basis(z[it->first], dindex[(size_t)k * nv + it->first]);
  return psi;
}

void ParameterContainer::getBasisParamDeg(int k, int *degs) {
  memcpy(degs, &dindex[(size_t)k * tnum_parameters], sizeof(int) * tnum_parameters);
}
void ParameterContainer::getBasisParamMaxDeg(int *pmax) {
  for (std::map<int, AbstractParameter *>::iterator it = pmap.begin(); it != pmap.end(); ++it)
    pmax[it->first] = it->second->getMaxDegree();
}

pspace_ctx *ParameterContainer::deviceContext() { ensureContext(); return ctx; }
void ParameterContainer::setDevice(int device) {
  if (device != device_) releaseDevice();
  device_ = device;
}

// ---- batched GPU entry points ------------------------------------------------------------------------------
void ParameterContainer::evalBasis(int npts, const scalar *z, scalar *psi) {
  ensureContext();
  const size_t w = PSPACE_SCALAR_WIDTH;
  const size_t zb = sizeof(double) * w * (size_t)npts * tnum_parameters, pb = sizeof(double) * w * (size_t)npts * tnum_basis_terms;
  void *dz = 0, *dp = 0;
  PS_MUST(pspace_cuda_malloc(&dz, zb));
  PS_MUST(pspace_cuda_malloc(&dp, pb));
  PS_MUST(pspace_cuda_memcpy(dz, z, zb, 1));
  PS_MUST(pspace_cuda_eval_basis_points(ctx, 0, tnum_basis_terms, npts, (const double *)dz, (double *)dp, 0));
  PS_MUST(pspace_cuda_memcpy(psi, dp, pb, 0));
  pspace_cuda_free(dz);
  pspace_cuda_free(dp);
}

void ParameterContainer::evaluateExpansion(int m, const scalar *V, scalar *U) {
  ensureContext();
  const size_t w = PSPACE_SCALAR_WIDTH;
  const size_t vb = sizeof(double) * w * (size_t)tnum_basis_terms * m, ub = sizeof(double) * w * (size_t)tnum_quadrature_points * m;
  void *dv = 0, *du = 0;
  PS_MUST(pspace_cuda_malloc(&dv, vb));
  PS_MUST(pspace_cuda_malloc(&du, ub));
  PS_MUST(pspace_cuda_memcpy(dv, V, vb, 1));
  PS_MUST(pspace_cuda_evaluate_expansion(ctx, 0, tnum_basis_terms, 0, tnum_quadrature_points, m, 0, (const double *)dv, (double *)du, 0));
  PS_MUST(pspace_cuda_memcpy(U, du, ub, 0));
  pspace_cuda_free(dv);
  pspace_cuda_free(du);
}

void ParameterContainer::adjointStates(int nnodes, int ndvpn, const scalar *v, const scalar *adj, scalar *UP) {
  ensureContext();
  const size_t w = PSPACE_SCALAR_WIDTH, m = (size_t)nnodes * ndvpn;
  const size_t vb = sizeof(double) * w * (size_t)tnum_basis_terms * m, ub = sizeof(double) * w * (size_t)tnum_quadrature_points * 2 * m;
  void *dv = 0, *da = 0, *du = 0;
  PS_MUST(pspace_cuda_malloc(&dv, vb));
  PS_MUST(pspace_cuda_malloc(&da, vb));
  PS_MUST(pspace_cuda_malloc(&du, ub));
  PS_MUST(pspace_cuda_memcpy(dv, v, vb, 1));
  PS_MUST(pspace_cuda_memcpy(da, adj, vb, 1));
  PS_MUST(pspace_cuda_adjoint_states(ctx, nnodes, ndvpn, 0, tnum_quadrature_points, (const double *)dv, (const double *)da, (double *)du, 0));
  PS_MUST(pspace_cuda_memcpy(UP, du, ub, 0));
  pspace_cuda_free(dv); pspace_cuda_free(da); pspace_cuda_free(du);
}

void ParameterContainer::adjointProduct(int dvLen, scalar scale, const scalar *G, scalar *dfdx) {
  ensureContext();
  const size_t w = PSPACE_SCALAR_WIDTH;
  const size_t gb = sizeof(double) * w * (size_t)tnum_quadrature_points * dvLen, db = sizeof(double) * w * (size_t)dvLen;
  void *dg = 0, *dd = 0;
  PS_MUST(pspace_cuda_malloc(&dg, gb));
  PS_MUST(pspace_cuda_malloc(&dd, db));
  PS_MUST(pspace_cuda_memcpy(dg, G, gb, 1));
  PS_MUST(pspace_cuda_memcpy(dd, dfdx, db, 1));            // the reference adds into dfdx (:606-608)
  PS_MUST(pspace_cuda_adjoint_product(ctx, dvLen, (const double *)&scale, 0, tnum_quadrature_points, (const double *)dg, (double *)dd, 1, 0));
  PS_MUST(pspace_cuda_memcpy(dfdx, dd, db, 0));
  pspace_cuda_free(dg); pspace_cuda_free(dd);
}

void ParameterContainer::projectVector(int m, const scalar *f, scalar *F) {
  ensureContext();
  pspace_project_args a;
  memset(&a, 0, sizeof a);
  a.k0 = 0; a.k1 = tnum_basis_terms; a.q0 = 0; a.q1 = tnum_quadrature_points; a.ncols = m;
  PS_MUST(pspace_cuda_project_vector_host(ctx, &a, (const double *)f, (double *)F));
}

void ParameterContainer::projectMatrixWindow(int m2, const scalar *J, int i0, int i1, int j0, int j1, scalar *C) {
  ensureContext();
  pspace_project_args a;
  memset(&a, 0, sizeof a);
  a.k0 = i0; a.k1 = i1; a.j0 = j0; a.j1 = j1; a.q0 = 0; a.q1 = tnum_quadrature_points; a.ncols = m2;
  PS_MUST(pspace_cuda_project_matrix_host(ctx, &a, (const double *)J, (double *)C));
}

void ParameterContainer::projectMatrix(int m2, const scalar *J, scalar *C) {
  projectMatrixWindow(m2, J, 0, tnum_basis_terms, 0, tnum_basis_terms, C);
}

int ParameterContainer::projectVectorDevice(int m, const scalar *d_f, scalar *d_F, long long q0, long long q1, void *stream) {
  ensureContext();
  pspace_project_args a;
  memset(&a, 0, sizeof a);
  a.k0 = 0; a.k1 = tnum_basis_terms; a.q0 = q0; a.q1 = q1; a.ncols = m;
  const size_t need = pspace_cuda_project_workspace(ctx, 0, &a);
  if (need > ws_bytes) { if (d_ws) pspace_cuda_free(d_ws); PS_MUST(pspace_cuda_malloc(&d_ws, need)); ws_bytes = need; }
  return pspace_cuda_project_vector(ctx, &a, (const double *)d_f, (double *)d_F, d_ws, ws_bytes, stream);
}

int ParameterContainer::projectMatrixDevice(int m2, const scalar *d_J, scalar *d_C, int i0, int i1, int j0, int j1,
                                            long long q0, long long q1, void *stream) {
  ensureContext();
  pspace_project_args a;
  memset(&a, 0, sizeof a);
  a.k0 = i0; a.k1 = i1; a.j0 = j0; a.j1 = j1; a.q0 = q0; a.q1 = q1; a.ncols = m2;
  const size_t need = pspace_cuda_project_workspace(ctx, 1, &a);
  if (need > ws_bytes) { if (d_ws) pspace_cuda_free(d_ws); PS_MUST(pspace_cuda_malloc(&d_ws, need)); ws_bytes = need; }
  return pspace_cuda_project_matrix(ctx, &a, (const double *)d_J, (double *)d_C, d_ws, ws_bytes, stream);
}
