// AbstractParameter, the three distribution parameters and ParameterFactory (host side).
//
// The 1-D arithmetic is the single host+device source pspace_b200/csrc/pspace_poly.cuh — the same templates
// the CUDA table kernels instantiate — so a per-call `param->basis(z, d)` / `param->quadrature(n, z, y, w)` and the
// device-built tables cannot drift apart.  (Reference: src/cpp/{Normal,Uniform,Exponential}Parameter.cpp:30-44,
// AbstractParameter.cpp, ParameterFactory.cpp:22-109.)
#include <cstdio>
#include <cstdlib>

#include "ParameterFactory.h"
#include "pspace_cuda.h"
#include "pspace_poly.cuh"

#define PSPACE_GAUSS_TABLE(name) static const double name[]
#include "gauss_tables.inc"

namespace {
#ifdef USE_COMPLEX
typedef cplx dev_scalar;
inline dev_scalar to_dev(const scalar &s) { return mkc(s.real(), s.imag()); }
inline scalar from_dev(const dev_scalar &s) { return scalar(s.re, s.im); }
#else
typedef double dev_scalar;
inline dev_scalar to_dev(const scalar &s) { return s; }
inline scalar from_dev(const dev_scalar &s) { return s; }
#endif

void rule(int kind, scalar p1, scalar p2, int npoints, scalar *z, scalar *y, scalar *w) {
  if (npoints < 1 || npoints > PSPACE_GAUSS_NMAX) {
    fprintf(stderr, "pspace: 1-D Gauss rules exist for 1..%d points, %d requested\n", PSPACE_GAUSS_NMAX, npoints);
    abort();
  }
  const double *X = kind == PSPACE_NORMAL ? PSPACE_HERMITE_X : kind == PSPACE_UNIFORM ? PSPACE_LEGENDRE_X : PSPACE_LAGUERRE_X;
  const double *Wt = kind == PSPACE_NORMAL ? PSPACE_HERMITE_W : kind == PSPACE_UNIFORM ? PSPACE_LEGENDRE_W : PSPACE_LAGUERRE_W;
  const int o = npoints * (npoints - 1) / 2;
  for (int i = 0; i < npoints; i++) {
    dev_scalar zz, yy, ww;
    ps_map_rule<dev_scalar>(kind, to_dev(p1), to_dev(p2), X[o + i], Wt[o + i], zz, yy, ww);
    z[i] = from_dev(zz); y[i] = from_dev(yy); w[i] = from_dev(ww);
  }
}
}  // namespace

AbstractParameter::AbstractParameter() : parameter_id(-1), dmax(0) {}
AbstractParameter::~AbstractParameter() {}
int AbstractParameter::getParameterID() { return parameter_id; }
int AbstractParameter::getMaxDegree() { return dmax; }
void AbstractParameter::setParameterID(int pid) { parameter_id = pid; }
void AbstractParameter::setMaxDegree(int d) { dmax = d; }

#define PSPACE_DEFINE_PARAMETER(Class, KIND, A, B)                                                        \
  Class::Class(int pid, scalar A##_, scalar B##_) : AbstractParameter(), A(A##_), B(B##_) {               \
    setParameterID(pid);                                                                                  \
  }                                                                                                       \
  Class::~Class() {}                                                                                      \
  void Class::quadrature(int npoints, scalar *z, scalar *y, scalar *w) { rule(KIND, A, B, npoints, z, y, w); } \
  scalar Class::basis(scalar z, int d) { return from_dev(ps_param_basis<dev_scalar>(KIND, to_dev(z), d)); } \
  int Class::kind() const { return KIND; }                                                                \
  void Class::distribution(scalar *first, scalar *second) const { *first = A; *second = B; }

PSPACE_DEFINE_PARAMETER(NormalParameter, PSPACE_NORMAL, mu, sigma)
PSPACE_DEFINE_PARAMETER(UniformParameter, PSPACE_UNIFORM, a, b)
PSPACE_DEFINE_PARAMETER(ExponentialParameter, PSPACE_EXPONENTIAL, mu, beta)

ParameterFactory::ParameterFactory() : next_parameter_id(0) {}
ParameterFactory::~ParameterFactory() {}

AbstractParameter *ParameterFactory::createNormalParameter(scalar mu, scalar sigma) {
  return (AbstractParameter *)new NormalParameter(next_parameter_id++, mu, sigma);
}
AbstractParameter *ParameterFactory::createUniformParameter(scalar a, scalar b) {
  return (AbstractParameter *)new UniformParameter(next_parameter_id++, a, b);
}
AbstractParameter *ParameterFactory::createExponentialParameter(scalar mu, scalar beta) {
  return (AbstractParameter *)new ExponentialParameter(next_parameter_id++, mu, beta);
}
AbstractParameter *ParameterFactory::createNormalParameter(scalar mu, scalar sigma, int dmax) {
  AbstractParameter *p = createNormalParameter(mu, sigma);
  p->setMaxDegree(dmax);
  return p;
}
AbstractParameter *ParameterFactory::createUniformParameter(scalar a, scalar b, int dmax) {
  AbstractParameter *p = createUniformParameter(a, b);
  p->setMaxDegree(dmax);
  return p;
}
AbstractParameter *ParameterFactory::createExponentialParameter(scalar mu, scalar beta, int dmax) {
  AbstractParameter *p = createExponentialParameter(mu, beta);
  p->setMaxDegree(dmax);
  return p;
}
