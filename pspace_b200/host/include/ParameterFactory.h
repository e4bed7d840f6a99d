// ParameterFactory: source of AbstractParameter objects for one ParameterContainer.
//
// Every create* call stamps the new parameter with the next id (0, 1, 2, ...); a container indexes its per-parameter
// arrays by that id, so use a fresh factory for each container.  The six create* signatures are the ones the
// reference declares (src/include/ParameterFactory.h:21-28) and its Cython wrapper binds (pspace/PSPACE.pxd:20-25);
// objects are heap-allocated and owned by the caller, as there.
#ifndef PSPACE_B200_PARAMETER_FACTORY_H
#define PSPACE_B200_PARAMETER_FACTORY_H

#include "DistributionParameters.h"

class ParameterFactory {
 public:
  ParameterFactory();
  ~ParameterFactory();

  // Gaussian y ~ N(mu, sigma^2): Gauss-Hermite rule, orthonormal Hermite basis.
  AbstractParameter *createNormalParameter(scalar mu, scalar sigma);
  AbstractParameter *createNormalParameter(scalar mu, scalar sigma, int dmax);

  // Uniform y ~ U(a, b): Gauss-Legendre rule, orthonormal shifted Legendre basis.
  AbstractParameter *createUniformParameter(scalar a, scalar b);
  AbstractParameter *createUniformParameter(scalar a, scalar b, int dmax);

  // Shifted exponential y = mu + beta * Exp(1): Gauss-Laguerre rule, Laguerre basis.
  AbstractParameter *createExponentialParameter(scalar mu, scalar beta);
  AbstractParameter *createExponentialParameter(scalar mu, scalar beta, int dmax);
  // (the 3-argument forms also record dmax, the degree ParameterContainer::initialize() expands the parameter to)

 private:
  int next_parameter_id;
};

#endif
