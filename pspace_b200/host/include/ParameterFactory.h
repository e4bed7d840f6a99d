// ParameterFactory — hands out parameters with consecutive ids 0,1,2,...  Interface = reference
// src/include/ParameterFactory.h:14-32 (bound by pspace/PSPACE.pxd:20-25).  Use ONE factory per container:
// ids keep counting across containers and a container indexes its arrays by id (reference
// ParameterContainer.cpp:134-137).
#ifndef PSPACE_B200_PARAMETER_FACTORY_H
#define PSPACE_B200_PARAMETER_FACTORY_H

#include "AbstractParameter.h"
#include "DistributionParameters.h"
#include "scalar.h"

class ParameterFactory {
 public:
  ParameterFactory();
  ~ParameterFactory();

  // without a maximum degree (the caller is expected to setMaxDegree or use initializeBasis/Quadrature)
  AbstractParameter *createNormalParameter(scalar mu, scalar sigma);
  AbstractParameter *createUniformParameter(scalar a, scalar b);
  AbstractParameter *createExponentialParameter(scalar mu, scalar beta);

  // with the maximum polynomial degree used by ParameterContainer::initialize()
  AbstractParameter *createNormalParameter(scalar mu, scalar sigma, int dmax);
  AbstractParameter *createUniformParameter(scalar a, scalar b, int dmax);
  AbstractParameter *createExponentialParameter(scalar mu, scalar beta, int dmax);

 private:
  int next_parameter_id;
};

#endif
