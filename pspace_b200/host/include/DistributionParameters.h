// The three ParameterFactory products.  The reference declares them in NormalParameter.h /
// UniformParameter.h / ExponentialParameter.h (src/include/*.h:10-22, private inheritance); those names are
// kept as forwarding headers.  Their 1-D arithmetic is NOT re-implemented here: both methods instantiate the
// single host+device source pspace_b200/csrc/pspace_poly.cuh that the CUDA table kernels use.
#ifndef PSPACE_B200_DISTRIBUTION_PARAMETERS_H
#define PSPACE_B200_DISTRIBUTION_PARAMETERS_H

#include "AbstractParameter.h"

class NormalParameter : AbstractParameter {
 public:
  NormalParameter(int pid, scalar mu, scalar sigma);
  ~NormalParameter();
  void quadrature(int npoints, scalar *z, scalar *y, scalar *w);   // Gauss-Hermite mapped to N(mu, sigma^2)
  scalar basis(scalar z, int d);                                  // He_d(z)/sqrt(d!)
  int kind() const;
  void distribution(scalar *first, scalar *second) const;
 private:
  scalar mu, sigma;
};

class UniformParameter : AbstractParameter {
 public:
  UniformParameter(int pid, scalar a, scalar b);
  ~UniformParameter();
  void quadrature(int npoints, scalar *z, scalar *y, scalar *w);   // Gauss-Legendre mapped to U(a, b)
  scalar basis(scalar z, int d);                                  // shifted Legendre on [0,1] * sqrt(2d+1)
  int kind() const;
  void distribution(scalar *first, scalar *second) const;
 private:
  scalar a, b;
};

class ExponentialParameter : AbstractParameter {
 public:
  ExponentialParameter(int pid, scalar mu, scalar beta);
  ~ExponentialParameter();
  void quadrature(int npoints, scalar *z, scalar *y, scalar *w);   // Gauss-Laguerre mapped to mu + beta*Exp(1)
  scalar basis(scalar z, int d);                                  // Laguerre L_d(z)
  int kind() const;
  void distribution(scalar *first, scalar *second) const;
 private:
  scalar mu, beta;
};

#endif
