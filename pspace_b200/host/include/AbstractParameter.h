// AbstractParameter — a probabilistically modelled parameter: it binds a 1-D Gauss rule to the matching
// orthonormal polynomial family.  Public interface = reference src/include/AbstractParameter.h:16-45
// (virtual slot order quadrature, basis; non-virtual destructor; id / max-degree accessors), which is what
// pspace/PSPACE.pxd:12-15 binds.  Additive: kind() / distribution() let a ParameterContainer describe the
// parameter to the device (include/pspace_cuda.h: pspace_cuda_create).
#ifndef PSPACE_B200_ABSTRACT_PARAMETER_H
#define PSPACE_B200_ABSTRACT_PARAMETER_H

#include <cstdio>
#include <list>
#include <map>
#include "scalar.h"

class AbstractParameter {
 public:
  AbstractParameter();
  ~AbstractParameter();

  // 1-D rule with npoints nodes: z standard-space nodes, y physical-space nodes, w weights (sum to 1)
  virtual void quadrature(int npoints, scalar *z, scalar *y, scalar *w) = 0;
  // orthonormal polynomial of degree d of this parameter's family at standard-space point z
  virtual scalar basis(scalar z, int d) = 0;

  int getParameterID();
  int getMaxDegree();
  void setParameterID(int pid);
  void setMaxDegree(int dmax);

  // --- additive (not in the reference) -------------------------------------------------------------
  // PSPACE_NORMAL / PSPACE_UNIFORM / PSPACE_EXPONENTIAL of include/pspace_cuda.h
  virtual int kind() const = 0;
  // the two distribution parameters in factory order: (mu,sigma) / (a,b) / (mu,beta)
  virtual void distribution(scalar *first, scalar *second) const = 0;

 private:
  int parameter_id;
  int dmax;
};

#endif
