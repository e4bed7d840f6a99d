// ParameterContainer — multivariate basis + tensor quadrature of a set of parameters.
//
// Public interface = reference src/include/ParameterContainer.h:17-41 (what pspace/PSPACE.pxd:30-50 and the
// C++ callers shaped like examples/stacs/cpp/TACSStochasticElement.cpp use); those methods keep their
// names, argument meaning and (absent) error convention.  Behind them everything that is more than a
// table lookup runs on the GPU through the C-ABI of include/pspace_cuda.h:
//   initializeBasis      -> pspace_cuda_initialize_basis      (multi-index kernel)
//   initializeQuadrature -> pspace_cuda_initialize_quadrature (1-D rules, 1-D basis tables, grid weights)
//   quadrature(q,zq,yq)  -> lookup in the host mirror of the device-built rules/weights
//   basis(k,z)           -> product of the parameters' basis(z_d, degree) as in the reference
// and the projection loops the reference leaves to its callers are batched methods (additive):
//   projectVector / projectMatrix / evalBasis (+ device-pointer and q-range forms).
// There is no CPU implementation of the batched methods: without a CUDA device they report and abort().
#ifndef PSPACE_B200_PARAMETER_CONTAINER_H
#define PSPACE_B200_PARAMETER_CONTAINER_H

#include <cstdio>
#include <map>
#include <vector>

#include "AbstractParameter.h"

struct pspace_ctx;   // include/pspace_cuda.h

class ParameterContainer {
 public:
  ParameterContainer(int basis_type = 0, int quadrature_type = 0);
  ~ParameterContainer();

  void addParameter(AbstractParameter *param);

  // per-call accessors (reference semantics)
  scalar quadrature(int q, scalar *zq, scalar *yq);   // returns the weight w_q
  scalar basis(int k, scalar *z);

  int getNumBasisTerms();
  int getNumParameters();
  int getNumQuadraturePoints();
  void getBasisParamDeg(int k, int *degs);
  void getBasisParamMaxDeg(int *pmax);

  void initialize();                            // pmax = dmax, nqpts = dmax + 1 for every parameter
  void initializeBasis(const int *pmax);
  void initializeQuadrature(const int *nqpts);

  // --- additive: batched GPU entry points ------------------------------------------------------------
  // 4-argument form of quadrature named by the project brief (the Fortran port has it:
  // src/fortran/parameter_container.f90:163-175)
  void quadrature(int q, scalar *zq, scalar *yq, scalar *wq);
  // psi[k*npts + p] = basis(k, z + p*nvars), all k; host buffers
  void evalBasis(int npts, const scalar *z, scalar *psi);
  // U[q*m + c] = sum_k psi_k(z_q) V[k*m + c]: the expansion with coefficient rows V evaluated at every quadrature
  // point (what getDeterministicStates of the reference caller recomputes per (i,j,q)); host buffers
  void evaluateExpansion(int m, const scalar *V, scalar *U);
  // F[k*m + c] = sum_q w_q psi_k(z_q) f[q*m + c]                      host buffers, all k, all q
  void projectVector(int m, const scalar *f, scalar *F);
  // C[(i*N + j)*m2 + c] = sum_q w_q psi_i(z_q) psi_j(z_q) J[q*m2 + c]  host buffers, all ordered pairs, all q
  void projectMatrix(int m2, const scalar *J, scalar *C);
  // pair window [i0,i1) x [j0,j1):  C[((i-i0)*(j1-j0) + (j-j0))*m2 + c]
  void projectMatrixWindow(int m2, const scalar *J, int i0, int i1, int j0, int j1, scalar *C);
  // adjoint-residual product (TACSStochasticElement::addAdjResProduct, examples/stacs/cpp/TACSStochasticElement.cpp:548-637)
  // around the caller's deterministic element; m = nnodes*ndvpn; v, adj = TACS-interleaved stochastic vectors [nnodes*N*ndvpn]:
  //   adjointStates : UP[q*2m + c] = u_q[c] (c < m), UP[q*2m + m + c] = psi_q[c]     (getDeterministicStates / -Adjoint, :53-120)
  //   adjointProduct: dfdx[n] += scale * sum_q w_q basis(0,z_q) G[q*dvLen + n]      (G = the element's integrand per point)
  void adjointStates(int nnodes, int ndvpn, const scalar *v, const scalar *adj, scalar *UP);
  void adjointProduct(int dvLen, scalar scale, const scalar *G, scalar *dfdx);
  // device-pointer forms for callers that keep J / C resident (q-range [q0,q1) = this rank's shard; d_J holds
  // rows q0..q1-1).  `stream` is a cudaStream_t.  Returns the C-ABI status (0 = ok).
  int projectVectorDevice(int m, const scalar *d_f, scalar *d_F, long long q0, long long q1, void *stream);
  int projectMatrixDevice(int m2, const scalar *d_J, scalar *d_C, int i0, int i1, int j0, int j1, long long q0,
                          long long q1, void *stream);
  pspace_ctx *deviceContext();                  // for direct use of include/pspace_cuda.h
  // GPU of this container (default: $PSPACE_DEVICE, else $LOCAL_RANK, else 0); call before initialize*().  One process
  // per GPU needs nothing; a process that drives several GPUs from threads gives each thread's container its device.
  void setDevice(int device);

 private:
  void ensureContext();
  void mirrorQuadrature();
  void releaseDevice();

  std::map<int, AbstractParameter *> pmap;      // id -> parameter, ascending id = product order
  int basis_type, quadrature_type;
  int tnum_parameters, tnum_basis_terms;
  long long tnum_quadrature_points;
  pspace_ctx *ctx;
  std::vector<int> dindex;                      // [N][nvars] mirror of the device degree table
  std::vector<int> nqpts_;                      // 1-D rule sizes
  std::vector<std::vector<scalar> > zp, yp;     // mirrors of the device-built 1-D rules
  std::vector<scalar> W;                        // mirror of the device-built grid weights
  int device_;                                  // -1 = from the environment
  void *d_ws;                                   // device scratch for the q-split of the device-pointer forms
  size_t ws_bytes;
};

#endif
