// TEST INFRASTRUCTURE — not part of the product.
//
// C-ABI shim over the UNMODIFIED reference classes (ParameterFactory /
// ParameterContainer, /root/reference/src/include/*.h).  It is compiled by
// oracle/Makefile together with the reference's own .cpp files, where they
// lie under /root/reference, into oracle/_ref/libpspace_ref{,_z}.so.
// It exists so that Python (ctypes) can drive the real reference in this
// container to (a) pin the restated oracle (oracle/pspace_oracle.c) and
// (b) generate the golden vectors under tests/golden/, and (c) serve as the
// "reference" CPU baseline of bench.py.
//
// The projection loops below restate, over the reference's own per-call
// primitives, the loops of the reference *caller*
//   examples/stacs/cpp/TACSStochasticElement.cpp:283-311 (vector: scale = basis(i,zq)*wq)
//   examples/stacs/cpp/TACSStochasticElement.cpp:380-414 (matrix: scale = basis(i,zq)*basis(j,zq)*wq)
// with the (un-vendored) TACS element call replaced by the synthetic
// per-point block  A[c] += scale*J_q[c].
//
// scalars cross the ABI as double (real build) or interleaved (re,im) pairs
// (USE_COMPLEX build), i.e. a `scalar*` is passed as `double*`.
#include <thread>
#include <vector>
#include <cstring>
#include "scalar.h"
#include "ParameterFactory.h"
#include "ParameterContainer.h"
#include "BasisHelper.h"

struct RefHandle {
  ParameterFactory *factory;      // one factory per container (ids restart at 0)
  ParameterContainer *pc;
  std::vector<AbstractParameter*> params;
  int nvars;
};

static inline scalar mk(const double *p, int i){
#ifdef USE_COMPLEX
  return scalar(p[2*i], p[2*i+1]);
#else
  return p[i];
#endif
}

extern "C" {

int refshim_is_complex(){
#ifdef USE_COMPLEX
  return 1;
#else
  return 0;
#endif
}

// kinds: 0 normal(mu,sigma) 1 uniform(a,b) 2 exponential(mu,beta)
void *refshim_create(int nvars, const int *kinds, const double *p1, const double *p2,
                     const int *dmax, int basis_type){
  RefHandle *h = new RefHandle();
  h->factory = new ParameterFactory();
  h->pc = new ParameterContainer(basis_type, 0);
  h->nvars = nvars;
  for (int d = 0; d < nvars; d++){
    AbstractParameter *p = 0;
    if (kinds[d] == 0)      p = h->factory->createNormalParameter(mk(p1,d), mk(p2,d), dmax[d]);
    else if (kinds[d] == 1) p = h->factory->createUniformParameter(mk(p1,d), mk(p2,d), dmax[d]);
    else                    p = h->factory->createExponentialParameter(mk(p1,d), mk(p2,d), dmax[d]);
    h->params.push_back(p);
    h->pc->addParameter(p);
  }
  return h;
}
void refshim_initialize(void *hh){ ((RefHandle*)hh)->pc->initialize(); }
void refshim_initialize_basis(void *hh, const int *pmax){ ((RefHandle*)hh)->pc->initializeBasis(pmax); }
void refshim_initialize_quadrature(void *hh, const int *nq){ ((RefHandle*)hh)->pc->initializeQuadrature(nq); }
int refshim_num_basis(void *hh){ return ((RefHandle*)hh)->pc->getNumBasisTerms(); }
int refshim_num_quad(void *hh){ return ((RefHandle*)hh)->pc->getNumQuadraturePoints(); }
int refshim_num_params(void *hh){ return ((RefHandle*)hh)->pc->getNumParameters(); }

// degrees out[N][nvars]
void refshim_degrees(void *hh, int *out){
  RefHandle *h = (RefHandle*)hh;
  int N = h->pc->getNumBasisTerms();
  for (int k = 0; k < N; k++) h->pc->getBasisParamDeg(k, out + (size_t)k*h->nvars);
}
void refshim_max_degrees(void *hh, int *out){ ((RefHandle*)hh)->pc->getBasisParamMaxDeg(out); }

// Z[Q][nvars], Y[Q][nvars], W[Q] (point-major, as quadrature(q,zq,yq) hands them out)
void refshim_grid(void *hh, double *Z, double *Y, double *W){
  RefHandle *h = (RefHandle*)hh;
  int Q = h->pc->getNumQuadraturePoints(), nv = h->nvars;
  scalar *z = (scalar*)Z, *y = (scalar*)Y, *w = (scalar*)W;
  for (int q = 0; q < Q; q++) w[q] = h->pc->quadrature(q, z + (size_t)q*nv, y + (size_t)q*nv);
}

// psi[k][p] = basis(k, z_p), z[npts][nvars]
void refshim_basis_at(void *hh, int npts, const double *Z, double *out){
  RefHandle *h = (RefHandle*)hh;
  int N = h->pc->getNumBasisTerms(), nv = h->nvars;
  scalar *z = (scalar*)Z, *o = (scalar*)out;
  for (int k = 0; k < N; k++)
    for (int p = 0; p < npts; p++) o[(size_t)k*npts + p] = h->pc->basis(k, z + (size_t)p*nv);
}

// 1-D rule / polynomial of one parameter of the container
void refshim_param_quadrature(void *hh, int pid, int npoints, double *z, double *y, double *w){
  ((RefHandle*)hh)->params[pid]->quadrature(npoints, (scalar*)z, (scalar*)y, (scalar*)w);
}
void refshim_param_basis(void *hh, int pid, int npts, const double *z, int d, double *out){
  RefHandle *h = (RefHandle*)hh;
  for (int p = 0; p < npts; p++) ((scalar*)out)[p] = h->params[pid]->basis(((const scalar*)z)[p], d);
}

// F[k][c] = sum_q (basis(k,zq)*wq) * f[q][c]      k in [k0,k1)   (TACSStochasticElement.cpp:283-311)
void refshim_project_vector(void *hh, int m, const double *f_, int k0, int k1, double *F_){
  RefHandle *h = (RefHandle*)hh;
  const scalar *f = (const scalar*)f_; scalar *F = (scalar*)F_;
  int Q = h->pc->getNumQuadraturePoints(), nv = h->nvars;
  std::vector<scalar> zq(nv), yq(nv);
  for (int i = k0; i < k1; i++){
    scalar *r = F + (size_t)(i-k0)*m;
    for (int c = 0; c < m; c++) r[c] = 0.0;
    for (int q = 0; q < Q; q++){
      scalar wq = h->pc->quadrature(q, zq.data(), yq.data());
      scalar scale = h->pc->basis(i, zq.data())*wq;
      const scalar *fq = f + (size_t)q*m;
      for (int c = 0; c < m; c++) r[c] += fq[c]*scale;
    }
  }
}

// U[q][c] = sum_k v[k][c]*basis(k,zq), k ascending  (getDeterministicStates, TACSStochasticElement.cpp:84-120)
void refshim_evaluate_expansion(void *hh, int m, const double *V_, int k0, int k1, long long q0, long long q1, double *U_){
  RefHandle *h = (RefHandle*)hh;
  const scalar *V = (const scalar*)V_; scalar *U = (scalar*)U_;
  int nv = h->nvars;
  std::vector<scalar> zq(nv), yq(nv);
  for (long long q = q0; q < q1; q++){
    h->pc->quadrature((int)q, zq.data(), yq.data());
    scalar *u = U + (size_t)(q-q0)*m;
    for (int c = 0; c < m; c++) u[c] = 0.0;
    for (int k = k0; k < k1; k++){
      scalar psikz = h->pc->basis(k, zq.data());
      const scalar *v = V + (size_t)(k-k0)*m;
      for (int c = 0; c < m; c++) u[c] += v[c]*psikz;
    }
  }
}

// C[(i-i0)][(j-j0)][c] = sum_q (basis(i,zq)*basis(j,zq)*wq) * J[q][c]   (TACSStochasticElement.cpp:380-414)
// rows i are partitioned over nthreads std::threads (container reads are thread-safe after initialize()).
static void matrix_rows(RefHandle *h, int m2, const scalar *J, int ia, int ib, int i0, int j0, int j1, scalar *C){
  int Q = h->pc->getNumQuadraturePoints(), nv = h->nvars;
  std::vector<scalar> zq(nv), yq(nv);
  int nj = j1 - j0;
  for (int i = ia; i < ib; i++){
    for (int j = j0; j < j1; j++){
      scalar *A = C + ((size_t)(i-i0)*nj + (j-j0))*m2;
      for (int c = 0; c < m2; c++) A[c] = 0.0;
      for (int q = 0; q < Q; q++){
        scalar wq = h->pc->quadrature(q, zq.data(), yq.data());
        scalar scale = h->pc->basis(i, zq.data())*h->pc->basis(j, zq.data())*wq;
        const scalar *Jq = J + (size_t)q*m2;
        for (int c = 0; c < m2; c++) A[c] += scale*Jq[c];
      }
    }
  }
}
void refshim_project_matrix(void *hh, int m2, const double *J_, int i0, int i1, int j0, int j1,
                            double *C_, int nthreads){
  RefHandle *h = (RefHandle*)hh;
  const scalar *J = (const scalar*)J_; scalar *C = (scalar*)C_;
  if (nthreads <= 1){ matrix_rows(h, m2, J, i0, i1, i0, j0, j1, C); return; }
  std::vector<std::thread> th;
  int rows = i1 - i0;
  for (int t = 0; t < nthreads; t++){
    int ia = i0 + (int)((long long)rows*t/nthreads), ib = i0 + (int)((long long)rows*(t+1)/nthreads);
    if (ib > ia) th.emplace_back(matrix_rows, h, m2, J, ia, ib, i0, j0, j1, C);
  }
  for (auto &t : th) t.join();
}

// (f3) moments of m function values per point over q in [q0,q1)  (TACSStochasticFunction.cpp:209-238 getFunctionValue):
// fmean += wq*basis(0,zq)*f ; ffmean += wq*basis(0,zq)*f*f ; variance = ffmean - fmean*fmean.   f holds rows q0..q1-1.
void refshim_moments(void *hh, int m, const double *f_, long long q0, long long q1, double *mean_, double *second_, double *var_){
  RefHandle *h = (RefHandle*)hh;
  const scalar *f = (const scalar*)f_; scalar *mean = (scalar*)mean_, *second = (scalar*)second_, *var = (scalar*)var_;
  std::vector<scalar> zq(h->nvars), yq(h->nvars);
  for (int c = 0; c < m; c++){ mean[c] = 0.0; second[c] = 0.0; }
  for (long long q = q0; q < q1; q++){
    scalar wq = h->pc->quadrature((int)q, zq.data(), yq.data());
    scalar wb = wq*h->pc->basis(0, zq.data());
    const scalar *fq = f + (size_t)(q-q0)*m;
    for (int c = 0; c < m; c++){ mean[c] += wb*fq[c]; second[c] += wb*fq[c]*fq[c]; }
  }
  for (int c = 0; c < m; c++) var[c] = second[c] - mean[c]*mean[c];
}

// (f4) the loop of TACSStochasticElement::addAdjResProduct (examples/stacs/cpp/TACSStochasticElement.cpp:548-637) over
// the reference's primitives: for j = 0 (:583), for q: wq = quadrature(q,zq,yq); wt = basis(j,zq)*wq; deterministic
// states / adjoint from the TACS-interleaved stochastic vectors (getDeterministicStates :84-120, getDeterministicAdjoint
// :53-82); the deterministic element (TACS is not vendored: a callback with the oracle's signature) adds into dfdxj with
// weight wt*scale; dfdx[n] += dfdxj[n].
typedef void (*adjres_element)(void *user, const double *scale, const double *psi, const double *u, const double *y,
                               int m, int nvars, int dvlen, double *dfdx);
void refshim_adj_res_product(void *hh, int nnodes, int ndvpn, int dvlen, const double *scale_, const double *v_,
                             const double *adj_, adjres_element elem, void *user, double *dfdx_){
  RefHandle *h = (RefHandle*)hh;
  const scalar scale = *(const scalar*)scale_;
  const scalar *v = (const scalar*)v_, *adj = (const scalar*)adj_;
  scalar *dfdx = (scalar*)dfdx_;
  const int nsterms = h->pc->getNumBasisTerms(), nsvpn = nsterms*ndvpn, nddof = nnodes*ndvpn;
  const int nqpts = h->pc->getNumQuadraturePoints();
  std::vector<scalar> zq(h->nvars), yq(h->nvars), uq(nddof), psiq(nddof), dfdxj(dvlen);
  for (int j = 0; j < 1; j++){
    for (int n = 0; n < dvlen; n++) dfdxj[n] = 0.0;
    for (int q = 0; q < nqpts; q++){
      scalar wq = h->pc->quadrature(q, zq.data(), yq.data());
      scalar wt = h->pc->basis(j, zq.data())*wq;
      for (int c = 0; c < nddof; c++){ uq[c] = 0.0; psiq[c] = 0.0; }
      for (int n = 0; n < nnodes; n++){
        for (int k = 0; k < nsterms; k++){
          scalar psikz = h->pc->basis(k, zq.data());
          int lptr = n*ndvpn, gptr = n*nsvpn + k*ndvpn;
          for (int d = 0; d < ndvpn; d++){
            uq[lptr+d] += v[gptr+d]*psikz;
            psiq[lptr+d] += adj[gptr+d]*psikz;
          }
        }
      }
      scalar ws = wt*scale;
      elem(user, (const double*)&ws, (const double*)psiq.data(), (const double*)uq.data(), (const double*)yq.data(),
           nddof, h->nvars, dvlen, (double*)dfdxj.data());
    }
    for (int n = 0; n < dvlen; n++) dfdx[n] += dfdxj[n];
  }
}

// (f2) basis-pair filter: the reference library's own BasisHelper::sparse (BasisHelper.cpp:76-88) per variable, combined as
// the caller's nonzero() does (TACSStochasticElement.cpp:7-21: product of the per-variable flags), on a degree table [N][nvars]
void refshim_pair_mask(int N, int nvars, const int *deg, const int *dmapf, unsigned char *mask){
  BasisHelper bh(0);
  std::vector<int> di(nvars), dj(nvars), dk(dmapf, dmapf + nvars);
  bool *flt = new bool[nvars];
  for (int i = 0; i < N; i++){
    for (int j = 0; j < N; j++){
      for (int d = 0; d < nvars; d++){ di[d] = deg[i*nvars + d]; dj[d] = deg[j*nvars + d]; }
      bh.sparse(nvars, di.data(), dj.data(), dk.data(), flt);
      int prod = 1;
      for (int d = 0; d < nvars; d++) prod *= flt[d] ? 1 : 0;
      mask[(size_t)i*N + j] = (unsigned char)prod;
    }
  }
  delete [] flt;
}

} // extern "C"
