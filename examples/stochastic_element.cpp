// A C++ caller shaped like the reference's TACSStochasticElement (examples/stacs/cpp/TACSStochasticElement.cpp),
// written against pspace_b200/host/include — the same three headers a reference caller includes.
//
// The deterministic element here is a toy 2-dof spring-mass-damper whose stiffness, damping and forcing depend on the
// random parameters y = (k, c, f0):   R(u; y) = [ m*udd0 + c*(ud0-ud1) + k*(u0-u1) - f0 ,  m*udd1 - c*(ud0-ud1) - k*(u0-u1) + k*u1 ]
// so its Jacobian block J(y) = alpha*dR/du + beta*dR/dud + gamma*dR/dudd is a 2x2 matrix, affine in y.
//
// What the reference does per element (addJacobian :380-438, addResidual :283-324) is N*N*Q resp. N*Q calls of the
// deterministic element with 2 basis() evaluations each.  Here the deterministic blocks are evaluated ONCE per
// quadrature point and the container projects them in one call each:
//     pc->projectMatrix(m*m, J, C)      C[(i*N+j)*m*m + r*m + c] = sum_q w_q psi_i(z_q) psi_j(z_q) J_q[r][c]
//     pc->projectVector(m, f, F)        F[k*m + c]               = sum_q w_q psi_k(z_q) f_q[c]
// `--check` also runs the reference's own (i,j,q) loops on the host, over pc->basis()/pc->quadrature(), and compares.
//
//   make -C examples && ./examples/stochastic_element --check            (needs a CUDA device)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ParameterContainer.h"
#include "ParameterFactory.h"

namespace {
const int NDOF = 2;

// deterministic residual and Jacobian at parameter values y = (k, c, f0)
void det_residual(const scalar *y, const scalar *u, const scalar *ud, const scalar *udd, scalar *res) {
  const scalar k = y[0], c = y[1], f0 = y[2], mass = 2.5;
  res[0] = mass * udd[0] + c * (ud[0] - ud[1]) + k * (u[0] - u[1]) - f0;
  res[1] = mass * udd[1] - c * (ud[0] - ud[1]) - k * (u[0] - u[1]) + k * u[1];
}
void det_jacobian(const scalar *y, double alpha, double beta, double gamma, scalar *A /* row-major 2x2 */) {
  const scalar k = y[0], c = y[1], mass = 2.5;
  A[0] = alpha * k + beta * c + gamma * mass;
  A[1] = -alpha * k - beta * c;
  A[2] = -alpha * k - beta * c;
  A[3] = alpha * (k + k) + beta * c + gamma * mass;
}
double mag(const scalar &x) { return std::abs(x); }
}  // namespace

int main(int argc, char **argv) {
  const bool check = argc > 1 && !strcmp(argv[1], "--check");
  ParameterFactory *factory = new ParameterFactory();
  AbstractParameter *pk = factory->createNormalParameter(10.0, 1.0, 3);        // stiffness
  AbstractParameter *pc_ = factory->createUniformParameter(0.1, 0.5, 2);      // damping
  AbstractParameter *pf = factory->createExponentialParameter(1.0, 0.25, 2);  // forcing
  ParameterContainer *pc = new ParameterContainer(1 /* complete basis */);
  pc->addParameter(pk);
  pc->addParameter(pc_);
  pc->addParameter(pf);
  pc->initialize();
  const int N = pc->getNumBasisTerms(), Q = pc->getNumQuadraturePoints(), nv = pc->getNumParameters();
  const int m = NDOF, m2 = m * m;
  const double alpha = 1.0, beta = 0.3, gamma = 0.05;
  printf("stochastic element: %d parameters, %d basis terms, %d quadrature points, %dx%d blocks\n", nv, N, Q, m, m);

  // a state expansion v[k][dof] to linearise about; deterministic states at every quadrature point in one call
  std::vector<scalar> v((size_t)N * m), uq((size_t)Q * m), zero((size_t)m, 0.0);
  for (int k = 0; k < N; k++)
    for (int d = 0; d < m; d++) v[(size_t)k * m + d] = 1.0 / (1.0 + k) + 0.1 * d;
  pc->evaluateExpansion(m, v.data(), uq.data());

  // deterministic blocks, once per quadrature point
  std::vector<scalar> J((size_t)Q * m2), f((size_t)Q * m), zq(nv), yq(nv);
  for (int q = 0; q < Q; q++) {
    pc->quadrature(q, zq.data(), yq.data());
    det_jacobian(yq.data(), alpha, beta, gamma, &J[(size_t)q * m2]);
    det_residual(yq.data(), &uq[(size_t)q * m], zero.data(), zero.data(), &f[(size_t)q * m]);
  }

  // the two projections, one call each
  std::vector<scalar> C((size_t)N * N * m2), F((size_t)N * m);
  pc->projectMatrix(m2, J.data(), C.data());
  pc->projectVector(m, f.data(), F.data());

  // scatter into the stochastic element's interleaved layouts exactly as the reference does
  // (res[n*nsvpn + i*ndvpn + d] :316-322;  mat[(ni*nsvpn + i*ndvpn + di)*nsdof + nj*nsvpn + j*ndvpn + dj] :417-432)
  const int nnodes = 2, ndvpn = 1, nsvpn = N * ndvpn, nsdof = N * m;
  std::vector<scalar> res((size_t)nsdof, 0.0), mat((size_t)nsdof * nsdof, 0.0);
  for (int i = 0; i < N; i++)
    for (int n = 0; n < nnodes; n++)
      for (int d = 0; d < ndvpn; d++) res[(size_t)n * nsvpn + i * ndvpn + d] += F[(size_t)i * m + n * ndvpn + d];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      for (int ni = 0; ni < nnodes; ni++)
        for (int di = 0; di < ndvpn; di++)
          for (int nj = 0; nj < nnodes; nj++)
            for (int dj = 0; dj < ndvpn; dj++)
              mat[(size_t)(ni * nsvpn + i * ndvpn + di) * nsdof + nj * nsvpn + j * ndvpn + dj] +=
                  C[((size_t)i * N + j) * m2 + (ni * ndvpn + di) * m + nj * ndvpn + dj];
  double sres = 0.0, smat = 0.0;
  for (size_t e = 0; e < res.size(); e++) sres += mag(res[e]);
  for (size_t e = 0; e < mat.size(); e++) smat += mag(mat[e]);
  printf("sum |res| = %.15e   sum |mat| = %.15e\n", sres, smat);

  int bad = 0;
  if (check) {
    // the reference's loops: for every (i, j) walk all quadrature points, scale = basis(i)*basis(j)*w
    double emax = 0.0, cmax = 0.0, fmax = 0.0, femax = 0.0;
    std::vector<scalar> A(m2), Aq(m2), r(m), rq(m), us(m);
    for (int i = 0; i < N; i++) {
      std::fill(r.begin(), r.end(), scalar(0.0));
      for (int q = 0; q < Q; q++) {
        const scalar wq = pc->quadrature(q, zq.data(), yq.data());
        const scalar scale = pc->basis(i, zq.data()) * wq;
        for (int d = 0; d < m; d++) {          // getDeterministicStates: u_q = sum_k psi_k(z_q) v_k
          us[d] = 0.0;
          for (int k = 0; k < N; k++) us[d] += pc->basis(k, zq.data()) * v[(size_t)k * m + d];
        }
        det_residual(yq.data(), us.data(), zero.data(), zero.data(), rq.data());
        for (int d = 0; d < m; d++) r[d] += rq[d] * scale;
      }
      for (int d = 0; d < m; d++) {
        fmax = std::max(fmax, mag(r[d]));
        femax = std::max(femax, mag(r[d] - F[(size_t)i * m + d]));
      }
      for (int j = 0; j < N; j++) {
        std::fill(A.begin(), A.end(), scalar(0.0));
        for (int q = 0; q < Q; q++) {
          const scalar wq = pc->quadrature(q, zq.data(), yq.data());
          const scalar scale = pc->basis(i, zq.data()) * pc->basis(j, zq.data()) * wq;
          det_jacobian(yq.data(), alpha, beta, gamma, Aq.data());
          for (int c = 0; c < m2; c++) A[c] += scale * Aq[c];
        }
        for (int c = 0; c < m2; c++) {
          cmax = std::max(cmax, mag(A[c]));
          emax = std::max(emax, mag(A[c] - C[((size_t)i * N + j) * m2 + c]));
        }
      }
    }
    printf("matrix projection: max |C - C_ref| / max |C_ref| = %.3e\n", emax / cmax);
    printf("vector projection: max |F - F_ref| / max |F_ref| = %.3e\n", femax / fmax);
    bad = !(emax / cmax < 1e-12 && femax / fmax < 1e-12);
    printf(bad ? "FAILED\n" : "PASSED (<= 1e-12)\n");
  }
  delete pc;
  delete factory;
  return bad;
}
