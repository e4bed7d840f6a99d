// (f4) Adjoint-residual product — TACSStochasticElement::addAdjResProduct (examples/stacs/cpp/TACSStochasticElement.cpp:548-637),
// the derivative path of the reference's callers (verified there by complex step, examples/stacs/smd/projection.cpp:228-236).
//
// The reference walks the quadrature points; at each it rebuilds the deterministic states u_q and the deterministic adjoint
// psi_q from the TACS-interleaved stochastic vectors with N calls of basis() per node (getDeterministicStates :84-120,
// getDeterministicAdjoint :53-82), lets the deterministic element add  wt*scale * g(psi_q, u_q, y_q)  with
// wt = basis(0, z_q) * w_q  (:590; only basis index j = 0 is projected on, :583), and adds the sum into dfdx (:606-608).
// Batched here as three steps around the caller's element:
//   1. pspace_cuda_adjoint_states   k_gather_pair puts v and adj side by side as ONE coefficient matrix [N][2m]; a single
//                                   expansion evaluation (expand.cu, factored operand: psi generated once for both) gives
//                                   UP[q][0..m) = u_q, UP[q][m..2m) = psi_q for every point of the range;
//   2. the caller's element evaluates G[q][n] = g_n(psi_q, u_q, y_q), n < dvLen, on the device (its code, not ours);
//   3. pspace_cuda_adjoint_product  dfdx[n] (+)= scale * sum_q (w_q psi_0(z_q)) G[q][n]: the weighted reduction over the
//                                   points (moments.cu, HBM bound) and a scalar axpy.
// Real and USE_COMPLEX: the complex build carries a 1e-30j perturbation through all three steps, so Im(.)/1e-30 of a mean
// reproduces the real build's adjoint derivative (the reference's verification flow).
#include <algorithm>
#include "ctx.h"

namespace {
template <int W>
__global__ void k_gather_pair(int N, int nnodes, int ndvpn, const double *v, const double *adj, double *V) {
  const long long m = (long long)nnodes * ndvpn, total = (long long)N * 2 * m;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / (2 * m)), c2 = (int)(e % (2 * m));
    const int c = c2 < m ? c2 : c2 - (int)m, n = c / ndvpn, d = c % ndvpn;
    const double *src = (c2 < m ? v : adj) + (size_t)W * ((long long)n * N * ndvpn + (long long)k * ndvpn + d);
#pragma unroll
    for (int w = 0; w < W; w++) V[W * e + w] = src[w];
  }
}
template <bool CPLX>
__global__ void k_scaled_add(int n, double sr, double si, const double *x, double *y, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!CPLX) y[i] = (accumulate ? y[i] : 0.0) + sr * x[i];
  else {
    const double xr = x[2 * i], xi = x[2 * i + 1];
    y[2 * i] = (accumulate ? y[2 * i] : 0.0) + (sr * xr - si * xi);
    y[2 * i + 1] = (accumulate ? y[2 * i + 1] : 0.0) + (sr * xi + si * xr);
  }
}
}  // namespace

extern "C" int pspace_cuda_adjoint_states(pspace_ctx *c, int nnodes, int ndvpn, long long q0, long long q1, const double *d_v,
                                          const double *d_adj, double *d_UP, void *stream) {
  if (!c || !d_v || !d_adj || (!d_UP && q1 > q0)) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->tables_ready) return ps_fail(PSPACE_ESTATE, "adjoint_states called before initializeBasis + initializeQuadrature");
  if (nnodes < 1 || ndvpn < 1 || q0 < 0 || q1 > c->Q || q1 < q0) return ps_fail(PSPACE_EINVAL, "range");
  PS_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long m = (long long)nnodes * ndvpn, n = (long long)c->N * 2 * m;
  double *V = nullptr;          // stream-ordered scratch: no cross-stream state in the context
  PS_CUDA(cudaMallocAsync(&V, sizeof(double) * c->width * (size_t)n, st));
  const unsigned nb = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 8);
  if (c->is_complex) k_gather_pair<2><<<nb, 256, 0, st>>>(c->N, nnodes, ndvpn, d_v, d_adj, V);
  else k_gather_pair<1><<<nb, 256, 0, st>>>(c->N, nnodes, ndvpn, d_v, d_adj, V);
  PS_LAUNCH_CHECK(c);
  const int rc = ps_expand(c, 0, c->N, q0, q1, (int)(2 * m), 0, V, d_UP, st);
  PS_CUDA(cudaFreeAsync(V, st));
  return rc;
}

extern "C" int pspace_cuda_adjoint_product(pspace_ctx *c, int dvlen, const double *h_scale, long long q0, long long q1,
                                           const double *d_G, double *d_dfdx, int accumulate, void *stream) {
  if (!c || !h_scale || !d_G || !d_dfdx) return ps_fail(PSPACE_EINVAL, "null argument");
  if (dvlen < 1 || q0 < 0 || q1 > c->Q || q1 <= q0) return ps_fail(PSPACE_EINVAL, "range");
  PS_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  // mean[n] = sum_q (w_q psi_0(z_q)) G[q][n]  (psi_0 == 1 exactly: every degree-0 factor is 1.0), then dfdx (+)= scale * mean
  double *mean = nullptr;
  PS_CUDA(cudaMallocAsync(&mean, sizeof(double) * c->width * (size_t)dvlen, st));
  int rc = pspace_cuda_moments(c, q0, q1, dvlen, d_G, mean, nullptr, nullptr, stream);
  if (rc) { cudaFreeAsync(mean, st); return rc; }
  const double sr = h_scale[0], si = c->is_complex ? h_scale[1] : 0.0;
  if (c->is_complex) k_scaled_add<true><<<(dvlen + 127) / 128, 128, 0, st>>>(dvlen, sr, si, mean, d_dfdx, accumulate);
  else k_scaled_add<false><<<(dvlen + 127) / 128, 128, 0, st>>>(dvlen, sr, si, mean, d_dfdx, accumulate);
  PS_LAUNCH_CHECK(c);
  PS_CUDA(cudaFreeAsync(mean, st));
  return PSPACE_OK;
}
