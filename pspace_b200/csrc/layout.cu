// (f2) TACS-interleaved layouts and the basis-pair sparsity filter — pure index work, bit-exact.
//
// The reference caller stores stochastic states / residuals / Jacobians node-major with the basis index between
// node and dof (examples/stacs/cpp/TACSStochasticElement.cpp):
//   vector   res[n*nsvpn + i*ndvpn + d]                     += F[i][n*ndvpn + d]           (:316-322)      nsvpn = N*ndvpn
//   matrix   mat[(ni*nsvpn + i*ndvpn + di)*nsdof + nj*nsvpn + j*ndvpn + dj]
//                                                         += A_ij[(ni*ndvpn+di)*nddof + nj*ndvpn+dj]  (:417-432)  nsdof = N*nddof
//   states   u_q gathers v[n*nsvpn + k*ndvpn + d]  (getDeterministicStates :84-120)
//   point quantity   quantity[d*nsterms + i]                                                        (:525)
// and may skip basis pairs whose Jacobian block is structurally zero:
//   nonzero(i,j) = prod_d ( |deg_i[d] - deg_j[d]| <= dmapf[d] )      (:7-21; BasisHelper::sparse BasisHelper.cpp:76-88)
// The projection kernels work on the blocked layouts F[N][m], C[N][N][m*m]; these kernels convert.  HBM bound:
// every element is read once and written (or read-modify-written) once.
#include "ctx.h"
#include <algorithm>

namespace {
template <int W>
__global__ void k_scatter_vector_tacs(int N, int nnodes, int ndvpn, const double *F, double *res, int accumulate) {
  const long long m = (long long)nnodes * ndvpn, total = (long long)N * m;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), c = (int)(e % m), n = c / ndvpn, d = c % ndvpn;
    const long long dst = (long long)n * N * ndvpn + (long long)i * ndvpn + d;
#pragma unroll
    for (int w = 0; w < W; w++) res[W * dst + w] = accumulate ? res[W * dst + w] + F[W * e + w] : F[W * e + w];
  }
}
template <int W>
__global__ void k_gather_vector_tacs(int N, int nnodes, int ndvpn, const double *v, double *V) {
  const long long m = (long long)nnodes * ndvpn, total = (long long)N * m;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / m), c = (int)(e % m), n = c / ndvpn, d = c % ndvpn;
    const long long src = (long long)n * N * ndvpn + (long long)k * ndvpn + d;
#pragma unroll
    for (int w = 0; w < W; w++) V[W * e + w] = v[W * src + w];
  }
}
// Destination order: a block walks rows of `mat`, its threads consecutive columns, so the read-modify-write of the
// big interleaved matrix (2/3 of the traffic) is fully coalesced; the source C is gathered in runs of ndvpn values.
template <int W>
__global__ void k_scatter_matrix_tacs(int N, int nnodes, int ndvpn, const double *C, double *mat, int accumulate) {
  const int m = nnodes * ndvpn, nsvpn = N * ndvpn, nsdof = N * m;
  const long long m2 = (long long)m * m;
  for (int row = blockIdx.x; row < nsdof; row += gridDim.x) {
    const int rr = row % nsvpn, i = rr / ndvpn, r = (row / nsvpn) * ndvpn + rr % ndvpn;
    const long long src_row = (long long)i * N * m2 + (long long)r * m;
    double *drow = mat + (size_t)W * row * nsdof;
    for (int col = threadIdx.x; col < nsdof; col += blockDim.x) {
      const int cr = col % nsvpn, j = cr / ndvpn, c = (col / nsvpn) * ndvpn + cr % ndvpn;
      const long long src = src_row + (long long)j * m2 + c;
#pragma unroll
      for (int w = 0; w < W; w++) drow[W * col + w] = accumulate ? drow[W * col + w] + C[W * src + w] : C[W * src + w];
    }
  }
}
// out[d*N + i] = F[i][d]   (evalPointQuantity layout)
template <int W>
__global__ void k_point_quantity_layout(int N, int m, const double *F, double *out) {
  const long long total = (long long)N * m;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), d = (int)(e % m);
#pragma unroll
    for (int w = 0; w < W; w++) out[W * ((long long)d * N + i) + w] = F[W * e + w];
  }
}
__global__ void k_pair_mask(int N, int nvars, const int *deg, const int *dmapf, unsigned char *mask) {
  const long long total = (long long)N * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / N), j = (int)(e % N);
    int ok = 1;
    for (int d = 0; d < nvars; d++) {
      int diff = deg[(size_t)i * nvars + d] - deg[(size_t)j * nvars + d];
      diff = diff < 0 ? -diff : diff;
      ok &= (diff <= dmapf[d]) ? 1 : 0;
    }
    mask[e] = (unsigned char)ok;
  }
}
unsigned blocks_for(long long n) { return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 32)); }
}  // namespace

extern "C" {
int pspace_cuda_scatter_vector_tacs(int N, int nnodes, int ndvpn, int is_complex, const double *d_F, double *d_res,
                                    int accumulate, void *stream) {
  if (N < 0 || nnodes < 1 || ndvpn < 1 || !d_F || !d_res) return ps_fail(PSPACE_EINVAL, "bad argument");
  const long long n = (long long)N * nnodes * ndvpn;
  if (n == 0) return PSPACE_OK;
  if (is_complex) k_scatter_vector_tacs<2><<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(N, nnodes, ndvpn, d_F, d_res, accumulate);
  else k_scatter_vector_tacs<1><<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(N, nnodes, ndvpn, d_F, d_res, accumulate);
  PS_LAUNCH_CHECK(nullptr);
  return PSPACE_OK;
}
int pspace_cuda_gather_vector_tacs(int N, int nnodes, int ndvpn, int is_complex, const double *d_v, double *d_V, void *stream) {
  if (N < 0 || nnodes < 1 || ndvpn < 1 || !d_v || !d_V) return ps_fail(PSPACE_EINVAL, "bad argument");
  const long long n = (long long)N * nnodes * ndvpn;
  if (n == 0) return PSPACE_OK;
  if (is_complex) k_gather_vector_tacs<2><<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(N, nnodes, ndvpn, d_v, d_V);
  else k_gather_vector_tacs<1><<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(N, nnodes, ndvpn, d_v, d_V);
  PS_LAUNCH_CHECK(nullptr);
  return PSPACE_OK;
}
int pspace_cuda_scatter_matrix_tacs(int N, int nnodes, int ndvpn, int is_complex, const double *d_C, double *d_mat,
                                    int accumulate, void *stream) {
  if (N < 0 || nnodes < 1 || ndvpn < 1 || !d_C || !d_mat) return ps_fail(PSPACE_EINVAL, "bad argument");
  const long long m = (long long)nnodes * ndvpn, n = (long long)N * N * m * m;
  if (n == 0) return PSPACE_OK;
  const unsigned rows_grid = (unsigned)std::min<long long>((long long)N * m, 148LL * 16);
  if (is_complex) k_scatter_matrix_tacs<2><<<rows_grid, 256, 0, (cudaStream_t)stream>>>(N, nnodes, ndvpn, d_C, d_mat, accumulate);
  else k_scatter_matrix_tacs<1><<<rows_grid, 256, 0, (cudaStream_t)stream>>>(N, nnodes, ndvpn, d_C, d_mat, accumulate);
  PS_LAUNCH_CHECK(nullptr);
  return PSPACE_OK;
}
int pspace_cuda_point_quantity_layout(int N, int m, int is_complex, const double *d_F, double *d_out, void *stream) {
  if (N < 0 || m < 1 || !d_F || !d_out) return ps_fail(PSPACE_EINVAL, "bad argument");
  const long long n = (long long)N * m;
  if (n == 0) return PSPACE_OK;
  if (is_complex) k_point_quantity_layout<2><<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(N, m, d_F, d_out);
  else k_point_quantity_layout<1><<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(N, m, d_F, d_out);
  PS_LAUNCH_CHECK(nullptr);
  return PSPACE_OK;
}
int pspace_cuda_pair_mask(const pspace_ctx *c, const int *dmapf, unsigned char *d_mask, void *stream) {
  if (!c || !dmapf || !d_mask) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->basis_ready) return ps_fail(PSPACE_ESTATE, "basis not initialised");
  PS_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int *d_f = nullptr;
  PS_CUDA(cudaMallocAsync(&d_f, sizeof(int) * c->nvars, st));
  PS_CUDA(cudaMemcpyAsync(d_f, dmapf, sizeof(int) * c->nvars, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaStreamSynchronize(st));   // dmapf is a caller stack array in practice
  k_pair_mask<<<blocks_for((long long)c->N * c->N), 256, 0, st>>>(c->N, c->nvars, c->d_deg, d_f, d_mask);
  PS_LAUNCH_CHECK(c);
  PS_CUDA(cudaFreeAsync(d_f, st));
  return PSPACE_OK;
}
}
