// (f3) Moments at the quadrature points — the step right after the projection in the reference's function classes:
//   mean[c]   = sum_q (w_q psi_0(z_q)) f[q][c]
//   second[c] = sum_q (w_q psi_0(z_q)) f[q][c] f[q][c]            variance = second - mean*mean
// (examples/stacs/cpp/TACSStochasticFunction.cpp:209-238; the collocation form E[f] = sum_q w_q f_q of
// examples/ouu/twobartruss/stochastic_two_bar_truss.py:99-114).  psi_0 is the degree-0 basis function (the first
// row of the graded degree table), which the orthonormal families evaluate to exactly 1.
// HBM bound: f is read once (8*ncols bytes per point).  Deterministic: fixed q-slices per block, fixed-order
// shared-memory reduction, slices added in order by k_moments_final.
#include "ctx.h"
#include <algorithm>

namespace {
constexpr int MT_X = 32, MT_Y = 8;     // 32 columns x 8 rows of quadrature points per block step

template <bool CPLX>
__global__ void __launch_bounds__(MT_X *MT_Y) k_moments_partial(const double *f, const double *W, long long q0, long long nq,
                                                               int ncols, long long rows_per_slice, double *part /*[slices][2][ncols*width]*/) {
  constexpr int WD = CPLX ? 2 : 1;
  __shared__ double sm[MT_Y][MT_X][2 * WD];
  const int tx = threadIdx.x % MT_X, ty = threadIdx.x / MT_X;
  const int col = blockIdx.y * MT_X + tx;
  const long long r0 = blockIdx.x * rows_per_slice, r1 = min(nq, r0 + rows_per_slice);
  double m_re = 0, m_im = 0, s_re = 0, s_im = 0;
  if (col < ncols) {
    for (long long r = r0 + ty; r < r1; r += MT_Y) {
      if (!CPLX) {
        const double w = W[q0 + r], v = f[r * ncols + col];
        const double t = w * v;          // (w_q*psi_0)*f with psi_0 == 1
        m_re += t;
        s_re += t * v;
      } else {
        const double wr = W[2 * (q0 + r)], wi = W[2 * (q0 + r) + 1];
        const double vr = f[2 * (r * ncols + col)], vi = f[2 * (r * ncols + col) + 1];
        const double tr = wr * vr - wi * vi, ti = wr * vi + wi * vr;
        m_re += tr; m_im += ti;
        s_re += tr * vr - ti * vi; s_im += tr * vi + ti * vr;
      }
    }
  }
  sm[ty][tx][0] = m_re; sm[ty][tx][WD] = s_re;
  if (CPLX) { sm[ty][tx][1] = m_im; sm[ty][tx][WD + 1] = s_im; }
  __syncthreads();
  if (ty == 0 && col < ncols) {
    double a[2 * WD];
#pragma unroll
    for (int k = 0; k < 2 * WD; k++) a[k] = sm[0][tx][k];
    for (int y = 1; y < MT_Y; y++)
#pragma unroll
      for (int k = 0; k < 2 * WD; k++) a[k] += sm[y][tx][k];
    double *dst = part + (size_t)blockIdx.x * 2 * ncols * WD;
#pragma unroll
    for (int k = 0; k < WD; k++) { dst[col * WD + k] = a[k]; dst[(size_t)ncols * WD + col * WD + k] = a[WD + k]; }
  }
}

template <bool CPLX>
__global__ void k_moments_final(const double *part, int slices, int ncols, double *mean, double *second, double *var) {
  constexpr int WD = CPLX ? 2 : 1;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double m[2] = {0, 0}, s[2] = {0, 0};
  for (int b = 0; b < slices; b++) {
    const double *src = part + (size_t)b * 2 * ncols * WD;
    for (int k = 0; k < WD; k++) { m[k] += src[col * WD + k]; s[k] += src[(size_t)ncols * WD + col * WD + k]; }
  }
  for (int k = 0; k < WD; k++) {
    if (mean) mean[col * WD + k] = m[k];
    if (second) second[col * WD + k] = s[k];
  }
  if (var) {
    if (!CPLX) var[col] = s[0] - m[0] * m[0];
    else { var[2 * col] = s[0] - (m[0] * m[0] - m[1] * m[1]); var[2 * col + 1] = s[1] - 2.0 * m[0] * m[1]; }
  }
}
}  // namespace

extern "C" int pspace_cuda_moments(const pspace_ctx *c, long long q0, long long q1, int ncols, const double *d_f,
                                   double *d_mean, double *d_second, double *d_var, void *stream) {
  if (!c || !d_f) return ps_fail(PSPACE_EINVAL, "null argument");
  if (!c->quad_ready) return ps_fail(PSPACE_ESTATE, "quadrature not initialised");
  if (q0 < 0 || q1 > c->Q || q1 <= q0 || ncols < 1) return ps_fail(PSPACE_EINVAL, "range");
  PS_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long nq = q1 - q0;
  const int ctiles = (ncols + MT_X - 1) / MT_X;
  int slices = (int)std::max<long long>(1, std::min<long long>((148LL * 8 + ctiles - 1) / ctiles, (nq + 63) / 64));
  const long long rows_per_slice = (nq + slices - 1) / slices;
  slices = (int)((nq + rows_per_slice - 1) / rows_per_slice);
  double *part = nullptr;
  PS_CUDA(cudaMallocAsync(&part, sizeof(double) * (size_t)slices * 2 * ncols * c->width, st));
  dim3 grid(slices, ctiles);
  if (c->is_complex) k_moments_partial<true><<<grid, MT_X * MT_Y, 0, st>>>(d_f, c->d_W, q0, nq, ncols, rows_per_slice, part);
  else k_moments_partial<false><<<grid, MT_X * MT_Y, 0, st>>>(d_f, c->d_W, q0, nq, ncols, rows_per_slice, part);
  PS_LAUNCH_CHECK(c);
  if (c->is_complex) k_moments_final<true><<<(ncols + 127) / 128, 128, 0, st>>>(part, slices, ncols, d_mean, d_second, d_var);
  else k_moments_final<false><<<(ncols + 127) / 128, 128, 0, st>>>(part, slices, ncols, d_mean, d_second, d_var);
  PS_LAUNCH_CHECK(c);
  PS_CUDA(cudaFreeAsync(part, st));
  return PSPACE_OK;
}
