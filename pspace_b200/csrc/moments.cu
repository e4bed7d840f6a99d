// (f3) Moments at the quadrature points — the step right after the projection in the reference's function classes:
//   mean[c]   = sum_q (w_q psi_0(z_q)) f[q][c]
//   second[c] = sum_q (w_q psi_0(z_q)) f[q][c] f[q][c]            variance = second - mean*mean
// (examples/stacs/cpp/TACSStochasticFunction.cpp:209-238; the collocation form E[f] = sum_q w_q f_q of
// examples/ouu/twobartruss/stochastic_two_bar_truss.py:99-114).  psi_0 is the degree-0 basis function (the first
// row of the graded degree table), which the orthonormal families evaluate to exactly 1.
// HBM bound: f is read once (8*ncols bytes per point).  Deterministic: fixed q-slices per block, fixed-order
// shared-memory reduction, slices added in a fixed order by k_moments_final (a warp per column).
#include "ctx.h"
#include <algorithm>

namespace {
constexpr int MT = 256;            // threads per block
constexpr int MT_COLS = 128;       // widest column tile; wider f is covered by blockIdx.y

// Thread t of a block owns column (t % ct) of the tile and the rows (t / ct) + k*R of the block's q-slice, so the
// threads of a block read one contiguous run of R rows per step (fully coalesced), 4 steps in flight.
template <bool CPLX>
__global__ void __launch_bounds__(MT) k_moments_partial(const double *f, const double *W, long long q0, long long nq,
                                                         int ncols, int ct, long long rows_per_slice,
                                                         double *part /*[slices][2][ncols*width]*/) {
  constexpr int WD = CPLX ? 2 : 1;
  __shared__ double sm[MT][2 * WD];
  const int R = MT / ct;                               // rows handled in parallel
  const int c_in = threadIdx.x % ct, rr = threadIdx.x / ct;
  const int col = blockIdx.y * ct + c_in;
  const long long r0 = blockIdx.x * rows_per_slice, r1 = min(nq, r0 + rows_per_slice);
  double m_re = 0, m_im = 0, s_re = 0, s_im = 0;
  const bool active = rr < R && col < ncols;
  if (active) {
    long long r = r0 + rr;
    for (; r + 3LL * R < r1; r += 4LL * R) {
      double w[4][2], v[4][2];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const long long row = r + (long long)u * R;
        if (!CPLX) { w[u][0] = W[q0 + row]; v[u][0] = f[row * ncols + col]; }
        else {
          w[u][0] = W[2 * (q0 + row)]; w[u][1] = W[2 * (q0 + row) + 1];
          v[u][0] = f[2 * (row * ncols + col)]; v[u][1] = f[2 * (row * ncols + col) + 1];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (!CPLX) { const double t = w[u][0] * v[u][0]; m_re += t; s_re += t * v[u][0]; }
        else {
          const double tr = w[u][0] * v[u][0] - w[u][1] * v[u][1], ti = w[u][0] * v[u][1] + w[u][1] * v[u][0];
          m_re += tr; m_im += ti;
          s_re += tr * v[u][0] - ti * v[u][1]; s_im += tr * v[u][1] + ti * v[u][0];
        }
      }
    }
    for (; r < r1; r += R) {
      if (!CPLX) { const double wv = W[q0 + r], v = f[r * ncols + col]; const double t = wv * v; m_re += t; s_re += t * v; }
      else {
        const double wr = W[2 * (q0 + r)], wi = W[2 * (q0 + r) + 1];
        const double vr = f[2 * (r * ncols + col)], vi = f[2 * (r * ncols + col) + 1];
        const double tr = wr * vr - wi * vi, ti = wr * vi + wi * vr;
        m_re += tr; m_im += ti;
        s_re += tr * vr - ti * vi; s_im += tr * vi + ti * vr;
      }
    }
  }
  sm[threadIdx.x][0] = m_re; sm[threadIdx.x][WD] = s_re;
  if (CPLX) { sm[threadIdx.x][1] = m_im; sm[threadIdx.x][WD + 1] = s_im; }
  __syncthreads();
  if (rr == 0 && col < ncols) {                        // fixed-order reduction over the R row groups
    double a[2 * WD];
#pragma unroll
    for (int k = 0; k < 2 * WD; k++) a[k] = sm[c_in][k];
    for (int y = 1; y < R; y++)
#pragma unroll
      for (int k = 0; k < 2 * WD; k++) a[k] += sm[y * ct + c_in][k];
    double *dst = part + (size_t)blockIdx.x * 2 * ncols * WD;
#pragma unroll
    for (int k = 0; k < WD; k++) { dst[col * WD + k] = a[k]; dst[(size_t)ncols * WD + col * WD + k] = a[WD + k]; }
  }
}

// one warp per column: lane l adds the slices l, l+32, ... in order, then a fixed xor-shuffle tree over the lanes
// (deterministic; the partials of a column are read by 32 lanes at once instead of one thread walking every slice)
template <bool CPLX>
__global__ void __launch_bounds__(128) k_moments_final(const double *part, int slices, int ncols, double *mean,
                                                        double *second, double *var) {
  constexpr int WD = CPLX ? 2 : 1;
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (col >= ncols) return;                    // warp-uniform
  double m[2] = {0, 0}, s[2] = {0, 0};
  for (int b = lane; b < slices; b += 32) {
    const double *src = part + (size_t)b * 2 * ncols * WD;
#pragma unroll
    for (int k = 0; k < WD; k++) { m[k] += src[col * WD + k]; s[k] += src[(size_t)ncols * WD + col * WD + k]; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int k = 0; k < WD; k++) {
      m[k] += __shfl_xor_sync(0xffffffffu, m[k], off);
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
    }
  if (lane != 0) return;
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
  const int ct = ncols < MT_COLS ? ncols : MT_COLS;          // column tile; R = 256/ct rows per step
  const int ctiles = (ncols + ct - 1) / ct;
  const long long rows_per_step = MT / ct;
  int slices = (int)std::max<long long>(1, std::min<long long>((148LL * 8 + ctiles - 1) / ctiles, (nq + 16 * rows_per_step - 1) / (16 * rows_per_step)));
  const long long rows_per_slice = (nq + slices - 1) / slices;
  slices = (int)((nq + rows_per_slice - 1) / rows_per_slice);
  double *part = nullptr;
  PS_CUDA(cudaMallocAsync(&part, sizeof(double) * (size_t)slices * 2 * ncols * c->width, st));
  dim3 grid(slices, ctiles);
  if (c->is_complex) k_moments_partial<true><<<grid, MT, 0, st>>>(d_f, c->d_W, q0, nq, ncols, ct, rows_per_slice, part);
  else k_moments_partial<false><<<grid, MT, 0, st>>>(d_f, c->d_W, q0, nq, ncols, ct, rows_per_slice, part);
  PS_LAUNCH_CHECK(c);
  if (c->is_complex) k_moments_final<true><<<(ncols + 3) / 4, 128, 0, st>>>(part, slices, ncols, d_mean, d_second, d_var);
  else k_moments_final<false><<<(ncols + 3) / 4, 128, 0, st>>>(part, slices, ncols, d_mean, d_second, d_var);
  PS_LAUNCH_CHECK(c);
  PS_CUDA(cudaFreeAsync(part, st));
  return PSPACE_OK;
}
