// (f1) Polynomial-chaos expansion evaluated at quadrature points — the step right before the projection in every
// caller:      U[q][c] = sum_k psi_k(z_q) V[k][c]
// Replaces getDeterministicStates / getDeterministicAdjoint of the reference caller
// (examples/stacs/cpp/TACSStochasticElement.cpp:53-120: for every (i,j,q) it recomputes u_q with N calls of
// pc->basis(k,zq) per node; pspace/core.py:788-792).  Here it is one (nq x N).(N x m) FP64 GEMM per call:
//   * left operand = the unweighted basis panel P[part][q][k] written by k_psi_panel (project.cu) — q-major,
//     so a DMMA A-fragment row is a quadrature point and the fragment k index is the basis index;
//   * right operand V[k][c] (N x m, row-major; complex: interleaved re/im columns), staged whole rows at a time;
//   * DMMA.8x8x4 accumulation, 16-byte cp.async double-buffered K pipeline, conflict-free pitches (== 4 mod 16);
//   * complex: U = P_re*[V] + P_im*[V'] with V'[k][2c] = -V[k][2c+1], V'[k][2c+1] = V[k][2c]  (as in project.cu).
// Bound: HBM — it streams the panel once (8*N*nq bytes, cfg5 2.4 GB) for 2*N*m FLOP per point.
#include "ctx.h"

namespace {
constexpr int XT = 256;          // 8 warps, each 16 quadrature rows x BN columns
constexpr int XBM = 128;         // quadrature points per CTA
constexpr int XKC = 32;          // basis rows per pipeline stage
constexpr int XPA = XKC + 4;     // A pitch  (== 4 mod 16)

struct ExpandParams {
  const double *P;               // panel [W][nq][np]  (np = padded row pitch)
  const double *V;               // [N][ncols_d]
  double *U;                     // [nq][ncols_d]
  long long nq;
  int np, nk, ncols_d, accumulate, vecV;
};

__device__ __forceinline__ void cpa16(void *smem, const void *g, int bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa8(void *smem, const void *g, int bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip(double x, unsigned mask) {
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

template <int WN, bool CPLX>
__global__ void __launch_bounds__(XT) k_expand(const __grid_constant__ ExpandParams p) {
  constexpr int BN = WN * 8, PB = BN + 4, W = CPLX ? 2 : 1;
  extern __shared__ __align__(16) unsigned char xsm[];
  double *As = reinterpret_cast<double *>(xsm);               // [2][W][XBM][XPA]
  double *Bs = As + 2 * W * XBM * XPA;                        // [2][XKC][PB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, kl = lane & 3;
  const long long q0 = blockIdx.x * (long long)XBM;
  const int col0 = blockIdx.y * BN;
  const int nstage = (p.nk + XKC - 1) / XKC;

  auto issue = [&](int s) {
    if (s < nstage) {
      const int k0 = s * XKC, buf = s & 1;
      // A: XBM rows x XKC doubles per plane, 16-byte chunks (np is even, k0 multiple of 32)
      for (int idx = tid; idx < W * XBM * (XKC / 2); idx += XT) {
        const int ch = idx % (XKC / 2), row = (idx / (XKC / 2)) % XBM, part = idx / ((XKC / 2) * XBM);
        const int k = k0 + ch * 2;
        int bytes = (p.np - k) * 8;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        if (q0 + row >= p.nq) bytes = 0;
        const double *src = bytes ? p.P + ((size_t)part * p.nq + q0 + row) * p.np + k : p.P;
        cpa16(As + ((size_t)(buf * W + part) * XBM + row) * XPA + ch * 2, src, bytes);
      }
      // B: XKC rows x BN columns of V
      if (p.vecV) {
        for (int idx = tid; idx < XKC * (BN / 2); idx += XT) {
          const int ch = idx % (BN / 2), row = idx / (BN / 2);
          const int col = col0 + ch * 2;
          int bytes = (p.ncols_d - col) * 8;
          bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
          if (k0 + row >= p.nk) bytes = 0;
          const double *src = bytes ? p.V + (size_t)(k0 + row) * p.ncols_d + col : p.V;
          cpa16(Bs + ((size_t)buf * XKC + row) * PB + ch * 2, src, bytes);
        }
      } else {
        for (int idx = tid; idx < XKC * BN; idx += XT) {
          const int ch = idx % BN, row = idx / BN;
          const int col = col0 + ch;
          const int bytes = (col < p.ncols_d && k0 + row < p.nk) ? 8 : 0;
          const double *src = bytes ? p.V + (size_t)(k0 + row) * p.ncols_d + col : p.V;
          cpa8(Bs + ((size_t)buf * XKC + row) * PB + ch, src, bytes);
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  double acc[2][WN][2];
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < WN; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  const unsigned sgn = (g & 1) ? 0u : 0x80000000u;

  issue(0);
  for (int s = 0; s < nstage; s++) {
    issue(s + 1);
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    const int buf = s & 1;
    const double *a_re = As + ((size_t)(buf * W) * XBM + warp * 16 + g) * XPA;
    const double *a_im = a_re + (size_t)XBM * XPA;
    const double *bs = Bs + (size_t)buf * XKC * PB + g;
#pragma unroll
    for (int ks = 0; ks < XKC / 4; ks++) {
      const int k = ks * 4 + kl;
      double ar[2], ai[2];
#pragma unroll
      for (int a = 0; a < 2; a++) {
        ar[a] = a_re[a * 8 * XPA + k];
        if (CPLX) ai[a] = a_im[a * 8 * XPA + k];
      }
#pragma unroll
      for (int b = 0; b < WN; b++) {
        const double bf = bs[k * PB + b * 8];
        double bw = 0.0;
        if (CPLX) bw = flip(bs[k * PB + b * 8 + ((g & 1) ? -1 : 1)], sgn);
#pragma unroll
        for (int a = 0; a < 2; a++) {
          dmma(acc[a][b][0], acc[a][b][1], ar[a], bf);
          if (CPLX) dmma(acc[a][b][0], acc[a][b][1], ai[a], bw);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const long long q = q0 + warp * 16 + a * 8 + g;
    if (q >= p.nq) continue;
#pragma unroll
    for (int b = 0; b < WN; b++) {
      const int col = col0 + b * 8 + 2 * kl;
      double *dst = p.U + q * p.ncols_d + col;
      if (col < p.ncols_d) dst[0] = p.accumulate ? dst[0] + acc[a][b][0] : acc[a][b][0];
      if (col + 1 < p.ncols_d) dst[1] = p.accumulate ? dst[1] + acc[a][b][1] : acc[a][b][1];
    }
  }
}

template <int WN, bool CPLX> int launch_expand(const pspace_ctx *c, const ExpandParams &p, cudaStream_t st) {
  constexpr int BN = WN * 8, W = CPLX ? 2 : 1;
  const size_t smem = sizeof(double) * (2 * W * XBM * XPA + 2 * XKC * (BN + 4));
  PS_CUDA(cudaFuncSetAttribute(k_expand<WN, CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((p.nq + XBM - 1) / XBM), (unsigned)((p.ncols_d + BN - 1) / BN));
  k_expand<WN, CPLX><<<grid, XT, smem, st>>>(p);
  PS_LAUNCH_CHECK(c);
  return PSPACE_OK;
}
}  // namespace

int ps_build_panel_rows(pspace_ctx *c, int r0, int nr, long long q0, long long q1, int weighted, double **d_panel,
                        int *nr_pad, cudaStream_t st);   // project.cu

int ps_expand(const pspace_ctx *c, int k0, int k1, long long q0, long long q1, int ncols, int accumulate,
              const double *d_V, double *d_U, cudaStream_t st) {
  if (!c->tables_ready) return ps_fail(PSPACE_ESTATE, "evaluate_expansion called before initializeBasis + initializeQuadrature");
  if (k0 < 0 || k1 > c->N || k1 < k0 || q0 < 0 || q1 > c->Q || q1 < q0 || ncols < 1) return ps_fail(PSPACE_EINVAL, "range");
  if (q1 == q0) return PSPACE_OK;
  const int ncols_d = ncols * c->width;
  if (k1 == k0) {
    if (!accumulate) PS_CUDA(cudaMemsetAsync(d_U, 0, sizeof(double) * (size_t)(q1 - q0) * ncols_d, st));
    return PSPACE_OK;
  }
  double *panel = nullptr;
  int np = 0;
  int rc = ps_build_panel_rows(const_cast<pspace_ctx *>(c), k0, k1 - k0, q0, q1, 0, &panel, &np, st);
  if (rc) return rc;
  ExpandParams p;
  p.P = panel; p.V = d_V; p.U = d_U; p.nq = q1 - q0; p.np = np; p.nk = k1 - k0; p.ncols_d = ncols_d;
  p.accumulate = accumulate;
  p.vecV = (ncols_d % 2 == 0 && (reinterpret_cast<size_t>(d_V) % 16) == 0) ? 1 : 0;
  const bool z = c->is_complex;
  // warp tile width WN*8 columns: fewest column tiles first (every column tile streams the panel again), then the
  // fewest padded columns
  static const int cand[] = {8, 6, 4, 3, 2, 1};
  const int min_tiles = (ncols_d + 63) / 64;
  int wn = 8;
  for (int k = 0; k < 6; k++)
    if ((ncols_d + cand[k] * 8 - 1) / (cand[k] * 8) == min_tiles) wn = cand[k];
  switch (wn) {
    case 8: rc = z ? launch_expand<8, true>(c, p, st) : launch_expand<8, false>(c, p, st); break;
    case 6: rc = z ? launch_expand<6, true>(c, p, st) : launch_expand<6, false>(c, p, st); break;
    case 4: rc = z ? launch_expand<4, true>(c, p, st) : launch_expand<4, false>(c, p, st); break;
    case 3: rc = z ? launch_expand<3, true>(c, p, st) : launch_expand<3, false>(c, p, st); break;
    case 2: rc = z ? launch_expand<2, true>(c, p, st) : launch_expand<2, false>(c, p, st); break;
    default: rc = z ? launch_expand<1, true>(c, p, st) : launch_expand<1, false>(c, p, st); break;
  }
  return rc;
}
