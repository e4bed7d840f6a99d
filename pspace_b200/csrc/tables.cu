// (2) 1-D Gauss rules + orthonormal 1-D basis tables, tensor grid, batched basis evaluation.
//
// Compiled with -fmad=false: every product and sum rounds once, as in the reference's x86-64 build, so
// the real-scalar tables, grids and basis values are bit-identical to the reference (Legendre d >= 5 goes
// through a double-double integer power instead of libm pow — see pspace_poly.cuh).
//
// Replaces (reference paths): {Normal,Uniform,Exponential}Parameter::quadrature (…Parameter.cpp:30-34) ->
// GaussianQuadrature::{hermite,legendre,laguerre}Quadrature (GaussianQuadrature.cpp:31-1483; the literal
// tables are data, embedded from gauss_tables.inc), QuadratureHelper::tensorProduct
// (QuadratureHelper.cpp:34-132), ParameterContainer::initializeQuadrature / quadrature / basis
// (ParameterContainer.cpp:116-192) and OrthogonalPolynomials::unit_* (OrthogonalPolynomials.cpp:42-92).
#include "ctx.h"
#include "pspace_poly.cuh"

#define PSPACE_GAUSS_TABLE(name) __device__ const double name[]
#include "gauss_tables.inc"

namespace {

struct TabParams {
  int nvars;
  int kinds[PSPACE_MAX_VARS], pmax[PSPACE_MAX_VARS], nq[PSPACE_MAX_VARS];
  int roff[PSPACE_MAX_VARS + 1], toff[PSPACE_MAX_VARS + 1];
  double p1[2 * PSPACE_MAX_VARS], p2[2 * PSPACE_MAX_VARS];
};

__device__ __forceinline__ void raw_rule(int kind, int n, int i, double &x, double &w) {
  int o = n * (n - 1) / 2 + i;
  if (kind == PSPACE_KIND_NORMAL) { x = PSPACE_HERMITE_X[o]; w = PSPACE_HERMITE_W[o]; }
  else if (kind == PSPACE_KIND_UNIFORM) { x = PSPACE_LEGENDRE_X[o]; w = PSPACE_LEGENDRE_W[o]; }
  else { x = PSPACE_LAGUERRE_X[o]; w = PSPACE_LAGUERRE_W[o]; }
}

// one block per parameter: rule (z,y,w)[nq] then T[alpha][iq] = unit_poly(z_iq, alpha), alpha <= pmax
template <typename S>
__global__ void k_rules_tables(TabParams p, double *zp, double *yp, double *wp, double *T, double *Tw, int with_tables) {
  typedef scalar_of<S> SO;
  const int d = blockIdx.x, n = p.nq[d], kind = p.kinds[d];
  __shared__ double zs[2 * PSPACE_MAX_NQ], ws[2 * PSPACE_MAX_NQ];
  S a = SO::load(p.p1, d), b = SO::load(p.p2, d);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double x, w;
    raw_rule(kind, n, i, x, w);
    S z, y, wo;
    ps_map_rule<S>(kind, a, b, x, w, z, y, wo);
    SO::store(zp, p.roff[d] + i, z);
    SO::store(yp, p.roff[d] + i, y);
    SO::store(wp, p.roff[d] + i, wo);
    SO::store(zs, i, z);
    SO::store(ws, i, wo);
  }
  __syncthreads();
  if (!with_tables) return;
  const int rows = p.pmax[d] + 1;
  for (int e = threadIdx.x; e < rows * n; e += blockDim.x) {
    int alpha = e / n, i = e % n;
    S v = ps_param_basis<S>(kind, SO::load(zs, i), alpha);
    SO::store(T, p.toff[d] + alpha * n + i, v);
    // weighted table for the factored one-pass kernels: prod_d Tw_d = w_q psi_k(z_q) up to rounding order
    SO::store(Tw, p.toff[d] + alpha * n + i, s_mul(v, SO::load(ws, i)));
  }
}

struct GridParams {
  int nvars, DP;
  int nq[PSPACE_MAX_VARS], roff[PSPACE_MAX_VARS + 1], toff[PSPACE_MAX_VARS + 1];
};

__device__ __forceinline__ void digits_of(const GridParams &g, long long q, int *dig) {
  for (int d = g.nvars - 1; d >= 0; d--) { dig[d] = (int)(q % g.nq[d]); q /= g.nq[d]; }
}

// digit table + weights for q in [0,Q):  W = ((w0*w1)*w2)...  left to right (QuadratureHelper.cpp:123)
template <typename S>
__global__ void k_grid_digits_weights(GridParams g, long long Q, const double *wp, uint8_t *qdig, double *W) {
  typedef scalar_of<S> SO;
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int dig[PSPACE_MAX_VARS];
  digits_of(g, q, dig);
  S w = SO::load(wp, g.roff[0] + dig[0]);
  for (int d = 1; d < g.nvars; d++) w = s_mul(w, SO::load(wp, g.roff[d] + dig[d]));
  SO::store(W, q, w);
  uint32_t packed[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int d = 0; d < g.nvars; d++) packed[d >> 2] |= (uint32_t)dig[d] << (8 * (d & 3));
  uint4 *dst = reinterpret_cast<uint4 *>(qdig + q * g.DP);
  dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  if (g.DP == 32) dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
}

// materialised grid columns: Z,Y [q1-q0][nvars] point-major, W [q1-q0]  (what quadrature(q,zq,yq) hands out).
// One thread per (point, variable) element so that consecutive threads write consecutive addresses.
template <typename S>
__global__ void k_tensor_grid(GridParams g, long long q0, long long q1, const double *zp, const double *yp,
                              const double *wp, const uint8_t *qdig, double *Z, double *Y, double *W) {
  typedef scalar_of<S> SO;
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = q1 - q0;
  if (e >= n * g.nvars) return;
  const long long r = e / g.nvars;
  const int d = (int)(e % g.nvars);
  const uint8_t *dg = qdig + (q0 + r) * g.DP;
  const int o = g.roff[d] + dg[d];
  if (Z) SO::store(Z, e, SO::load(zp, o));
  if (Y) SO::store(Y, e, SO::load(yp, o));
  if (W && d == 0) {
    S w = SO::load(wp, g.roff[0] + dg[0]);
    for (int k = 1; k < g.nvars; k++) w = s_mul(w, SO::load(wp, g.roff[k] + dg[k]));
    SO::store(W, r, w);
  }
}

// Batched basis evaluation at quadrature points: psi[k][q] = ((1*T_0[a_0][i_0])*T_1[a_1][i_1])*...
// (ParameterContainer.cpp:184-192 product order).  Block = 8 warps; the 1-D tables are staged in shared
// memory; a warp owns a tile of 32 consecutive points (coalesced 256 B stores along q) and walks the
// block's KB basis rows eight at a time (eight independent product chains; their per-variable table-row
// offsets are staged once per block, variable-major, so one chain step costs two broadcast LDS.128 + one
// digit read for eight table lookups).
constexpr int EB_WARPS = 8, EB_KB = 32;
template <typename S>
__global__ void __launch_bounds__(EB_WARPS * 32) k_eval_basis_grid(const __grid_constant__ GridParams g, int tsz,
                                                                   const double *T, const int *deg, int k0, int k1,
                                                                   long long q0, long long q1, const uint8_t *qdig,
                                                                   double *psi) {
  typedef scalar_of<S> SO;
  extern __shared__ __align__(16) double sm[];
  double *Ts = sm;                                            // tsz scalars
  int *rowoff = reinterpret_cast<int *>(sm + (((size_t)tsz * SO::width + 1) & ~(size_t)1));   // [nvars][EB_KB]
  for (int i = threadIdx.x; i < tsz * SO::width; i += blockDim.x) Ts[i] = T[i];
  const int kb0 = k0 + blockIdx.y * EB_KB;
  for (int e = threadIdx.x; e < EB_KB * g.nvars; e += blockDim.x) {
    const int d = e / EB_KB, kk = kb0 + e % EB_KB;
    rowoff[e] = (kk < k1) ? g.toff[d] + deg[(size_t)kk * g.nvars + d] * g.nq[d] : 0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nqr = q1 - q0;
  long long r = (blockIdx.x * (long long)EB_WARPS + warp) * 32 + lane;
  if (r >= nqr) return;
  // table offset of (variable d, my point's digit) in my own shared-memory slot: dynamic indexing without local memory
  unsigned short *mybase = reinterpret_cast<unsigned short *>(rowoff + EB_KB * g.nvars) + threadIdx.x * PSPACE_MAX_VARS;
  const uint8_t *dg = qdig + (q0 + r) * g.DP;
  for (int d = 0; d < g.nvars; d++) mybase[d] = (unsigned short)dg[d];
  const int nk = min(EB_KB, k1 - kb0);
  int kk = 0;
  for (; kk + 8 <= nk; kk += 8) {                      // eight independent product chains
    S v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = SO::from(1.0);
    for (int d = 0; d < g.nvars; d++) {
      const int dd = mybase[d];
      const int4 ra = *reinterpret_cast<const int4 *>(rowoff + d * EB_KB + kk);
      const int4 rb = *reinterpret_cast<const int4 *>(rowoff + d * EB_KB + kk + 4);
      v[0] = s_mul(v[0], SO::load(Ts, ra.x + dd));
      v[1] = s_mul(v[1], SO::load(Ts, ra.y + dd));
      v[2] = s_mul(v[2], SO::load(Ts, ra.z + dd));
      v[3] = s_mul(v[3], SO::load(Ts, ra.w + dd));
      v[4] = s_mul(v[4], SO::load(Ts, rb.x + dd));
      v[5] = s_mul(v[5], SO::load(Ts, rb.y + dd));
      v[6] = s_mul(v[6], SO::load(Ts, rb.z + dd));
      v[7] = s_mul(v[7], SO::load(Ts, rb.w + dd));
    }
#pragma unroll
    for (int c = 0; c < 8; c++) SO::store(psi, (size_t)(kb0 + kk + c - k0) * nqr + r, v[c]);
  }
  for (; kk < nk; kk++) {
    S v = SO::from(1.0);
    for (int d = 0; d < g.nvars; d++) v = s_mul(v, SO::load(Ts, rowoff[d * EB_KB + kk] + mybase[d]));
    SO::store(psi, (size_t)(kb0 + kk - k0) * nqr + r, v);
  }
}

// Arbitrary points z[npts][nvars]: per block PB points; the 1-D polynomial values of every (point,
// variable, degree) are computed once with the reference formulas into shared memory, then multiplied.
constexpr int EP_PB = 8;
template <typename S>
__global__ void __launch_bounds__(256) k_eval_basis_points(TabParams p, const int *deg, int k0, int k1,
                                                           long long npts, const double *z, double *psi) {
  typedef scalar_of<S> SO;
  extern __shared__ double sm[];   // vals[EP_PB][sum_d (pmax_d+1)] scalars
  int poff[PSPACE_MAX_VARS + 1];
  poff[0] = 0;
  for (int d = 0; d < p.nvars; d++) poff[d + 1] = poff[d] + p.pmax[d] + 1;
  const int per = poff[p.nvars];
  const long long pb0 = blockIdx.x * (long long)EP_PB;
  for (int e = threadIdx.x; e < EP_PB * per; e += blockDim.x) {
    int pp = e / per, c = e % per;
    long long pt = pb0 + pp;
    if (pt >= npts) continue;
    int d = 0;
    while (c >= poff[d + 1]) d++;
    int alpha = c - poff[d];
    S v = ps_param_basis<S>(p.kinds[d], SO::load(z, pt * p.nvars + d), alpha);
    SO::store(sm, e, v);
  }
  __syncthreads();
  const int nk = k1 - k0;
  for (int e = threadIdx.x; e < EP_PB * nk; e += blockDim.x) {
    int pp = e % EP_PB, kk = e / EP_PB;
    long long pt = pb0 + pp;
    if (pt >= npts) continue;
    S v = SO::from(1.0);
    for (int d = 0; d < p.nvars; d++)
      v = s_mul(v, SO::load(sm, pp * per + poff[d] + deg[(size_t)(k0 + kk) * p.nvars + d]));
    SO::store(psi, (size_t)kk * npts + pt, v);
  }
}

void fill_tab(const pspace_ctx *c, TabParams &p) {
  p.nvars = c->nvars;
  for (int d = 0; d < c->nvars; d++) { p.kinds[d] = c->kinds[d]; p.pmax[d] = c->pmax[d]; p.nq[d] = c->nq[d]; }
  for (int d = 0; d <= c->nvars; d++) { p.roff[d] = c->roff[d]; p.toff[d] = c->toff[d]; }
  for (int i = 0; i < 2 * c->nvars; i++) { p.p1[i] = c->p1[i]; p.p2[i] = c->p2[i]; }
}
void fill_grid(const pspace_ctx *c, GridParams &g) {
  g.nvars = c->nvars; g.DP = c->DP;
  for (int d = 0; d < c->nvars; d++) g.nq[d] = c->nq[d];
  for (int d = 0; d <= c->nvars; d++) { g.roff[d] = c->roff[d]; g.toff[d] = c->toff[d]; }
}
}  // namespace

// (re)build the 1-D rules and, when the basis is known too, the T tables
int ps_build_tables(pspace_ctx *c, cudaStream_t st) {
  if (!c->quad_ready) return PSPACE_OK;
  const int w = c->width;
  int sumnq = 0;
  for (int d = 0; d < c->nvars; d++) { c->roff[d] = sumnq; sumnq += c->nq[d]; }
  c->roff[c->nvars] = sumnq;
  c->tsz = 0;
  const bool with_tables = c->basis_ready;
  for (int d = 0; d < c->nvars; d++) { c->toff[d] = c->tsz; if (with_tables) c->tsz += (c->pmax[d] + 1) * c->nq[d]; }
  c->toff[c->nvars] = c->tsz;
  const size_t need_rules = sizeof(double) * w * sumnq, need_T = sizeof(double) * w * (c->tsz > 0 ? c->tsz : 1);
  if (need_rules > c->cap_rules) {
    if (c->d_zp) { PS_CUDA(cudaFree(c->d_zp)); PS_CUDA(cudaFree(c->d_yp)); PS_CUDA(cudaFree(c->d_wp)); c->d_zp = c->d_yp = c->d_wp = nullptr; }
    c->cap_rules = 0;
    PS_CUDA(cudaMalloc(&c->d_zp, need_rules));
    PS_CUDA(cudaMalloc(&c->d_yp, need_rules));
    PS_CUDA(cudaMalloc(&c->d_wp, need_rules));
    c->cap_rules = need_rules;
  }
  if (with_tables && need_T > c->cap_T) {
    if (c->d_T) { PS_CUDA(cudaFree(c->d_T)); PS_CUDA(cudaFree(c->d_Tw)); c->d_T = c->d_Tw = nullptr; }
    c->cap_T = 0;
    PS_CUDA(cudaMalloc(&c->d_T, need_T));
    PS_CUDA(cudaMalloc(&c->d_Tw, need_T));
    c->cap_T = need_T;
  }
  TabParams p;
  fill_tab(c, p);
  if (c->is_complex) k_rules_tables<cplx><<<c->nvars, 128, 0, st>>>(p, c->d_zp, c->d_yp, c->d_wp, c->d_T, c->d_Tw, with_tables);
  else k_rules_tables<double><<<c->nvars, 128, 0, st>>>(p, c->d_zp, c->d_yp, c->d_wp, c->d_T, c->d_Tw, with_tables);
  PS_LAUNCH_CHECK(c);
  c->tables_ready = with_tables;
  return PSPACE_OK;
}

// digit table + weights over the whole grid (device-resident: Q*DP bytes + Q scalars)
int ps_build_grid(pspace_ctx *c, cudaStream_t st) {
  const int w = c->width;
  c->DP = c->nvars <= 16 ? 16 : 32;
  const size_t need_dig = (size_t)c->Q * c->DP, need_W = sizeof(double) * w * (size_t)c->Q;
  if (need_dig > c->cap_qdig) {
    if (c->d_qdig) { PS_CUDA(cudaFree(c->d_qdig)); c->d_qdig = nullptr; }
    c->cap_qdig = 0;
    PS_CUDA(cudaMalloc(&c->d_qdig, need_dig));
    c->cap_qdig = need_dig;
  }
  if (need_W > c->cap_W) {
    if (c->d_W) { PS_CUDA(cudaFree(c->d_W)); c->d_W = nullptr; }
    c->cap_W = 0;
    PS_CUDA(cudaMalloc(&c->d_W, need_W));
    c->cap_W = need_W;
  }
  GridParams g;
  fill_grid(c, g);
  unsigned nb = (unsigned)((c->Q + 255) / 256);
  if (c->is_complex) k_grid_digits_weights<cplx><<<nb, 256, 0, st>>>(g, c->Q, c->d_wp, c->d_qdig, c->d_W);
  else k_grid_digits_weights<double><<<nb, 256, 0, st>>>(g, c->Q, c->d_wp, c->d_qdig, c->d_W);
  PS_LAUNCH_CHECK(c);
  return PSPACE_OK;
}

int ps_tensor_grid(const pspace_ctx *c, long long q0, long long q1, double *Z, double *Y, double *W, cudaStream_t st) {
  if (q1 <= q0) return PSPACE_OK;
  GridParams g;
  fill_grid(c, g);
  unsigned nb = (unsigned)(((q1 - q0) * c->nvars + 255) / 256);
  if (c->is_complex) k_tensor_grid<cplx><<<nb, 256, 0, st>>>(g, q0, q1, c->d_zp, c->d_yp, c->d_wp, c->d_qdig, Z, Y, W);
  else k_tensor_grid<double><<<nb, 256, 0, st>>>(g, q0, q1, c->d_zp, c->d_yp, c->d_wp, c->d_qdig, Z, Y, W);
  PS_LAUNCH_CHECK(c);
  return PSPACE_OK;
}

int ps_eval_basis_grid(const pspace_ctx *c, int k0, int k1, long long q0, long long q1, double *psi, cudaStream_t st) {
  if (k1 <= k0 || q1 <= q0) return PSPACE_OK;
  GridParams g;
  fill_grid(c, g);
  dim3 grid((unsigned)((q1 - q0 + EB_WARPS * 32 - 1) / (EB_WARPS * 32)), (unsigned)((k1 - k0 + EB_KB - 1) / EB_KB));
  size_t smem = sizeof(double) * (((size_t)c->width * c->tsz + 1) & ~(size_t)1) + sizeof(int) * EB_KB * c->nvars +
                sizeof(unsigned short) * EB_WARPS * 32 * PSPACE_MAX_VARS;
  if (c->is_complex) {
    PS_CUDA(cudaFuncSetAttribute(k_eval_basis_grid<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_eval_basis_grid<cplx><<<grid, EB_WARPS * 32, smem, st>>>(g, c->tsz, c->d_T, c->d_deg, k0, k1, q0, q1, c->d_qdig, psi);
  } else {
    PS_CUDA(cudaFuncSetAttribute(k_eval_basis_grid<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_eval_basis_grid<double><<<grid, EB_WARPS * 32, smem, st>>>(g, c->tsz, c->d_T, c->d_deg, k0, k1, q0, q1, c->d_qdig, psi);
  }
  PS_LAUNCH_CHECK(c);
  return PSPACE_OK;
}

int ps_eval_basis_points(const pspace_ctx *c, int k0, int k1, long long npts, const double *z, double *psi, cudaStream_t st) {
  if (k1 <= k0 || npts <= 0) return PSPACE_OK;
  TabParams p;
  fill_tab(c, p);
  int per = 0;
  for (int d = 0; d < c->nvars; d++) per += c->pmax[d] + 1;
  size_t smem = sizeof(double) * c->width * EP_PB * per;
  unsigned nb = (unsigned)((npts + EP_PB - 1) / EP_PB);
  if (c->is_complex) {
    PS_CUDA(cudaFuncSetAttribute(k_eval_basis_points<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_eval_basis_points<cplx><<<nb, 256, smem, st>>>(p, c->d_deg, k0, k1, npts, z, psi);
  } else {
    PS_CUDA(cudaFuncSetAttribute(k_eval_basis_points<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_eval_basis_points<double><<<nb, 256, smem, st>>>(p, c->d_deg, k0, k1, npts, z, psi);
  }
  PS_LAUNCH_CHECK(c);
  return PSPACE_OK;
}
