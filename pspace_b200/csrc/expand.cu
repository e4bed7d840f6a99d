// (f1) Polynomial-chaos expansion evaluated at quadrature points — the step right before the projection in every
// caller:      U[q][c] = sum_k psi_k(z_q) V[k][c]
// Replaces getDeterministicStates / getDeterministicAdjoint of the reference caller
// (examples/stacs/cpp/TACSStochasticElement.cpp:53-120: for every (i,j,q) it recomputes u_q with N calls of
// pc->basis(k,zq) per node; pspace/core.py:788-792).  Here it is one (nq x N).(N x m) FP64 GEMM per call.
//
// k_expand_f (default; V TMA-addressable = even real column count, 16-byte aligned): factored operand, see factored.cuh.
//   psi_k(z_q) = P_k[q div L] * S_k[q mod L].  Persistent CTAs walk (256-point tile, column tile) pairs; the K loop runs
//   over chunks of KC basis rows through a TMA + mbarrier ring.  Per stage the producer warps write S (KC rows x L tail
//   points) and P (KC rows x the prefix blocks of the tile) into shared memory — products of the L1-resident 1-D tables in
//   the reference's order within each factor — and one elected thread fetches the V rows with TMA.  Consumer warps
//   (32 points x BN columns each): A-fragment element (row = point, k = basis row) = S * P, ONE DMUL feeding WN DMMA.8x8x4;
//   the accumulators are stored straight to U.  Nothing of size N x Q is ever written: algorithmic bytes 8*(Q*m + N*m),
//   2*N*m FLOP per point, bound by the FP64 tensor pipe.
// k_expand (operands TMA cannot address, or option force_panel): the round-1 kernel that streams the unweighted basis
//   panel written by k_psi_panel (project.cu); the q-range is walked in slabs so that the panel never exceeds 1 GiB.
//   * complex: U = P_re*[V] + P_im*[V'] with V'[k][2c] = -V[k][2c+1], V'[k][2c+1] = V[k][2c]  (as in project.cu).
#include "ctx.h"
#include "factored.cuh"
#include <algorithm>
#include <atomic>
#include <stdio.h>

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

// ======================================= factored-operand kernel ===========================================
using namespace psf;
constexpr int XFT = 384, XNCW = 8, XNPT = 128;

struct XFParams {
  double *U;
  const double *T;             // unweighted 1-D tables
  const int *deg;
  const uint8_t *qdig;
  long long q0, q1, Qtot, tiles_q;
  int k0, nk, ncols_d, nvars, DP, tsz;
  int toff[PSPACE_MAX_VARS], nq[PSPACE_MAX_VARS];
  int dsplit, L, npfx;
  int resident;                // 1: the tail table S of ALL basis rows (and their table offsets) stays in shared memory
  int tiles_n, nkc, stages, stage_bytes;
  int accumulate;
  unsigned zero;
};

template <int WM_, int WN_, bool CPLX_, int KC_> struct XTile {
  static constexpr int WM = WM_, WN = WN_;
  static constexpr bool CPLX = CPLX_;
  static constexpr int W = CPLX ? 2 : 1;
  static constexpr int BQ = XNCW * WM * 8, BN = WN * 8, LDB = BN + 4;
  static constexpr int KC = KC_;                   // basis rows per pipeline stage (128, 64 or 32)
  static constexpr int KCP = KC + 8;               // == 8 (mod 16): conflict-free 16-byte S reads (rows g, g+1 -> banks 0-15, 16-31)
  static constexpr int TPR = XNPT / KC;            // producer threads per basis row
  static constexpr int V_BYTES = KC * LDB * 8;
  static_assert(LDB % 16 == 4 || LDB % 16 == 12, "V pitch");
  // shared-memory header: table offsets of the rows, the 1-D tables, (resident mode) S of all rows
  __host__ __device__ static size_t eo_bytes(int nvars, int nkc, int res) {
    return ((size_t)(res ? nkc * KC : TPR * KC) * nvars * 2 + 15) & ~(size_t)15;
  }
  __host__ __device__ static size_t tab_bytes(int tsz) { return tsz <= psf::TAB_SMEM_MAX ? (((size_t)tsz * W * 8 + 15) & ~(size_t)15) : 0; }   // S_res behind it is read with 16-byte loads
  __host__ __device__ static size_t sres_bytes(int L, int nkc, int res) { return res ? (size_t)W * L * (nkc * KC + 8) * 8 : 0; }
  __host__ __device__ static size_t header_bytes(int nvars, int tsz, int L, int nkc, int res) {
    return (eo_bytes(nvars, nkc, res) + tab_bytes(tsz) + sres_bytes(L, nkc, res) + 127) / 128 * 128;
  }
  __host__ __device__ static int stage_bytes(int L, int npfx, int res) {
    return (V_BYTES + W * ((res ? 0 : L) + npfx) * KCP * 8 + 127) / 128 * 128;
  }
};

// Column of basis row r in the S / P tables: inside each group of 8 rows, rows kl and kl + 4 — the A-fragment elements one thread
// needs for two consecutive k-steps — sit side by side, so that one 16-byte load serves both k-steps.
__device__ __forceinline__ int xf_col(int r) { return (r & ~7) | ((r & 3) << 1) | ((r >> 2) & 1); }

template <class TL>
__global__ void __launch_bounds__(XFT, 1) k_expand_f(const __grid_constant__ XFParams p, const __grid_constant__ CUtensorMap tmapV) {
  constexpr int WM = TL::WM, WN = TL::WN, W = TL::W, BQ = TL::BQ, BN = TL::BN, LDB = TL::LDB, KC = TL::KC, KCP = TL::KCP;
  constexpr bool CPLX = TL::CPLX;
  constexpr bool PAIRED = CPLX && (WN % 2) == 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int RES = p.resident, NR = p.nkc * KC, NRP = NR + 8;
  unsigned short *eo_all = reinterpret_cast<unsigned short *>(smem);        // resident: [nvars][NR]; else [TPR][nvars][KC]
  double *Ts = reinterpret_cast<double *>(smem + TL::eo_bytes(p.nvars, p.nkc, RES));
  double *S_res = reinterpret_cast<double *>(smem + TL::eo_bytes(p.nvars, p.nkc, RES) + TL::tab_bytes(p.tsz));   // [W][L][NRP]
  unsigned char *ring = smem + TL::header_bytes(p.nvars, p.tsz, p.L, p.nkc, RES);
  const int STAGES = p.stages;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(ring + (size_t)STAGES * p.stage_bytes);
  unsigned long long *empty = full + STAGES;
  const int S_OFF = TL::V_BYTES, P_OFF = TL::V_BYTES + (RES ? 0 : W * p.L * KCP * 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ntiles = p.tiles_q * p.tiles_n;
  const bool tab_smem = p.tsz <= TAB_SMEM_MAX;
  if (tab_smem)
    for (int i = tid; i < p.tsz * W; i += XFT) Ts[i] = p.T[i];
  const double *T = tab_smem ? Ts : p.T;
  const int nq_tail = p.nq[p.nvars - 1], nq_pfx = p.dsplit > 0 ? p.nq[p.dsplit - 1] : 1;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1 + XNPT); mbar_init(&empty[s], XNCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (RES) {
    // once per CTA: table offsets and the tail table of every basis row (a thread owns a row; it reads the tables from
    // global memory here — the shared-memory copy is being written by other threads)
    for (int r = tid; r < NR; r += XFT) {
      const bool valid = r < p.nk;
      for (int d = 0; d < p.nvars; d++)
        eo_all[d * NR + r] = (unsigned short)(valid ? p.toff[d] + __ldg(p.deg + (size_t)(p.k0 + r) * p.nvars + d) * p.nq[d] : 0);
      run_products<CPLX>(p.T, eo_all + r, NR, p.qdig, p.DP, 0, p.L, 1, p.Qtot, p.dsplit, p.nvars, nq_tail, valid, S_res + xf_col(r), NRP, p.L);
    }
  }
  __syncthreads();

  if (warp >= XNCW) {
    // =========================== PRODUCERS ================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(56));
    const int pt = tid - XNCW * 32, r = pt % KC, sub = pt / KC;
    if (pt == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmapV) : "memory");
    // entries of a stage: npfx prefix values, then (streamed mode) L tail values per row; the TPR threads of a row split them
    const int ne = p.npfx + (RES ? 0 : p.L), per = (ne + TL::TPR - 1) / TL::TPR;
    const int e_lo = sub * per, e_hi = min(ne, e_lo + per);
    int s = 0;
    unsigned ph = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tn = (int)(tile % p.tiles_n);
      const long long tq = tile / p.tiles_n;
      const long long blk0 = (p.q0 + tq * BQ) / p.L;                  // prefix block of the tile's first point
      {
        // digit rows of the NEXT tile's prefix blocks into L1 ahead of use (one row per thread)
        const long long tnext = tile + gridDim.x;
        if (tnext < ntiles && pt < p.npfx) {
          const long long qn = ((p.q0 + (tnext / p.tiles_n) * BQ) / p.L + pt) * p.L;
          if (qn < p.Qtot) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p.qdig + (size_t)qn * p.DP));
        }
      }
      for (int kc = 0; kc < p.nkc; kc++) {
        mbar_wait(&empty[s], ph ^ 1u);
        unsigned char *st = ring + (size_t)s * p.stage_bytes;
        if (pt == 0) {
          mbar_arrive_expect_tx(&full[s], TL::V_BYTES);
          tma_load_2d(st, &tmapV, &full[s], tn * BN, kc * KC);
        }
        const int kr = kc * KC + r;
        const bool valid = kr < p.nk;
        const unsigned short *eo;
        int eos;
        if (RES) { eo = eo_all + kr; eos = NR; }
        else {
          unsigned short *mine = eo_all + (size_t)sub * p.nvars * KC + r;      // my own copy of the row's offsets
          if (valid)
            for (int d = 0; d < p.nvars; d++)
              mine[d * KC] = (unsigned short)(p.toff[d] + __ldg(p.deg + (size_t)(p.k0 + kr) * p.nvars + d) * p.nq[d]);
          eo = mine; eos = KC;
        }
        double *Sst = reinterpret_cast<double *>(st + S_OFF), *Pst = reinterpret_cast<double *>(st + P_OFF);
        if (e_lo < p.npfx) {
          const int n = min(e_hi, p.npfx) - e_lo;
          run_products<CPLX>(T, eo, eos, p.qdig, p.DP, blk0 + e_lo, n, p.L, p.Qtot, 0, p.dsplit, nq_pfx, valid,
                             Pst + e_lo * KCP + xf_col(r), KCP, p.npfx);
        }
        if (e_hi > p.npfx) {
          const int t0 = max(e_lo, p.npfx) - p.npfx, n = e_hi - p.npfx - t0;
          run_products<CPLX>(T, eo, eos, p.qdig, p.DP, t0, n, 1, p.Qtot, p.dsplit, p.nvars, nq_tail, valid,
                             Sst + t0 * KCP + xf_col(r), KCP, p.L);
        }
        mbar_arrive(&full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // =========================== CONSUMERS ================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(216));
    const int g = lane >> 2, kl = lane & 3;
    const unsigned sgn = (g & 1) ? 0u : 0x80000000u;
    const int spitch = RES ? NRP : KCP;
    auto acc_pair = [&](const double (&ac)[WM][WN][2], int a, int idx, double &v0, double &v1, int &cc) {
      if (PAIRED) { const int b2 = idx >> 1, e = idx & 1; v0 = ac[a][2 * b2][e]; v1 = ac[a][2 * b2 + (WN > 1 ? 1 : 0)][e]; cc = 16 * b2 + 4 * kl + 2 * e; }
      else { v0 = ac[a][idx][0]; v1 = ac[a][idx][1]; cc = 8 * idx + 2 * kl; }
    };
    int s = 0;
    unsigned ph = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long tq = p.tiles_n == 1 ? tile : tile / p.tiles_n;
      const int tn = (int)(tile - tq * p.tiles_n);
      const long long qt0 = p.q0 + tq * BQ, blk0 = qt0 / p.L;
      int soff[WM], poff[WM];
      {
        // 32-bit index arithmetic relative to the tile's first prefix block (one 64-bit division per tile, above)
        const int r0 = (int)(qt0 - blk0 * p.L);                      // tile's first point within its block
        const int rmax = (int)min((long long)BQ - 1, p.q1 - 1 - qt0);   // rows past the range are computed on a valid point and not stored
#pragma unroll
        for (int a = 0; a < WM; a++) {
          const int r = r0 + min(warp * (WM * 8) + a * 8 + g, rmax);
          const int b = r / p.L;
          soff[a] = (r - b * p.L) * spitch + 2 * kl;
          poff[a] = b * KCP + 2 * kl;
        }
      }
      double acc[WM][WN][2];
#pragma unroll
      for (int a = 0; a < WM; a++)
#pragma unroll
        for (int b = 0; b < WN; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

      for (int kc = 0; kc < p.nkc; kc++) {
        mbar_wait(&full[s], ph);
        const unsigned char *st = ring + (size_t)s * p.stage_bytes;
        const double *bs = reinterpret_cast<const double *>(st) + kl * LDB + g;
        const double *bs2 = reinterpret_cast<const double *>(st) + kl * LDB + 2 * g;
        const double *Sr = RES ? S_res + kc * KC : reinterpret_cast<const double *>(st + S_OFF);
        const double *Si = Sr + (size_t)p.L * spitch;
        const double *Pr = reinterpret_cast<const double *>(st + P_OFF), *Pi = Pr + (size_t)p.npfx * KCP;
        const int ksteps = (min(KC, p.nk - kc * KC) + 3) >> 2;
        int depi = 0;
        // one step = TWO k-steps.  Real scalars: the S and P elements of both come from one 16-byte load each (see xf_col);
        // the complex build (twice the fragment registers) reads them per k-step from the same layout.
        auto kpair = [&](const int kp) {
          double are[2][WM], aim[CPLX ? 2 : 1][CPLX ? WM : 1];
          if (!CPLX) {
#pragma unroll
            for (int a = 0; a < WM; a++) {
              const double2 sr = *reinterpret_cast<const double2 *>(Sr + soff[a] + kp * 8);
              const double2 pr = *reinterpret_cast<const double2 *>(Pr + poff[a] + kp * 8);
              are[0][a] = sr.x * pr.x; are[1][a] = sr.y * pr.y;
            }
          }
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int ks = 2 * kp + e;
            if (CPLX) {
#pragma unroll
              for (int a = 0; a < WM; a++) {
                const double sr = Sr[soff[a] + kp * 8 + e], pr = Pr[poff[a] + kp * 8 + e];
                const double si = Si[soff[a] + kp * 8 + e], pi = Pi[poff[a] + kp * 8 + e];
                are[e][a] = sr * pr - si * pi;
                aim[CPLX ? e : 0][CPLX ? a : 0] = sr * pi + si * pr;
              }
            }
            double bfr[WN], bsw[WN];
            if (PAIRED) {
#pragma unroll
              for (int b2 = 0; b2 < WN / 2; b2++) {
                const double2 v = *reinterpret_cast<const double2 *>(bs2 + ks * 4 * LDB + 16 * b2);
                bfr[2 * b2] = v.x; bfr[2 * b2 + 1] = v.y;
                if (CPLX) { bsw[2 * b2] = flip_sign(v.y, 0x80000000u); bsw[2 * b2 + 1] = v.x; }
              }
            } else {
#pragma unroll
              for (int b = 0; b < WN; b++) {
                bfr[b] = bs[ks * 4 * LDB + b * 8];
                if (CPLX) bsw[b] = flip_sign(bs[ks * 4 * LDB + b * 8 + ((g & 1) ? -1 : 1)], sgn);
              }
            }
#pragma unroll
            for (int a = 0; a < WM; a++)
#pragma unroll
              for (int b = 0; b < WN; b++) {
                dmma884(acc[a][b][0], acc[a][b][1], are[e][a], bfr[b]);
                if (CPLX) dmma884(acc[a][b][0], acc[a][b][1], aim[CPLX ? e : 0][CPLX ? a : 0], bsw[b]);
              }
            depi = __double2loint(are[e][WM - 1]) ^ __double2loint(bfr[WN - 1]);
          }
        };
        const int kpairs = (ksteps + 1) >> 1;          // an odd k-step count runs one more k-step of zero rows
        if (kpairs == KC / 8) {
#pragma unroll
          for (int kp = 0; kp < KC / 8; kp++) kpair(kp);      // full chunk: one straight-line block, loads hoisted across k-steps
        } else if (kpairs == KC / 16) {
#pragma unroll
          for (int kp = 0; kp < KC / 16; kp++) kpair(kp);     // half chunk, also straight-line (cfg5: 286 rows = 4 x 64 + 30)
        } else {
#pragma unroll 2
          for (int kp = 0; kp < kpairs; kp++) kpair(kp);
        }
        // release the stage: the arrive's address depends on the last fragment registers read from it
        const unsigned dep = (unsigned)depi & p.zero;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s], dep);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }

#pragma unroll
      for (int a = 0; a < WM; a++) {
        const long long qrel = tq * BQ + warp * (WM * 8) + a * 8 + g;
        if (p.q0 + qrel >= p.q1) continue;
#pragma unroll
        for (int b = 0; b < WN; b++) {
          double v0, v1; int cc;
          acc_pair(acc, a, b, v0, v1, cc);
          const int col = tn * BN + cc;
          double *dst = p.U + (size_t)qrel * p.ncols_d + col;
          if (col + 1 < p.ncols_d && ((reinterpret_cast<size_t>(dst) & 15) == 0)) {
            if (p.accumulate) { const double2 o = *reinterpret_cast<double2 *>(dst); v0 += o.x; v1 += o.y; }
            *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
          } else {
            if (col < p.ncols_d) dst[0] = p.accumulate ? dst[0] + v0 : v0;
            if (col + 1 < p.ncols_d) dst[1] = p.accumulate ? dst[1] + v1 : v1;
          }
        }
      }
    }
  }
}

// Tail length = a product of trailing rule sizes.  A (stage rows KC, tail length, resident?) candidate is rated by the table
// products a producer thread makes per stage — the prefix entries of a tile, plus, unless the tail table of all rows fits in
// shared memory next to two stages ("resident"), the L tail entries; an entry of a run costs ~1 + (variables - 1)/rule size
// products — relative to the consumers' k-steps per stage.
struct XFChoice { int ds, L, npfx, res; double score; };
template <class TL> XFChoice plan_expand_f(const pspace_ctx *c, int nk) {
  const int nkc = (nk + TL::KC - 1) / TL::KC;
  const size_t cap = (size_t)227 * 1024 - 1024 - 256;
  XFChoice best = {-1, 1, 0, 0, 0.0};
  long long L = 1;
  for (int ds = c->nvars; ds >= 0; ds--) {
    if (ds < c->nvars) L *= c->nq[ds];
    if (L > 4096) break;
    const int npfx = (int)((TL::BQ + L - 2) / L) + 1;
    const double c_pfx = npfx * (ds > 0 ? 1.0 + (double)(ds - 1) / c->nq[ds - 1] : 0.0);
    const double c_tail = (double)L * (ds < c->nvars ? 1.0 + (double)(c->nvars - ds - 1) / c->nq[c->nvars - 1] : 0.0);
    for (int res = 1; res >= 0; res--) {
      const size_t need = TL::header_bytes(c->nvars, c->tsz, (int)L, nkc, res) + 2 * (size_t)TL::stage_bytes((int)L, npfx, res);
      if (need > cap || (res && (size_t)nkc * TL::KC > 8192)) continue;
      const double score = (c_pfx + (res ? 0.0 : c_tail) + 4.0) / TL::TPR / (TL::KC / 4);
      if (best.ds < 0 || score < best.score) best = {ds, (int)L, npfx, res, score};
    }
  }
  return best;
}

template <class TL> int launch_expand_f(const pspace_ctx *c, XFParams &p, const double *d_V, cudaStream_t st) {
  p.nkc = (p.nk + TL::KC - 1) / TL::KC;
  const size_t cap = (size_t)227 * 1024 - 1024 - 256;
  const XFChoice ch = plan_expand_f<TL>(c, p.nk);
  if (ch.ds < 0) return ps_fail(PSPACE_EINVAL, "evaluate_expansion: no factorisation fits shared memory");
  p.dsplit = ch.ds; p.L = ch.L; p.npfx = ch.npfx; p.resident = ch.res;
  p.stage_bytes = TL::stage_bytes(p.L, p.npfx, p.resident);
  const size_t head = TL::header_bytes(c->nvars, c->tsz, p.L, p.nkc, p.resident);
  p.stages = (int)std::max<size_t>(2, std::min<size_t>(6, (cap - head) / p.stage_bytes));
  const size_t smem = head + (size_t)p.stages * p.stage_bytes + 2 * p.stages * 8 + 1024;
  p.tiles_q = (p.q1 - p.q0 + TL::BQ - 1) / TL::BQ;
  p.tiles_n = (p.ncols_d + TL::BN - 1) / TL::BN;
  CUtensorMap tmV;
  {
    unsigned long long dims[2] = {(unsigned long long)p.ncols_d, (unsigned long long)p.nk};
    unsigned long long str[1] = {(unsigned long long)p.ncols_d * 8};
    unsigned box[2] = {(unsigned)TL::LDB, (unsigned)TL::KC};
    int rc = ps_make_tmap(&tmV, d_V, 2, dims, str, box);
    if (rc) return rc;
  }
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  const int grid = (int)std::max<long long>(1, std::min<long long>(nsm, p.tiles_q * p.tiles_n));
  static std::atomic<bool> attr_dev[64];        // per instantiation and device; setting it twice is harmless
  if (!attr_dev[c->device & 63].load(std::memory_order_acquire)) {
    PS_CUDA(cudaFuncSetAttribute(k_expand_f<TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_dev[c->device & 63].store(true, std::memory_order_release);
  }
  k_expand_f<TL><<<grid, XFT, smem, st>>>(p, tmV);
  PS_LAUNCH_CHECK(c);
  pspace_ctx *mc = const_cast<pspace_ctx *>(c);
  char buf[256];
  snprintf(buf, sizeof buf, "%sexpand %dx%d factored operand: tail=%d vars (L=%d%s) prefix slots=%d KC=%d stages=%d ctas=%d", TL::CPLX ? "z " : "",
           TL::BQ, TL::BN, c->nvars - p.dsplit, p.L, p.resident ? ", tail table resident" : "", p.npfx, TL::KC, p.stages, grid);
  mc->plan = buf;
  mc->plan_flop = 2.0 * (double)p.nk * (double)p.ncols_d * (double)(p.q1 - p.q0) * (TL::CPLX ? 2.0 : 1.0);
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
  const bool z = c->is_complex;
  const bool tma_ok = (ncols_d % 2) == 0 && (reinterpret_cast<size_t>(d_V) % 16) == 0;
  if (tma_ok && !c->force_panel && !c->force_generic && c->tsz + PSPACE_MAX_NQ < 65536) {
    XFParams f;
    f.U = d_U; f.T = c->d_T; f.deg = c->d_deg; f.qdig = c->d_qdig;
    f.q0 = q0; f.q1 = q1; f.Qtot = c->Q;
    f.k0 = k0; f.nk = k1 - k0; f.ncols_d = ncols_d; f.nvars = c->nvars; f.DP = c->DP; f.tsz = c->tsz;
    for (int d = 0; d < PSPACE_MAX_VARS; d++) { f.toff[d] = d < c->nvars ? c->toff[d] : 0; f.nq[d] = d < c->nvars ? c->nq[d] : 1; }
    f.accumulate = accumulate; f.zero = 0;
    // column tile: fewest tiles first (each regenerates the operand), then the fewest padded columns
    static const int cand[] = {6, 4, 3, 2, 1};
    const int min_tiles = (ncols_d + 47) / 48;
    int wn = 6;
    for (int k = 0; k < 5; k++)
      if ((ncols_d + cand[k] * 8 - 1) / (cand[k] * 8) == min_tiles) wn = cand[k];
    if (z && wn == 3) wn = 4;       // complex tiles hold whole (re,im) pairs: even WN
    if (z && wn == 1) wn = 2;
    // stage rows KC: the candidate whose best factorisation loads the producers least (ties: the larger stage)
#define PS_XF(WM_, WN_, Z_)                                                                                   \
  {                                                                                                            \
    const XFChoice a128 = plan_expand_f<XTile<WM_, WN_, Z_, 128>>(c, f.nk), a64 = plan_expand_f<XTile<WM_, WN_, Z_, 64>>(c, f.nk), \
                   a32 = plan_expand_f<XTile<WM_, WN_, Z_, 32>>(c, f.nk);                                     \
    if (a128.ds >= 0 && (a64.ds < 0 || a128.score <= a64.score) && (a32.ds < 0 || a128.score <= a32.score))    \
      return launch_expand_f<XTile<WM_, WN_, Z_, 128>>(c, f, d_V, st);                                         \
    if (a64.ds >= 0 && (a32.ds < 0 || a64.score <= a32.score)) return launch_expand_f<XTile<WM_, WN_, Z_, 64>>(c, f, d_V, st); \
    return launch_expand_f<XTile<WM_, WN_, Z_, 32>>(c, f, d_V, st);                                            \
  }
    // narrow real tiles (<= 24 columns) take 512 points per CTA when that still gives every SM two tiles: the per-tile
    // prologue/epilogue, where all eight consumer warps leave the tensor pipe idle together, is paid half as often and a
    // V fragment feeds 8 row tiles instead of 4 (cfg5: 0.510 -> see profiles/r02_ops.md)
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
    const bool tall = !z && wn <= 3 && (q1 - q0) >= 2LL * nsm * 512;
    switch (wn) {
      case 6: if (z) PS_XF(4, 6, true) else PS_XF(4, 6, false)
      case 4: if (z) PS_XF(4, 4, true) else PS_XF(4, 4, false)
      case 3: if (tall) PS_XF(8, 3, false) else PS_XF(4, 3, false)
      case 2: if (z) PS_XF(4, 2, true) else if (tall) PS_XF(8, 2, false) else PS_XF(4, 2, false)
      default: if (tall) PS_XF(8, 1, false) else PS_XF(4, 1, false)
    }
#undef PS_XF
  }
  // panel path, in q-slabs of at most 1 GiB of panel
  const long long np_est = ((k1 - k0) + 1) & ~1;
  long long slab = ((1LL << 30) / ((long long)c->width * np_est * 8)) / XBM * XBM;
  if (slab < XBM) slab = XBM;
  static const int cand[] = {8, 6, 4, 3, 2, 1};
  const int min_tiles = (ncols_d + 63) / 64;
  int wn = 8;
  for (int k = 0; k < 6; k++)
    if ((ncols_d + cand[k] * 8 - 1) / (cand[k] * 8) == min_tiles) wn = cand[k];
  int rc = PSPACE_OK;
  for (long long qs = q0; qs < q1 && rc == PSPACE_OK; qs += slab) {
    const long long qe = std::min(q1, qs + slab);
    double *panel = nullptr;
    int np = 0;
    rc = ps_build_panel_rows(const_cast<pspace_ctx *>(c), k0, k1 - k0, qs, qe, 0, &panel, &np, st);
    if (rc) return rc;
    ExpandParams p;
    p.P = panel; p.V = d_V; p.U = d_U + (size_t)(qs - q0) * ncols_d; p.nq = qe - qs; p.np = np; p.nk = k1 - k0; p.ncols_d = ncols_d;
    p.accumulate = accumulate;
    p.vecV = tma_ok ? 1 : 0;
    switch (wn) {
      case 8: rc = z ? launch_expand<8, true>(c, p, st) : launch_expand<8, false>(c, p, st); break;
      case 6: rc = z ? launch_expand<6, true>(c, p, st) : launch_expand<6, false>(c, p, st); break;
      case 4: rc = z ? launch_expand<4, true>(c, p, st) : launch_expand<4, false>(c, p, st); break;
      case 3: rc = z ? launch_expand<3, true>(c, p, st) : launch_expand<3, false>(c, p, st); break;
      case 2: rc = z ? launch_expand<2, true>(c, p, st) : launch_expand<2, false>(c, p, st); break;
      default: rc = z ? launch_expand<1, true>(c, p, st) : launch_expand<1, false>(c, p, st); break;
    }
  }
  ps_panels_mark_used(const_cast<pspace_ctx *>(c), st);
  const_cast<pspace_ctx *>(c)->plan = "expand: basis panel streamed (round-1 kernel)";
  return rc;
}
