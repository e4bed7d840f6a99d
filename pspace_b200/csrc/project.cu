// (3) Galerkin projection — the hot path.
//
//   matrix:  C[i][j][c] = sum_q (w_q psi_i(z_q) psi_j(z_q)) J[q][c]      TACSStochasticElement.cpp:380-414
//   vector:  F[k][c]    = sum_q (w_q psi_k(z_q)) f[q][c]                 TACSStochasticElement.cpp:283-311
//
// i.e. the dense contraction (N^2 x Q).(Q x m^2) with a GENERATED left operand.  The reference evaluates it as
// N^2*Q scalar calls of basis()/quadrature(); here it is one FP64 tensor-core GEMM.  Common to both kernels below:
//
//  * math: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  Measured on this B200 (profiles/r01_fp64_peak.md): DMMA sustains
//    37.0 TFLOP/s = 99.5 % of 148 SM x 64 FMA/clk x 1.965 GHz, the DFMA pipe 33.5; both are served by the same FP64
//    datapath (mixing them adds nothing), DMMA needs 1/8 of the issue slots, so DMMA it is.  tcgen05 has no FP64 kind.
//  * a CTA tile is BI basis rows i x 8 basis rows j (= BM = 8*BI ordered pairs) x BN real columns; each lane forms its
//    A-fragment element a[row=(i,j)][k=q] = psi_i(q) * (w_q psi_j(q)) with ONE DMUL that feeds WN DMMAs — the pair
//    products (N^2 x Q) are never materialised.
//  * shared-memory tiles have a pitch == 4 (mod 16) doubles so the DMMA fragment reads (k = lane%4, n = lane/4) hit
//    16 distinct banks per half-warp.
//  * USE_COMPLEX: J, psi and C are interleaved (re,im); C = A_re*[B] + A_im*[B'] with B'[k][2c] = -B[k][2c+1],
//    B'[k][2c+1] = B[k][2c]  (4 real FMAs per complex FMA, naive complex multiply so complex-step imaginary parts survive).
//  * the q-range [q0,q1) is a call argument (multi-GPU shard / upload slab).
//
// k_project_ws (fast path, TMA-addressable operands): persistent, warp-specialised, TMA + mbarrier; psi comes from basis
//   panels written once per call by k_psi_panel; work = data-parallel full waves + stream-K remainder + k_fixup.  See the
//   comment block above it.
// k_project (generic path: odd column counts / unaligned operands): all 8 warps stage psi in-kernel per 32-point chunk as
//   products of the shared-memory 1-D tables T_d[alpha][i_d] selected by the mixed-radix digits of q (reference product
//   order, ParameterContainer.cpp:184-192), J streams in with 16/8-byte cp.async (LDGSTS) through a STAGES-deep ring;
//   the q-range is cut into `ksplit` slices over blockIdx.y whose partial windows k_reduce_slices adds in slice order.
#include "ctx.h"
#include <algorithm>
#include <atomic>
#include <cuda.h>
#include <stdio.h>

namespace {

int sm_count(int device);

constexpr int NT = 256;   // 8 warps
constexpr int KQ = 32;    // quadrature points per pipeline stage (8 DMMA k-steps)
constexpr int ZPAD = PSPACE_MAX_NQ;   // zero scalars appended to the staged tables (target of invalid rows)

enum { MODE_MATRIX = 0, MODE_VECTOR = 1 };

struct ProjParams {
  const double *in;        // J or f rows q0..q1-1, pitch ld_in doubles
  long long ld_in;
  double *out;             // C window / F rows, or the split workspace
  long long stride_i, stride_j, slice_stride;   // in doubles
  int i0, ni, j0, nj;      // matrix: window; vector: j0/nj = basis range, ni = 1
  long long q0, q1;
  int ncols_d;             // real columns (= ncols * width)
  int nvars, tsz, DP;
  const double *T;
  const int *deg;
  const uint8_t *qdig;
  const double *W;
  int toff[PSPACE_MAX_VARS], nq[PSPACE_MAX_VARS];
  int tiles_n, tiles_j, tiles_i;
  long long chunks_per_slice;
  int accumulate, vec16;
  int fast;                 // host side only: 1 = TMA stream-K kernel, 0 = single-launch cp.async kernel
  double *sk_ws;           // stream-K partial tiles: [grid][2][BM*BN]   (fast path only)
  const unsigned char *mask;   // optional basis-pair filter [N][N] (matrix mode): 0 = structurally zero block
  int mask_ld;                 // = N
  const int *tile_list;        // fast path with a mask: ids of the tiles that contain at least one kept pair ...
  const int *tile_count;       // ... and how many there are (device scalars: no host round trip)
  int symmetric;               // matrix mode, square window: compute pairs j >= i only and mirror the block to (j,i)
  // EPI_PEER: receive buffers of all ranks (peer-mapped), layout [src rank][owned pair tile][BM rows][ncols_pad]
  double *peer_recv[PSPACE_MAX_PEERS];
  int peer_world, peer_rank, peer_lp;      // ranks, my rank, pair tiles owned per rank (owner = pair tile / peer_lp)
  long long peer_ncols_pad;                // tiles_n * BN
  unsigned lag_ns;                         // RAG instantiation of k_project_ws: see there (0)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gptr, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gptr), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gptr, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gptr), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double x, unsigned mask) {   // x * (+-1) on the integer pipe
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

template <int WM_, int WN_, int WGM_, int WGN_, bool CPLX_, int MODE_, int STAGES_, int MINB_>
struct Tile {
  static constexpr int WM = WM_, WN = WN_, WGM = WGM_, WGN = WGN_, MODE = MODE_, STAGES = STAGES_, MINB = MINB_;
  static constexpr bool CPLX = CPLX_;
  static constexpr int W = CPLX ? 2 : 1;
  static constexpr int BM = WGM * WM * 8;          // rows (ordered pairs, or basis rows in vector mode)
  static constexpr int BN = WGN * WN * 8;          // real columns
  static constexpr int BI = BM / 8;                // matrix mode: i rows per tile (8 j rows)
  static constexpr int LDB = BN + 4;               // == 4 or 12 (mod 16): conflict-free B fragments
  static constexpr int BIP = BI + 2;               // i-side pitch (even)
  static constexpr int NJ = (MODE == MODE_MATRIX) ? 8 : BM;    // weighted-side rows
  static constexpr int PJP = NJ + 4;               // == 12 (mod 16) resp. 4 (mod 16)
  static constexpr int NE = (MODE == MODE_MATRIX) ? BI + 8 : BM;   // staged basis rows
  static_assert(WGM * WGN * 32 == NT, "8 warps");
  static_assert(BN % 8 == 0 && (LDB % 16 == 4 || LDB % 16 == 12), "B pitch");
  static size_t smem_bytes(int tsz, int nvars) {
    size_t d = (size_t)(tsz + ZPAD) * W + (size_t)STAGES * KQ * LDB + (size_t)2 * W * KQ * PJP;
    if (MODE == MODE_MATRIX) d += (size_t)2 * W * KQ * BIP;
    size_t b = d * 8 + 32 + (size_t)NE * nvars * 2 + (size_t)NT * 32;
    return (b + 15) & ~(size_t)15;
  }
};

template <class TL>
__global__ void __launch_bounds__(NT, TL::MINB) k_project(const __grid_constant__ ProjParams p) {
  constexpr int WM = TL::WM, WN = TL::WN, WGN = TL::WGN, W = TL::W, BM = TL::BM, BN = TL::BN, BI = TL::BI;
  constexpr int LDB = TL::LDB, BIP = TL::BIP, PJP = TL::PJP, NE = TL::NE, STAGES = TL::STAGES;
  constexpr bool CPLX = TL::CPLX;
  constexpr bool MAT = TL::MODE == MODE_MATRIX;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *Ts = reinterpret_cast<double *>(smem_raw);                 // (tsz + ZPAD) * W
  double *Bs = Ts + (size_t)(p.tsz + ZPAD) * W;                      // STAGES * KQ * LDB
  if ((((size_t)(p.tsz + ZPAD) * W) & 1) != 0) Bs += 1;              // keep 16-byte alignment (slack is in the size)
  double *Pj = Bs + (size_t)STAGES * KQ * LDB;                       // [2][W][KQ][PJP]
  double *Pi = Pj + (size_t)2 * W * KQ * PJP;                        // [2][W][KQ][BIP]   (matrix mode)
  unsigned short *eoff = reinterpret_cast<unsigned short *>(Pi + (MAT ? (size_t)2 * W * KQ * BIP : 0));  // [NE][nvars]
  unsigned char *digs = reinterpret_cast<unsigned char *>(                                              // [NT][32], 16-B aligned
      (reinterpret_cast<size_t>(eoff + (size_t)NE * p.nvars) + 15) & ~(size_t)15);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wgm = warp / WGN, wgn = warp % WGN;
  const int g = lane >> 2, kl = lane & 3;

  // ---- tile coordinates --------------------------------------------------------------------------
  int t = blockIdx.x;
  const int tn = t % p.tiles_n; t /= p.tiles_n;
  const int tj = t % p.tiles_j; t /= p.tiles_j;
  const int ti = t;
  const int ibase = p.i0 + ti * BI;                       // matrix: first i of the tile
  const int jbase = p.j0 + tj * TL::NJ;                   // first weighted-side basis row
  const int iend = p.i0 + p.ni, jend = p.j0 + p.nj;
  const int col0 = tn * BN;

  // ---- q-slice of this CTA -----------------------------------------------------------------------
  const long long nchunk_all = (p.q1 - p.q0 + KQ - 1) / KQ;
  const long long c_begin = blockIdx.y * p.chunks_per_slice;
  const long long c_end = min(nchunk_all, c_begin + p.chunks_per_slice);
  const int nchunk = (int)max(0LL, c_end - c_begin);

  // ---- one-time staging: 1-D tables and per-row table offsets ------------------------------------
  for (int i = tid; i < p.tsz * W; i += NT) Ts[i] = p.T[i];
  for (int i = tid; i < ZPAD * W; i += NT) Ts[p.tsz * W + i] = 0.0;
  for (int e = tid; e < NE * p.nvars; e += NT) {
    const int r = e / p.nvars, d = e % p.nvars;
    int k;
    bool ok;
    if (MAT) {
      if (r < BI) { k = ibase + r; ok = k < iend; }
      else { k = jbase + (r - BI); ok = k < jend; }
    } else { k = jbase + r; ok = k < jend; }
    // invalid rows point at the zero pad for variable 0 (product becomes 0), anywhere valid otherwise
    int off = ok ? p.toff[d] + p.deg[(size_t)k * p.nvars + d] * p.nq[d] : (d == 0 ? p.tsz : p.toff[d]);
    eoff[e] = (unsigned short)off;
  }

  // ---- producers ---------------------------------------------------------------------------------
  auto issue_B = [&](int c) {   // chunk c (relative to this slice) -> stage c % STAGES
    if (c < nchunk) {
      double *dst = Bs + (size_t)(c % STAGES) * KQ * LDB;
      const long long r0 = (c_begin + c) * KQ;            // row of J relative to q0
      const long long nrows = p.q1 - p.q0;
      if (p.vec16) {
        for (int idx = tid; idx < KQ * (BN / 2); idx += NT) {
          const int row = idx / (BN / 2), ch = idx % (BN / 2);
          const int col = col0 + ch * 2;
          int bytes = (p.ncols_d - col) * 8;
          bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
          if (r0 + row >= nrows) bytes = 0;
          const double *src = bytes ? p.in + (r0 + row) * p.ld_in + col : p.in;
          cp_async16(dst + row * LDB + ch * 2, src, bytes);
        }
      } else {
        for (int idx = tid; idx < KQ * BN; idx += NT) {
          const int row = idx / BN, ch = idx % BN;
          const int col = col0 + ch;
          int bytes = (col < p.ncols_d && r0 + row < nrows) ? 8 : 0;
          const double *src = bytes ? p.in + (r0 + row) * p.ld_in + col : p.in;
          cp_async8(dst + row * LDB + ch, src, bytes);
        }
      }
    }
    cp_async_commit();
  };

  // digits + weight of "my" quadrature point of a chunk (lane = point, warp = row group)
  uint4 dg0 = make_uint4(0, 0, 0, 0), dg1 = make_uint4(0, 0, 0, 0);
  double wre = 0.0, wim = 0.0;
  auto fetch_point = [&](int c) {
    if (c < nchunk) {
      long long q = p.q0 + (c_begin + c) * KQ + lane;
      const bool ok = q < p.q1;
      if (!ok) q = p.q1 - 1;
      const uint4 *src = reinterpret_cast<const uint4 *>(p.qdig + q * p.DP);
      dg0 = __ldg(src);
      if (p.DP == 32) dg1 = __ldg(src + 1);
      if (CPLX) { wre = ok ? __ldg(p.W + 2 * q) : 0.0; wim = ok ? __ldg(p.W + 2 * q + 1) : 0.0; }
      else wre = ok ? __ldg(p.W + q) : 0.0;
    }
  };

  unsigned char *mydig = digs + tid * 32;
  auto stage_psi = [&](int buf) {
    // publish my digits to my private smem slot (dynamic byte indexing without local memory)
    *reinterpret_cast<uint4 *>(mydig) = dg0;
    if (p.DP == 32) *reinterpret_cast<uint4 *>(mydig + 16) = dg1;
    double *pj = Pj + (size_t)buf * W * KQ * PJP;
    double *pi = Pi + (size_t)buf * W * KQ * BIP;
    for (int e = warp; e < NE; e += NT / 32) {
      const unsigned short *eo = eoff + e * p.nvars;
      double vr = 1.0, vi = 0.0;
      if (!CPLX) {
        for (int d = 0; d < p.nvars; d++) vr *= Ts[eo[d] + mydig[d]];
      } else {
        for (int d = 0; d < p.nvars; d++) {
          const int o = 2 * (eo[d] + mydig[d]);
          const double tr = Ts[o], tm = Ts[o + 1];
          const double nr = vr * tr - vi * tm, ni = vr * tm + vi * tr;
          vr = nr; vi = ni;
        }
      }
      const bool weighted = !MAT || e >= BI;
      if (weighted) {
        if (!CPLX) vr *= wre;
        else { const double nr = vr * wre - vi * wim, ni = vr * wim + vi * wre; vr = nr; vi = ni; }
        const int r = MAT ? e - BI : e;
        pj[lane * PJP + r] = vr;
        if (CPLX) pj[KQ * PJP + lane * PJP + r] = vi;
      } else {
        pi[lane * BIP + e] = vr;
        if (CPLX) pi[KQ * BIP + lane * BIP + e] = vi;
      }
    }
  };

  // ---- accumulators ------------------------------------------------------------------------------
  double acc[WM][WN][2];
#pragma unroll
  for (int a = 0; a < WM; a++)
#pragma unroll
    for (int b = 0; b < WN; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  const unsigned sgn = (g & 1) ? 0u : 0x80000000u;   // complex: even real-column (re part) takes -B_im

  // ---- pipeline ----------------------------------------------------------------------------------
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue_B(s);
  fetch_point(0);
  __syncthreads();   // tables + eoff visible

  for (int c = 0; c < nchunk; c++) {
    const int buf = c & 1;
    stage_psi(buf);
    fetch_point(c + 1);
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue_B(c + STAGES - 1);

    const double *bs = Bs + (size_t)(c % STAGES) * KQ * LDB + wgn * WN * 8 + g;
    const double *pj = Pj + (size_t)buf * W * KQ * PJP;
    const double *pi = Pi + (size_t)buf * W * KQ * BIP;
#pragma unroll
    for (int ks = 0; ks < KQ / 4; ks++) {
      const int k = ks * 4 + kl;
      double are[WM], aim[WM];
      if (MAT) {
        const double ajr = pj[k * PJP + g];
        const double aji = CPLX ? pj[KQ * PJP + k * PJP + g] : 0.0;
#pragma unroll
        for (int a = 0; a < WM; a++) {
          const double air = pi[k * BIP + wgm * WM + a];
          if (!CPLX) are[a] = air * ajr;
          else {
            const double aii = pi[KQ * BIP + k * BIP + wgm * WM + a];
            are[a] = air * ajr - aii * aji;
            aim[a] = air * aji + aii * ajr;
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < WM; a++) {
          are[a] = pj[k * PJP + (wgm * WM + a) * 8 + g];
          if (CPLX) aim[a] = pj[KQ * PJP + k * PJP + (wgm * WM + a) * 8 + g];
        }
      }
      double bfr[WN], bsw[WN];
#pragma unroll
      for (int b = 0; b < WN; b++) {
        bfr[b] = bs[k * LDB + b * 8];
        if (CPLX) bsw[b] = flip_sign(bs[k * LDB + b * 8 + ((g & 1) ? -1 : 1)], sgn);
      }
#pragma unroll
      for (int a = 0; a < WM; a++)
#pragma unroll
        for (int b = 0; b < WN; b++) {
          dmma884(acc[a][b][0], acc[a][b][1], are[a], bfr[b]);
          if (CPLX) dmma884(acc[a][b][0], acc[a][b][1], aim[a], bsw[b]);
        }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: lane owns rows g of each 8-row tile, columns 2*kl, 2*kl+1 of each 8-column tile ----
  double *outp = p.out + (long long)blockIdx.y * p.slice_stride;
#pragma unroll
  for (int a = 0; a < WM; a++) {
    long long rowoff;
    bool rok;
    bool keep = true;
    if (MAT) {
      const int i = ibase + wgm * WM + a, j = jbase + g;
      rok = i < iend && j < jend;
      rowoff = (long long)(i - p.i0) * p.stride_i + (long long)(j - p.j0) * p.stride_j;
      if (rok && p.mask) keep = p.mask[(size_t)i * p.mask_ld + j] != 0;
    } else {
      const int k = jbase + (wgm * WM + a) * 8 + g;
      rok = k < jend;
      rowoff = (long long)(k - p.j0) * p.stride_j;
    }
    if (!rok) continue;
    if (!keep && p.accumulate) continue;
#pragma unroll
    for (int b = 0; b < WN; b++) {
      const int col = col0 + wgn * WN * 8 + b * 8 + 2 * kl;
      double *dst = outp + rowoff + col;
      double v0 = keep ? acc[a][b][0] : 0.0, v1 = keep ? acc[a][b][1] : 0.0;
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


// =====================================================================================================
// Fast path: TMA + mbarrier, warp-specialised (used whenever the right operand is TMA-addressable, i.e. a
// 16-byte aligned base and an even number of real columns).
//
// Two kernels:
//  k_psi_panel   writes the basis panels of the call once:  Pi[part][q][i] = psi_i(z_q)   (i in the window)
//                                                           Pj[part][q][j] = w_q psi_j(z_q)
//                (q-major, basis row contiguous; part = re/im plane).  Products of the shared-memory 1-D
//                tables in the reference's order.  (ni+nj) x nq x 8 B — 4.8 GB for cfg5, written at HBM
//                speed in ~1 ms, i.e. 0.05 % of the projection; the pair products psi_i*psi_j (N^2 x Q) are
//                never materialised.
//  k_project_ws  warps 0-7 (two per SM sub-partition) = CONSUMERS (setmaxnreg.inc): wait full[s], build their
//                A fragments a[(i,j)][q] = Pi[q][i] * Pj[q][j] with ONE DMUL each, issue DMMA.8x8x4, release
//                empty[s].   warps 8-11 = PRODUCER warpgroup (setmaxnreg.dec): one elected thread arms full[s]
//                with expect_tx and issues THREE TMA tile loads per stage (cp.async.bulk.tensor): the J tile
//                (KQ rows x BN+4 columns) and the two panel tiles (KQ x BI+2, KQ x 12).  The over-fetched
//                columns are what gives each shared-memory tile its bank-conflict-free pitch with a dense
//                TMA box; out-of-range rows/columns are zero-filled by the TMA unit, which also handles the
//                ragged last chunk and ragged windows.  The producer executes no FP64 instruction, so the
//                FP64/DMMA pipe is shared only with the consumers' own fragment DMULs (an earlier version
//                that formed psi in the producer warps lost 10 % to DMUL-vs-DMMA pipe contention:
//                profiles/r01_project_ws_b).
// =====================================================================================================
constexpr int NT_WS = 384;
constexpr int NPROD = 128;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, unsigned long long *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, unsigned long long *bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <class TL> struct WsLayout {
  static constexpr int W = TL::W;
  // quadrature points per pipeline stage of the fast kernel: 64 for real scalars (16 k-steps = 384 DMMAs per warp between
  // two chunk boundaries, halving the per-chunk overhead of barrier wait + first fragment loads), 32 for complex (already
  // 384 DMMAs per chunk)
  static constexpr int KQW = TL::CPLX ? 32 : 64;
  static constexpr int B_BYTES = KQW * TL::LDB * 8;
  static constexpr int PJ_BYTES = W * KQW * TL::PJP * 8;
  static constexpr int PI_BYTES = (TL::MODE == MODE_MATRIX) ? W * KQW * TL::BIP * 8 : 0;
  // every TMA destination must be 128-byte aligned
  static constexpr int PJ_OFF = ((B_BYTES + 127) / 128) * 128;
  static constexpr int PI_OFF = ((PJ_OFF + PJ_BYTES + 127) / 128) * 128;
  static constexpr int STAGE_BYTES = ((PI_OFF + PI_BYTES + 127) / 128) * 128;
  static constexpr int FIT = (220 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 6 ? 6 : (FIT < 2 ? 2 : FIT);
  static constexpr unsigned TX_BYTES = B_BYTES + PJ_BYTES + PI_BYTES;
  static size_t smem_bytes() { return (size_t)STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024; }
};

struct PanelParams {
  int nvars, tsz, DP, width;
  const double *T;
  const int *deg;
  const uint8_t *qdig;
  const double *W;
  int toff[PSPACE_MAX_VARS], nq[PSPACE_MAX_VARS];
  long long q0, q1;
  int r0, nr, nr_pad;       // basis rows [r0, r0+nr), panel row pitch nr_pad (pad rows are written as zeros)
  int weighted;
  int dsplit, tail_stride;  // variables [dsplit, nvars) are the "tail": tail_stride = prod of their rule sizes
  int pts;                  // quadrature points per block (<= PSI_PTS; fewer for small grids, so that they still fill the machine)
  double *out;              // [width][q1-q0][nr_pad]
  double *out_w;            // not null: `out` gets the unweighted panel and out_w the weighted one (same rows) in one pass
};

// Basis panel P[part][q][row] = psi_row(z_q) (x w_q when weighted), product order ((1*T_0)*T_1)*... as everywhere.
// A thread owns one basis row (its per-variable table offsets live in registers) and walks PSI_PTS consecutive
// quadrature points of the block; the threads of a warp store 32 consecutive rows of one point (coalesced).  The
// last variable is the fastest index of the tensor grid (QuadratureHelper.cpp:34-132), so the partial product over
// the leading variables [0, dsplit) is the same for tail_stride consecutive points: it is kept in a register and only
// the tail factors are multiplied per point — same left-to-right order, hence bit-identical values, with
// (nvars - dsplit) + dsplit / tail_stride shared-memory table reads per value instead of nvars.
constexpr int PSI_PTS = 128, PSI_MAXT = 512, PSI_TAIL = 4;
template <bool CPLX>
__global__ void __launch_bounds__(PSI_MAXT) k_psi_panel(const __grid_constant__ PanelParams p) {
  constexpr int W = CPLX ? 2 : 1;
  extern __shared__ __align__(16) unsigned char psm[];
  double *Ts = reinterpret_cast<double *>(psm);                        // 1-D tables, tsz scalars
  double *Ws = Ts + ((p.tsz * W + 1) & ~1);                            // weights of the block's points
  uint8_t *Dg = reinterpret_cast<uint8_t *>(Ws + PSI_PTS * W);         // digit rows of the block's points [pts][DP]
  const long long nq = p.q1 - p.q0;
  const long long rb0 = blockIdx.x * (long long)p.pts;
  const int npts = (int)min((long long)p.pts, nq - rb0);
  for (int i = threadIdx.x; i < p.tsz * W; i += blockDim.x) Ts[i] = p.T[i];
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.qdig + (p.q0 + rb0) * p.DP);
    uint4 *dst = reinterpret_cast<uint4 *>(Dg);
    for (int i = threadIdx.x; i < npts * (p.DP / 16); i += blockDim.x) dst[i] = src[i];
    if (p.weighted || p.out_w)
      for (int i = threadIdx.x; i < npts * W; i += blockDim.x) Ws[i] = p.W[(p.q0 + rb0) * W + i];
  }
  __syncthreads();
  const int row = blockIdx.y * blockDim.x + threadIdx.x;
  if (row >= p.nr_pad) return;
  const bool live = row < p.nr;
  int ro[PSPACE_MAX_VARS];          // table offset of (variable d, this row's degree), digit 0 — statically indexed
  int tro[PSI_TAIL];                // the same for the tail variables
#pragma unroll
  for (int k = 0; k < PSI_TAIL; k++) tro[k] = 0;
  {
    const int *dk = p.deg + (size_t)(p.r0 + (live ? row : 0)) * p.nvars;
#pragma unroll
    for (int d = 0; d < PSPACE_MAX_VARS; d++) {
      ro[d] = 0;
      if (d < p.nvars) {
        ro[d] = p.toff[d] + dk[d] * p.nq[d];
#pragma unroll
        for (int k = 0; k < PSI_TAIL; k++)
          if (d == p.dsplit + k) tro[k] = ro[d];
      }
    }
  }
  const int ntail = p.nvars - p.dsplit;
  int cnt = (int)((p.q0 + rb0) % p.tail_stride);
  double pr = 1.0, pi = 0.0;
  double *out_re = p.out + (size_t)rb0 * p.nr_pad + row;
  double *out_im = p.out + (size_t)(nq + rb0) * p.nr_pad + row;
  for (int t = 0; t < npts; t++) {
    const uint8_t *dg = Dg + t * p.DP;
    if (t == 0 || cnt == 0) {       // block-uniform: a new prefix
      pr = 1.0; pi = 0.0;
#pragma unroll
      for (int d = 0; d < PSPACE_MAX_VARS; d++) {
        if (d < p.dsplit) {
          const int o = ro[d] + dg[d];
          if (!CPLX) pr *= Ts[o];
          else {
            const double tr = Ts[2 * o], tm = Ts[2 * o + 1];
            const double nr = pr * tr - pi * tm, ni = pr * tm + pi * tr;
            pr = nr; pi = ni;
          }
        }
      }
    }
    if (++cnt == p.tail_stride) cnt = 0;
    double vr = pr, vi = pi;
#pragma unroll
    for (int k = 0; k < PSI_TAIL; k++) {
      if (k < ntail) {
        const int o = tro[k] + dg[p.dsplit + k];
        if (!CPLX) vr *= Ts[o];
        else {
          const double tr = Ts[2 * o], tm = Ts[2 * o + 1];
          const double nr = vr * tr - vi * tm, ni = vr * tm + vi * tr;
          vr = nr; vi = ni;
        }
      }
    }
    if (!live) { vr = 0.0; vi = 0.0; }
    if (p.out_w) {                  // both panels of a square window: the unweighted value goes out before the weight is applied
      out_re[(size_t)t * p.nr_pad] = vr;
      if (CPLX) out_im[(size_t)t * p.nr_pad] = vi;
    }
    if (p.weighted || p.out_w) {
      if (!CPLX) vr *= Ws[t];
      else {
        const double wre = Ws[2 * t], wim = Ws[2 * t + 1];
        const double nr = vr * wre - vi * wim, ni = vr * wim + vi * wre;
        vr = nr; vi = ni;
      }
    }
    if (!live) { vr = 0.0; vi = 0.0; }
    double *o_re = p.out_w ? p.out_w + (out_re - p.out) : out_re;
    o_re[(size_t)t * p.nr_pad] = vr;
    if (CPLX) { double *o_im = p.out_w ? p.out_w + (out_im - p.out) : out_im; o_im[(size_t)t * p.nr_pad] = vi; }
  }
}

// EPI selects the epilogue / tile enumeration; separate instantiations so that the dense kernel's code generation is
// not perturbed by the other variants (a run-time mask branch in the epilogue cost 1.3 %):
//   EPI_DENSE  every tile of the window, result stored to C
//   EPI_MASKED a basis-pair filter or the i<->j symmetry: device tile list + masked / mirrored stores
//   EPI_PEER   multi-GPU fused reduce-scatter: the finished partial tile is PUSHED over NVLink into the receive slot of
//              the rank that owns the tile (peer-mapped symmetric memory), overlapping the exchange with the math of
//              the following tiles; k_peer_reduce_gather finishes the collective
enum { EPI_DENSE = 0, EPI_MASKED = 1, EPI_PEER = 2 };
// k-steps [0, nks) of one stage, rolled: the last, partly filled chunk of a tile in the RAG instantiation of k_project_ws.
// Same arithmetic, in the same order, as the unrolled loop of the kernel.
template <class TL>
__device__ __forceinline__ void ws_tail_ksteps(double (&acc)[TL::WM][TL::WN][2], const int nks, const double *bs, const double *bs2,
                                               const double *pj, const double *pi, const int wgm, const int g, const int kl,
                                               const unsigned sgn) {
  typedef WsLayout<TL> LY;
  constexpr int WM = TL::WM, WN = TL::WN, LDB = TL::LDB, BIP = TL::BIP, PJP = TL::PJP, KQ = LY::KQW;
  constexpr bool CPLX = TL::CPLX, MAT = TL::MODE == MODE_MATRIX, PAIRED = CPLX && (WN % 2) == 0;
#pragma unroll 1
  for (int ks = 0; ks < nks; ks++) {
    const int k = ks * 4 + kl;
    double are[WM], aim[WM];
    if (MAT) {
      const double ajr = pj[k * PJP + g];
      const double aji = CPLX ? pj[KQ * PJP + k * PJP + g] : 0.0;
#pragma unroll
      for (int a = 0; a < WM; a++) {
        const double air = pi[k * BIP + wgm * WM + a];
        if (!CPLX) are[a] = air * ajr;
        else {
          const double aii = pi[KQ * BIP + k * BIP + wgm * WM + a];
          are[a] = air * ajr - aii * aji;
          aim[a] = air * aji + aii * ajr;
        }
      }
    } else {
#pragma unroll
      for (int a = 0; a < WM; a++) {
        are[a] = pj[k * PJP + (wgm * WM + a) * 8 + g];
        if (CPLX) aim[a] = pj[KQ * PJP + k * PJP + (wgm * WM + a) * 8 + g];
      }
    }
    double bfr[WN], bsw[WN];
    if (PAIRED) {
#pragma unroll
      for (int b2 = 0; b2 < WN / 2; b2++) {
        const double2 v = *reinterpret_cast<const double2 *>(bs2 + k * LDB + 16 * b2);
        bfr[2 * b2] = v.x; bfr[2 * b2 + 1] = v.y;
        if (CPLX) { bsw[2 * b2] = flip_sign(v.y, 0x80000000u); bsw[2 * b2 + 1] = v.x; }
      }
    } else {
#pragma unroll
      for (int b = 0; b < WN; b++) {
        bfr[b] = bs[k * LDB + b * 8];
        if (CPLX) bsw[b] = flip_sign(bs[k * LDB + b * 8 + ((g & 1) ? -1 : 1)], sgn);
      }
    }
#pragma unroll
    for (int a = 0; a < WM; a++)
#pragma unroll
      for (int b = 0; b < WN; b++) {
        dmma884(acc[a][b][0], acc[a][b][1], are[a], bfr[b]);
        if (CPLX) dmma884(acc[a][b][0], acc[a][b][1], aim[a], bsw[b]);
      }
  }
}

template <class TL, int EPI, bool RAG = false>
__global__ void __launch_bounds__(NT_WS, 1) k_project_ws(const __grid_constant__ ProjParams p,
                                                         const __grid_constant__ CUtensorMap tmapB,
                                                         const __grid_constant__ CUtensorMap tmapI,
                                                         const __grid_constant__ CUtensorMap tmapJ) {
  typedef WsLayout<TL> LY;
  constexpr int WM = TL::WM, WN = TL::WN, WGN = TL::WGN, W = TL::W, BN = TL::BN, BI = TL::BI;
  constexpr int LDB = TL::LDB, BIP = TL::BIP, PJP = TL::PJP, STAGES = LY::STAGES;
  constexpr bool CPLX = TL::CPLX;
  constexpr bool MAT = TL::MODE == MODE_MATRIX;
  constexpr int KQ = LY::KQW;          // shadows the generic kernel's chunk size

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *stage_base = smem_raw;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)STAGES * LY::STAGE_BYTES);
  unsigned long long *empty = full + STAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int iend = p.i0 + p.ni, jend = p.j0 + p.nj;

  // ---- work decomposition: data-parallel full waves + stream-K remainder --------------------------------------
  // Tiles are numbered tn fastest, then tj, then ti.  With G = gridDim.x persistent CTAs, CTA g owns whole tiles
  // g, g+G, ... of the first floor(ntiles/G) waves (all q, result stored straight to C), and a contiguous range
  // [r0,r1) of the remaining rem_tiles*nchunk (tile, chunk) units.  That range is shorter than one tile's
  // nchunk, so it touches at most two tiles; those partial accumulators go to sk_ws[g][0..1] and k_fixup adds
  // the pieces of each remainder tile in CTA order (deterministic).  No wave-quantisation tail.
  const int G = gridDim.x, gcta = blockIdx.x;
  const int nchunk = (int)((p.q1 - p.q0 + KQ - 1) / KQ);
  constexpr bool MASKED = EPI == EPI_MASKED;
  const int ntiles = MASKED ? *p.tile_count : p.tiles_i * p.tiles_j * p.tiles_n;
  const int nfull = ntiles / G;
  const long long rem_units = (long long)(ntiles - nfull * G) * nchunk;
  const long long r0 = rem_units * gcta / G, r1 = rem_units * (gcta + 1) / G;
  const int nseg = nfull + (r1 > r0 ? 1 : 0) + ((r1 > r0 && r1 > (r0 / nchunk + 1) * (long long)nchunk) ? 1 : 0);
  auto segment = [&](int sidx, int &tile, int &c0, int &c1, int &slot) {
    if (sidx < nfull) { tile = sidx * G + gcta; if (MASKED) tile = p.tile_list[tile]; c0 = 0; c1 = nchunk; slot = -1; return; }
    const long long ta = r0 / nchunk;
    if (sidx == nfull) {
      tile = nfull * G + (int)ta; c0 = (int)(r0 - ta * nchunk); c1 = (int)min((long long)nchunk, r1 - ta * nchunk); slot = 0;
    } else {
      tile = nfull * G + (int)ta + 1; c0 = 0; c1 = (int)(r1 - (ta + 1) * nchunk); slot = 1;
    }
    if (MASKED) tile = p.tile_list[tile];           // position in the compacted list -> tile id
  };

  constexpr int NCW = (NT_WS - NPROD) / 32;     // consumer warps 0..7, producer warpgroup = warps 8..11
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCW) {
    // =========================== PRODUCER warpgroup: TMA only =========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(40));
    if (tid == NCW * 32) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmapB) : "memory");
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmapJ) : "memory");
      if (MAT) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmapI) : "memory");
      int it = 0;                                     // running chunk counter across segments -> stage / phase
      for (int sidx = 0; sidx < nseg; sidx++) {
        int tile, c0, c1, slot;
        segment(sidx, tile, c0, c1, slot);
        const int tn = tile % p.tiles_n, tj = (tile / p.tiles_n) % p.tiles_j, ti = tile / (p.tiles_n * p.tiles_j);
        for (int c = c0; c < c1; c++, it++) {
          const int s = it % STAGES;
          const unsigned n = (unsigned)(it / STAGES);
          mbar_wait(&empty[s], (n & 1u) ^ 1u);
          // Stage-release ordering BY CONSTRUCTION.  A consumer warp arrives on empty[s] when it has ISSUED the last
          // fragment load of the chunk in stage s — ptxas places the arrive right there, ahead of the DMMAs that consume
          // the load — so that arrive alone does not prove the load has returned.  The refill therefore also waits for
          // the release of the FOLLOWING chunk (it - STAGES + 1): a warp reaches that arrive only after it has left the
          // spin-wait on the following chunk's `full` barrier, a branch that program order places after every DMMA of
          // the chunk in stage s; a DMMA cannot issue before its operand registers are written, so by then every fragment
          // load from stage s has returned.  Cost: the ring prefetches STAGES - 2 chunks ahead instead of STAGES - 1
          // (measured: none), and the consumer loop — whose ptxas schedule is fragile — is untouched.
          if (it >= STAGES) {
            const int j = it - STAGES + 1;
            mbar_wait(&empty[j % STAGES], (unsigned)(j / STAGES) & 1u);
          }
          unsigned char *st = stage_base + (size_t)s * LY::STAGE_BYTES;
          const int row = c * KQ;                         // quadrature row relative to q0
          mbar_arrive_expect_tx(&full[s], LY::TX_BYTES);
          tma_load_2d(st, &tmapB, &full[s], tn * BN, row);
          tma_load_3d(st + LY::PJ_OFF, &tmapJ, &full[s], tj * TL::NJ, row, 0);
          if (MAT) tma_load_3d(st + LY::PI_OFF, &tmapI, &full[s], ti * BI, row, 0);
        }
      }
    }
  } else {
    // =========================== CONSUMER warpgroups ==================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(232));
    const int cw = warp;
    const int wgm = cw / WGN, wgn = cw % WGN;
    const int g = lane >> 2, kl = lane & 3;
    const unsigned sgn = (g & 1) ? 0u : 0x80000000u;
    constexpr bool PAIRED = CPLX && (WN % 2) == 0;   // measured: +2.5 % on the complex build (the neighbour column comes for free), -0.7 % on the real dense kernel
    // accumulator pair `idx` (0..WN-1) of row tile a: two values in adjacent physical columns, first column cc (relative
    // to the warp's column span)
    auto acc_pair = [&](const double (&ac)[WM][WN][2], int a, int idx, double &v0, double &v1, int &cc) {
      if (PAIRED) { const int b2 = idx >> 1, e = idx & 1; v0 = ac[a][2 * b2][e]; v1 = ac[a][2 * b2 + (WN > 1 ? 1 : 0)][e]; cc = 16 * b2 + 4 * kl + 2 * e; }
      else { v0 = ac[a][idx][0]; v1 = ac[a][idx][1]; cc = 8 * idx + 2 * kl; }
    };
    const int nks_last = (int)((p.q1 - p.q0 - (long long)(nchunk - 1) * KQ + 3) / 4);   // k-steps of a tile's last chunk that hold points
    int it = 0;
    for (int sidx = 0; sidx < nseg; sidx++) {
    int tile, c0, c1, slot;
    segment(sidx, tile, c0, c1, slot);
    const int tn = tile % p.tiles_n, tj = (tile / p.tiles_n) % p.tiles_j, ti = tile / (p.tiles_n * p.tiles_j);
    const int ibase = p.i0 + ti * BI, jbase = p.j0 + tj * TL::NJ, col0 = tn * BN;
    double acc[WM][WN][2];
#pragma unroll
    for (int a = 0; a < WM; a++)
#pragma unroll
      for (int b = 0; b < WN; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    // RAG and EPI_PEER instantiations only: optional start-of-segment delay of consumer warps 4-7 (p.lag_ns, 0 = none: the default).
    // Measured: no gain from the delay itself, but with this statement in place ptxas gives the RAG kernel's
    // main loop the 6257-cycle schedule instead of the 6338-cycle one (cfg3 71.1 -> 70.5 ms per element), and the fused
    // multi-GPU kernel 6254 instead of 6308; the dense RAG = false kernel and the masked one (which it would slow: 6388) do
    // not contain it and keep their code.  tests/test_host_cpu.py guards both counts.
    if ((RAG || EPI == EPI_PEER) && p.lag_ns != 0 && cw >= NCW / 2 && c1 - c0 >= 4) asm volatile("nanosleep.u32 %0;\n" ::"r"(p.lag_ns));

    const int c1m = (RAG && c1 == nchunk) ? c1 - 1 : c1;     // RAG: the tile's last chunk is handled after the loop
    for (int c = c0; c < c1m; c++, it++) {
      const int s = it % STAGES;
      const unsigned n = (unsigned)(it / STAGES);
      mbar_wait(&full[s], n & 1u);
      const unsigned char *st = stage_base + (size_t)s * LY::STAGE_BYTES;
      const double *bs = reinterpret_cast<const double *>(st) + wgn * WN * 8 + g;
      const double *bs2 = reinterpret_cast<const double *>(st) + wgn * WN * 8 + 2 * g;
      const double *pj = reinterpret_cast<const double *>(st + LY::PJ_OFF);
      const double *pi = reinterpret_cast<const double *>(st + LY::PI_OFF);
      // RAG instantiation (launched when the point count is not a multiple of the stage): the last chunk of a tile runs
      // only the k-steps that hold points — the rows TMA zero-filled add exact zeros, so the result keeps its bits
      // (cfg3: 1296 points = 20.25 stages, 3.6 % of the DMMAs were such zeros).  The tail is a separate rolled loop
      // (ws_tail_ksteps) after the chunk loop so that the unrolled loop below, whose ptxas schedule is fragile, stays as it is.
#pragma unroll
      for (int ks = 0; ks < KQ / 4; ks++) {
        const int k = ks * 4 + kl;
        double are[WM], aim[WM];
        if (MAT) {
          const double ajr = pj[k * PJP + g];
          const double aji = CPLX ? pj[KQ * PJP + k * PJP + g] : 0.0;
#pragma unroll
          for (int a = 0; a < WM; a++) {
            const double air = pi[k * BIP + wgm * WM + a];
            if (!CPLX) are[a] = air * ajr;
            else {
              const double aii = pi[KQ * BIP + k * BIP + wgm * WM + a];
              are[a] = air * ajr - aii * aji;
              aim[a] = air * aji + aii * ajr;
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < WM; a++) {
            are[a] = pj[k * PJP + (wgm * WM + a) * 8 + g];
            if (CPLX) aim[a] = pj[KQ * PJP + k * PJP + (wgm * WM + a) * 8 + g];
          }
        }
        double bfr[WN], bsw[WN];
        if (PAIRED) {
          // one 16-byte load = physical columns (2g, 2g+1) of a 16-column group: the even one feeds column tile 2*b2, the
          // odd one tile 2*b2+1 (column tiles are just a labelling of the columns; the epilogue undoes it).  Halves the
          // B-fragment loads; in the complex build the neighbour column is the other half of the same load.
#pragma unroll
          for (int b2 = 0; b2 < WN / 2; b2++) {
            const double2 v = *reinterpret_cast<const double2 *>(bs2 + k * LDB + 16 * b2);
            bfr[2 * b2] = v.x; bfr[2 * b2 + 1] = v.y;
            if (CPLX) { bsw[2 * b2] = flip_sign(v.y, 0x80000000u); bsw[2 * b2 + 1] = v.x; }
          }
        } else {
#pragma unroll
          for (int b = 0; b < WN; b++) {
            bfr[b] = bs[k * LDB + b * 8];
            if (CPLX) bsw[b] = flip_sign(bs[k * LDB + b * 8 + ((g & 1) ? -1 : 1)], sgn);
          }
        }
#pragma unroll
        for (int a = 0; a < WM; a++)
#pragma unroll
          for (int b = 0; b < WN; b++) {
            dmma884(acc[a][b][0], acc[a][b][1], are[a], bfr[b]);
            if (CPLX) dmma884(acc[a][b][0], acc[a][b][1], aim[a], bsw[b]);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);      // see the producer: a stage is refilled one release LATER than this one
    }
    if (RAG && c1m < c1) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (unsigned)(it / STAGES) & 1u);
      const unsigned char *st = stage_base + (size_t)s * LY::STAGE_BYTES;
      ws_tail_ksteps<TL>(acc, nks_last, reinterpret_cast<const double *>(st) + wgn * WN * 8 + g,
                         reinterpret_cast<const double *>(st) + wgn * WN * 8 + 2 * g, reinterpret_cast<const double *>(st + LY::PJ_OFF),
                         reinterpret_cast<const double *>(st + LY::PI_OFF), wgm, g, kl, sgn);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      it++;
    }

    // ---- epilogue of this segment ----------------------------------------------------------------
    // All 8 consumer warps meet here before any of them starts storing.  Reason: a stage is refilled by TMA once all
    // warps have arrived on empty[s], and ptxas issues that arrive right after the last fragment LDS is *issued*.  If a
    // warp that is still in the tile's last chunks had its LDS queued behind the epilogue stores of warps that are
    // already done (seen with the slow remote NVLink stores of EPI_PEER: 1e-10 glitches in that warp's last column
    // tile), the arrive and the refill could overtake the load.  With this barrier no store is in flight while any warp
    // of the CTA still reads fragments of the tile; after it, the last-arriving warp of the next tile's first chunks
    // has completed its own stores (EPI_PEER: __threadfence_system below), so the load/store queue is empty again.
    asm volatile("bar.sync 1, %0;\n" ::"n"(NT_WS - NPROD) : "memory");
    if (EPI == EPI_PEER && slot < 0) {
      // whole tile finished on this rank: push the partial block into the owner's receive slot (NVLink stores)
      const int pt = tile / p.tiles_n;
      const int owner = min(p.peer_world - 1, pt / p.peer_lp), lp = pt - owner * p.peer_lp;
      double *base = p.peer_recv[owner] + ((size_t)(p.peer_rank * p.peer_lp + lp) * TL::BM) * p.peer_ncols_pad + (size_t)tn * BN;
#pragma unroll
      for (int a = 0; a < WM; a++)
#pragma unroll
        for (int b = 0; b < WN; b++) {
          double v0, v1; int cw_;
          acc_pair(acc, a, b, v0, v1, cw_);
          const int r = (wgm * WM + a) * 8 + g, cc = wgn * WN * 8 + cw_;
          *reinterpret_cast<double2 *>(base + (size_t)r * p.peer_ncols_pad + cc) = make_double2(v0, v1);
        }
      __threadfence_system();      // peer stores must have landed at the owner before this kernel is seen as complete
    } else if (slot >= 0) {
      // partial tile -> stream-K workspace, dense [BM][BN], no guards (k_fixup applies them)
      double *wsp = p.sk_ws + ((size_t)gcta * 2 + slot) * (TL::BM * BN);
#pragma unroll
      for (int a = 0; a < WM; a++)
#pragma unroll
        for (int b = 0; b < WN; b++) {
          double v0, v1; int cw_;
          acc_pair(acc, a, b, v0, v1, cw_);
          const int r = (wgm * WM + a) * 8 + g, cc = wgn * WN * 8 + cw_;
          *reinterpret_cast<double2 *>(wsp + (size_t)r * BN + cc) = make_double2(v0, v1);
        }
    } else {
#pragma unroll
    for (int a = 0; a < WM; a++) {
      long long rowoff, mirroff = -1;
      bool rok;
      bool keep = true;
      if (MAT) {
        const int i = ibase + wgm * WM + a, j = jbase + g;
        rok = i < iend && j < jend;
        rowoff = (long long)(i - p.i0) * p.stride_i + (long long)(j - p.j0) * p.stride_j;
        if (MASKED && rok) keep = p.symmetric ? (j >= i) : (p.mask[(size_t)i * p.mask_ld + j] != 0);
        if (MASKED && p.symmetric && rok && j > i)      // mirrored block (j,i) of the same window
          mirroff = (long long)(j - p.i0) * p.stride_i + (long long)(i - p.j0) * p.stride_j;
      } else {
        const int k = jbase + (wgm * WM + a) * 8 + g;
        rok = k < jend;
        rowoff = (long long)(k - p.j0) * p.stride_j;
      }
      if (!rok) continue;
      if (MASKED && !keep && (p.accumulate || p.symmetric)) continue;   // filtered pair: nothing to add / written by its mirror
#pragma unroll
      for (int b = 0; b < WN; b++) {
        double a0, a1; int cw_;
        acc_pair(acc, a, b, a0, a1, cw_);
        const int col = col0 + wgn * WN * 8 + cw_;
        double *dst = p.out + rowoff + col;
        double v0 = (!MASKED || keep) ? a0 : 0.0, v1 = (!MASKED || keep) ? a1 : 0.0;
        if (col + 1 < p.ncols_d && ((reinterpret_cast<size_t>(dst) & 15) == 0)) {
          if (p.accumulate) { const double2 o = *reinterpret_cast<double2 *>(dst); v0 += o.x; v1 += o.y; }
          *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
        } else {
          if (col < p.ncols_d) dst[0] = p.accumulate ? dst[0] + v0 : v0;
          if (col + 1 < p.ncols_d) dst[1] = p.accumulate ? dst[1] + v1 : v1;
        }
        if (MASKED && mirroff >= 0) {
          double *md = p.out + mirroff + col;
          double w0 = a0, w1 = a1;
          if (col + 1 < p.ncols_d && ((reinterpret_cast<size_t>(md) & 15) == 0)) {
            if (p.accumulate) { const double2 o = *reinterpret_cast<double2 *>(md); w0 += o.x; w1 += o.y; }
            *reinterpret_cast<double2 *>(md) = make_double2(w0, w1);
          } else {
            if (col < p.ncols_d) md[0] = p.accumulate ? md[0] + w0 : w0;
            if (col + 1 < p.ncols_d) md[1] = p.accumulate ? md[1] + w1 : w1;
          }
        }
      }
    }
    }
    }   // segments
  }
}

// With a pair mask: ids of the tiles (tn fastest, tj, ti) that hold at least one kept pair, in increasing id order.
// Single block: the tile count is small (cfg5: 3888) and the order must be reproducible.
__global__ void __launch_bounds__(256) k_tile_list(const unsigned char *mask, int mask_ld, int symmetric, int i0, int ni, int j0,
                                                   int nj, int tiles_i, int tiles_j, int tiles_n, int BI, int *list, int *count) {
  __shared__ int base;
  __shared__ int wsum[8];
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int npt = tiles_i * tiles_j;                      // pair tiles; each expands to tiles_n column tiles
  for (int t0 = 0; t0 < npt; t0 += 256) {
    const int pt = t0 + threadIdx.x;
    int any = 0;
    if (pt < npt) {
      const int ti = pt / tiles_j, tj = pt % tiles_j;
      for (int a = 0; a < BI && !any; a++) {
        const int i = i0 + ti * BI + a;
        if (i >= i0 + ni) break;
        for (int b = 0; b < 8; b++) {
          const int j = j0 + tj * 8 + b;
          if (j < j0 + nj && (symmetric ? (j >= i) : (mask[(size_t)i * mask_ld + j] != 0))) { any = 1; break; }
        }
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, any);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; w++) off += wsum[w];
    off += __popc(bal & ((1u << lane) - 1));
    if (any) for (int tn = 0; tn < tiles_n; tn++) list[off * tiles_n + tn] = pt * tiles_n + tn;
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < 8; w++) tot += wsum[w]; base += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base * tiles_n;
}

// Adds the stream-K pieces of each remainder tile in CTA order and stores the tile (guards, accumulate).
// grid = (remainder tiles, element slices of a tile).  Geometry arguments restate k_project_ws's decomposition.
constexpr int FIXUP_MAX_G = 1024;
__global__ void __launch_bounds__(256) k_fixup(const __grid_constant__ ProjParams p, int G, int BM, int BN, int BI, int NJ,
                                               int matrix, int kq) {
  const int nchunk = (int)((p.q1 - p.q0 + kq - 1) / kq);
  const int ntiles = p.tile_list ? *p.tile_count : p.tiles_i * p.tiles_j * p.tiles_n;
  const int nfull = ntiles / G;
  const long long rem_units = (long long)(ntiles - nfull * G) * nchunk;
  const int rt = blockIdx.x;
  if (rt >= ntiles - nfull * G) return;          // launched with the worst-case G-1 blocks when a tile list is used
  const int tile = p.tile_list ? p.tile_list[nfull * G + rt] : nfull * G + rt;
  const int tn = tile % p.tiles_n, tj = (tile / p.tiles_n) % p.tiles_j, ti = tile / (p.tiles_n * p.tiles_j);
  const int ibase = p.i0 + ti * BI, jbase = p.j0 + tj * NJ, col0 = tn * BN;
  const int iend = p.i0 + p.ni, jend = p.j0 + p.nj;
  const long long u0 = (long long)rt * nchunk, u1 = u0 + nchunk;
  // first CTA whose range reaches past u0:  r1_g = rem_units*(g+1)/G > u0
  int g_lo = (int)((u0 * G) / rem_units);
  while (g_lo > 0 && rem_units * g_lo / G > u0) g_lo--;
  while (rem_units * (g_lo + 1) / G <= u0) g_lo++;
  // the workspace slots that hold pieces of this tile, in CTA order (block-uniform, computed once)
  __shared__ int s_src[FIXUP_MAX_G];
  __shared__ int s_n;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int g = g_lo; g < G && n < FIXUP_MAX_G; g++) {
      const long long a0 = rem_units * g / G, a1 = rem_units * (g + 1) / G;
      if (a0 >= u1) break;
      if (a1 <= a0) continue;
      s_src[n++] = g * 2 + ((a0 / nchunk == rt) ? 0 : 1);
    }
    s_n = n;
  }
  __syncthreads();
  const int npieces = s_n;
  // blockIdx.y splits the tile's elements, so that a few big remainder tiles still fill the machine
  bool pushed = false;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < BM * BN; e += gridDim.y * blockDim.x) {
    const int r = e / BN, cc = e % BN;
    double sum = 0.0;
#pragma unroll 4
    for (int k = 0; k < npieces; k++) sum += p.sk_ws[(size_t)s_src[k] * ((size_t)BM * BN) + e];
    if (p.peer_world > 0) {          // fused multi-GPU mode: the summed partial tile goes to its owner's receive slot
      const int pt = tile / p.tiles_n;
      const int owner = min(p.peer_world - 1, pt / p.peer_lp), lp = pt - owner * p.peer_lp;
      p.peer_recv[owner][((size_t)(p.peer_rank * p.peer_lp + lp) * BM + r) * p.peer_ncols_pad + (size_t)tn * BN + cc] = sum;
      pushed = true;
      continue;
    }
    const int col = col0 + cc;
    if (col >= p.ncols_d) continue;
    long long rowoff, mirroff = -1;
    bool rok;
    if (matrix) {
      const int i = ibase + r / 8, j = jbase + (r & 7);
      rok = i < iend && j < jend;
      rowoff = (long long)(i - p.i0) * p.stride_i + (long long)(j - p.j0) * p.stride_j;
      if (rok && p.symmetric) {
        if (j < i) continue;
        if (j > i) mirroff = (long long)(j - p.i0) * p.stride_i + (long long)(i - p.j0) * p.stride_j;
      } else if (rok && p.mask && p.mask[(size_t)i * p.mask_ld + j] == 0) { if (p.accumulate) continue; sum = 0.0; }
    } else {
      const int k = jbase + r;
      rok = k < jend;
      rowoff = (long long)(k - p.j0) * p.stride_j;
    }
    if (!rok) continue;
    double *dst = p.out + rowoff + col;
    *dst = p.accumulate ? *dst + sum : sum;
    if (mirroff >= 0) { double *md = p.out + mirroff + col; *md = p.accumulate ? *md + sum : sum; }
  }
  if (pushed) __threadfence_system();      // peer stores of this thread must have landed at their owners before the kernel completes
}

// Second half of the fused collective.  Every rank owns peer_lp consecutive pair tiles; after all ranks' pushes have
// landed (cross-rank barrier), the owner adds the `world` partial blocks of each owned element in rank order
// (deterministic) and stores the total into the output of EVERY rank (peer-mapped): reduce-scatter + all-gather.
struct PeerGatherParams {
  const double *recv;                     // my receive buffer [world][peer_lp][BM][ncols_pad]
  double *peer_out[PSPACE_MAX_PEERS];     // the C windows of all ranks
  int world, rank, lp, BM, BI, tiles_j, npt;
  long long ncols_pad, stride_i, stride_j;
  int i0, ni, j0, nj, ncols_d, accumulate;
};
__global__ void __launch_bounds__(256) k_peer_reduce_gather(const __grid_constant__ PeerGatherParams p) {
  const int lp = blockIdx.x, pt = p.rank * p.lp + lp;
  if (pt >= p.npt) return;
  const int ti = pt / p.tiles_j, tj = pt % p.tiles_j;
  const size_t slot_elems = (size_t)p.lp * p.BM * p.ncols_pad;
  const int half = (int)(p.ncols_pad / 2);                 // ncols_pad is even (BN is a multiple of 8)
  for (int e = threadIdx.x; e < p.BM * half; e += blockDim.x) {
    const int r = e / half, c2 = (e % half) * 2;
    const int i = p.i0 + ti * p.BI + r / 8, j = p.j0 + tj * 8 + (r & 7);
    if (i >= p.i0 + p.ni || j >= p.j0 + p.nj || c2 >= p.ncols_d) continue;
    const size_t off = ((size_t)lp * p.BM + r) * p.ncols_pad + c2;
    double2 s = *reinterpret_cast<const double2 *>(p.recv + off);
    for (int w = 1; w < p.world; w++) {
      const double2 v = *reinterpret_cast<const double2 *>(p.recv + w * slot_elems + off);
      s.x += v.x; s.y += v.y;
    }
    const long long o = (long long)(i - p.i0) * p.stride_i + (long long)(j - p.j0) * p.stride_j + c2;
    const bool two = c2 + 1 < p.ncols_d;
    for (int w = 0; w < p.world; w++) {
      double *dst = p.peer_out[w] + o;
      if (p.accumulate) { dst[0] += s.x; if (two) dst[1] += s.y; }
      else if (two && ((reinterpret_cast<size_t>(dst) & 15) == 0)) *reinterpret_cast<double2 *>(dst) = s;
      else { dst[0] = s.x; if (two) dst[1] = s.y; }
    }
  }
  __threadfence_system();
}

// TMA descriptors (fp64, no swizzle, zero fill out of bounds).  dims/box innermost first.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn resolve_encode_tiled() {
  void *sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
  return (EncodeTiledFn)sym;
}
int make_tmap(CUtensorMap *tm, const double *base, int rank, const unsigned long long *dims,
              const unsigned long long *strides_bytes /*rank-1*/, const unsigned *box) {
  static const EncodeTiledFn fn = resolve_encode_tiled();     // C++11 magic static: initialised once, thread-safe
  if (!fn) return ps_fail(PSPACE_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<double *>(base), gdim, gstr, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ps_fail(PSPACE_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PSPACE_OK;
}

// One basis panel (rows [r0, r0+nr) over [q0,q1), weighted or not) in its context slot: grow-only buffer, rebuilt
// unless the cache is on and the slot already holds exactly this panel.
// slot_prepare: buffer of the slot (grown if needed), ordered after its last use; *build = the panel has to be (re)written.
int slot_prepare(pspace_ctx *c, bool weighted, int r0, int nr, long long q0, long long q1, double **d_out, int *nr_pad,
                 cudaStream_t st, bool *build) {
  pspace_ctx::PanelSlot &s = weighted ? c->panel_w : c->panel_u;
  const int W = c->width;
  const long long nq = q1 - q0;
  *nr_pad = (nr + 1) & ~1;
  *build = false;
  const size_t need = (size_t)W * nq * *nr_pad * 8;
  // order this call after the last use of the slot on whatever stream that was
  if (s.ev_pending) PS_CUDA(cudaStreamWaitEvent(st, s.ev, 0));
  if (need > s.bytes) {
    if (s.d) {
      if (s.ev_pending) PS_CUDA(cudaEventSynchronize(s.ev));
      PS_CUDA(cudaStreamSynchronize(st));
      PS_CUDA(cudaFree(s.d)); s.d = nullptr; s.bytes = 0;
    }
    s.valid = false;
    cudaError_t e = cudaMalloc(&s.d, need);
    if (e != cudaSuccess) return ps_fail(PSPACE_ENOMEM, "basis panel of %zu bytes: %s", need, cudaGetErrorString(e));
    s.bytes = need;
  }
  *d_out = s.d;
  if (need == 0) return PSPACE_OK;
  if (c->panel_cache && s.valid && s.r0 == r0 && s.nr == nr && s.q0 == q0 && s.q1 == q1) return PSPACE_OK;
  s.valid = false;
  *build = true;
  return PSPACE_OK;
}

// One launch of k_psi_panel for rows [r0, r0+nr) over [q0,q1): the unweighted panel into d_u and/or the weighted one into d_w.
int panel_launch(pspace_ctx *c, int r0, int nr, long long q0, long long q1, double *d_u, double *d_w, cudaStream_t st) {
  const int W = c->width;
  const long long nq = q1 - q0;
  PanelParams pp;
  pp.nvars = c->nvars; pp.tsz = c->tsz; pp.DP = c->DP; pp.width = W;
  pp.T = c->d_T; pp.deg = c->d_deg; pp.qdig = c->d_qdig; pp.W = c->d_W;
  for (int d = 0; d < c->nvars; d++) { pp.toff[d] = c->toff[d]; pp.nq[d] = c->nq[d]; }
  pp.q0 = q0; pp.q1 = q1;
  ps_choose_tail(c, PSI_TAIL, &pp.dsplit, &pp.tail_stride);
  pp.r0 = r0; pp.nr = nr; pp.nr_pad = (nr + 1) & ~1;
  pp.weighted = (d_w && !d_u) ? 1 : 0;
  pp.out = d_u ? d_u : d_w;
  pp.out_w = (d_u && d_w) ? d_w : nullptr;
  const size_t smem = (size_t)(((c->tsz * W + 1) & ~1) + PSI_PTS * W) * 8 + (size_t)PSI_PTS * c->DP;
  const int nt = std::min(PSI_MAXT, (pp.nr_pad + 31) & ~31);
  // points per block: 128 for big grids (long runs amortise the prefix products); small grids get shorter runs so that at
  // least ~4 blocks per SM exist (cfg2: 3125 points x 126 rows was 25 blocks walking 128 points each: 35 us per panel)
  const long long row_blocks = (pp.nr_pad + nt - 1) / nt;
  int pts = PSI_PTS;
  while (pts > 16 && ((nq + pts - 1) / pts) * row_blocks < 4LL * sm_count(c->device)) pts >>= 1;
  pp.pts = pts;
  const dim3 grid((unsigned)((nq + pts - 1) / pts), (unsigned)row_blocks);
  if (smem > 48 * 1024) {   // only then is an opt-in needed (tables of more than ~5000 entries); never on the BASELINE configs
    if (c->is_complex) PS_CUDA(cudaFuncSetAttribute(k_psi_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else PS_CUDA(cudaFuncSetAttribute(k_psi_panel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (c->is_complex) k_psi_panel<true><<<grid, nt, smem, st>>>(pp);
  else k_psi_panel<false><<<grid, nt, smem, st>>>(pp);
  PS_LAUNCH_CHECK(c);
  return PSPACE_OK;
}

void slot_commit(pspace_ctx *c, bool weighted, int r0, int nr, long long q0, long long q1) {
  pspace_ctx::PanelSlot &s = weighted ? c->panel_w : c->panel_u;
  s.r0 = r0; s.nr = nr; s.q0 = q0; s.q1 = q1;
  s.valid = true;
}

// One basis panel (rows [r0, r0+nr) over [q0,q1), weighted or not) in its context slot: grow-only buffer, rebuilt
// unless the cache is on and the slot already holds exactly this panel.
int ensure_panel(pspace_ctx *c, bool weighted, int r0, int nr, long long q0, long long q1, double **d_out, int *nr_pad,
                 cudaStream_t st) {
  bool build = false;
  int rc = slot_prepare(c, weighted, r0, nr, q0, q1, d_out, nr_pad, st, &build);
  if (rc || !build) return rc;
  if ((rc = panel_launch(c, r0, nr, q0, q1, weighted ? nullptr : *d_out, weighted ? *d_out : nullptr, st))) return rc;
  slot_commit(c, weighted, r0, nr, q0, q1);
  return PSPACE_OK;
}

// basis panels of a projection call: Pi (unweighted, matrix mode only) and Pj (weighted)
int build_panels(pspace_ctx *c, const ProjParams &p, bool matrix, double **d_pi, double **d_pj, int *ni_pad,
                 int *nj_pad, cudaStream_t st) {
  *ni_pad = 0;
  *d_pi = nullptr;
  if (!matrix) return ensure_panel(c, true, p.j0, p.nj, p.q0, p.q1, d_pj, nj_pad, st);
  bool bi = false, bj = false;
  int rc = slot_prepare(c, false, p.i0, p.ni, p.q0, p.q1, d_pi, ni_pad, st, &bi);
  if (rc) return rc;
  if ((rc = slot_prepare(c, true, p.j0, p.nj, p.q0, p.q1, d_pj, nj_pad, st, &bj))) return rc;
  if (bi && bj && p.i0 == p.j0 && p.ni == p.nj) {
    // square window (the reference's all-pairs loop): both panels hold the same rows - one pass writes psi and w*psi
    if ((rc = panel_launch(c, p.i0, p.ni, p.q0, p.q1, *d_pi, *d_pj, st))) return rc;
  } else {
    if (bi && (rc = panel_launch(c, p.i0, p.ni, p.q0, p.q1, *d_pi, nullptr, st))) return rc;
    if (bj && (rc = panel_launch(c, p.j0, p.nj, p.q0, p.q1, nullptr, *d_pj, st))) return rc;
  }
  if (bi) slot_commit(c, false, p.i0, p.ni, p.q0, p.q1);
  if (bj) slot_commit(c, true, p.j0, p.nj, p.q0, p.q1);
  return PSPACE_OK;
}

// sums the ksplit partial windows in slice order (deterministic) into the output
__global__ void k_reduce_slices(const double *ws, long long slice_stride, int ksplit, double *out, long long n,
                                int accumulate) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    double s = ws[i];
    for (int k = 1; k < ksplit; k++) s += ws[i + k * slice_stride];
    out[i] = accumulate ? out[i] + s : s;
  }
}

// ---- host-side planning ----------------------------------------------------------------------------
struct Variant {
  int id, BM, BN, BI, NJ, minb;
  const char *name;
};

struct Plan {
  int variant_idx;
  int tiles_i, tiles_j, tiles_n, ksplit;
  long long chunks_per_slice, nchunk;
  size_t ws_bytes;          // what the caller must provide (max over the two kernels' needs)
  size_t ws_generic, ws_streamk;
  int sk_grid, sk_rem_tiles;   // fast path: persistent CTAs, tiles left to the stream-K remainder
  long long out_elems;
};

// Launches one projection with tile configuration TL.  `fast` = TMA warp-specialised persistent kernel
// (stream-K, writes C directly + k_fixup); otherwise the generic kernel over (tiles, ksplit) + k_reduce_slices.
template <class TL> int launch(const pspace_ctx *c, ProjParams &p, const Plan &pl, const pspace_project_args *a,
                               double *d_out, void *d_ws, cudaStream_t st) {
  pspace_ctx *mc = const_cast<pspace_ctx *>(c);
  if (p.fast) {
    typedef WsLayout<TL> LY;
    constexpr bool MAT = TL::MODE == MODE_MATRIX;
    const size_t smem = LY::smem_bytes();
    static_assert(LY::STAGES * LY::STAGE_BYTES + 2048 <= 227 * 1024, "stage ring exceeds shared memory");
    double *d_pi = nullptr, *d_pj = nullptr;
    int ni_pad = 0, nj_pad = 0;
    int rc = build_panels(mc, p, MAT, &d_pi, &d_pj, &ni_pad, &nj_pad, st);
    if (rc) return rc;
    const unsigned long long nq = (unsigned long long)(p.q1 - p.q0);
    CUtensorMap tmB, tmI, tmJ;
    {
      unsigned long long dims[2] = {(unsigned long long)p.ncols_d, nq}, str[1] = {(unsigned long long)p.ld_in * 8};
      unsigned box[2] = {(unsigned)TL::LDB, (unsigned)LY::KQW};
      if ((rc = make_tmap(&tmB, p.in, 2, dims, str, box))) return rc;
    }
    {
      unsigned long long dims[3] = {(unsigned long long)nj_pad, nq, (unsigned long long)TL::W};
      unsigned long long str[2] = {(unsigned long long)nj_pad * 8, (unsigned long long)nj_pad * 8 * nq};
      unsigned box[3] = {(unsigned)TL::PJP, (unsigned)LY::KQW, (unsigned)TL::W};
      if ((rc = make_tmap(&tmJ, d_pj, 3, dims, str, box))) return rc;
    }
    if (MAT) {
      unsigned long long dims[3] = {(unsigned long long)ni_pad, nq, (unsigned long long)TL::W};
      unsigned long long str[2] = {(unsigned long long)ni_pad * 8, (unsigned long long)ni_pad * 8 * nq};
      unsigned box[3] = {(unsigned)TL::BIP, (unsigned)LY::KQW, (unsigned)TL::W};
      if ((rc = make_tmap(&tmI, d_pi, 3, dims, str, box))) return rc;
    } else tmI = tmJ;
    static std::atomic<bool> attr_ws_dev[64];   // per instantiation and device (function attributes are per device);
    std::atomic<bool> &attr_ws = attr_ws_dev[c->device & 63];   // setting an attribute twice is harmless, the flag is atomic
    if (!attr_ws.load(std::memory_order_acquire)) {
      PS_CUDA(cudaFuncSetAttribute(k_project_ws<TL, EPI_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      if (MAT) {
        PS_CUDA(cudaFuncSetAttribute(k_project_ws<TL, EPI_DENSE, MAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        PS_CUDA(cudaFuncSetAttribute(k_project_ws<TL, MAT ? EPI_MASKED : EPI_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        PS_CUDA(cudaFuncSetAttribute(k_project_ws<TL, MAT ? EPI_PEER : EPI_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      }
      attr_ws.store(true, std::memory_order_release);
    }
    p.out = d_out;
    p.accumulate = a->accumulate;
    p.sk_ws = (double *)d_ws;
    // dense matrix launches whose last stage is at most 3/4 full and is >= 0.4 % of a tile's k-loop: kernel with the trimmed last chunk
    const long long tail_pts = (long long)(nq % LY::KQW);
    const bool ragged = !c->no_ragged && tail_pts != 0 && tail_pts <= LY::KQW * 3 / 4 && (LY::KQW - tail_pts) * 250 >= (long long)nq;
    int fix_blocks = pl.sk_rem_tiles;
    int *d_list = nullptr;
    if (p.mask || p.symmetric) {
      // tile list on the device; with a mask untouched tiles must read as zero (symmetric mode writes every block)
      const int ntl = p.tiles_i * p.tiles_j * p.tiles_n;
      PS_CUDA(cudaMallocAsync(&d_list, sizeof(int) * (ntl + 1), st));
      k_tile_list<<<1, 256, 0, st>>>(p.mask, p.mask_ld, p.symmetric, p.i0, p.ni, p.j0, p.nj, p.tiles_i, p.tiles_j, p.tiles_n, TL::BI,
                                     d_list + 1, d_list);
      PS_LAUNCH_CHECK(c);
      p.tile_list = d_list + 1;
      p.tile_count = d_list;
      if (!a->accumulate && !p.symmetric) PS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * pl.out_elems, st));
      fix_blocks = pl.sk_grid - 1;     // remainder tile count is only known on the device
    }
    if (MAT && p.peer_world > 0) k_project_ws<TL, MAT ? EPI_PEER : EPI_DENSE><<<pl.sk_grid, NT_WS, smem, st>>>(p, tmB, tmI, tmJ);
    else if (MAT && (p.mask || p.symmetric)) k_project_ws<TL, MAT ? EPI_MASKED : EPI_DENSE><<<pl.sk_grid, NT_WS, smem, st>>>(p, tmB, tmI, tmJ);
    else if (MAT && ragged) k_project_ws<TL, EPI_DENSE, MAT><<<pl.sk_grid, NT_WS, smem, st>>>(p, tmB, tmI, tmJ);
    else k_project_ws<TL, EPI_DENSE><<<pl.sk_grid, NT_WS, smem, st>>>(p, tmB, tmI, tmJ);
    PS_LAUNCH_CHECK(c);
    if (fix_blocks > 0) {
      // ~2 blocks per SM in total, at least 256 elements per block
      const int fix_split = std::max(1, std::min((TL::BM * TL::BN) / 256, (2 * pl.sk_grid + fix_blocks - 1) / fix_blocks));
      k_fixup<<<dim3(fix_blocks, fix_split), 256, 0, st>>>(p, pl.sk_grid, TL::BM, TL::BN, TL::BI, TL::NJ, MAT ? 1 : 0, LY::KQW);
      PS_LAUNCH_CHECK(c);
    }
    if (d_list) PS_CUDA(cudaFreeAsync(d_list, st));
    ps_panels_mark_used(mc, st);
    mc->plan_kernel = "tma+mbarrier warp-specialised, persistent stream-K";
    mc->plan_chunk_pts = LY::KQW;
    return PSPACE_OK;
  }
  size_t smem = TL::smem_bytes(c->tsz, c->nvars);
  if (smem > 227 * 1024) return ps_fail(PSPACE_EINVAL, "1-D tables too large for the staged kernel (%zu B shared memory)", smem);
  static std::atomic<bool> attr_set_dev[64];   // per instantiation and device
  std::atomic<bool> &attr_set = attr_set_dev[c->device & 63];
  if (!attr_set.load(std::memory_order_acquire)) {
    PS_CUDA(cudaFuncSetAttribute(k_project<TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set.store(true, std::memory_order_release);
  }
  const bool split = pl.ksplit > 1;
  p.out = split ? (double *)d_ws : d_out;
  p.slice_stride = split ? pl.out_elems : 0;
  p.accumulate = split ? 0 : a->accumulate;
  dim3 grid((unsigned)((long long)pl.tiles_i * pl.tiles_j * pl.tiles_n), (unsigned)pl.ksplit);
  k_project<TL><<<grid, NT, smem, st>>>(p);
  PS_LAUNCH_CHECK(c);
  if (split) {
    const long long n = pl.out_elems;
    const unsigned nb = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 16);
    k_reduce_slices<<<nb, 256, 0, st>>>((const double *)d_ws, pl.out_elems, pl.ksplit, d_out, n, a->accumulate);
    PS_LAUNCH_CHECK(c);
  }
  mc->plan_kernel = "cp.async all-warps";
  mc->plan_chunk_pts = KQ;
  return PSPACE_OK;
}

//                         WM WN WGM WGN cplx   mode         stages minb
// "wide-warp" tiles: each consumer warp owns 16 pairs x ALL columns of the CTA tile, so one fragment DMUL feeds
// WN = 12 (8) DMMAs; measured 2.6 % faster on cfg5 than the 4x2 warp grid of the same CTA tile (ids 5, 6).
typedef Tile<2, 12, 8, 1, false, MODE_MATRIX, 4, 1> M96;    // 128 pairs x 96 cols   (m*m = 576 = 6 x 96)
typedef Tile<2, 8, 8, 1, false, MODE_MATRIX, 4, 1> M64;     // 128 x 64              (m*m = 64)
typedef Tile<4, 2, 8, 1, false, MODE_MATRIX, 4, 1> M16;     // 256 x 16
typedef Tile<4, 1, 8, 1, false, MODE_MATRIX, 4, 1> M8;      // 256 x 8               (scalar J)
typedef Tile<4, 6, 4, 2, false, MODE_MATRIX, 4, 1> M96G;    // 128 x 96, 4x2 warp grid
typedef Tile<4, 4, 4, 2, false, MODE_MATRIX, 4, 1> M64G;    // 128 x 64, 4x2 warp grid
typedef Tile<2, 12, 8, 1, true, MODE_MATRIX, 4, 1> ZM96;
typedef Tile<2, 8, 8, 1, true, MODE_MATRIX, 4, 1> ZM64;
typedef Tile<4, 2, 8, 1, true, MODE_MATRIX, 4, 1> ZM16;
typedef Tile<4, 1, 8, 1, true, MODE_MATRIX, 4, 1> ZM8;
typedef Tile<4, 6, 4, 2, true, MODE_MATRIX, 4, 1> ZM96G;
typedef Tile<4, 4, 4, 2, true, MODE_MATRIX, 4, 1> ZM64G;
typedef Tile<2, 4, 8, 1, false, MODE_VECTOR, 4, 1> V32;     // 128 basis rows x 32 cols
typedef Tile<2, 1, 8, 1, false, MODE_VECTOR, 4, 1> V8;
typedef Tile<2, 3, 8, 1, false, MODE_VECTOR, 4, 1> V24;     // 128 x 24 (m = 24: one column tile, no padded columns)
typedef Tile<2, 4, 8, 1, true, MODE_VECTOR, 4, 1> ZV32;
typedef Tile<2, 1, 8, 1, true, MODE_VECTOR, 4, 1> ZV8;
typedef Tile<2, 3, 8, 1, true, MODE_VECTOR, 4, 1> ZV24;

const Variant MAT_VARIANTS[] = {{1, M96::BM, M96::BN, M96::BI, 8, 1, "dmma 128x96"},
                                {2, M64::BM, M64::BN, M64::BI, 8, 1, "dmma 128x64"},
                                {3, M16::BM, M16::BN, M16::BI, 8, 1, "dmma 256x16"},
                                {4, M8::BM, M8::BN, M8::BI, 8, 1, "dmma 256x8"},
                                {5, M96G::BM, M96G::BN, M96G::BI, 8, 1, "dmma 128x96 (4x2 warps)"},
                                {6, M64G::BM, M64G::BN, M64G::BI, 8, 1, "dmma 128x64 (4x2 warps)"}};
const Variant VEC_VARIANTS[] = {{1, V32::BM, V32::BN, 0, V32::BM, 1, "dmma vec 128x32"},
                                {2, V8::BM, V8::BN, 0, V8::BM, 1, "dmma vec 128x8"},
                                {3, V24::BM, V24::BN, 0, V24::BM, 1, "dmma vec 128x24"}};

int sm_count(int device) {
  static std::atomic<int> cached[64];
  if (device < 64) { const int v = cached[device].load(std::memory_order_relaxed); if (v) return v; }
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (device < 64) cached[device].store(n, std::memory_order_relaxed);
  return n;
}

int make_plan(const pspace_ctx *c, int is_matrix, const pspace_project_args *a, Plan &pl) {
  if (!c->tables_ready) return ps_fail(PSPACE_ESTATE, "project called before initializeBasis + initializeQuadrature");
  if (a->ncols < 1) return ps_fail(PSPACE_EINVAL, "ncols must be >= 1");
  if (a->q0 < 0 || a->q1 > c->Q || a->q1 < a->q0) return ps_fail(PSPACE_EINVAL, "q range [%lld,%lld) outside [0,%lld)", a->q0, a->q1, c->Q);
  if (a->k0 < 0 || a->k1 > c->N || a->k1 < a->k0) return ps_fail(PSPACE_EINVAL, "basis range [%d,%d) outside [0,%d)", a->k0, a->k1, c->N);
  if (is_matrix && (a->j0 < 0 || a->j1 > c->N || a->j1 < a->j0)) return ps_fail(PSPACE_EINVAL, "j range [%d,%d) outside [0,%d)", a->j0, a->j1, c->N);
  const int ncols_d = a->ncols * c->width;
  const Variant *vs = is_matrix ? MAT_VARIANTS : VEC_VARIANTS;
  const int nv = is_matrix ? 6 : 3;
  const int nauto = is_matrix ? 4 : 3;   // variants considered by the automatic choice
  int best = -1;
  if (a->variant > 0) {
    for (int i = 0; i < nv; i++) if (vs[i].id == a->variant) best = i;
    if (best < 0) return ps_fail(PSPACE_EINVAL, "unknown tile variant %d", a->variant);
  } else {
    double bestcost = 0;
    int besttn = 0;
    for (int i = 0; i < nauto; i++) {
      // padded work; ties go to the earlier (wider) variant
      const int tn = (ncols_d + vs[i].BN - 1) / vs[i].BN;
      double padn = (double)tn * vs[i].BN;
      double padm;
      if (is_matrix) padm = (double)((a->k1 - a->k0 + vs[i].BI - 1) / vs[i].BI) * vs[i].BI * ((a->j1 - a->j0 + 7) / 8) * 8;
      else padm = (double)((a->k1 - a->k0 + vs[i].BM - 1) / vs[i].BM) * vs[i].BM;
      double cost = padn * padm;
      // the vector projection is bound by streaming the basis panel, which every column tile reads again:
      // fewest column tiles first, padded work second
      const bool better = is_matrix ? cost < bestcost * 0.999
                                    : (tn < besttn || (tn == besttn && cost < bestcost * 0.999));
      if (best < 0 || better) { best = i; bestcost = cost; besttn = tn; }
    }
  }
  const Variant &v = vs[best];
  pl.variant_idx = best;
  pl.tiles_n = (ncols_d + v.BN - 1) / v.BN;
  if (is_matrix) {
    pl.tiles_i = (a->k1 - a->k0 + v.BI - 1) / v.BI;
    pl.tiles_j = (a->j1 - a->j0 + 7) / 8;
    pl.out_elems = (long long)(a->k1 - a->k0) * (a->j1 - a->j0) * ncols_d;
  } else {
    pl.tiles_i = 1;
    pl.tiles_j = (a->k1 - a->k0 + v.BM - 1) / v.BM;
    pl.out_elems = (long long)(a->k1 - a->k0) * ncols_d;
  }
  pl.nchunk = (a->q1 - a->q0 + KQ - 1) / KQ;
  long long ntiles = (long long)pl.tiles_i * pl.tiles_j * pl.tiles_n;
  int ks = 1;
  if (a->ksplit > 0) ks = a->ksplit;
  else if (ntiles > 0) {
    // fill ~2 waves of the machine, but keep >= 8 chunks (256 points) per slice
    const long long target = 2LL * sm_count(c->device) * v.minb;
    if (ntiles < target) ks = (int)std::min<long long>((target + ntiles - 1) / ntiles, std::max<long long>(1, pl.nchunk / 8));
  }
  if (ks > pl.nchunk) ks = (int)std::max<long long>(1, pl.nchunk);
  if (ks > 65535) ks = 65535;
  pl.chunks_per_slice = ks > 0 ? (pl.nchunk + ks - 1) / ks : 0;
  if (pl.chunks_per_slice > 0) ks = (int)((pl.nchunk + pl.chunks_per_slice - 1) / pl.chunks_per_slice);
  pl.ksplit = std::max(1, ks);
  pl.ws_generic = pl.ksplit > 1 ? (size_t)pl.ksplit * pl.out_elems * sizeof(double) : 0;
  // fast path (possible when the real column count is even; the operand's alignment is only known at launch)
  pl.sk_grid = (int)std::max<long long>(1, std::min<long long>(sm_count(c->device), ntiles * pl.nchunk));
  pl.sk_rem_tiles = (int)(ntiles % pl.sk_grid);
  pl.ws_streamk = (ncols_d % 2 == 0 && (pl.sk_rem_tiles > 0 || (is_matrix && (a->d_pair_mask || a->symmetric)))) ? (size_t)pl.sk_grid * 2 * v.BM * v.BN * sizeof(double) : 0;
  pl.ws_bytes = std::max(pl.ws_generic, pl.ws_streamk);
  return PSPACE_OK;
}
}  // namespace

void ps_panels_mark_used(pspace_ctx *c, cudaStream_t st) {
  pspace_ctx::PanelSlot *slots[2] = {&c->panel_u, &c->panel_w};
  for (pspace_ctx::PanelSlot *s : slots) {
    if (!s->d) continue;
    if (!s->ev && cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming) != cudaSuccess) { s->ev = nullptr; continue; }
    if (cudaEventRecord(s->ev, st) == cudaSuccess) s->ev_pending = true;
  }
}

int ps_make_tmap(CUtensorMap *tm, const double *base, int rank, const unsigned long long *dims,
                 const unsigned long long *strides_bytes, const unsigned *box) {
  return make_tmap(tm, base, rank, dims, strides_bytes, box);
}

// Unweighted / weighted basis panel of rows [r0, r0+nr) over [q0,q1) in the context's panel slot (used by expand.cu)
int ps_build_panel_rows(pspace_ctx *c, int r0, int nr, long long q0, long long q1, int weighted, double **d_panel,
                        int *nr_pad, cudaStream_t st) {
  return ensure_panel(c, weighted != 0, r0, nr, q0, q1, d_panel, nr_pad, st);
}

size_t ps_project_workspace(const pspace_ctx *c, int is_matrix, const pspace_project_args *a) {
  Plan pl;
  if (!is_matrix) {
    // which vector kernel runs depends on the operand's alignment, known only at the call — size for either
    // (`variant` ids above 3 name tiles of the factored-operand kernel only)
    pspace_project_args b = *a;
    if (b.variant > 3) b.variant = 0;
    const size_t w_old = make_plan(c, 0, &b, pl) == PSPACE_OK ? pl.ws_bytes : 0;
    return std::max(w_old, ps_vproject_workspace(c, a));
  }
  if (make_plan(c, is_matrix, a, pl) != PSPACE_OK) return 0;
  return pl.ws_bytes;
}

int ps_project(const pspace_ctx *c, int is_matrix, const pspace_project_args *a, const double *d_in, double *d_out,
               void *d_ws, size_t ws_bytes, cudaStream_t st) {
  return ps_project_ex(c, is_matrix, a, d_in, d_out, d_ws, ws_bytes, st, nullptr);
}

// geometry of the fused multi-GPU exchange: doubles each rank must provide as its receive buffer
int ps_peer_layout(const pspace_ctx *c, const pspace_project_args *a, int world, long long *recv_elems) {
  Plan pl;
  int rc = make_plan(c, 1, a, pl);
  if (rc) return rc;
  if (world < 1 || world > PSPACE_MAX_PEERS) return ps_fail(PSPACE_EINVAL, "world %d outside 1..%d", world, PSPACE_MAX_PEERS);
  const Variant &v = MAT_VARIANTS[pl.variant_idx];
  const long long npt = (long long)pl.tiles_i * pl.tiles_j, lp = (npt + world - 1) / world;
  *recv_elems = (long long)world * lp * v.BM * ((long long)pl.tiles_n * v.BN);
  return PSPACE_OK;
}

int ps_peer_reduce_gather(const pspace_ctx *c, const pspace_project_args *a, const double *d_recv, double *const *peer_out,
                          int world, int rank, cudaStream_t st) {
  Plan pl;
  int rc = make_plan(c, 1, a, pl);
  if (rc) return rc;
  const Variant &v = MAT_VARIANTS[pl.variant_idx];
  PeerGatherParams g;
  g.recv = d_recv;
  for (int w = 0; w < PSPACE_MAX_PEERS; w++) g.peer_out[w] = w < world ? peer_out[w] : nullptr;
  g.world = world; g.rank = rank;
  g.npt = pl.tiles_i * pl.tiles_j;
  g.lp = (g.npt + world - 1) / world;
  g.BM = v.BM; g.BI = v.BI; g.tiles_j = pl.tiles_j;
  g.ncols_pad = (long long)pl.tiles_n * v.BN;
  const int ncols_d = a->ncols * c->width;
  g.i0 = a->k0; g.ni = a->k1 - a->k0; g.j0 = a->j0; g.nj = a->j1 - a->j0;
  g.stride_j = ncols_d; g.stride_i = (long long)g.nj * ncols_d;
  g.ncols_d = ncols_d; g.accumulate = a->accumulate;
  if (g.lp > 0) {
    k_peer_reduce_gather<<<g.lp, 256, 0, st>>>(g);
    PS_LAUNCH_CHECK(c);
  }
  return PSPACE_OK;
}

int ps_project_ex(const pspace_ctx *c, int is_matrix, const pspace_project_args *a, const double *d_in, double *d_out,
                  void *d_ws, size_t ws_bytes, cudaStream_t st, const PeerInfo *peer) {
  Plan pl;
  // vector mode: `variant` may name a tile of the factored-operand kernel (ids 1..8), which make_plan does not know
  const bool to_vf = !is_matrix && !peer && ps_vproject_usable(c, a, d_in);
  pspace_project_args a_chk = *a;
  if (to_vf) a_chk.variant = 0;
  int rc = make_plan(c, is_matrix, &a_chk, pl);
  if (rc) return rc;
  pspace_ctx *mc = const_cast<pspace_ctx *>(c);
  const int ncols_d = a->ncols * c->width;
  if (pl.out_elems == 0) return PSPACE_OK;
  if (pl.nchunk == 0 && peer) return ps_fail(PSPACE_EINVAL, "fused peer projection: this rank's quadrature shard is empty");
  if (pl.nchunk == 0) {   // empty q range: the projection is zero
    if (!a->accumulate) PS_CUDA(cudaMemsetAsync(d_out, 0, pl.out_elems * sizeof(double), st));
    return PSPACE_OK;
  }
  // vector projection: the factored-operand kernel (vproject.cu) takes every TMA-addressable call; the kernels below
  // remain for odd column counts / unaligned operands (generic, in-kernel psi) and as the measured round-1 alternative
  if (to_vf) return ps_vproject(c, a, d_in, d_out, d_ws, ws_bytes, st);
  if (pl.ws_bytes > ws_bytes || (pl.ws_bytes && !d_ws))
    return ps_fail(PSPACE_EINVAL, "workspace of %zu bytes needed, %zu given", pl.ws_bytes, ws_bytes);

  ProjParams p;
  p.in = d_in;
  p.ld_in = ncols_d;
  p.q0 = a->q0; p.q1 = a->q1;
  p.ncols_d = ncols_d;
  p.nvars = c->nvars; p.tsz = c->tsz; p.DP = c->DP;
  p.T = c->d_T; p.deg = c->d_deg; p.qdig = c->d_qdig; p.W = c->d_W;
  for (int d = 0; d < c->nvars; d++) { p.toff[d] = c->toff[d]; p.nq[d] = c->nq[d]; }
  p.tiles_n = pl.tiles_n; p.tiles_j = pl.tiles_j; p.tiles_i = pl.tiles_i;
  p.chunks_per_slice = pl.chunks_per_slice;
  p.vec16 = ((ncols_d % 2) == 0 && (reinterpret_cast<size_t>(d_in) % 16) == 0) ? 1 : 0;
  // Kernel choice.  The TMA stream-K kernel needs a 16-byte addressable operand and costs two panel launches + a fix-up
  // on top of its own; below ~1 GFLOP of window the single-launch kernel (basis values formed in the kernel) is faster
  // (profiles/r02_small_problems.md: 64 x 64 x 64 points x 64 columns 14 us against 37 us).  force_generic: 1 = always
  // the single-launch kernel, -1 = the TMA kernel wherever it can run (tests), 0 = by size.
  p.fast = (p.vec16 && c->force_generic <= 0) ? 1 : 0;
  if (p.fast && c->force_generic == 0 && !c->force_panel && !peer && !(is_matrix && a->symmetric)) {   // symmetric: keep the bitwise mirror
    const double rows = is_matrix ? (double)(a->k1 - a->k0) * (a->j1 - a->j0) : (double)(a->k1 - a->k0);
    if (2.0 * rows * (double)(a->q1 - a->q0) * ncols_d * c->width < 1e9) p.fast = 0;
  }
  if (is_matrix) {
    p.i0 = a->k0; p.ni = a->k1 - a->k0; p.j0 = a->j0; p.nj = a->j1 - a->j0;
    p.stride_j = ncols_d; p.stride_i = (long long)p.nj * ncols_d;
  } else {
    p.i0 = 0; p.ni = 1; p.j0 = a->k0; p.nj = a->k1 - a->k0;
    p.stride_j = ncols_d; p.stride_i = 0;
  }
  p.out = d_out; p.slice_stride = 0; p.accumulate = a->accumulate; p.sk_ws = nullptr;
  p.mask = is_matrix ? a->d_pair_mask : nullptr; p.mask_ld = c->N; p.tile_list = nullptr; p.tile_count = nullptr;
  p.symmetric = 0;
  p.peer_world = 0; p.peer_rank = 0; p.peer_lp = 1; p.peer_ncols_pad = 0; p.lag_ns = 0;
  for (int w = 0; w < PSPACE_MAX_PEERS; w++) p.peer_recv[w] = nullptr;
  if (peer) {
    if (!is_matrix || a->d_pair_mask || a->symmetric || a->accumulate)
      return ps_fail(PSPACE_EINVAL, "fused peer projection: matrix mode without mask / symmetric / accumulate");
    if (!p.fast) return ps_fail(PSPACE_EINVAL, "fused peer projection needs a TMA-addressable operand (even column count, 16-byte aligned)");
    const Variant &pv = MAT_VARIANTS[pl.variant_idx];
    const int npt = pl.tiles_i * pl.tiles_j;
    p.peer_world = peer->world; p.peer_rank = peer->rank;
    p.peer_lp = (npt + peer->world - 1) / peer->world;
    p.peer_ncols_pad = (long long)pl.tiles_n * pv.BN;
    for (int w = 0; w < peer->world; w++) p.peer_recv[w] = peer->recv[w];
  }
  if (is_matrix && a->symmetric) {
    if (a->k0 != a->j0 || a->k1 != a->j1) return ps_fail(PSPACE_EINVAL, "symmetric projection needs a square window (i range == j range)");
    if (a->d_pair_mask) return ps_fail(PSPACE_EINVAL, "symmetric and d_pair_mask cannot be combined");
    // only the TMA kernel implements it; operands it cannot address take the dense generic kernel (same result)
    p.symmetric = p.fast;
  }

  const Variant &v = (is_matrix ? MAT_VARIANTS : VEC_VARIANTS)[pl.variant_idx];
  const int key = (is_matrix ? 0 : 100) + (c->is_complex ? 10 : 0) + v.id;
  switch (key) {
    case 1: rc = launch<M96>(c, p, pl, a, d_out, d_ws, st); break;
    case 2: rc = launch<M64>(c, p, pl, a, d_out, d_ws, st); break;
    case 3: rc = launch<M16>(c, p, pl, a, d_out, d_ws, st); break;
    case 4: rc = launch<M8>(c, p, pl, a, d_out, d_ws, st); break;
    case 5: rc = launch<M96G>(c, p, pl, a, d_out, d_ws, st); break;
    case 6: rc = launch<M64G>(c, p, pl, a, d_out, d_ws, st); break;
    case 11: rc = launch<ZM96>(c, p, pl, a, d_out, d_ws, st); break;
    case 12: rc = launch<ZM64>(c, p, pl, a, d_out, d_ws, st); break;
    case 13: rc = launch<ZM16>(c, p, pl, a, d_out, d_ws, st); break;
    case 14: rc = launch<ZM8>(c, p, pl, a, d_out, d_ws, st); break;
    case 15: rc = launch<ZM96G>(c, p, pl, a, d_out, d_ws, st); break;
    case 16: rc = launch<ZM64G>(c, p, pl, a, d_out, d_ws, st); break;
    case 101: rc = launch<V32>(c, p, pl, a, d_out, d_ws, st); break;
    case 102: rc = launch<V8>(c, p, pl, a, d_out, d_ws, st); break;
    case 103: rc = launch<V24>(c, p, pl, a, d_out, d_ws, st); break;
    case 111: rc = launch<ZV32>(c, p, pl, a, d_out, d_ws, st); break;
    case 112: rc = launch<ZV8>(c, p, pl, a, d_out, d_ws, st); break;
    case 113: rc = launch<ZV24>(c, p, pl, a, d_out, d_ws, st); break;
    default: rc = ps_fail(PSPACE_EINVAL, "no kernel for key %d", key);
  }
  if (rc) return rc;
  // plan record (bench / DESIGN): executed FLOP = 2 * rows * cols * q per real FMA, x2 again for the two
  // DMMA passes of the complex build (= 8 FLOP per complex FMA)
  const double rows = is_matrix ? (double)p.ni * p.nj : (double)p.nj;
  mc->plan_flop = 2.0 * rows * (double)ncols_d * (double)(a->q1 - a->q0) * (c->is_complex ? 2.0 : 1.0);   // useful (all ordered pairs)
  char buf[256];
  snprintf(buf, sizeof buf, "%s%s [%s] tiles=%dx%dx%d chunks=%lld ctas=%d streamk_tiles=%d ksplit(generic)=%d", c->is_complex ? "z " : "",
           v.name, c->plan_kernel, pl.tiles_i, pl.tiles_j, pl.tiles_n,
           (long long)((a->q1 - a->q0 + c->plan_chunk_pts - 1) / c->plan_chunk_pts), pl.sk_grid, pl.sk_rem_tiles, pl.ksplit);
  mc->plan = buf;
  return PSPACE_OK;
}
