// (3a) Vector (residual-type) Galerkin projection, factored-operand kernel.
//
//      F[k][c] = sum_q (w_q psi_k(z_q)) f[q][c]                      TACSStochasticElement.cpp:283-311 (addResidual),
//                                                                    :196-238 (getInitConditions), :487-530 (evalPointQuantity)
//
// the dense (N x Q).(Q x m) contraction with a GENERATED left operand, one pass over f.  Round 1 wrote the weighted
// operand to HBM first (k_psi_panel: 2.4 GB for cfg5, 68.7 GB for cfg4) and streamed it back once — 25x the algorithmic
// traffic and HBM-bound at 0.31 of the FP64 peak.  Here the operand never leaves the SM (see factored.cuh):
//
//   w_q psi_k(z_q) = P_k[q div L] * S_k[q mod L]     with the weighted 1-D tables Tw_d = w_d T_d
//
// and P is constant over a prefix block of L consecutive points, so it leaves the inner loop altogether:
//
//   F[k][c] = sum_blocks P_k[b] * G_b[k][c],      G_b[k][c] = sum_{t in block b} S_k[t] f[b*L + t][c]
//
//   * work is cut into CHUNKS that never cross a block: a block is split at multiples of 64 points inside the block
//     (cfg5 / cfg4: L = 64, one chunk per block), the first and last block of the call's q-range are clipped to it;
//   * S (rows of the CTA's row tile x L tail points) is written to shared memory once per work segment by the producer
//     warps; per chunk they write ONE prefix value per row and a chunk descriptor, next to the f rows that one elected
//     producer thread fetches with TMA (cp.async.bulk.tensor, mbarrier expect_tx, zero-filled ragged edges);
//   * consumer warps: per chunk a fresh accumulator G; the DMMA A fragment is S itself (shared-memory reads at
//     compile-time offsets from one base: no multiply, no index map), B = f; after the chunk acc += P * G
//     (WM*WN*2 DFMA per lane and chunk instead of one DMUL per A element and k-step);
//   * warp grid RG row groups x KG k-groups: with KG = 2 the two warps of an SM sub-partition take alternate chunks and add
//     their accumulators through shared memory at the end of a segment — this is what makes a 96-row tile possible
//     (286 basis rows = 3 x 96 - 2, instead of 3 x 128 - 98);
//   * work = (row tile, column tile, chunk) units cut into G equal contiguous ranges (stream-K, one persistent CTA per
//     SM): full tiles are stored straight to F, partial ones go to a workspace and k_vfixup adds the pieces of each tile
//     in a fixed order (deterministic, no atomics);
//   * a stage is released with an mbarrier arrive whose address depends on the stage's last fragment registers.
//
// Algorithmic work: 2*m FLOP per (q,k) (x4 complex), bytes 8*(Q*m + N*m).  Bound: FP64 tensor pipe (DMMA).
#include "ctx.h"
#include "factored.cuh"
#include <algorithm>
#include <atomic>
#include <stdio.h>

namespace {
using namespace psf;

constexpr int VNPT = 128;      // producer threads (4 warps); consumer warps = RG*KG per tile type
constexpr int VKQ = 64;        // quadrature points per pipeline stage

struct VParams {
  double *out;                 // F rows [nk][ncols_d]
  double *sk_ws;               // stream-K partial tiles [G][2][BM*BN]
  const double *Tw;            // weighted 1-D tables
  const int *deg;
  const uint8_t *qdig;
  long long q0, q1, Qtot;
  int k0, nk, ncols_d, nvars, DP, tsz;
  int toff[PSPACE_MAX_VARS], nq[PSPACE_MAX_VARS];
  int dsplit, L, LP;           // tail = variables [dsplit, nvars); L = prod of their rule sizes; LP = S rows incl. zero pad
  int tiles_r, tiles_n, nchunk;
  int stages, stage_bytes;     // pipeline ring (host-sized)
  int accumulate;
  unsigned zero;               // run-time 0 (stage-release dependency, see mbar_arrive)
};

template <int WM_, int WN_, int RG_, int KG_, bool CPLX_> struct VTile {
  static constexpr int WM = WM_, WN = WN_, RG = RG_, KG = KG_;
  static constexpr bool CPLX = CPLX_;
  static constexpr int W = CPLX ? 2 : 1;
  static constexpr int BM = RG * WM * 8, BN = WN * 8;
  static constexpr int LDB = BN + 4;          // == 4 or 12 (mod 16): conflict-free B fragments
  static constexpr int SP = BM + 4;           // == 4 (mod 16): conflict-free S reads (k = lane%4 -> consecutive tail points)
  static constexpr int B_BYTES = VKQ * LDB * 8;
  static constexpr int NCW = RG * KG, NT = NCW * 32 + VNPT;     // consumer warps, threads per CTA
  static constexpr int CREG = NCW <= 8 ? 216 : (NCW <= 12 ? 136 : 104);   // setmaxnreg of the consumers
  static_assert(RG % 4 == 0 && KG >= 1 && KG <= 4, "row groups spread over the 4 sub-partitions");
  static_assert(LDB % 16 == 4 || LDB % 16 == 12, "B pitch");
  // fixed shared-memory header: S planes, row offsets, accumulator exchange
  __host__ __device__ static size_t s_bytes(int L) { return (size_t)W * L * SP * 8; }
  __host__ __device__ static size_t eo_bytes(int nvars) { return ((size_t)nvars * BM * 2 + 15) & ~(size_t)15; }
  __host__ __device__ static size_t xchg_bytes() { return (size_t)(KG - 1) * RG * 32 * WM * WN * 2 * 8; }
  __host__ __device__ static int stage_bytes() { return (B_BYTES + W * SP * 8 + 16 + 127) / 128 * 128; }
  __host__ __device__ static size_t tab_bytes(int tsz) { return tsz <= psf::TAB_SMEM_MAX ? (size_t)tsz * W * 8 : 0; }
  __host__ __device__ static size_t header_bytes(int L, int nvars, int tsz) {
    return (s_bytes(L) + eo_bytes(nvars) + xchg_bytes() + tab_bytes(tsz) + 127) / 128 * 128;
  }
};

// stream-K decomposition shared by the kernel and the fix-up: CTA g owns whole tiles g, g+G, ... of the first
// floor(ntiles/G) waves and the contiguous range [r0,r1) of the remaining rem_tiles*nchunk (tile, chunk) units
struct SkWork {
  int G, nchunk, ntiles, nfull;
  long long rem_units;
  __host__ __device__ SkWork(int G_, int ntiles_, int nchunk_) : G(G_), nchunk(nchunk_), ntiles(ntiles_) {
    nfull = ntiles / G;
    rem_units = (long long)(ntiles - nfull * G) * nchunk;
  }
  __host__ __device__ long long lo(int g) const { return rem_units * g / G; }
  __host__ __device__ int nseg(int g) const {
    const long long r0 = lo(g), r1 = lo(g + 1);
    return nfull + (r1 > r0 ? 1 : 0) + ((r1 > r0 && r1 > (r0 / nchunk + 1) * (long long)nchunk) ? 1 : 0);
  }
  __host__ __device__ void segment(int g, int sidx, int &tile, int &c0, int &c1, int &slot) const {
    if (sidx < nfull) { tile = sidx * G + g; c0 = 0; c1 = nchunk; slot = -1; return; }
    const long long r0 = lo(g), r1 = lo(g + 1), ta = r0 / nchunk;
    if (sidx == nfull) {
      tile = nfull * G + (int)ta; c0 = (int)(r0 - ta * nchunk);
      c1 = (int)(r1 - ta * nchunk < nchunk ? r1 - ta * nchunk : nchunk); slot = 0;
    } else { tile = nfull * G + (int)ta + 1; c0 = 0; c1 = (int)(r1 - (ta + 1) * nchunk); slot = 1; }
  }
};

// Chunks never cross a prefix block: block b of L points is cut at multiples of 64 inside the block (cpb chunks per
// block); the first and the last block of the call's q-range are clipped to it.  Chunk c of the range -> block, tail range.
struct ChunkMap {
  long long q0, q1, b0, b1;
  int L, cpb, r0, e1;
  __host__ __device__ ChunkMap(long long q0_, long long q1_, int L_) : q0(q0_), q1(q1_), L(L_) {
    cpb = (L + VKQ - 1) / VKQ;
    b0 = q0 / L; r0 = (int)(q0 - b0 * L);
    b1 = q1 > q0 ? (q1 - 1) / L : b0; e1 = q1 > q0 ? (int)((q1 - 1) - b1 * L) + 1 : r0;
  }
  __host__ __device__ long long nchunk() const { return q1 > q0 ? (b1 - b0 + 1) * cpb : 0; }
  // tail range [tlo, thi) of sub-chunk j of block b (may be empty at the clipped ends).  Sub-chunks start at the block's
  // first point INSIDE the range (r0 in the first block) and are 64 long, so only the last one of a block can have a
  // length that is not a multiple of 4 — its last k-step then reads the zero pad of S (block end) or rows of f past the
  // range, which TMA fills with zeros (clipped end); it never reaches into the next sub-chunk.
  __host__ __device__ void range(long long b, int j, int &tlo, int &thi) const {
    tlo = (b == b0 ? r0 : 0) + j * VKQ;
    if (tlo > L) tlo = L;
    thi = tlo + VKQ < L ? tlo + VKQ : L;
    if (b == b1 && thi > e1) thi = e1;
    if (thi < tlo) thi = tlo;
  }
};

// S[t][row] (and the zero pad rows [L, LP)) of one row tile: thread -> (row, part), part p of nthreads/BM covers the tail
// points [p*per, (p+1)*per).
template <bool CPLX>
__device__ __forceinline__ void fill_tail_rows(const VParams &p, const double *T, const unsigned short *eo, double *S, int LP, int SP,
                                            int BM, int nthreads, int tid, int tr) {
  const int nparts = nthreads / BM, row = tid % BM, part = tid / BM;
  if (part >= nparts) return;
  const int per = (p.L + nparts - 1) / nparts, t0 = part * per, n = min(p.L, t0 + per) - t0;
  const bool valid = tr * BM + row < p.nk;
  if (n > 0)
    psf::run_products<CPLX>(T, eo + row, BM, p.qdig, p.DP, t0, n, 1, p.Qtot, p.dsplit, p.nvars, p.nq[p.nvars - 1], valid, S + t0 * SP + row, SP, LP);
  if (part == nparts - 1)
    for (int t = p.L; t < LP; t++) { S[t * SP + row] = 0.0; if (CPLX) S[(LP + t) * SP + row] = 0.0; }
}

template <class TL>
__global__ void __launch_bounds__(TL::NT, 1) k_project_vf(const __grid_constant__ VParams p, const __grid_constant__ CUtensorMap tmapB) {
  constexpr int WM = TL::WM, WN = TL::WN, RG = TL::RG, KG = TL::KG, W = TL::W, BM = TL::BM, BN = TL::BN;
  constexpr int LDB = TL::LDB, SP = TL::SP;
  constexpr bool CPLX = TL::CPLX;
  constexpr bool PAIRED = CPLX && (WN % 2) == 0;
  constexpr int VNCW = TL::NCW, VT = TL::NT;
  // the consumers help to write the tail table in the real build; the complex kernels have no registers to spare for it
  // (measured: -20 % on cfg4 when the fill code shares the consumer's register allocation)
  constexpr bool ALLFILL = !CPLX;

  extern __shared__ __align__(1024) unsigned char smem[];
  const int LP = p.LP;                                                                     // S rows incl. the zero pad
  double *S = reinterpret_cast<double *>(smem);                                             // [W][LP][SP]
  unsigned short *eo = reinterpret_cast<unsigned short *>(smem + TL::s_bytes(LP));          // [nvars][BM]
  double *xchg = reinterpret_cast<double *>(smem + TL::s_bytes(LP) + TL::eo_bytes(p.nvars));
  double *Ts = reinterpret_cast<double *>(smem + TL::s_bytes(LP) + TL::eo_bytes(p.nvars) + TL::xchg_bytes());
  unsigned char *ring = smem + TL::header_bytes(LP, p.nvars, p.tsz);
  const int STAGES = p.stages;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(ring + (size_t)STAGES * p.stage_bytes);
  unsigned long long *empty = full + STAGES;
  unsigned long long *s_full = empty + STAGES, *s_free = s_full + 1, *eo_full = s_full + 2;
  constexpr int P_OFF = TL::B_BYTES, D_OFF = TL::B_BYTES + W * SP * 8;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SkWork wk(gridDim.x, p.tiles_r * p.tiles_n, p.nchunk);
  const ChunkMap cm(p.q0, p.q1, p.L);
  const int gcta = blockIdx.x, nseg = wk.nseg(gcta);

  const bool tab_smem = p.tsz <= TAB_SMEM_MAX;
  if (tab_smem)
    for (int i = tid; i < p.tsz * W; i += VT) Ts[i] = p.Tw[i];
  const double *T = tab_smem ? Ts : p.Tw;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1 + VNPT); mbar_init(&empty[s], RG); }
    mbar_init(s_full, ALLFILL ? VT : VNPT);
    mbar_init(s_free, VNCW);
    mbar_init(eo_full, VNPT);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // The tail table S of a segment's row tile is written by ALL warps (the consumers have nothing else to do until it
  // exists): fill_tail_rows, called before the accumulators of the segment come to life.
  if (warp >= VNCW) {
    // =========================== PRODUCERS: TMA + the factor tables =====================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(56));
    const int pt = tid - VNCW * 32;                  // 0..127 = row of the tile
    if (pt == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmapB) : "memory");
    int s = 0;
    unsigned ph = 0;
    for (int sidx = 0; sidx < nseg; sidx++) {
      int tile, c0, c1, slot;
      wk.segment(gcta, sidx, tile, c0, c1, slot);
      const int tn = tile % p.tiles_n, tr = tile / p.tiles_n;
      const int k = p.k0 + tr * BM + pt;
      const bool valid = pt < BM && k < p.k0 + p.nk;
      // ---- per segment: row offsets and the tail table S of this row tile (rows [L, LP) are the zero pad) ----
      if (sidx > 0) mbar_wait(s_free, (unsigned)(sidx - 1) & 1u);
      if (pt < BM)
        for (int d = 0; d < p.nvars; d++)
          eo[d * BM + pt] = (unsigned short)(valid ? p.toff[d] + __ldg(p.deg + (size_t)k * p.nvars + d) * p.nq[d] : 0);
      if (ALLFILL) {
        mbar_arrive(eo_full);                        // the row offsets of the tile are in place: everyone fills S
        mbar_wait(eo_full, (unsigned)sidx & 1u);
        fill_tail_rows<CPLX>(p, T, eo, S, LP, SP, BM, VT, tid, tr);
      } else {
        fill_tail_rows<CPLX>(p, T, eo, S, LP, SP, BM, VNPT, pt, tr);
      }
      mbar_arrive(s_full);
      // ---- per stage: f rows of the chunk by TMA, the prefix value P of the chunk's block, the chunk descriptor ----
      // block = hi*nql + lo (lo = digit of the last prefix variable); the product over the other prefix variables is
      // cached per hi — it changes once every nql*L points
      const int nql = p.dsplit > 0 ? p.nq[p.dsplit - 1] : 1;
      const int eo_last = p.dsplit > 0 ? (int)eo[(p.dsplit - 1) * BM + (pt < BM ? pt : 0)] : 0;
      long long blk = cm.b0 + c0 / cm.cpb;
      int j = c0 % cm.cpb;
      long long hi = blk / nql;
      int lo = (int)(blk - hi * nql);
      long long c_hi = -1;
      double cbr = 1.0, cbi = 0.0;
      for (int c = c0; c < c1; c++) {
        mbar_wait(&empty[s], ph ^ 1u);
        unsigned char *st = ring + (size_t)s * p.stage_bytes;
        int tlo, thi;
        cm.range(blk, j, tlo, thi);
        if (pt == 0) {
          const long long row = thi > tlo ? blk * cm.L + tlo - p.q0 : 0;     // first f row of the chunk (relative to q0)
          mbar_arrive_expect_tx(&full[s], TL::B_BYTES);
          tma_load_2d(st, &tmapB, &full[s], tn * BN, (int)row);
          *reinterpret_cast<int2 *>(st + D_OFF) = make_int2(tlo * SP, (thi - tlo + 3) >> 2);   // S row offset, k-steps
        }
        if (pt < BM) {
          const long long qb = blk * cm.L;                                   // first point of the block
          double vr = 0.0, vi = 0.0;
          if (valid && qb < p.Qtot) {
            if (p.dsplit > 0) {
              if (hi != c_hi) {
                const long long qh = hi * nql * cm.L;
                cbr = 1.0; cbi = 0.0;
                table_product<CPLX>(T, eo + pt, BM, p.qdig + (size_t)qh * p.DP, 0, p.dsplit - 1, cbr, cbi);
                c_hi = hi;
                const long long qn = qh + (long long)nql * cm.L;             // digit row of the next run: into L1 ahead of use
                if (qn < p.Qtot) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p.qdig + (size_t)qn * p.DP));
              }
              table_mul<CPLX>(T, eo_last + lo, cbr, cbi, vr, vi);
            } else vr = 1.0;
          }
          double *Pst = reinterpret_cast<double *>(st + P_OFF);
          Pst[pt] = vr;
          if (CPLX) Pst[SP + pt] = vi;
        }
        mbar_arrive(&full[s]);
        if (++j == cm.cpb) { j = 0; blk++; if (++lo == nql) { lo = 0; hi++; } }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // =========================== CONSUMERS ===============================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(TL::CREG));
    const int rg = warp % RG, kg = warp / RG;
    const int g = lane >> 2, kl = lane & 3;
    const unsigned sgn = (g & 1) ? 0u : 0x80000000u;
    const int rbase = rg * WM * 8 + g;
    auto acc_pair = [&](const double (&ac)[WM][WN][2], int a, int idx, double &v0, double &v1, int &cc) {
      if (PAIRED) { const int b2 = idx >> 1, e = idx & 1; v0 = ac[a][2 * b2][e]; v1 = ac[a][2 * b2 + (WN > 1 ? 1 : 0)][e]; cc = 16 * b2 + 4 * kl + 2 * e; }
      else { v0 = ac[a][idx][0]; v1 = ac[a][idx][1]; cc = 8 * idx + 2 * kl; }
    };
    int s = 0, it = 0;
    unsigned ph = 0;
    for (int sidx = 0; sidx < nseg; sidx++) {
      int tile, c0, c1, slot;
      wk.segment(gcta, sidx, tile, c0, c1, slot);
      const int tn = tile % p.tiles_n, tr = tile / p.tiles_n;
      // every consumer warp has left the previous segment (s_free) and the row offsets exist (eo_full): write my share of S
      if (ALLFILL) {
        if (sidx > 0) mbar_wait(s_free, (unsigned)(sidx - 1) & 1u);
        mbar_wait(eo_full, (unsigned)sidx & 1u);
        fill_tail_rows<CPLX>(p, T, eo, S, LP, SP, BM, VT, tid, tr);
        mbar_arrive(s_full);
      }
      mbar_wait(s_full, (unsigned)sidx & 1u);
      double acc[WM][WN][2];
#pragma unroll
      for (int a = 0; a < WM; a++)
#pragma unroll
        for (int b = 0; b < WN; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

      for (int c = c0; c < c1; c++, it++) {
        const bool mine = KG == 1 || (it % KG) == kg;
        if (mine) {
          mbar_wait(&full[s], ph);
          const unsigned char *st = ring + (size_t)s * p.stage_bytes;
          const int2 desc = *reinterpret_cast<const int2 *>(st + D_OFF);
          const double *bs = reinterpret_cast<const double *>(st) + kl * LDB + g;
          const double *bs2 = reinterpret_cast<const double *>(st) + kl * LDB + 2 * g;
          const double *Sr = S + desc.x + kl * SP + rbase;                    // S rows tlo + k of my rows
          const double *Si = Sr + (size_t)LP * SP;
          // G = sum over the chunk's points of S (x) f — the block's prefix value P is applied once, below
          double G[WM][WN][2];
#pragma unroll
          for (int a = 0; a < WM; a++)
#pragma unroll
            for (int b = 0; b < WN; b++) G[a][b][0] = G[a][b][1] = 0.0;
          int depi = 0;
          auto kstep = [&](int ks) {
            double are[WM], aim[WM];
#pragma unroll
            for (int a = 0; a < WM; a++) {
              are[a] = Sr[ks * 4 * SP + a * 8];
              if (CPLX) aim[a] = Si[ks * 4 * SP + a * 8];
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
                dmma884(G[a][b][0], G[a][b][1], are[a], bfr[b]);
                if (CPLX) dmma884(G[a][b][0], G[a][b][1], aim[a], bsw[b]);
              }
            depi = __double2loint(bfr[WN - 1]) ^ __double2loint(are[WM - 1]);
          };
          if (desc.y == VKQ / 4) {
#pragma unroll
            for (int ks = 0; ks < VKQ / 4; ks++) kstep(ks);
          } else {
#pragma unroll 2
            for (int ks = 0; ks < desc.y; ks++) kstep(ks);
          }
          // acc += P (x) G : one prefix value per basis row of the lane
          const double *Pst = reinterpret_cast<const double *>(st + P_OFF) + rbase;
#pragma unroll
          for (int a = 0; a < WM; a++) {
            const double pr = Pst[a * 8];
            if (!CPLX) {
#pragma unroll
              for (int b = 0; b < WN; b++) { acc[a][b][0] = fma(pr, G[a][b][0], acc[a][b][0]); acc[a][b][1] = fma(pr, G[a][b][1], acc[a][b][1]); }
            } else {
              const double pi = Pst[SP + a * 8];
#pragma unroll
              for (int idx = 0; idx < WN; idx++) {
                // (re, im) of one complex column = the two values acc_pair hands out together
                double *gr, *gi, *ar, *ai;
                if (PAIRED) { const int b2 = idx >> 1, e = idx & 1; gr = &G[a][2 * b2][e]; gi = &G[a][2 * b2 + 1][e]; ar = &acc[a][2 * b2][e]; ai = &acc[a][2 * b2 + 1][e]; }
                else { gr = &G[a][idx][0]; gi = &G[a][idx][1]; ar = &acc[a][idx][0]; ai = &acc[a][idx][1]; }
                *ar += pr * *gr - pi * *gi;
                *ai += pr * *gi + pi * *gr;
              }
            }
          }
          // release the stage: the arrive's address depends on the last fragment registers read from it
          const unsigned dep = ((unsigned)depi ^ (unsigned)__double2loint(Pst[0])) & p.zero;
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s], dep);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }

      // ---- end of segment -------------------------------------------------------------------------
      if (KG > 1) {
        // the k-groups hold partial sums of the same tile: groups 1.. hand their accumulators to group 0 (added in group order)
        if (kg > 0) {
          double *x = xchg + (((size_t)(kg - 1) * RG + rg) * 32 + lane) * (WM * WN * 2);
#pragma unroll
          for (int a = 0; a < WM; a++)
#pragma unroll
            for (int b = 0; b < WN; b++) { x[(a * WN + b) * 2] = acc[a][b][0]; x[(a * WN + b) * 2 + 1] = acc[a][b][1]; }
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(VNCW * 32) : "memory");
        if (kg == 0) {
#pragma unroll
          for (int k2 = 1; k2 < KG; k2++) {
            const double *x = xchg + (((size_t)(k2 - 1) * RG + rg) * 32 + lane) * (WM * WN * 2);
#pragma unroll
            for (int a = 0; a < WM; a++)
#pragma unroll
              for (int b = 0; b < WN; b++) { acc[a][b][0] += x[(a * WN + b) * 2]; acc[a][b][1] += x[(a * WN + b) * 2 + 1]; }
          }
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(VNCW * 32) : "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);            // this warp no longer reads S of the segment
      if (KG > 1 && kg > 0) continue;
      if (slot >= 0) {
        double *wsp = p.sk_ws + ((size_t)gcta * 2 + slot) * (BM * BN);
#pragma unroll
        for (int a = 0; a < WM; a++)
#pragma unroll
          for (int b = 0; b < WN; b++) {
            double v0, v1; int cc;
            acc_pair(acc, a, b, v0, v1, cc);
            *reinterpret_cast<double2 *>(wsp + (size_t)(rbase + a * 8) * BN + cc) = make_double2(v0, v1);
          }
      } else {
#pragma unroll
        for (int a = 0; a < WM; a++) {
          const int r = tr * BM + rbase + a * 8;
          if (r >= p.nk) continue;
#pragma unroll
          for (int b = 0; b < WN; b++) {
            double v0, v1; int cc;
            acc_pair(acc, a, b, v0, v1, cc);
            const int col = tn * BN + cc;
            double *dst = p.out + (size_t)r * p.ncols_d + col;
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
}

// Adds the stream-K pieces of each remainder tile and stores the tile.  grid = (remainder tiles, element slices); a block
// owns 64 elements x 4 piece lanes: lane j adds the pieces j, j+4, j+8, ... in CTA order and the four partial sums are
// combined as (s0 + s1) + (s2 + s3) — a fixed order, so the result is reproducible run to run.
constexpr int VFX_ELEMS = 32, VFX_LANES = 256 / VFX_ELEMS;
__global__ void __launch_bounds__(256) k_vfixup(const __grid_constant__ VParams p, int G, int BM, int BN) {
  const SkWork wk(G, p.tiles_r * p.tiles_n, p.nchunk);
  const int rt = blockIdx.x, tile = wk.nfull * G + rt;
  const int tn = tile % p.tiles_n, tr = tile / p.tiles_n;
  const long long u0 = (long long)rt * wk.nchunk, u1 = u0 + wk.nchunk;
  __shared__ int s_src[1024];
  __shared__ short s_flag[1024];     // per CTA of the projection: 0 = no piece of this tile, 1 = its slot 0, 2 = its slot 1
  __shared__ int s_n;
  __shared__ double part[VFX_LANES][VFX_ELEMS];
  // pieces of this tile: every thread tests CTAs (one 64-bit division each, in parallel); thread 0 then compacts the flags in
  // CTA order, which fixes the order of the sum (the serial search with two divisions per step took most of this kernel's time)
  for (int g = threadIdx.x; g < G && g < 1024; g += blockDim.x) {
    const long long a0 = wk.lo(g), a1 = wk.lo(g + 1);
    short f = 0;
    if (a1 > a0 && a0 < u1 && a1 > u0) {
      // the CTA's range [a0, a1) touches at most two tiles: slot 0 = the tile holding a0, slot 1 = the next one
      f = (a0 / wk.nchunk == rt) ? 1 : 2;
    }
    s_flag[g] = f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int g = 0; g < G && g < 1024; g++)
      if (s_flag[g]) s_src[n++] = g * 2 + (s_flag[g] - 1);
    s_n = n;
  }
  __syncthreads();
  const int npieces = s_n, el = threadIdx.x & (VFX_ELEMS - 1), pl = threadIdx.x / VFX_ELEMS;
  const int e = blockIdx.y * VFX_ELEMS + el;
  double s0 = 0.0, s1 = 0.0;
  if (e < BM * BN) {
    const size_t tsz = (size_t)BM * BN;
    int k = pl;
    for (; k + VFX_LANES < npieces; k += 2 * VFX_LANES) {       // two independent loads in flight per step
      const double v0 = p.sk_ws[(size_t)s_src[k] * tsz + e], v1 = p.sk_ws[(size_t)s_src[k + VFX_LANES] * tsz + e];
      s0 += v0; s1 += v1;
    }
    if (k < npieces) s0 += p.sk_ws[(size_t)s_src[k] * tsz + e];
  }
  part[pl][el] = s0 + s1;
  __syncthreads();
  if (pl == 0 && e < BM * BN) {
    const int r = tr * BM + e / BN, col = tn * BN + e % BN;
    if (r < p.nk && col < p.ncols_d) {
      double tot = 0.0;
#pragma unroll
      for (int l = 0; l < VFX_LANES; l += 2) tot += part[l][el] + part[l + 1][el];      // fixed order: deterministic
      double *dst = p.out + (size_t)r * p.ncols_d + col;
      *dst = p.accumulate ? *dst + tot : tot;
    }
  }
}

struct VVariant { int id, BM, BN; const char *name; };
//                      WM WN RG KG
typedef VTile<2, 1, 8, 1, false> V128x8;
typedef VTile<2, 3, 8, 1, false> V128x24;
typedef VTile<2, 4, 8, 1, false> V128x32;
typedef VTile<2, 6, 8, 1, false> V128x48;
typedef VTile<3, 1, 4, 2, false> V96x8;
typedef VTile<3, 3, 4, 2, false> V96x24;
typedef VTile<3, 4, 4, 2, false> V96x32;
typedef VTile<3, 6, 4, 2, false> V96x48;
typedef VTile<3, 3, 4, 3, false> V96x24k3;
typedef VTile<3, 3, 4, 4, false> V96x24k4;
typedef VTile<2, 3, 8, 2, false> V128x24k2;
typedef VTile<2, 2, 8, 1, true> ZV128x16;
typedef VTile<2, 4, 8, 1, true> ZV128x32;
typedef VTile<2, 6, 8, 1, true> ZV128x48;
typedef VTile<3, 2, 4, 2, true> ZV96x16;
typedef VTile<3, 4, 4, 2, true> ZV96x32;
typedef VTile<3, 6, 4, 2, true> ZV96x48;
const VVariant V_REAL[] = {{1, 128, 8, "128x8"}, {2, 128, 24, "128x24"}, {3, 128, 32, "128x32"}, {4, 128, 48, "128x48"},
                           {5, 96, 8, "96x8 (2 k-groups)"}, {6, 96, 24, "96x24 (2 k-groups)"}, {7, 96, 32, "96x32 (2 k-groups)"},
                           {8, 96, 48, "96x48 (2 k-groups)"}, {9, 96, 24, "96x24 (3 k-groups)"}, {10, 96, 24, "96x24 (4 k-groups)"},
                           {11, 128, 24, "128x24 (2 k-groups)"}};
const VVariant V_CPLX[] = {{1, 128, 16, "128x16"}, {2, 128, 32, "128x32"}, {3, 128, 48, "128x48"},
                           {4, 96, 16, "96x16 (2 k-groups)"}, {5, 96, 32, "96x32 (2 k-groups)"}, {6, 96, 48, "96x48 (2 k-groups)"}};

struct VPlan {
  int vi;                        // index into V_REAL / V_CPLX
  int tiles_r, tiles_n, nchunk, grid, rem_tiles;
  int dsplit, L, LP, stages, stage_bytes;
  size_t smem, ws_bytes;
};

int sm_count_v(int device) {
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

// tile variant (fewest column tiles first — every column tile regenerates the operand —, then the least padded work) and
// the workspace bound, both independent of the tail length
int plan_v(const pspace_ctx *c, const pspace_project_args *a, VPlan &pl) {
  const int ncols_d = a->ncols * c->width, nk = a->k1 - a->k0;
  const VVariant *vs = c->is_complex ? V_CPLX : V_REAL;
  const int nv = c->is_complex ? 6 : 8, nv_all = c->is_complex ? 6 : 11;
  int best = -1;
  if (a->variant > 0) {
    for (int i = 0; i < nv_all; i++) if (vs[i].id == a->variant) best = i;
    if (best < 0) return ps_fail(PSPACE_EINVAL, "unknown vector tile variant %d", a->variant);
  } else {
    long long best_tn = 0, best_cost = 0;
    for (int i = 0; i < nv; i++) {
      const long long tn = (ncols_d + vs[i].BN - 1) / vs[i].BN, trr = (nk + vs[i].BM - 1) / vs[i].BM;
      const long long cost = tn * vs[i].BN * trr * vs[i].BM;
      if (best < 0 || tn < best_tn || (tn == best_tn && cost < best_cost)) { best = i; best_tn = tn; best_cost = cost; }
    }
  }
  pl.vi = best;
  pl.tiles_n = (ncols_d + vs[best].BN - 1) / vs[best].BN;
  pl.tiles_r = (nk + vs[best].BM - 1) / vs[best].BM;
  // stream-K partial tiles: 2 per persistent CTA (an upper bound that does not depend on the chunking)
  pl.ws_bytes = (size_t)sm_count_v(c->device) * 2 * vs[best].BM * vs[best].BN * sizeof(double);
  return PSPACE_OK;
}

// Tail length L = a product of trailing rule sizes: the candidate with the longest average chunk (chunks are cut at
// multiples of 64 inside a prefix block, so avg = L / ceil(L/64)) whose tail table leaves room for 3 pipeline stages.
template <class TL> int size_ring(const pspace_ctx *c, const pspace_project_args *a, VPlan &pl) {
  const size_t cap = (size_t)227 * 1024 - 1024 - 256;
  int best_ds = -1;
  long long best_L = 1, L = 1;
  double best_avg = -1.0;
  for (int ds = c->nvars - 1; ds >= 0; ds--) {
    L *= c->nq[ds];
    if (L > 8192) break;
    // zero pad rows after the table: the last k-step of a block's last chunk reads up to 3 rows past L - 1 unless both L
    // and the range's offset into its first block are multiples of 4 (then only a clipped END is ragged, and there the
    // extra rows of f lie past the range and come back as zeros from TMA)
    const int pad = ((L % 4) != 0 || ((a->q0 % L) % 4) != 0) ? 4 : 0;
    const int LP = (int)L + pad, cpb = (int)((L + VKQ - 1) / VKQ);
    const size_t need = TL::header_bytes(LP, c->nvars, c->tsz) + 3 * (size_t)TL::stage_bytes() + 256;
    if (need > cap) break;
    const double avg = (double)L / cpb;
    if (avg > best_avg + 1e-9) { best_avg = avg; best_ds = ds; best_L = L; }
  }
  if (best_ds < 0) return ps_fail(PSPACE_EINVAL, "vector projection: the tail table of one rule does not fit shared memory");
  pl.dsplit = best_ds; pl.L = (int)best_L;
  pl.LP = (int)best_L + (((best_L % 4) != 0 || ((a->q0 % best_L) % 4) != 0) ? 4 : 0);
  pl.stage_bytes = TL::stage_bytes();
  const size_t head = TL::header_bytes(pl.LP, c->nvars, c->tsz);
  pl.stages = (int)std::max<size_t>(2, std::min<size_t>(8, (cap - head - 256) / pl.stage_bytes));
  pl.smem = head + (size_t)pl.stages * pl.stage_bytes + (2 * pl.stages + 3) * 8 + 1024;
  const long long nch = ChunkMap(a->q0, a->q1, pl.L).nchunk();
  if (nch > 0x7fffffffLL) return ps_fail(PSPACE_EINVAL, "vector projection: %lld chunks", nch);
  pl.nchunk = (int)nch;
  const long long ntiles = (long long)pl.tiles_r * pl.tiles_n;
  pl.grid = (int)std::max<long long>(1, std::min<long long>(sm_count_v(c->device), ntiles * pl.nchunk));
  pl.rem_tiles = (int)(ntiles % pl.grid);
  return PSPACE_OK;
}

template <class TL> int launch_v(const pspace_ctx *c, VParams &p, const VPlan &pl, const double *d_in, cudaStream_t st) {
  CUtensorMap tmB;
  {
    unsigned long long dims[2] = {(unsigned long long)p.ncols_d, (unsigned long long)(p.q1 - p.q0)};
    unsigned long long str[1] = {(unsigned long long)p.ncols_d * 8};
    unsigned box[2] = {(unsigned)TL::LDB, (unsigned)VKQ};
    int rc = ps_make_tmap(&tmB, d_in, 2, dims, str, box);
    if (rc) return rc;
  }
  if (pl.smem > (size_t)227 * 1024) return ps_fail(PSPACE_EINVAL, "vector projection: %zu bytes of shared memory", pl.smem);
  static std::atomic<bool> attr_dev[64];        // per instantiation and device; setting it twice is harmless
  if (!attr_dev[c->device & 63].load(std::memory_order_acquire)) {
    PS_CUDA(cudaFuncSetAttribute(k_project_vf<TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_dev[c->device & 63].store(true, std::memory_order_release);
  }
  k_project_vf<TL><<<pl.grid, TL::NT, pl.smem, st>>>(p, tmB);
  PS_LAUNCH_CHECK(c);
  if (pl.rem_tiles > 0) {
    k_vfixup<<<dim3(pl.rem_tiles, (TL::BM * TL::BN + VFX_ELEMS - 1) / VFX_ELEMS), 256, 0, st>>>(p, pl.grid, TL::BM, TL::BN);
    PS_LAUNCH_CHECK(c);
  }
  return PSPACE_OK;
}
}  // namespace

// workspace of the factored vector projection (0 when every tile is a full one)
size_t ps_vproject_workspace(const pspace_ctx *c, const pspace_project_args *a) {
  VPlan pl;
  if (a->k1 <= a->k0 || a->q1 <= a->q0 || plan_v(c, a, pl) != PSPACE_OK) return 0;
  return pl.ws_bytes;
}

// true when the factored kernel can take the call (TMA-addressable f: even real column count, 16-byte aligned base)
bool ps_vproject_usable(const pspace_ctx *c, const pspace_project_args *a, const double *d_in) {
  const int ncols_d = a->ncols * c->width;
  return !c->force_generic && !c->force_panel && (ncols_d % 2) == 0 && (reinterpret_cast<size_t>(d_in) % 16) == 0 && c->d_Tw &&
         c->tsz + PSPACE_MAX_NQ < 65536;
}

int ps_vproject(const pspace_ctx *c, const pspace_project_args *a, const double *d_in, double *d_out, void *d_ws,
                size_t ws_bytes, cudaStream_t st) {
  VPlan pl;
  int rc = plan_v(c, a, pl);
  if (rc) return rc;
  if (pl.ws_bytes > ws_bytes || (pl.ws_bytes && !d_ws))
    return ps_fail(PSPACE_EINVAL, "workspace of %zu bytes needed, %zu given", pl.ws_bytes, ws_bytes);
  VParams p;
  p.out = d_out; p.sk_ws = (double *)d_ws;
  p.Tw = c->d_Tw; p.deg = c->d_deg; p.qdig = c->d_qdig;
  p.q0 = a->q0; p.q1 = a->q1; p.Qtot = c->Q;
  p.k0 = a->k0; p.nk = a->k1 - a->k0; p.ncols_d = a->ncols * c->width; p.nvars = c->nvars; p.DP = c->DP; p.tsz = c->tsz;
  for (int d = 0; d < PSPACE_MAX_VARS; d++) { p.toff[d] = d < c->nvars ? c->toff[d] : 0; p.nq[d] = d < c->nvars ? c->nq[d] : 1; }
  p.tiles_r = pl.tiles_r; p.tiles_n = pl.tiles_n;
  p.accumulate = a->accumulate; p.zero = 0;
  const VVariant &v = (c->is_complex ? V_CPLX : V_REAL)[pl.vi];
  const int key = (c->is_complex ? 100 : 0) + v.id;
#define PS_VCASE(K, TL)                                                                              \
  case K:                                                                                            \
    if ((rc = size_ring<TL>(c, a, pl))) break;                                                       \
    p.dsplit = pl.dsplit; p.L = pl.L; p.LP = pl.LP; p.stages = pl.stages; p.stage_bytes = pl.stage_bytes; \
    p.nchunk = pl.nchunk;                                                                            \
    rc = launch_v<TL>(c, p, pl, d_in, st);                                                           \
    break;
  switch (key) {
    PS_VCASE(1, V128x8) PS_VCASE(2, V128x24) PS_VCASE(3, V128x32) PS_VCASE(4, V128x48)
    PS_VCASE(5, V96x8) PS_VCASE(6, V96x24) PS_VCASE(7, V96x32) PS_VCASE(8, V96x48)
    PS_VCASE(9, V96x24k3) PS_VCASE(10, V96x24k4) PS_VCASE(11, V128x24k2)
    PS_VCASE(101, ZV128x16) PS_VCASE(102, ZV128x32) PS_VCASE(103, ZV128x48)
    PS_VCASE(104, ZV96x16) PS_VCASE(105, ZV96x32) PS_VCASE(106, ZV96x48)
    default: rc = ps_fail(PSPACE_EINVAL, "no vector kernel for key %d", key);
  }
#undef PS_VCASE
  if (rc) return rc;
  pspace_ctx *mc = const_cast<pspace_ctx *>(c);
  mc->plan_kernel = "factored operand (tail table in shared memory, prefix applied per block), tma + mbarrier warp-specialised, persistent stream-K";
  mc->plan_chunk_pts = VKQ;
  mc->plan_flop = 2.0 * (double)p.nk * (double)p.ncols_d * (double)(a->q1 - a->q0) * (c->is_complex ? 2.0 : 1.0);
  char buf[320];
  snprintf(buf, sizeof buf, "%sdmma vec %s [%s] tiles=%dx%d chunks=%d ctas=%d streamk_tiles=%d tail=%d vars (L=%d) stages=%d scratch=%zu",
           c->is_complex ? "z " : "", v.name, c->plan_kernel, pl.tiles_r, pl.tiles_n, pl.nchunk, pl.grid, pl.rem_tiles,
           c->nvars - pl.dsplit, pl.L, pl.stages, pl.ws_bytes);
  mc->plan = buf;
  return PSPACE_OK;
}
