// (1) Multi-index basis construction — integer kernel, bit-exact with the reference.
//
// Replaces BasisHelper::basisDegrees (reference src/cpp/BasisHelper.cpp:33-66) and its ten hand-unrolled
// variants (:98-354 tensor, :393-584 complete) plus the ArrayList buckets (ArrayList.cpp:41-58).
// Rule restated for any nvars: enumerate tuples (a_0..a_{n-1}), 0 <= a_d <= pmax[d], row-major with the
// LAST variable fastest (:329-339); the complete space keeps sum(a) <= max_d pmax[d] (:375-383,:561-565);
// kept tuples are appended to the bucket of their total degree in encounter order; buckets are
// concatenated by increasing degree (:342-347).  I.e. output row of tuple t = offset[deg(t)] +
// #{kept t' < t with deg(t') = deg(t)}: a STABLE multi-bucket partition of the enumeration.
//
// B200 mapping: the enumeration index space [0, prod(1+pmax)) is cut into contiguous blocks of
// TILE = 256 threads x ITEMS tuples; pass 1 builds a per-block histogram over degrees, a tiny scan turns it
// into per-(block, degree) bases, pass 2 recomputes the tuples and writes each kept one to
// base + stable intra-block rank (warp match/ballot ranks + per-warp running counters).  Work is pure
// integer, HBM traffic = 4*nvars*N bytes written; for cfg5 (4^10 tuples -> 286 rows) it is microseconds.
#include "ctx.h"

namespace {
constexpr int NT = 256;            // threads per block
constexpr int ITEMS = 16;          // consecutive 32-tuple warp-steps per warp
constexpr int MAXDEG = PSPACE_MAX_VARS * PSPACE_MAX_NQ + 1;   // 641 > any sum of degrees

struct EnumParams {
  int nvars, complete, maxp, ndeg;
  int radix[PSPACE_MAX_VARS];      // 1 + pmax[d]
  long long total;                 // prod radix
};

__device__ __forceinline__ int decode(const EnumParams &p, long long t, int *tuple) {
  int s = 0;
  for (int d = p.nvars - 1; d >= 0; d--) {
    int r = p.radix[d];
    int a = (int)(t % r);
    t /= r;
    tuple[d] = a;
    s += a;
  }
  return s;
}

// tuples of one block: warp w owns [blk*TILE + w*32*ITEMS, +32*ITEMS), step i covers 32 consecutive tuples,
// so (warp, step, lane) order == enumeration order.
__global__ void __launch_bounds__(NT) k_count(EnumParams p, unsigned *blk_hist /*[nblk][ndeg]*/) {
  extern __shared__ unsigned hist[];   // [ndeg]
  for (int i = threadIdx.x; i < p.ndeg; i += NT) hist[i] = 0;
  __syncthreads();
  const long long tile = (long long)NT * ITEMS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long base = blockIdx.x * tile + (long long)warp * 32 * ITEMS;
  int tuple[PSPACE_MAX_VARS];
  for (int it = 0; it < ITEMS; it++) {
    long long t = base + it * 32 + lane;
    if (t < p.total) {
      int s = decode(p, t, tuple);
      if (!p.complete || s <= p.maxp) atomicAdd(&hist[s], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.ndeg; i += NT) blk_hist[(size_t)blockIdx.x * p.ndeg + i] = hist[i];
}

// exclusive scan: first over blocks within a degree, then degree offsets.  One block; ndeg*nblk is small
// next to the enumeration itself (nblk = total/4096).
__global__ void k_scan(int ndeg, int nblk, unsigned *blk_hist, unsigned *deg_off /*[ndeg+1]*/) {
  __shared__ unsigned tot[MAXDEG + 1];
  for (int s = threadIdx.x; s < ndeg; s += blockDim.x) {
    unsigned run = 0;
    for (int b = 0; b < nblk; b++) {
      unsigned c = blk_hist[(size_t)b * ndeg + s];
      blk_hist[(size_t)b * ndeg + s] = run;
      run += c;
    }
    tot[s] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int s = 0; s < ndeg; s++) { deg_off[s] = run; run += tot[s]; }
    deg_off[ndeg] = run;
  }
}

__global__ void __launch_bounds__(NT) k_scatter(EnumParams p, const unsigned *blk_base, const unsigned *deg_off,
                                               int *out, long long capacity_rows) {
  extern __shared__ unsigned sm[];            // warp_cnt[8][ndeg]
  unsigned *warp_cnt = sm;
  const int nwarp = NT / 32;
  for (int i = threadIdx.x; i < nwarp * p.ndeg; i += NT) warp_cnt[i] = 0;
  __syncthreads();
  const long long tile = (long long)NT * ITEMS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long base = blockIdx.x * tile + (long long)warp * 32 * ITEMS;
  int tuple[PSPACE_MAX_VARS];
  // pass A: per-warp totals per degree
  for (int it = 0; it < ITEMS; it++) {
    long long t = base + it * 32 + lane;
    int s = -1;
    if (t < p.total) { s = decode(p, t, tuple); if (p.complete && s > p.maxp) s = -1; }
    unsigned peers = __match_any_sync(0xffffffffu, s);
    if (s >= 0 && (peers & ((1u << lane) - 1)) == 0) warp_cnt[warp * p.ndeg + s] += __popc(peers);  // leader adds
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over warps per degree (in place)
  for (int s = threadIdx.x; s < p.ndeg; s += NT) {
    unsigned run = 0;
    for (int w = 0; w < nwarp; w++) { unsigned c = warp_cnt[w * p.ndeg + s]; warp_cnt[w * p.ndeg + s] = run; run += c; }
  }
  __syncthreads();
  // pass B: scatter with stable ranks
  for (int it = 0; it < ITEMS; it++) {
    long long t = base + it * 32 + lane;
    int s = -1;
    if (t < p.total) { s = decode(p, t, tuple); if (p.complete && s > p.maxp) s = -1; }
    unsigned peers = __match_any_sync(0xffffffffu, s);
    unsigned before = __popc(peers & ((1u << lane) - 1));
    unsigned wbase = 0;
    if (s >= 0) wbase = warp_cnt[warp * p.ndeg + s];
    __syncwarp();
    if (s >= 0) {
      long long row = (long long)deg_off[s] + blk_base[(size_t)blockIdx.x * p.ndeg + s] + wbase + before;
      if (row < capacity_rows)
        for (int d = 0; d < p.nvars; d++) out[row * p.nvars + d] = tuple[d];
      if (before == 0) warp_cnt[warp * p.ndeg + s] = wbase + __popc(peers);   // leader advances the warp cursor
    }
    __syncwarp();
  }
}

int make_params(int nvars, const int *pmax, int basis_type, EnumParams &p) {
  if (nvars < 1 || nvars > PSPACE_MAX_VARS) return ps_fail(PSPACE_EINVAL, "nvars %d outside 1..%d", nvars, PSPACE_MAX_VARS);
  p.nvars = nvars;
  p.total = 1;
  p.maxp = 0;
  int sump = 0;
  for (int d = 0; d < nvars; d++) {
    if (pmax[d] < 0 || pmax[d] >= PSPACE_MAX_NQ) return ps_fail(PSPACE_EINVAL, "pmax[%d]=%d outside 0..%d", d, pmax[d], PSPACE_MAX_NQ - 1);
    p.radix[d] = pmax[d] + 1;
    if (p.total > (1LL << 40) / p.radix[d]) return ps_fail(PSPACE_EINVAL, "enumeration space prod(1+pmax) exceeds 2^40");
    p.total *= p.radix[d];
    sump += pmax[d];
    if (pmax[d] > p.maxp) p.maxp = pmax[d];
  }
  // uniVariateBasisDegrees ignores the basis type (BasisHelper.cpp:40-41)
  p.complete = (basis_type != 0 && nvars > 1) ? 1 : 0;
  p.ndeg = (p.complete ? p.maxp : sump) + 1;
  return PSPACE_OK;
}
}  // namespace

// Number of terms without enumerating on the device: exact DP over variables on the host
// (count of tuples with given total degree) — used only to size buffers.
int ps_basis_count_host(int nvars, const int *pmax, int basis_type, long long *nbasis) {
  EnumParams p;
  int rc = make_params(nvars, pmax, basis_type, p);
  if (rc) return rc;
  int sump = 0;
  for (int d = 0; d < nvars; d++) sump += pmax[d];
  std::vector<long long> cnt(sump + 1, 0), nxt(sump + 1, 0);
  cnt[0] = 1;
  for (int d = 0; d < nvars; d++) {
    std::fill(nxt.begin(), nxt.end(), 0);
    for (int s = 0; s <= sump; s++)
      if (cnt[s]) for (int a = 0; a <= pmax[d] && s + a <= sump; a++) nxt[s + a] += cnt[s];
    cnt.swap(nxt);
  }
  long long tot = 0;
  for (int s = 0; s <= sump; s++) if (!p.complete || s <= p.maxp) tot += cnt[s];
  *nbasis = tot;
  return PSPACE_OK;
}

int ps_basis_degrees_device(int nvars, const int *pmax, int basis_type, int *nbasis, int *d_out,
                            long long capacity_rows, cudaStream_t st, long long *launches) {
  EnumParams p;
  int rc = make_params(nvars, pmax, basis_type, p);
  if (rc) return rc;
  const long long tile = (long long)NT * ITEMS;
  long long nblk_ll = (p.total + tile - 1) / tile;
  if (nblk_ll > 0x7fffffffLL) return ps_fail(PSPACE_EINVAL, "enumeration too large");
  int nblk = (int)nblk_ll;
  unsigned *d_hist = nullptr, *d_off = nullptr;
  PS_CUDA(cudaMallocAsync(&d_hist, sizeof(unsigned) * (size_t)nblk * p.ndeg, st));
  PS_CUDA(cudaMallocAsync(&d_off, sizeof(unsigned) * (p.ndeg + 1), st));
  k_count<<<nblk, NT, sizeof(unsigned) * p.ndeg, st>>>(p, d_hist);
  PS_LAUNCH_CHECK(nullptr);
  k_scan<<<1, 256, 0, st>>>(p.ndeg, nblk, d_hist, d_off);
  PS_LAUNCH_CHECK(nullptr);
  k_scatter<<<nblk, NT, sizeof(unsigned) * (NT / 32) * p.ndeg, st>>>(p, d_hist, d_off, d_out, capacity_rows);
  PS_LAUNCH_CHECK(nullptr);
  if (launches) *launches += 3;
  unsigned total = 0;
  PS_CUDA(cudaMemcpyAsync(&total, d_off + p.ndeg, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  PS_CUDA(cudaFreeAsync(d_hist, st));
  PS_CUDA(cudaFreeAsync(d_off, st));
  *nbasis = (int)total;
  if ((long long)total > capacity_rows) return ps_fail(PSPACE_EINVAL, "degree table needs %u rows, capacity %lld", total, capacity_rows);
  return PSPACE_OK;
}
