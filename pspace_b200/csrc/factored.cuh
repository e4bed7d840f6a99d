// Shared device helpers of the FACTORED-OPERAND kernels (vproject.cu: vector projection, expand.cu: expansion).
//
// Both are one-pass GEMMs whose generated operand psi_k(z_q) (N x Q) would cost more HBM traffic than the whole
// algorithmic input if it were materialised (the round-1 panels: 2.4 GB for cfg5, 68.7 GB for cfg4).  The tensor
// grid makes it separable: with the trailing variables [dsplit, nvars) as the "tail" and L = prod of their rule
// sizes (the last variable is the fastest grid index, QuadratureHelper.cpp:108-124),
//
//        psi_k(z_q) = P_k[q div L] * S_k[q mod L],     P_k[b] = prod_{d <  dsplit} T_d[alpha_d(k)][digit_d(b*L)]
//                                                       S_k[s] = prod_{d >= dsplit} T_d[alpha_d(k)][digit_d(s)]
//
// (weighted rows use the tables Tw_d = w_d T_d, so that w_q psi_k(z_q) is the same two-factor product).  The producer
// warps of a CTA write the few P and S entries a pipeline stage needs into shared memory; a consumer lane forms each
// DMMA A-fragment element with ONE DMUL from two shared-memory reads.  Nothing of size N x Q ever exists.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace psf {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// `dep` is a run-time zero derived from the last fragment registers of the stage: the arrive cannot issue before the
// loads that produced them have returned, so a stage is never released ahead of its last read.
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar, unsigned dep = 0) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar) + dep) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PSF_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PSF_DONE;\n"
      "bra PSF_WAIT;\n"
      "PSF_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, unsigned long long *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double x, unsigned mask) {   // x * (+-1) on the integer pipe
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

// prod_{d in [d0,d1)} T[eo[d*eo_stride] + dg[d]]   — reference product order (ascending parameter id,
// ParameterContainer.cpp:184-192), seeded with (vr,vi).  eo = per-variable table offsets of one basis row (shared
// memory, u16), dg = digit row of a grid point (global, broadcast across the warp), T = 1-D tables (generic pointer:
// the shared-memory copy when the tables are small, else global / L1).
template <bool CPLX>
__device__ __forceinline__ void table_product(const double *T, const unsigned short *eo, int eo_stride,
                                              const uint8_t *__restrict__ dg, int d0, int d1, double &vr, double &vi) {
#pragma unroll 4
  for (int d = d0; d < d1; d++) {
    const int o = (int)eo[d * eo_stride] + (int)__ldg(dg + d);
    if (!CPLX) vr *= T[o];
    else {
      const double2 t = reinterpret_cast<const double2 *>(T)[o];
      const double nr = vr * t.x - vi * t.y, ni = vr * t.y + vi * t.x;
      vr = nr; vi = ni;
    }
  }
}
template <bool CPLX>
__device__ __forceinline__ void table_mul(const double *T, int o, double br, double bi, double &vr, double &vi) {
  if (!CPLX) { vr = br * T[o]; vi = 0.0; }
  else {
    const double2 t = reinterpret_cast<const double2 *>(T)[o];
    vr = br * t.x - bi * t.y; vi = br * t.y + bi * t.x;
  }
}

// Factor entries of one basis row for a RUN of consecutive mixed-radix indices i0 .. i0+n-1 over the variables [d0,d1)
// (d1-1 fastest):  out[e] = prod_{d in [d0,d1)} T[eo[d] + digit_d(i0+e)].  The product over [d0,d1-1) is kept while the
// last digit runs through its rule (it is recomputed when the digit wraps to 0), so an entry costs ~1 + (d1-d0-1)/nq_last
// table reads.  Index i is grid point i*scale (scale = L for prefix blocks, 1 for tail points); its digit row comes
// from the global digit table.  Entries whose point lies at or past pt_end, and rows that are not valid, are zero.
template <bool CPLX>
__device__ __forceinline__ void run_products(const double *T, const unsigned short *eo, int eos, const uint8_t *__restrict__ qdig,
                                             int DP, long long i0, int n, long long scale, long long pt_end, int d0, int d1,
                                             int nq_last, bool valid, double *out, int out_stride, int plane) {
  int lo = (d1 > d0) ? (int)(i0 % nq_last) : 0;
  double br = 1.0, bi = 0.0;
  bool have = false;
  for (int e = 0; e < n; e++) {
    const long long pt = (i0 + e) * scale;
    double vr = 0.0, vi = 0.0;
    if (valid && pt < pt_end) {
      if (d1 > d0) {
        if (!have || lo == 0) {
          br = 1.0; bi = 0.0;
          table_product<CPLX>(T, eo, eos, qdig + (size_t)pt * DP, d0, d1 - 1, br, bi);
          have = true;
        }
        table_mul<CPLX>(T, (int)eo[(d1 - 1) * eos] + lo, br, bi, vr, vi);
      } else vr = 1.0;
    }
    out[e * out_stride] = vr;
    if (CPLX) out[(plane + e) * out_stride] = vi;
    if (++lo == nq_last) lo = 0;
  }
}

// the 1-D tables go to shared memory when they are at most this many doubles (cfg5: 160); larger ones stay in L1/L2
constexpr int TAB_SMEM_MAX = 1024;

}  // namespace psf

// project.cu: TMA descriptor (fp64, no swizzle, zero fill out of bounds); dims/box innermost first
int ps_make_tmap(CUtensorMap *tm, const double *base, int rank, const unsigned long long *dims,
                 const unsigned long long *strides_bytes /*rank-1*/, const unsigned *box);
