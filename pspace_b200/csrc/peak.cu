// Measurement helpers for bench.py: FP64 peak (roofline denominator, SURVEY.md §8d — MEASURED_PEAKS.json has
// no FP64 figure) and the synthetic splitmix64 block generator.
#include "ctx.h"

namespace {
template <int NACC>
__global__ void __launch_bounds__(256) k_peak_dfma(double *out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_peak_dmma(double *out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__global__ void k_fill_synthetic(double *out, long long n, unsigned long long seed, unsigned long long offset) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    unsigned long long x = splitmix64(seed ^ (offset + (unsigned long long)i));
    out[i] = (double)(x >> 11) * 0x1.0p-53 - 0.5;
  }
}
}  // namespace

extern "C" int pspace_cuda_fp64_peak(int device, int mode, double ms_target, double *tflops) {
  if (!tflops) return ps_fail(PSPACE_EINVAL, "null argument");
  PS_CUDA(cudaSetDevice(device));
  int sms = 0;
  PS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  double *out = nullptr;
  PS_CUDA(cudaMalloc(&out, 8));
  cudaEvent_t e0, e1;
  PS_CUDA(cudaEventCreate(&e0));
  PS_CUDA(cudaEventCreate(&e1));
  const int grid = sms * 4, iters = 20000;
  const double flop_per_launch = mode == 0 ? 2.0 * 16 * iters * 256.0 * grid        // 16 DFMA chains / thread
                                           : 2.0 * 256 * 8 * (double)iters * 8.0 * grid;   // 8 DMMA (256 FMA) / warp
  auto run = [&]() {
    if (mode == 0) k_peak_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    else k_peak_dmma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
  };
  run();
  PS_CUDA(cudaDeviceSynchronize());
  float ms = 0;
  PS_CUDA(cudaEventRecord(e0));
  run();
  PS_CUDA(cudaEventRecord(e1));
  PS_CUDA(cudaEventSynchronize(e1));
  PS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  int reps = (int)(ms_target / (ms > 1e-3 ? ms : 1e-3));
  if (reps < 1) reps = 1;
  PS_CUDA(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) run();
  PS_CUDA(cudaEventRecord(e1));
  PS_CUDA(cudaEventSynchronize(e1));
  PS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = flop_per_launch * reps / (ms * 1e-3) * 1e-12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return PSPACE_OK;
}

extern "C" int pspace_cuda_fill_synthetic(double *d_out, long long n, unsigned long long seed, unsigned long long offset,
                                          void *stream) {
  if (!d_out && n > 0) return ps_fail(PSPACE_EINVAL, "null argument");
  if (n <= 0) return PSPACE_OK;
  unsigned nb = (unsigned)((n + 255) / 256 < 148LL * 32 ? (n + 255) / 256 : 148LL * 32);
  k_fill_synthetic<<<nb, 256, 0, (cudaStream_t)stream>>>(d_out, n, seed, offset);
  PS_LAUNCH_CHECK(nullptr);
  return PSPACE_OK;
}
