// How much DMMA throughput survives when DMULs that FEED the DMMA A operand share the FP64 datapath?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_mix tools/dmma_mix_microbench.cu && /tmp/dmma_mix
// One CTA of 8 warps per SM (2 warps per sub-partition, like the consumer warps of the projection kernels).  Each
// iteration = NA fragment DMULs (a = s*p, operands rotate so they are not loop invariant) then NA*NB DMMAs.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NA, int NB, int MULS>   // MULS: DMULs per A fragment (0 = none, 1 = real, 4 = complex-like)
__global__ void __launch_bounds__(256, 1) k_mix(double *out, int iters, double s0, double p0) {
  double c0[NA][NB], c1[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; a++)
#pragma unroll
    for (int b = 0; b < NB; b++) { c0[a][b] = threadIdx.x * 1e-3 + a; c1[a][b] = b; }
  double s[NA], p[NA], bf[NB];
#pragma unroll
  for (int a = 0; a < NA; a++) { s[a] = s0 + a; p[a] = p0 + threadIdx.x; }
#pragma unroll
  for (int b = 0; b < NB; b++) bf[b] = 1.0 + b * 1e-9;
  for (int it = 0; it < iters; ++it) {
    double af[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) {
      af[a] = s[a];
      if (MULS >= 1) af[a] = s[a] * p[a];
      if (MULS >= 4) af[a] = af[a] - p[a] * s[(a + 1) % NA] + s[a] * 1.0000001 + p[a] * 0.9999999;   // ~3 more FP64 ops
      s[a] = __longlong_as_double(__double_as_longlong(s[a]) ^ (long long)(it & 1));       // integer pipe: not loop invariant
    }
#pragma unroll
    for (int a = 0; a < NA; a++)
#pragma unroll
      for (int b = 0; b < NB; b++) dmma(c0[a][b], c1[a][b], af[a], bf[b]);
  }
  double t = 0;
#pragma unroll
  for (int a = 0; a < NA; a++)
#pragma unroll
    for (int b = 0; b < NB; b++) t += c0[a][b] + c1[a][b];
  if (t == 123.456) out[0] = t;
}

template <int NA, int NB, int MULS> void run(const char *name, int threads = 256) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 200000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_mix<NA, NB, MULS><<<sms, threads>>>(out, 1000, 1.0, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_mix<NA, NB, MULS><<<sms, threads>>>(out, iters, 1.0, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 256 * NA * NB * (double)iters * (threads / 32) * sms;
  printf("%-40s %8.3f ms  %6.2f TFLOP/s (DMMA only)\n", name, ms, flop / (ms * 1e-3) * 1e-12);
  cudaFree(out);
}

int main() {
  run<1, 3, 0>("1x3 = 3 accumulators, 2 warps/SMSP");
  run<2, 3, 0>("2x3 = 6 accumulators, 2 warps/SMSP");
  run<1, 3, 0>("1x3 = 3 accumulators, 1 warp/SMSP", 128);
  run<2, 3, 0>("2x3 = 6 accumulators, 1 warp/SMSP", 128);
  run<3, 3, 0>("3x3 = 9 accumulators, 1 warp/SMSP", 128);
  run<4, 3, 0>("4x3 = 12 accumulators, 1 warp/SMSP", 128);
  run<2, 12, 0>("2x12 = 24 accumulators, 1 warp/SMSP", 128);
  run<3, 3, 0>("3x3 DMMA, no DMUL");
  run<3, 3, 1>("3x3 DMMA, 1 DMUL per A (1:3)");
  run<3, 3, 4>("3x3 DMMA, ~4 FP64 ops per A");
  run<4, 3, 0>("4x3 DMMA, no DMUL");
  run<4, 3, 1>("4x3 DMMA, 1 DMUL per A (1:3)");
  run<2, 12, 0>("2x12 DMMA, no DMUL");
  run<2, 12, 1>("2x12 DMMA, 1 DMUL per A (1:12)");
  run<2, 6, 1>("2x6 DMMA, 1 DMUL per A (1:6)");
  run<3, 6, 4>("3x6 DMMA, ~4 FP64 ops per A (complex 1:12)");
  return 0;
}
