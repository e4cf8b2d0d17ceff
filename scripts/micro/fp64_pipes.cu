// Microbenchmark: do the FP64 FMA pipe (DFMA) and the FP64 tensor-core path (DMMA, mma.m8n8k4.f64)
// run concurrently on sm_100a, or do they share one pipe?  Three kernels: all warps DFMA, all warps
// DMMA, and half the warps each; reports FMA/clk/SM of each kind.  It decides whether a dense
// compound-matrix formulation of the Givens blocks could add throughput on top of the DFMA pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_pipes scripts/micro/fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// mode: 0 all DFMA, 1 all DMMA, 2 even warps DFMA / odd warps DMMA
__global__ void pipes_kernel(double *out, double a, double b, int iters, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 1 || (mode == 2 && (warp & 1));
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
  if (do_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) dmma(x[i], x[i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  const int sms = p.multiProcessorCount;
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 20000, threads = 512, blocks = sms * 2;
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    pipes_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9, 100, mode);
    cudaEventRecord(e0);
    pipes_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9, iters, mode);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double w_fma = mode == 0 ? warps : (mode == 1 ? 0 : warps / 2);
    const double w_mma = warps - w_fma;
    // per warp and iteration: 8 DFMA x 32 lanes = 256 FMA; 4 DMMA x (8*8*4) = 1024 FMA
    const double fma_dfma = w_fma * 256.0 * iters, fma_dmma = w_mma * 1024.0 * iters;
    const double clk = ms * 1e-3 * ghz * 1e9;
    printf("{\"mode\": \"%s\", \"ms\": %.3f, \"dfma_fma_per_clk_per_sm\": %.1f, \"dmma_fma_per_clk_per_sm\": %.1f, "
           "\"total_tflops\": %.2f}\n",
           mode == 0 ? "dfma" : (mode == 1 ? "dmma" : "half_half"), ms, fma_dfma / clk / sms,
           fma_dmma / clk / sms, 2 * (fma_dfma + fma_dmma) / (ms * 1e-3) / 1e12);
  }
  return 0;
}
