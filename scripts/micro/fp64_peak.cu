// Microbenchmark: FP64 FMA throughput (DFMA/clk/SM, TFLOP/s) and dependent-issue latency on the
// device it runs on.  The fused Givens kernel is bounded by this pipe once all rotations of a
// spin sector are fused into one sweep, so this is its second roofline denominator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak scripts/micro/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int blocks_per_sm, int threads, int sms, double clock_ghz) {
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma_kernel<ILP><<<sms * blocks_per_sm, threads>>>(out, 1.0000001, 1e-9, 100);
  cudaEventRecord(e0);
  dfma_kernel<ILP><<<sms * blocks_per_sm, threads>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)sms * blocks_per_sm * threads * ILP * (double)iters;
  double per_clk_sm = fma / (ms * 1e-3) / (clock_ghz * 1e9) / sms;
  double cycles_per_iter = (ms * 1e-3) * clock_ghz * 1e9 / iters;
  printf("{\"ilp\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"dfma_per_clk_per_sm\": %.1f, \"cycles_per_iter\": %.1f}\n",
         ILP, blocks_per_sm * threads / 32, ms, 2 * fma / (ms * 1e-3) / 1e12, per_clk_sm, cycles_per_iter);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_max\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
  int sms = p.multiProcessorCount;
  run<1>(1, 32, sms, ghz);     // latency: one warp per SM, one chain -> cycles_per_iter = DFMA latency
  run<2>(1, 32, sms, ghz);
  run<4>(1, 32, sms, ghz);
  run<8>(1, 32, sms, ghz);
  run<1>(1, 128, sms, ghz);    // one warp per SMSP
  run<4>(1, 128, sms, ghz);
  run<8>(1, 128, sms, ghz);
  run<8>(1, 512, sms, ghz);    // the fused kernel's occupancy
  run<8>(2, 512, sms, ghz);
  run<16>(2, 1024, sms, ghz);
  return 0;
}
