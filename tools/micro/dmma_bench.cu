// Micro-benchmark: FP64 tensor-core (mma.sync m8n8k4 f64) vs DFMA throughput on the current GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dmma_kernel(int iters, double* sink) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0}, c4[2] = {0, 0}, c5[2] = {0, 0}, c6[2] = {0, 0}, c7[2] = {0, 0};
  for (int i = 0; i < iters; ++i) {
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
    MMA(c0) MMA(c1) MMA(c2) MMA(c3) MMA(c4) MMA(c5) MMA(c6) MMA(c7)
  }
  double s = c0[0] + c1[0] + c2[0] + c3[0] + c4[0] + c5[0] + c6[0] + c7[0] + c0[1] + c1[1] + c2[1] + c3[1] + c4[1] + c5[1] + c6[1] + c7[1];
  if (s == 1.2345) sink[0] = s;
}

__global__ void __launch_bounds__(256) dfma_kernel(int iters, double* sink) {
  double a0 = 1.0 + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 1.2345) sink[0] = s;
}

// both at once: alternate DMMA and DFMA to see whether they share a pipe
__global__ void __launch_bounds__(256) mixed_kernel(int iters, double* sink) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double a0 = 1.0 + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    MMA(c0) a0 = fma(a0, m, c); a1 = fma(a1, m, c);
    MMA(c1) a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    MMA(c2) a4 = fma(a4, m, c); a5 = fma(a5, m, c);
    MMA(c3) a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = c0[0] + c1[0] + c2[0] + c3[0] + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 1.2345) sink[0] = s;
}

template <class K> float timeit(K k, int iters, double* sink, int blocks) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, 256>>>(iters / 8, sink);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<<<blocks, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* sink; cudaMalloc(&sink, 8);
  const int blocks = sms * 8, iters = 20000;
  const double warps = (double)blocks * 8;
  float t1 = timeit(dmma_kernel, iters, sink, blocks);
  float t2 = timeit(dfma_kernel, iters, sink, blocks);
  float t3 = timeit(mixed_kernel, iters, sink, blocks);
  // one m8n8k4 = 256 FMA = 512 flop per warp instruction
  printf("DMMA m8n8k4: %.3f ms -> %.2f TFLOP/s\n", t1, warps * 8.0 * iters * 512.0 / (t1 * 1e-3) / 1e12);
  printf("DFMA       : %.3f ms -> %.2f TFLOP/s\n", t2, warps * 32.0 * 8.0 * iters * 2.0 / (t2 * 1e-3) / 1e12);
  printf("mixed (4 DMMA + 8 DFMA per iter): %.3f ms -> DMMA part %.2f + DFMA part %.2f TFLOP/s\n", t3,
         warps * 4.0 * iters * 512.0 / (t3 * 1e-3) / 1e12, warps * 32.0 * 8.0 * iters * 2.0 / (t3 * 1e-3) / 1e12);
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
