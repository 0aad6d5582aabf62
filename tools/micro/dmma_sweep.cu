// Micro-benchmark: how many warps x independent accumulator chains the FP64 tensor pipe needs to saturate, with the
// transform's real inner loop (one shared-memory A load per step, one DMMA + one recurrence DFMA per chain).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_sweep dmma_sweep.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int NT, bool REC>
__global__ void __launch_bounds__(512, 1) loop_kernel(int iters, double* sink) {
  __shared__ double gs[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) gs[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, kq = lane & 3, nq = lane >> 2;
  double bc[NT], bp[NT], tc[NT], c[NT][2];
#pragma unroll
  for (int t = 0; t < NT; ++t) { bc[t] = 0.1 * (t + 1) + lane * 1e-3; bp[t] = 0.05 * t; tc[t] = 1.9 + 1e-3 * t; c[t][0] = c[t][1] = 0; }
  const double* ga = gs + kq * 8 + nq;
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
    const double a = ga[(i & 127) * 32];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(bc[t]));
      if (REC) { const double bn = fma(tc[t], bc[t], -bp[t]); bp[t] = bc[t]; bc[t] = bn; }
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < NT; ++t) s += c[t][0] + c[t][1] + bc[t];
  if (s == 1.2345) sink[0] = s;
}

template <int NT, bool REC> void run(int warps, int sms, double* sink) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  loop_kernel<NT, REC><<<sms, warps * 32>>>(iters / 10, sink);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); loop_kernel<NT, REC><<<sms, warps * 32>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double dmma = (double)sms * warps * NT * iters;
  printf("warps/SM %2d chains %d rec %d : %.3f ms  %.2f TFLOP/s (DMMA only)  %.1f ns per step per warp\n", warps, NT, (int)REC, best,
         dmma * 512.0 / (best * 1e-3) / 1e12, best * 1e6 / iters);
}

int main() {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* sink; cudaMalloc(&sink, 8);
  for (int w : {4, 8, 12, 16}) {
    run<1, true>(w, sms, sink); run<2, true>(w, sms, sink); run<3, true>(w, sms, sink); run<4, true>(w, sms, sink);
    run<6, true>(w, sms, sink); run<8, true>(w, sms, sink); run<8, false>(w, sms, sink);
  }
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
