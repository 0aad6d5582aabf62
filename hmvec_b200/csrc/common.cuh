// common.cuh -- shared helpers for the hmvec_b200 CUDA kernels (sm_100a, FP64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include "../../include/hmvec_b200.h"

namespace hmv {

// thread-local error text returned by hmv_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(HMV_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return HMV_OK;
}

#define HMV_REQUIRE(cond, ...)                          \
  do {                                                  \
    if (!(cond)) return hmv::fail(HMV_E_ARG, __VA_ARGS__); \
  } while (0)

// device-side bounds checks of the debug build (make EXTRA=-DHMV_DEBUG; compute-sanitizer is not available on the pool):
// a failed check prints its location and traps, which the next CUDA call reports as an error
#ifdef HMV_DEBUG
#define HMV_DEV_ASSERT(cond)                                                                         \
  do {                                                                                               \
    if (!(cond)) {                                                                                   \
      printf("HMV_DEV_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, \
             (int)blockIdx.x, (int)threadIdx.x);                                                     \
      __trap();                                                                                      \
    }                                                                                                \
  } while (0)
#else
#define HMV_DEV_ASSERT(cond) ((void)0)
#endif

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* smem_warp /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem_warp[w] = v;
  __syncthreads();
  double r = (threadIdx.x < nw) ? smem_warp[threadIdx.x] : 0.0;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) smem_warp[0] = r;
  }
  __syncthreads();
  r = smem_warp[0];
  return r;
}

// trapezoid weight of sample i on a (non-uniform) grid x[0..n): trapz(y,x) == sum_i y_i * w_i
__device__ __forceinline__ double trapz_weight(const double* __restrict__ x, int i, int n) {
  if (n < 2) return 0.0;
  const double lo = (i > 0) ? x[i - 1] : x[0];
  const double hi = (i < n - 1) ? x[i + 1] : x[n - 1];
  return 0.5 * (hi - lo);
}

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ----------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// producer-side wait: back off between polls so the spinning lane does not eat the consumers' issue slots
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  for (;;) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace hmv
