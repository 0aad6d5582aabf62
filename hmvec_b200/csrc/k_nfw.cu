// k_nfw.cu -- K2: analytic truncated-NFW Fourier profile u(k|M,z) (reference hmvec.py:339-353).
// One thread per (z,M,k) element; FP64-pipe bound (Si/Ci + sincos), coalesced 8 B/element store.
#include "common.cuh"
#include "sici.cuh"

namespace hmv {

constexpr int NFW_T = 256, NFW_KPT = 4;  // 1024 k per block

__global__ void __launch_bounds__(NFW_T) uk_nfw_kernel(int nm, int nk, int ldk, int ktiles,
                                                        const double* __restrict__ zs,
                                                        const double* __restrict__ ks,
                                                        const double* __restrict__ cs,
                                                        const double* __restrict__ rvir,
                                                        double* __restrict__ uk) {
  const long long row = blockIdx.x / ktiles;          // row = z*nm + m
  const int kt = blockIdx.x - (int)(row * ktiles);
  const int z = (int)(row / nm);
  const double c = cs[row];
  const double a = rvir[row] / c * (1.0 + zs[z]);     // x = k * rs * (1+z), hmvec.py:342,349
  const double ln1pc = log1p(c);
  const double inv_mc = 1.0 / (ln1pc - c / (1.0 + c));  // hmvec.py:348
  double* out = uk + row * (long long)ldk;
  const int k0 = kt * (NFW_T * NFW_KPT) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < NFW_KPT; ++i) {
    const int k = k0 + i * NFW_T;
    if (k < nk) out[k] = nfw_bracket(ks[k] * a, c, ln1pc) * inv_mc;
  }
}

__global__ void sici_test_kernel(int n, const double* __restrict__ x, double* __restrict__ si,
                                 double* __restrict__ ci) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sici(x[i], si[i], ci[i]);
}

}  // namespace hmv
using namespace hmv;

extern "C" int hmv_uk_nfw(int nz, int nm, int nk, int ldk, const double* zs_d, const double* ks_d,
                          const double* cs_d, const double* rvir_d, double* uk_d, void* stream) {
  HMV_REQUIRE(nz > 0 && nm > 0 && nk > 0 && ldk >= nk, "hmv_uk_nfw: bad sizes (nz=%d nm=%d nk=%d ldk=%d)", nz, nm, nk, ldk);
  HMV_REQUIRE(zs_d && ks_d && cs_d && rvir_d && uk_d, "hmv_uk_nfw: null pointer");
  const int ktiles = cdiv(nk, NFW_T * NFW_KPT);
  const long long blocks = (long long)nz * nm * ktiles;
  if (blocks > 2147483647LL) return fail(HMV_E_LIMIT, "hmv_uk_nfw: %lld blocks exceeds the 2^31-1 grid limit", blocks);
  uk_nfw_kernel<<<(unsigned)blocks, NFW_T, 0, (cudaStream_t)stream>>>(nm, nk, ldk, ktiles, zs_d, ks_d, cs_d, rvir_d, uk_d);
  return check_launch("uk_nfw_kernel");
}

// test hook: elementwise Si/Ci of the device routine (x > 0)
extern "C" int hmv_sici_test(int n, const double* x_d, double* si_d, double* ci_d, void* stream) {
  HMV_REQUIRE(n > 0 && x_d && si_d && ci_d, "hmv_sici_test: bad arguments");
  sici_test_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x_d, si_d, ci_d);
  return check_launch("sici_test_kernel");
}
