// k_nfw.cu -- K2: analytic truncated-NFW Fourier profile u(k|M,z) (reference hmvec.py:339-353).
//
// The reference evaluates  u = [sin x (Si X - Si x) - sin(cx)/X + cos x (Ci X - Ci x)]/m_c,  X=(1+c)x, x = k r_s (1+z),
// with two scipy.special.sici calls per (z,M,k).  That closed form is the integral
//        u(x; c) = (1/m_c) int_0^c  t/(1+t)^2  sinc(x t) dt ,
// so for x c <= 16 (60-95% of every halo's k-range) this kernel sums its Maclaurin series in y = (x c)^2,
//        u = sum_n A_n y^n ,   A_n = (-1)^n c^2 Itilde_n / ((2n+1)! m_c) ,  Itilde_n = int_0^1 s^(2n+1)/(1+cs)^2 ds ,
// whose per-halo coefficients come from a small pre-pass (three-term recurrence in the moment order for c >= 1.5,
// 64-point Gauss-Legendre below) -- 5 to 39 FMAs per element instead of two Si/Ci pairs, three sincos and a log.
// Truncation + cancellation error of the series is < 1e-11 relative (worst at x c = 16).  Beyond x c = 16 the
// Si/Ci form is evaluated directly with the device routines in sici.cuh.
//
// One CTA per halo row (z,M); each warp walks 256-wide k chunks (8 elements per lane sharing every coefficient load,
// warp-uniform term count) and writes the row once, coalesced: 8 B/element of algorithmic traffic.
#include "common.cuh"
#include "nfw_device.cuh"

namespace hmv {

constexpr int NFW_T = 256, NFW_E = 8, NFW_CH = 32 * NFW_E;

// max of ks over each NFW_CH-wide chunk: lets both passes classify a chunk with one load
__global__ void nfw_chunkmax_kernel(int nk, const double* __restrict__ ks, double* __restrict__ kcmax) {
  const int chunk = blockIdx.x, lane = threadIdx.x;
  double m = 0.0;
  for (int k = chunk * NFW_CH + lane; k < min(nk, (chunk + 1) * NFW_CH); k += 32) m = fmax(m, ks[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) kcmax[chunk] = m;
}

// A CTA first sweeps the chunks that lie entirely in the series regime, then the remaining ones (some element with
// x c > 16), which need the Si/Ci routines.  MODE 0/1 run only one of the two parts (kept for profiling them apart);
// the library launches MODE 2, both in one kernel, so that the store-bound and the FP64-bound halves overlap.
template <int MODE>   // 0 series only, 1 tail only, 2 both in one launch
__global__ void __launch_bounds__(NFW_T, MODE ? 3 : 4) uk_nfw_kernel(int nk, int ldk, const double* __restrict__ ks,
                                                        const double* __restrict__ coef,
                                                        const double* __restrict__ kcmax, double kmax,
                                                        double* __restrict__ uk) {
  __shared__ __align__(16) double A[NFW_NREC];
  const long long row = blockIdx.x;                   // row = z*nm + m
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* rec = coef + row * NFW_NREC;
  const double ac = __ldg(rec + 44);
  if (MODE == 1) {   // rows whose whole k-range is in the series regime have nothing to do here
    if (kmax * ac <= NFW_XC_MAX) return;
  }
  if (threadIdx.x < NFW_NREC) A[threadIdx.x] = __ldg(rec + threadIdx.x);
  __syncthreads();
  double* out = uk + row * (long long)ldk;
  const int nchunks = (nk + NFW_CH - 1) / NFW_CH;
  if (MODE != 1) {
    for (int chunk = warp; chunk < nchunks; chunk += NFW_T / 32) {
      const double xcm = __ldg(kcmax + chunk) * ac;
      if (xcm > NFW_XC_MAX) continue;                 // left to the tail pass
      const int kbase = chunk * NFW_CH + lane;
      const int nt = nfw_terms(xcm);                  // warp-uniform trip count
      double y[NFW_E], u[NFW_E];
#pragma unroll
      for (int e = 0; e < NFW_E; ++e) {
        const double xc = __ldg(ks + min(kbase + 32 * e, nk - 1)) * ac;
        y[e] = xc * xc;
      }
      const double top = A[nt - 1];
#pragma unroll
      for (int e = 0; e < NFW_E; ++e) u[e] = top;
      for (int i = nt - 2; i >= 1; i -= 2) {
        const double2 a2 = *reinterpret_cast<const double2*>(A + i - 1);
#pragma unroll
        for (int e = 0; e < NFW_E; ++e) u[e] = fma(u[e], y[e], a2.y);
#pragma unroll
        for (int e = 0; e < NFW_E; ++e) u[e] = fma(u[e], y[e], a2.x);
      }
#pragma unroll
      for (int e = 0; e < NFW_E; ++e) {
        const int k = kbase + 32 * e;
        if (k < nk) out[k] = u[e];
      }
    }
  }
  if (MODE != 0 && kmax * ac > NFW_XC_MAX) {
    // every warp visits every tail chunk and takes its own 32-wide slice of it: the Si/Ci work of a row is spread
    // evenly over the CTA's warps however few chunks are in the tail
    static_assert(NFW_T / 32 == NFW_E, "one slice per warp");
    const double c = A[42], a = A[43], ln1pc = A[45], inv_mc = A[46];
    for (int cb = 0; cb < nchunks; cb += 32) {        // 32 chunks per ballot: visit only the tail chunks
      const int cc = cb + lane;
      unsigned tail = __ballot_sync(0xffffffffu, cc < nchunks && __ldg(kcmax + min(cc, nchunks - 1)) * ac > NFW_XC_MAX);
      while (tail) {
        const int chunk = cb + __ffs(tail) - 1;
        tail &= tail - 1;
        const int k = chunk * NFW_CH + 32 * warp + lane;
        if (k < nk) {
          const double kk = __ldg(ks + k), xc = kk * ac;
          out[k] = (xc <= NFW_XC_MAX) ? nfw_horner(A, nfw_terms(xc), xc * xc) : nfw_bracket(kk * a, c, ln1pc) * inv_mc;
        }
      }
    }
  }
}

__global__ void sici_test_kernel(int n, const double* __restrict__ x, double* __restrict__ si,
                                 double* __restrict__ ci) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sici(x[i], si[i], ci[i]);
}

}  // namespace hmv
using namespace hmv;

extern "C" long long hmv_uk_nfw_ws_doubles(int nz, int nm, int nk) {
  if (nz <= 0 || nm <= 0 || nk <= 0) return 0;
  return (long long)nz * nm * NFW_NREC + (nk + NFW_CH - 1) / NFW_CH;   // per-halo records + per-chunk max(k)
}

extern "C" int hmv_uk_nfw(int nz, int nm, int nk, int ldk, const double* zs_d, const double* ks_d,
                          double kmax, const double* cs_d, const double* rvir_d, double* ws_d, double* uk_d,
                          void* stream) {
  HMV_REQUIRE(nz > 0 && nm > 0 && nk > 0 && ldk >= nk, "hmv_uk_nfw: bad sizes (nz=%d nm=%d nk=%d ldk=%d)", nz, nm, nk, ldk);
  HMV_REQUIRE(zs_d && ks_d && cs_d && rvir_d && ws_d && uk_d, "hmv_uk_nfw: null pointer");
  const long long rows = (long long)nz * nm;
  if (rows > 2147483647LL) return fail(HMV_E_LIMIT, "hmv_uk_nfw: %lld halo rows exceed the 2^31-1 grid limit", rows);
  cudaStream_t st = (cudaStream_t)stream;
  nfw_record_kernel<<<cdiv(rows, 128), 128, 0, st>>>(nz, nm, zs_d, cs_d, rvir_d, ws_d);
  int rc = check_launch("nfw_record_kernel");
  if (rc) return rc;
  double* kcmax = ws_d + rows * NFW_NREC;
  nfw_chunkmax_kernel<<<(nk + NFW_CH - 1) / NFW_CH, 32, 0, st>>>(nk, ks_d, kcmax);
  rc = check_launch("nfw_chunkmax_kernel");
  if (rc) return rc;
  uk_nfw_kernel<2><<<(unsigned)rows, NFW_T, 0, st>>>(nk, ldk, ks_d, ws_d, kcmax, kmax, uk_d);
  return check_launch("uk_nfw_kernel<both>");
}

// test hook: elementwise Si/Ci of the device routine (x > 0)
extern "C" int hmv_sici_test(int n, const double* x_d, double* si_d, double* ci_d, void* stream) {
  HMV_REQUIRE(n > 0 && x_d && si_d && ci_d, "hmv_sici_test: bad arguments");
  sici_test_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x_d, si_d, ci_d);
  return check_launch("sici_test_kernel");
}
