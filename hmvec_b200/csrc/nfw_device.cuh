// nfw_device.cuh -- device routines shared by the NFW cube kernel (k_nfw.cu) and the fused spectra kernel
// (k_power.cu): Maclaurin-series coefficients of u_NFW in (x c)^2, the term-count rule, Horner evaluation, and
// (through sici.cuh) the Si/Ci form used beyond x c = 16.  See k_nfw.cu for the derivation.
#pragma once
#include "sici.cuh"
#include "gl64.inc"

namespace hmv {

constexpr int NFW_NMAX = 42;             // series coefficients per halo
constexpr int NFW_NREC = 48;             // doubles per halo record: A[0..42), c, a = r_s (1+z), a*c, ln(1+c), 1/m_c, 0
constexpr double NFW_XC_MAX = 16.0;      // series regime: x c <= 16

// A[0..NFW_NMAX): u_NFW(x; c) = sum_n A[n] (x c)^(2n), A_n = (-1)^n c^2 Itilde_n / ((2n+1)! m_c)
__device__ __forceinline__ void nfw_series_coefficients(double c, double mc, double* __restrict__ A) {
  const double pref = c * c / mc;
  if (c >= 1.5) {
    // Ktilde_m = (1/c^(m+1)) int_0^c t^m/(1+t)^2 dt :  K_m = 1/(c^2 (m-1)) - (2/c) K_(m-1) - K_(m-2)/c^2
    // (upward recurrence: the homogeneous solutions (-1/c)^m (a + b m) decay relative to K_m when c > 1)
    const double ic = 1.0 / c, ic2 = ic * ic;
    double k0 = 1.0 / (1.0 + c), k1 = mc * ic2;
    double sf = 1.0;                      // (-1)^n / (2n+1)!
    A[0] = pref * k1;
    for (int n = 1; n < NFW_NMAX; ++n) {
      const int m = 2 * n;                // even moment, then the odd one we need
      const double ke = ic2 / (double)(m - 1) - 2.0 * ic * k1 - ic2 * k0;
      const double ko = ic2 / (double)m - 2.0 * ic * ke - ic2 * k1;
      k0 = ke; k1 = ko;
      sf = -sf / ((double)(2 * n) * (double)(2 * n + 1));
      A[n] = pref * sf * ko;
    }
  } else {
    // Gauss-Legendre: node-major so that s^(2n+1) is a running product (no pow), one accumulator per moment
    double acc[NFW_NMAX];
#pragma unroll
    for (int n = 0; n < NFW_NMAX; ++n) acc[n] = 0.0;
    for (int q = 0; q < 64; ++q) {
      const double s = c_gl64_s[q], d = 1.0 + c * s, s2 = s * s;
      const double base = c_gl64_w[q] / (d * d);
      double pw = s;
#pragma unroll
      for (int n = 0; n < NFW_NMAX; ++n) {
        acc[n] = fma(base, pw, acc[n]);
        pw *= s2;
      }
    }
    double sf = 1.0;
#pragma unroll
    for (int n = 0; n < NFW_NMAX; ++n) {
      if (n > 0) sf = -sf / ((double)(2 * n) * (double)(2 * n + 1));
      A[n] = pref * sf * acc[n];
    }
  }
}

// per-halo NFW record rec[row][NFW_NREC] (row = z*nm + m): everything a kernel needs to evaluate u_NFW(k) for that halo
static __global__ void __launch_bounds__(128) nfw_record_kernel(int nz, int nm, const double* __restrict__ zs,
                                                                 const double* __restrict__ cs,
                                                                 const double* __restrict__ rvir,
                                                                 double* __restrict__ rec) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= (long long)nz * nm) return;
  const int z = (int)(row / nm);
  const double c = cs[row];
  const double ln1pc = log1p(c), mc = ln1pc - c / (1.0 + c);        // hmvec.py:348
  double* r = rec + row * NFW_NREC;
  nfw_series_coefficients(c, mc, r);
  const double a = rvir[row] / c * (1.0 + zs[z]);                     // x = k * rs * (1+z), hmvec.py:342,349
  r[42] = c; r[43] = a; r[44] = a * c; r[45] = ln1pc; r[46] = 1.0 / mc; r[47] = 0.0;
}

// odd term count n with y^n/(2n+1)! < 1e-19 (y = xc^2): tabulated at the low end, linear bound above
__device__ __forceinline__ int nfw_terms(double xc) {
  const float xf = (float)xc;
  const int n = xf < 0.03f ? 5 : xf < 0.3f ? 7 : xf < 1.0f ? 10 : min(NFW_NMAX - 1, (int)(1.8f * xf + 10.5f));
  return n | 1;
}

// sum_{i<n} A[i] y^i for odd n, coefficients fetched as aligned pairs
__device__ __forceinline__ double nfw_horner(const double* __restrict__ A, int n, double y) {
  double u = A[n - 1];
  for (int i = n - 2; i >= 1; i -= 2) {
    const double2 a2 = *reinterpret_cast<const double2*>(A + i - 1);
    u = fma(u, y, a2.y);
    u = fma(u, y, a2.x);
  }
  return u;
}

}  // namespace hmv
