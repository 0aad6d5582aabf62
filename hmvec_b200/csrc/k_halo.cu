// k_halo.cu -- K0: per-halo geometry (Duffy c, r_vir), mass-definition secant solve, GNFW shape parameters.
// Reference arithmetic: hmvec.py:68-73, 111-115, 627-628 (geometry); :748-798 (mdelta_from_mdelta);
// :215-249, 278-316, 800-802, 856-860, 918-927 (Battaglia fits).
#include "common.cuh"

namespace hmv {

__device__ __forceinline__ double nfw_mc(double c) { return log1p(c) - c / (1.0 + c); }  // hmvec.py:737

__global__ void halo_geometry_kernel(int nz, int nm, const double* __restrict__ zs, const double* __restrict__ ms,
                                     const double* __restrict__ drho1, double A, double alpha, double beta, double h,
                                     double* __restrict__ cs, double* __restrict__ rvir) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nz * nm) return;
  const int z = (int)(idx / nm), m = (int)(idx - (long long)z * nm);
  const double M = ms[m];
  cs[idx] = A * pow(h * M / 2.0e12, alpha) * pow(1.0 + zs[z], beta);
  rvir[idx] = cbrt(3.0 * M / (4.0 * M_PI * drho1[z]));
}

// Secant iteration in x = ln M2 on  g(x) = M1/mc(C1) - e^x / mc(C2(x)),  C2 = C1 (e^(x-lnM1) r)^(1/3),
// with scipy.optimize.newton's start pair (p0 = ln M1, p1 = p0 (1+dx) + dx, dx = eps^0.33).
__global__ void mdelta_kernel(int nz, int nm, const double* __restrict__ ms, const double* __restrict__ cs,
                              const double* __restrict__ drho1, const double* __restrict__ drho2,
                              double* __restrict__ m2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nz * nm) return;
  const int z = (int)(idx / nm), m = (int)(idx - (long long)z * nm);
  const double M1 = ms[m], C1 = cs[idx], r = drho1[z] / drho2[z];
  const double lnM1 = log(M1), lhs = M1 / nfw_mc(C1);
  auto g = [&](double x) { return lhs - exp(x) / nfw_mc(C1 * cbrt(exp(x - lnM1) * r)); };
  const double dx = 6.7685935259123e-06;  // (2^-52)^0.33
  double p0 = lnM1, p1 = p0 * (1.0 + dx) + (p0 >= 0 ? dx : -dx);
  double q0 = g(p0), q1 = g(p1), p = p1;
  for (int it = 0; it < 60; ++it) {
    if (q1 == q0) { p = 0.5 * (p1 + p0); break; }
    const double dp = q1 * (p1 - p0) / (q1 - q0);
    p = p1 - dp;
    if (fabs(dp) < 1e-13 * fmax(1.0, fabs(p))) break;
    p0 = p1; q0 = q1; p1 = p; q1 = g(p1);
  }
  m2[idx] = exp(p);
}

struct Fit9 { double v[9]; };

__device__ __forceinline__ double plaw(double m200, double opz, const double* t) {
  return t[0] * pow(m200 / 1.0e14, t[1]) * pow(opz, t[2]);  // hmvec.py:800-802
}

__global__ void gnfw_params_kernel(int kind, int nz, int nm, const double* __restrict__ zs,
                                   const double* __restrict__ m200c, const double* __restrict__ rvir,
                                   const double* __restrict__ rhocrit, const double* __restrict__ hofz, Fit9 fit,
                                   double gamma, double pres_alpha, double amp_const, double pref,
                                   double* __restrict__ rs, double* __restrict__ cmax, double* __restrict__ xc,
                                   double* __restrict__ alpha, double* __restrict__ expo, double* __restrict__ amp,
                                   double* __restrict__ outscale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nz * nm) return;
  const int z = (int)(idx / nm);
  const double opz = 1.0 + zs[z], m200 = m200c[idx], rhoc = rhocrit[z];
  const double r200 = cbrt(3.0 * m200 / (4.0 * M_PI * 200.0 * rhoc));  // hmvec.py:225
  const double q0 = plaw(m200, opz, fit.v), q1 = plaw(m200, opz, fit.v + 3), q2 = plaw(m200, opz, fit.v + 6);
  if (kind == 0) {  // density: rho0 (q0) cancels against the mass norm; x^g (1+x^a)^(-(b+g)/a)
    const double rsv = 0.5 * r200;  // hmvec.py:247
    rs[idx] = rsv;
    cmax[idx] = rvir[idx] / rsv;
    xc[idx] = 1.0;
    alpha[idx] = q1;
    expo[idx] = (q2 + gamma) / q1;
    amp[idx] = 1.0;
    outscale[idx] = 1.0;
  } else {  // pressure: P0 (x/xc)^g (1+(x/xc)^a)^(-b), hmvec.py:918-927, scale :316
    rs[idx] = r200;
    cmax[idx] = rvir[idx] / r200;
    xc[idx] = q1;
    alpha[idx] = pres_alpha;
    expo[idx] = q2;
    amp[idx] = amp_const * m200 * rhoc / (2.0 * r200) * q0;
    outscale[idx] = pref * (r200 * r200 * r200) * (opz * opz / hofz[z]);
  }
}

}  // namespace hmv
using namespace hmv;

extern "C" int hmv_halo_geometry(int nz, int nm, const double* zs_d, const double* ms_d, const double* drho1_d,
                                 double duffy_A, double duffy_alpha, double duffy_beta, double h, double* cs_d,
                                 double* rvir_d, void* stream) {
  HMV_REQUIRE(nz > 0 && nm > 0, "hmv_halo_geometry: sizes must be positive");
  HMV_REQUIRE(zs_d && ms_d && drho1_d && cs_d && rvir_d, "hmv_halo_geometry: null pointer");
  const long long n = (long long)nz * nm;
  halo_geometry_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(nz, nm, zs_d, ms_d, drho1_d, duffy_A,
                                                                        duffy_alpha, duffy_beta, h, cs_d, rvir_d);
  return check_launch("halo_geometry_kernel");
}

extern "C" int hmv_mdelta(int nz, int nm, const double* ms_d, const double* cs_d, const double* drho1_d,
                          const double* drho2_d, double* m2_d, void* stream) {
  HMV_REQUIRE(nz > 0 && nm > 0, "hmv_mdelta: sizes must be positive");
  HMV_REQUIRE(ms_d && cs_d && drho1_d && drho2_d && m2_d, "hmv_mdelta: null pointer");
  const long long n = (long long)nz * nm;
  mdelta_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(nz, nm, ms_d, cs_d, drho1_d, drho2_d, m2_d);
  return check_launch("mdelta_kernel");
}

extern "C" int hmv_gnfw_params(int kind, int nz, int nm, const double* zs_d, const double* m200c_d,
                               const double* rvir_d, const double* rhocrit_d, const double* hofz_d,
                               const double* fit9_h, double gamma, double pres_alpha, double amp_const, double pref,
                               double* rs_d, double* cmax_d, double* xc_d, double* alpha_d, double* expo_d,
                               double* amp_d, double* outscale_d, void* stream) {
  HMV_REQUIRE(kind == 0 || kind == 1, "hmv_gnfw_params: kind must be 0 (density) or 1 (pressure)");
  HMV_REQUIRE(nz > 0 && nm > 0, "hmv_gnfw_params: sizes must be positive");
  HMV_REQUIRE(zs_d && m200c_d && rvir_d && rhocrit_d && hofz_d && fit9_h, "hmv_gnfw_params: null input pointer");
  HMV_REQUIRE(rs_d && cmax_d && xc_d && alpha_d && expo_d && amp_d && outscale_d, "hmv_gnfw_params: null output");
  Fit9 f;
  for (int i = 0; i < 9; ++i) f.v[i] = fit9_h[i];
  const long long n = (long long)nz * nm;
  gnfw_params_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(kind, nz, nm, zs_d, m200c_d, rvir_d, rhocrit_d,
                                                                      hofz_d, f, gamma, pres_alpha, amp_const, pref,
                                                                      rs_d, cmax_d, xc_d, alpha_d, expo_d, amp_d,
                                                                      outscale_d);
  return check_launch("gnfw_params_kernel");
}
