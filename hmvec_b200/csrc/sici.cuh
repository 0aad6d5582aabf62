// sici.cuh -- FP64 device sine/cosine integrals for the analytic NFW profile (reference hmvec.py:349-352
// calls scipy.special.sici).  Own derivation (tools/gen_sici_tables.py): Maclaurin series for x <= 4,
// auxiliary functions f,g through 128 uniform-segment degree-6 polynomials in s = 16/x^2 for x > 4.
// Max error vs 50-digit mpmath: 2e-16 relative on f,g; 4e-16 relative on Si, 2e-15 absolute on Ci (x<=4).
#pragma once
#include "sici_tables.inc"

namespace hmv {

#define HMV_EULER 0.57721566490153286061

// S(z) with Si(x) = x S(x^2);  C(z) with Ci(x) = euler + ln x + x^2 C(x^2)      (x <= 4)
__device__ __forceinline__ void sici_series(double z, double& S, double& C) {
  double s = c_si_mac[HMV_SICI_NMAC - 1], c = c_ci_mac[HMV_SICI_NMAC - 1];
#pragma unroll
  for (int i = HMV_SICI_NMAC - 2; i >= 0; --i) {
    s = fma(s, z, c_si_mac[i]);
    c = fma(c, z, c_ci_mac[i]);
  }
  S = s;
  C = c;
}

// sin and cos of 0 <= x < ~1e6 for the Si/Ci tail (x = c k r_s(1+z) reaches ~1e4): two-constant Cody-Waite
// reduction by pi/2 with FMAs, then the kernels sin r = r PS(r^2), cos r = PC(r^2) on |r| <= pi/4 (own Chebyshev
// fits, 1.1e-16 absolute; tools/gen_sici_tables.py).  About 30 instructions; CUDA's sincospi costs ~80 here
// because it also serves huge and special arguments.
__device__ __forceinline__ void sincos_cw(double x, double& sn, double& cs) {
  const double q = rint(x * 0.63661977236758134308);           // x * 2/pi
  double r = fma(-q, 1.5707963267948966, x);                   // pi/2 = hi + lo
  r = fma(-q, 6.123233995736766e-17, r);
  const int iq = (int)q;
  const double w = r * r;
  double ps = c_sin_k[7], pc = c_cos_k[8];
#pragma unroll
  for (int i = 6; i >= 0; --i) ps = fma(ps, w, c_sin_k[i]);
#pragma unroll
  for (int i = 7; i >= 0; --i) pc = fma(pc, w, c_cos_k[i]);
  const double sr = r * ps;
  double s0 = (iq & 1) ? pc : sr;
  double c0 = (iq & 1) ? -sr : pc;
  if (iq & 2) { s0 = -s0; c0 = -c0; }
  sn = s0;
  cs = c0;
}

// f(x), g(x) for x > 4 from the reciprocal rx = 1/x:  Si = pi/2 - f cos x - g sin x,  Ci = f sin x - g cos x.
// s = 16/x^2 in (0,1] is cut into HMV_SICI_NSEG uniform segments (index = one multiply + float->int); each holds
// degree-HMV_SICI_DEG polynomials for F = x f and G = x^2 g whose coefficient pairs are read as double2 (L1).
__device__ __forceinline__ void sici_fg_r(double rx, double& f, double& g) {
  const double rx2 = rx * rx;
  const double t = (16.0 * HMV_SICI_NSEG) * rx2;
  const int seg = min(HMV_SICI_NSEG - 1, (int)t);
  const double u = fma(2.0, t - (double)seg, -1.0);
  const double2* co = g_sici_FG + seg * (HMV_SICI_DEG + 1);
  double2 c = __ldg(co + HMV_SICI_DEG);
  double F = c.x, G = c.y;
#pragma unroll
  for (int i = HMV_SICI_DEG - 1; i >= 0; --i) {
    c = __ldg(co + i);
    F = fma(F, u, c.x);
    G = fma(G, u, c.y);
  }
  f = F * rx;
  g = G * rx2;
}

__device__ __forceinline__ void sici_fg(double x, double& f, double& g) { sici_fg_r(1.0 / x, f, g); }

// g(x) alone (x > 4), from rx = 1/x
__device__ __forceinline__ double sici_g_r(double rx) {
  const double rx2 = rx * rx;
  const double t = (16.0 * HMV_SICI_NSEG) * rx2;
  const int seg = min(HMV_SICI_NSEG - 1, (int)t);
  const double u = fma(2.0, t - (double)seg, -1.0);
  const double2* co = g_sici_FG + seg * (HMV_SICI_DEG + 1);
  double G = __ldg(co + HMV_SICI_DEG).y;
#pragma unroll
  for (int i = HMV_SICI_DEG - 1; i >= 0; --i) G = fma(G, u, __ldg(co + i).y);
  return G * rx2;
}

// f(X), g(X) for X >= 64 from their asymptotic series in w = 1/X^2 (7 and 8 terms: the first omitted terms are
// 8.7e10 w^7 and 3.6e14 w^8, < 5e-15 relative at X = 64): no table, no conversions
__device__ __forceinline__ void sici_fg_far(double rX, double& f, double& g) {
  const double w = rX * rX;
  double F = 479001600.0, G = -1307674368000.0;
  F = fma(F, w, -3628800.0);  G = fma(G, w, 6227020800.0);
  F = fma(F, w, 40320.0);     G = fma(G, w, -39916800.0);
  F = fma(F, w, -720.0);      G = fma(G, w, 362880.0);
  F = fma(F, w, 24.0);        G = fma(G, w, -5040.0);
  F = fma(F, w, -2.0);        G = fma(G, w, 120.0);
  F = fma(F, w, 1.0);         G = fma(G, w, -6.0);
  G = fma(G, w, 1.0);
  f = F * rX;
  g = G * w;
}

// General-purpose pair (used by tests through hmv_sici_test): Si(x), Ci(x) for x > 0.
__device__ __forceinline__ void sici(double x, double& si, double& ci) {
  if (x <= 4.0) {
    double S, C;
    const double z = x * x;
    sici_series(z, S, C);
    si = x * S;
    ci = HMV_EULER + log(x) + z * C;
  } else {
    double f, g, sn, cs;
    sici_fg(x, f, g);
    sincos(x, &sn, &cs);
    si = M_PI_2 - f * cs - g * sn;
    ci = f * sn - g * cs;
  }
}

// 1/x for positive normal x to <= 1 ulp: hardware seed (rcp.approx.ftz.f64, ~2^-23) + two Newton steps.  IEEE
// division spends ~20 instructions on correct rounding and special operands that the tail never meets.
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// m_c * u_NFW(x; c) = sin x (Si X - Si x) - sin(c x)/X + cos x (Ci X - Ci x),  X = (1+c) x   (hmvec.py:352)
// Three regimes; in the two asymptotic ones the Si/Ci differences are combined analytically:
//   x > 4      :  f(X) sin(cx) - g(X) cos(cx) + g(x) - sin(cx)/X         (exact identity, one sincos)
//   x <= 4 < X :  series at x, f/g at X
//   X <= 4     :  series at both, Ci X - Ci x = ln(1+c) + Z C(Z) - z C(z)
// the x > 4 branch when also X >= 64 (every element beyond the polynomial range of k_nfw.cu): f, g at X from the
// asymptotic series, only g at x from the table
__device__ __forceinline__ double nfw_bracket_far(double x, double c) {
  const double X = (1.0 + c) * x;
  const double rX = rcp_fast(X);
  double fX, gX, scx, ccx;
  sici_fg_far(rX, fX, gX);
  const double gx = sici_g_r((1.0 + c) * rX);
  sincos_cw(c * x, scx, ccx);
  return fX * scx - gX * ccx + gx - scx * rX;
}

__device__ __forceinline__ double nfw_bracket(double x, double c, double ln1pc) {
  const double X = (1.0 + c) * x;
  const double rX = rcp_fast(X);
  if (x > 4.0) {
    double fX, gX, fx, gx, scx, ccx;
    sici_fg_r(rX, fX, gX);
    sici_fg_r((1.0 + c) * rX, fx, gx);     // 1/x = (1+c)/X: one reciprocal serves both arguments
    sincos_cw(c * x, scx, ccx);
    return fX * scx - gX * ccx + gx - scx * rX;
  }
  double sx, cx, S, C;
  sincos_cw(x, sx, cx);
  const double z = x * x;
  sici_series(z, S, C);
  if (X > 4.0) {
    double fX, gX, sX, cX;
    sici_fg_r(rX, fX, gX);
    sincos_cw(X, sX, cX);
    const double siX = M_PI_2 - fX * cX - gX * sX, ciX = fX * sX - gX * cX;
    const double six = x * S, cix = HMV_EULER + log(x) + z * C;
    const double scx = sX * cx - cX * sx;  // sin(X - x) = sin(c x)
    return sx * (siX - six) - scx * rX + cx * (ciX - cix);
  }
  double S2, C2;
  const double Z = X * X;
  sici_series(Z, S2, C2);
  const double dsi = X * S2 - x * S, dci = ln1pc + (Z * C2 - z * C);
  return sx * dsi - sin(c * x) * rX + cx * dci;
}

}  // namespace hmv
