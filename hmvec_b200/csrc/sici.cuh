// sici.cuh -- FP64 device sine/cosine integrals for the analytic NFW profile (reference hmvec.py:349-352
// calls scipy.special.sici).  Own derivation (tools/gen_sici_tables.py): Maclaurin series for x <= 4,
// auxiliary functions f,g through piecewise degree-13 polynomials in s = 16/x^2 for x > 4.
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

__device__ __forceinline__ int sici_segment(double s) {
  return (s > c_seg_edge[1]) + (s > c_seg_edge[2]) + (s > c_seg_edge[3]) + (s > c_seg_edge[4]) +
         (s > c_seg_edge[5]) + (s > c_seg_edge[6]) + (s > c_seg_edge[7]);
}

// f(x), g(x) for x > 4:  Si = pi/2 - f cos x - g sin x,  Ci = f sin x - g cos x
__device__ __forceinline__ void sici_fg(double x, double& f, double& g) {
  const double rx = 1.0 / x, rx2 = rx * rx, s = 16.0 * rx2;
  const int seg = sici_segment(s);
  const double u = (s - c_seg_mid[seg]) * c_seg_iscale[seg];
  const double* cf = c_F + seg * (HMV_SICI_DEG + 1);
  const double* cg = c_G + seg * (HMV_SICI_DEG + 1);
  double F = cf[HMV_SICI_DEG], G = cg[HMV_SICI_DEG];
#pragma unroll
  for (int i = HMV_SICI_DEG - 1; i >= 0; --i) {
    F = fma(F, u, cf[i]);
    G = fma(G, u, cg[i]);
  }
  f = F * rx;
  g = G * rx2;
}

__device__ __forceinline__ double sici_g(double x) {
  const double rx = 1.0 / x, rx2 = rx * rx, s = 16.0 * rx2;
  const int seg = sici_segment(s);
  const double u = (s - c_seg_mid[seg]) * c_seg_iscale[seg];
  const double* cg = c_G + seg * (HMV_SICI_DEG + 1);
  double G = cg[HMV_SICI_DEG];
#pragma unroll
  for (int i = HMV_SICI_DEG - 1; i >= 0; --i) G = fma(G, u, cg[i]);
  return G * rx2;
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

// m_c * u_NFW(x; c) = sin x (Si X - Si x) - sin(c x)/X + cos x (Ci X - Ci x),  X = (1+c) x   (hmvec.py:352)
// Three regimes; in the two asymptotic ones the Si/Ci differences are combined analytically:
//   x > 4      :  f(X) sin(cx) - g(X) cos(cx) + g(x) - sin(cx)/X         (exact identity, one sincos)
//   x <= 4 < X :  series at x, f/g at X
//   X <= 4     :  series at both, Ci X - Ci x = ln(1+c) + Z C(Z) - z C(z)
__device__ __forceinline__ double nfw_bracket(double x, double c, double ln1pc) {
  const double X = (1.0 + c) * x;
  if (x > 4.0) {
    double fX, gX, scx, ccx;
    sici_fg(X, fX, gX);
    const double gx = sici_g(x);
    sincos(c * x, &scx, &ccx);
    return fX * scx - gX * ccx + gx - scx / X;
  }
  double sx, cx, S, C;
  sincos(x, &sx, &cx);
  const double z = x * x;
  sici_series(z, S, C);
  if (X > 4.0) {
    double fX, gX, sX, cX;
    sici_fg(X, fX, gX);
    sincos(X, &sX, &cX);
    const double siX = M_PI_2 - fX * cX - gX * sX, ciX = fX * sX - gX * cX;
    const double six = x * S, cix = HMV_EULER + log(x) + z * C;
    const double scx = sX * cx - cX * sx;  // sin(X - x) = sin(c x)
    return sx * (siX - six) - scx / X + cx * (ciX - cix);
  }
  double S2, C2;
  const double Z = X * X;
  sici_series(Z, S2, C2);
  const double dsi = X * S2 - x * S, dci = ln1pc + (Z * C2 - z * C);
  return sx * dsi - sin(c * x) / X + cx * dci;
}

}  // namespace hmv
