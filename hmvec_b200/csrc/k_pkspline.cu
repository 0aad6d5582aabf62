// k_pkspline.cu -- P(z,k) on a (z, k) grid from the tensor-product B-spline of a matter-power interpolator: the
// step in front of the halo-model path when the linear power comes from CAMB / CLASS instead of the EH98 fit
// (reference cosmology.py:227-229 `_get_matter_power`, :353-374 `P_lin`, :376-382 `P_lin_slow`, all of which call
// PK.P(zs, ks, grid=True) of a scipy RectBivariateSpline in (z, ln k): utils.py:95-103, CAMB's
// get_matter_power_interpolator).  The spline FIT stays on the host (FITPACK, O(table)); this kernel is FITPACK's
// evaluation (fpbisp + fpbspl): clamp the argument to the knot range, locate the knot interval, de Boor-Cox basis
// values of degree kx / ky, and the (kx+1)(ky+1) tensor sum, then sign*exp() for log-interpolated spectra.
//
// One thread per (z,k): the two basis evaluations are ~60 flops and are recomputed per point rather than staged --
// the whole LARGE grid (200 x 10000) is 2e6 points, so the kernel is launch/latency bound (tens of microseconds).
#include "common.cuh"

namespace hmv {

constexpr int PKS_MAXDEG = 5;

// interval l with t[l] <= x < t[l+1], l in [k, n-k-2]  (fpbisp: "if(arg.lt.tx(l1) .or. l.eq.nkx1) go to ...")
__device__ __forceinline__ int pks_interval(const double* __restrict__ t, int n, int k, double x) {
  int lo = k, hi = n - k - 2;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (x >= t[mid]) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// fpbspl: the k+1 non-zero B-splines of degree k at x, knot interval l.  Same operations in the same order as FITPACK's
// recurrence (h_i <- h_i + f (t_li - x), h_{i+1} <- f (x - t_lj)), written with a carried term and fully unrolled to
// the maximum degree so that h[] stays in registers (no dynamically indexed local arrays).
__device__ __forceinline__ void pks_basis(const double* __restrict__ t, int k, double x, int l, double (&h)[PKS_MAXDEG + 1]) {
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= PKS_MAXDEG; ++j) h[j] = 0.0;
#pragma unroll
  for (int j = 1; j <= PKS_MAXDEG; ++j) {
    if (j <= k) {
      double carry = 0.0;
#pragma unroll
      for (int i = 0; i < j; ++i) {
        const double tr = t[l + 1 + i], tl = t[l + 1 + i - j];
        const double f = h[i] / (tr - tl);
        h[i] = carry + f * (tr - x);
        carry = f * (x - tl);
      }
      h[j] = carry;
    }
  }
}

__global__ void pk_spline_kernel(int nz, int nk, const double* __restrict__ zs, const double* __restrict__ ks,
                                 int nx, int ny, int kx, int ky, const double* __restrict__ tx,
                                 const double* __restrict__ ty, const double* __restrict__ c, int islog, double scale,
                                 double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
  if (k >= nk) return;
  const double xz = fmin(fmax(zs[z], tx[kx]), tx[nx - kx - 1]);
  const double yk = fmin(fmax(log(ks[k]), ty[ky]), ty[ny - ky - 1]);
  const int lx = pks_interval(tx, nx, kx, xz), ly = pks_interval(ty, ny, ky, yk);
  double wx[PKS_MAXDEG + 1], wy[PKS_MAXDEG + 1];
  pks_basis(tx, kx, xz, lx, wx);
  pks_basis(ty, ky, yk, ly, wy);
  const int ncy = ny - ky - 1;
  double sp = 0.0;
#pragma unroll
  for (int i = 0; i <= PKS_MAXDEG; ++i) {
    if (i <= kx) {
      const double* cr = c + (long long)(lx - kx + i) * ncy + (ly - ky);
#pragma unroll
      for (int j = 0; j <= PKS_MAXDEG; ++j)
        if (j <= ky) sp += cr[j] * wx[i] * wy[j];                     // fpbisp's summation order
    }
  }
  out[(long long)z * nk + k] = scale * (islog ? exp(sp) : sp);
}


// P(z,k) = D2[z] * v[k]: the separable EH98 linear power of accuracy='low' (cosmology.py:391-402) formed on the
// device from its two O(nz) and O(nk) host-side factors
__global__ void outer_kernel(int nz, int nk, const double* __restrict__ a, const double* __restrict__ b,
                             double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
  if (k < nk) out[(long long)z * nk + k] = a[z] * b[k];
}

// out = a + b (P = P1h + P2h, hmvec.py:500-502)
__global__ void sum2_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b,
                            double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}


// ---- Eisenstein & Hu 1998 transfer function (CDM + baryons with acoustic oscillations, eqs. 2-24; the zero-baryon
// shape fit, eqs. 28-31, when wiggles == 0) and the k-dependent factor of the separable accuracy='low' power
//      v(k) = pref k (k/kp)^(ns-1) T(k)^2            (reference cosmology.py:391-402 with Tk of :404-504)
// evaluated on the device: the host then only supplies D(z)^2 and the scalars (a numpy EH98 on 2 x 10000 wavenumbers
// costs the host 2 ms per HaloModel, more than a rank's whole device step on eight GPUs).
__global__ void eh98_factor_kernel(int nk, const double* __restrict__ ks, double h, double omch2, double ombh2,
                                   double omm0, int wiggles, double pref, double kp, double ns, double tcmb,
                                   double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nk) return;
  const double kmpc = ks[i], k = kmpc / h;                 // h/Mpc
  const double wm = omch2 + ombh2, wb = ombh2, fb = wb / wm, fc = omch2 / wm;
  const double t2 = (tcmb / 2.7) * (tcmb / 2.7);
  const double k_eq = 7.46e-2 * wm / t2 / h;               // eq. 3
  const double z_eq = 2.50e4 * wm / (t2 * t2);             // eq. 2
  const double zb1 = 0.313 * pow(wm, -0.419) * (1.0 + 0.607 * pow(wm, 0.674));
  const double zb2 = 0.238 * pow(wm, 0.223);
  const double z_d = 1291.0 * pow(wm, 0.251) / (1.0 + 0.659 * pow(wm, 0.828)) * (1.0 + zb1 * pow(wb, zb2));   // eq. 4
  const double Rd = 31.5 * wb / (t2 * t2) * (1.0e3 / z_d);  // eq. 5
  const double Req = 31.5 * wb / (t2 * t2) * (1.0e3 / z_eq);
  const double s = 2.0 / (3.0 * k_eq) * sqrt(6.0 / Req) * log((sqrt(1.0 + Rd) + sqrt(Req + Rd)) / (1.0 + sqrt(Req)));   // eq. 6
  const double k_silk = 1.6 * pow(wb, 0.52) * pow(wm, 0.73) * (1.0 + pow(10.4 * wm, -0.95)) / h;                        // eq. 7
  double T;
  if (!wiggles) {
    const double ag = 1.0 - 0.328 * log(431.0 * wm) * fb + 0.38 * log(22.3 * wm) * fb * fb;                              // eq. 31
    const double ks4 = (0.43 * k * s) * (0.43 * k * s) * (0.43 * k * s) * (0.43 * k * s);
    const double geff = omm0 * h * (ag + (1.0 - ag) / (1.0 + ks4));                                                      // eq. 30
    const double q = k * t2 / geff;
    const double L = log(2.0 * M_E + 1.8 * q);
    const double Cq = 14.2 + 731.0 / (1.0 + 62.5 * q);
    T = L / (L + Cq * q * q);                                                                                            // eq. 29
  } else {
    auto T0 = [&](double alpha, double beta) {              // eqs. 10, 19, 20
      const double q = k / (13.41 * k_eq);
      const double L = log(M_E + 1.8 * beta * q);
      const double Cq = 14.2 / alpha + 386.0 / (1.0 + 69.9 * pow(q, 1.08));
      return L / (L + Cq * q * q);
    };
    const double a1 = pow(46.9 * wm, 0.670) * (1.0 + pow(32.1 * wm, -0.532));   // eqs. 11, 12
    const double a2 = pow(12.0 * wm, 0.424) * (1.0 + pow(45.0 * wm, -0.582));
    const double alpha_c = pow(a1, -fb) * pow(a2, -(fb * fb * fb));
    const double b1 = 0.944 / (1.0 + pow(458.0 * wm, -0.708));
    const double b2 = pow(0.395 * wm, -0.0266);
    const double beta_c = 1.0 / (1.0 + b1 * (pow(fc, b2) - 1.0));
    const double ksf = k * s / 5.4;
    const double f = 1.0 / (1.0 + ksf * ksf * ksf * ksf);   // eq. 18
    const double Tc = f * T0(1.0, beta_c) + (1.0 - f) * T0(alpha_c, beta_c);     // eq. 17
    const double y = (1.0 + z_eq) / (1.0 + z_d);
    const double sq = sqrt(1.0 + y);
    const double G = y * (-6.0 * sq + (2.0 + 3.0 * y) * log((sq + 1.0) / (sq - 1.0)));   // eq. 15
    const double alpha_b = 2.07 * k_eq * s * pow(1.0 + Rd, -0.75) * G;          // eq. 14
    const double beta_node = 8.41 * pow(wm, 0.435);         // eq. 23
    const double bn = beta_node / (k * s);
    const double s_t = s / cbrt(1.0 + bn * bn * bn);        // eq. 22
    const double beta_b = 0.5 + fb + (3.0 - 2.0 * fb) * sqrt((17.2 * wm) * (17.2 * wm) + 1.0);   // eq. 24
    const double bb = beta_b / (k * s);
    const double x = k * s_t;
    const double sinc = (x == 0.0) ? 1.0 : sin(x) / x;
    const double ks2 = (k * s / 5.2) * (k * s / 5.2);
    const double Tb = (T0(1.0, 1.0) / (1.0 + ks2) + alpha_b / (1.0 + bb * bb * bb) * exp(-pow(k / k_silk, 1.4))) * sinc;   // eq. 21
    T = fb * Tb + fc * Tc;                                  // eq. 16
  }
  out[i] = pref * pow(kmpc / kp, ns - 1.0) * kmpc * T * T;
}

}  // namespace hmv
using namespace hmv;

extern "C" int hmv_pk_spline(int nz, int nk, const double* zs_d, const double* ks_d, int nx, int ny, int kx, int ky,
                             const double* tx_d, const double* ty_d, const double* c_d, int islog, double scale,
                             double* out_d, void* stream) {
  HMV_REQUIRE(nz > 0 && nk > 0, "hmv_pk_spline: bad sizes (nz=%d nk=%d)", nz, nk);
  HMV_REQUIRE(kx >= 1 && kx <= PKS_MAXDEG && ky >= 1 && ky <= PKS_MAXDEG, "hmv_pk_spline: degrees must be 1..%d", PKS_MAXDEG);
  HMV_REQUIRE(nx >= 2 * (kx + 1) && ny >= 2 * (ky + 1), "hmv_pk_spline: too few knots (nx=%d ny=%d)", nx, ny);
  HMV_REQUIRE(zs_d && ks_d && tx_d && ty_d && c_d && out_d, "hmv_pk_spline: null pointer");
  HMV_REQUIRE(nz <= 65535, "hmv_pk_spline: nz=%d exceeds the grid limit 65535", nz);
  dim3 grid(cdiv(nk, 128), nz);
  pk_spline_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(nz, nk, zs_d, ks_d, nx, ny, kx, ky, tx_d, ty_d, c_d, islog,
                                                            scale, out_d);
  return check_launch("pk_spline_kernel");
}

extern "C" int hmv_eh98_factor(int nk, const double* ks_d, double h, double omch2, double ombh2, double omm0,
                               int wiggles, double pref, double kp, double ns, double* out_d, void* stream) {
  HMV_REQUIRE(nk > 0 && ks_d && out_d, "hmv_eh98_factor: bad arguments");
  HMV_REQUIRE(h > 0 && omch2 > 0 && ombh2 > 0, "hmv_eh98_factor: need h, omch2, ombh2 > 0");
  eh98_factor_kernel<<<cdiv(nk, 128), 128, 0, (cudaStream_t)stream>>>(nk, ks_d, h, omch2, ombh2, omm0, wiggles, pref, kp, ns,
                                                                       2.726, out_d);
  return check_launch("eh98_factor_kernel");
}

extern "C" int hmv_outer(int nz, int nk, const double* a_d, const double* b_d, double* out_d, void* stream) {
  HMV_REQUIRE(nz > 0 && nk > 0 && nz <= 65535, "hmv_outer: bad sizes (nz=%d nk=%d)", nz, nk);
  HMV_REQUIRE(a_d && b_d && out_d, "hmv_outer: null pointer");
  dim3 grid(cdiv(nk, 256), nz);
  outer_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nz, nk, a_d, b_d, out_d);
  return check_launch("outer_kernel");
}

extern "C" int hmv_sum2(long long n, const double* a_d, const double* b_d, double* out_d, void* stream) {
  HMV_REQUIRE(n > 0 && a_d && b_d && out_d, "hmv_sum2: bad arguments");
  sum2_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, a_d, b_d, out_d);
  return check_launch("sum2_kernel");
}
