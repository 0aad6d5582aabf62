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
