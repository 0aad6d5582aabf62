// k_transform.cu -- K1: batched numerical profile transform fused with on-the-fly GNFW evaluation and the
// interpolation onto the target wavenumbers.  Replaces generic_profile_fft + fft_integral + _interp_loop
// (reference fft.py:35-115) and rho_gas_generic_x / P_e_generic_x (hmvec.py:856-860, 918-927).
//
// One CTA owns HB consecutive-mass halos of one redshift.  The (z,M,x) profile cube never exists: samples are
// evaluated chunk by chunk into shared memory, the sine sums
//        U_j = step * sum_n x_n y_n sin(2 pi j n / N)          ( == -Im rfft(x*y) * step, fft.py:49 )
// are accumulated only for the bins j the target k-range needs (bin-skipping; the theta-cut bounds n), with the
// twiddle advanced by a rotation recurrence re-seeded exactly (sincospi of a reduced integer phase) every chunk,
// then u_j = U_j/kt_j/mnorm is linearly interpolated onto ks from shared memory (direct index j=floor(k/kout_1),
// no search) and written once, coalesced.
#include "common.cuh"

namespace hmv {

constexpr int TT = 256;    // threads per CTA == samples per chunk
constexpr int NCH = TT;

struct TParams {
  int nz, nm, nk, ldk, N, J, JS, nmg, do_mass_norm;
  double gamma, dx, step, kt1, kmax;
  const double *zs, *ks, *rs, *cmax, *xc, *alpha, *expo, *amp, *outscale;
  double* uk;
};

template <int HB>
__global__ void __launch_bounds__(TT) profile_transform_kernel(const TParams p) {
  extern __shared__ double smem[];
  double* Us = smem;                          // [HB][JS]
  double* gs = Us + (size_t)HB * p.JS;        // [NCH][HB]
  double* red = gs + NCH * HB;                // [32]
  __shared__ double h_cmax[HB], h_lxc[HB], h_alpha[HB], h_expo[HB], h_amp[HB], h_a[HB], h_oscale[HB];
  __shared__ int h_valid[HB];

  const int tid = threadIdx.x;
  const int z = blockIdx.x / p.nmg;
  const int mg = p.nmg - 1 - (blockIdx.x - z * p.nmg);  // heavy (large-M, many-bin) groups are scheduled first
  const int m0 = mg * HB;
  const double opz = 1.0 + p.zs[z];

  if (tid < HB) {
    const int m = m0 + tid;
    const bool ok = m < p.nm;
    const long long r = (long long)z * p.nm + (ok ? m : p.nm - 1);
    h_valid[tid] = ok;
    h_cmax[tid] = ok ? p.cmax[r] : -1.0;
    h_lxc[tid] = log(p.xc[r]);
    h_alpha[tid] = p.alpha[r];
    h_expo[tid] = p.expo[r];
    h_amp[tid] = p.amp[r];
    h_a[tid] = p.rs[r] * opz;                 // kout_j = kt_j / rs / (1+z)      (fft.py:92)
    h_oscale[tid] = p.outscale ? p.outscale[r] : 1.0;
  }
  for (int i = tid; i < HB * p.JS; i += TT) Us[i] = 0.0;
  __syncthreads();

  // sample and bin bounds shared by the HB halos of this CTA
  double cmx = -1.0, amax = 0.0;
#pragma unroll
  for (int h = 0; h < HB; ++h) {
    cmx = fmax(cmx, h_cmax[h]);
    if (h_valid[h]) amax = fmax(amax, h_a[h]);
  }
  int nb = (cmx > 0.0) ? (int)fmin((double)p.N, floor(cmx / p.dx) + 2.0) : 0;
  const int jn = (int)fmin((double)p.J, floor(p.kmax * amax / p.kt1) + 2.0);

  const double twoN = 2.0 / (double)p.N;
  double msum[HB];
#pragma unroll
  for (int h = 0; h < HB; ++h) msum[h] = 0.0;

  for (int n0 = 0; n0 < nb; n0 += NCH) {
    {  // ---- evaluate x*y for sample n0+tid of every halo (theta-cut: x <= cmax, fft.py:79-81) ----
      const int n = n0 + tid;
      const double x = (double)(n + 1) * p.dx;
      const double lx = log(x);
      const double w = (n == 0 || n == p.N - 1) ? 0.5 * p.dx : p.dx;  // np.trapz weights on xs (fft.py:84)
#pragma unroll
      for (int h = 0; h < HB; ++h) {
        double v = 0.0;
        if (n < p.N && x <= h_cmax[h]) {
          const double lt = lx - h_lxc[h];
          // amp * t^gamma * (1+t^alpha)^(-expo)
          const double rho = h_amp[h] * exp(p.gamma * lt - h_expo[h] * log1p(exp(h_alpha[h] * lt)));
          v = x * rho;
          msum[h] = fma(w * x, v, msum[h]);
        }
        gs[tid * HB + h] = v;
      }
    }
    __syncthreads();
    const int nlen = min(NCH, nb - n0);
    for (int j = tid + 1; j <= jn; j += TT) {
      double s, c, S, C;
      sincospi(twoN * (double)(((long long)j * n0) % p.N), &s, &c);  // exact phase of sample n0
      sincospi(twoN * (double)j, &S, &C);
      double acc[HB];
#pragma unroll
      for (int h = 0; h < HB; ++h) acc[h] = 0.0;
      for (int nn = 0; nn < nlen; ++nn) {
        const double* g = gs + nn * HB;
#pragma unroll
        for (int h = 0; h < HB; ++h) acc[h] = fma(g[h], s, acc[h]);
        const double s2 = fma(s, C, c * S);
        c = fma(c, C, -s * S);
        s = s2;
      }
#pragma unroll
      for (int h = 0; h < HB; ++h) Us[(size_t)h * p.JS + j] += acc[h];
    }
    __syncthreads();
  }

  // ---- mass norm (fft.py:83-87) and u_j = U_j / kt_j / mnorm (fft.py:91) ----
  double scale[HB];
#pragma unroll
  for (int h = 0; h < HB; ++h) {
    const double mn = p.do_mass_norm ? block_sum(msum[h], red) : 1.0;
    scale[h] = p.step / mn;
  }
  for (int j = tid + 1; j <= jn; j += TT) {
    const double ikt = 1.0 / ((double)j * p.kt1);
#pragma unroll
    for (int h = 0; h < HB; ++h) Us[(size_t)h * p.JS + j] *= scale[h] * ikt;
  }
  __syncthreads();

  // ---- interpolation onto the target ks (fft.py:102-107): hold u_1 below bin 1, zero above bin J ----
#pragma unroll 1
  for (int h = 0; h < HB; ++h) {
    if (!h_valid[h]) continue;
    const double* U = Us + (size_t)h * p.JS;
    const double kout1 = p.kt1 / h_a[h];
    const double inv = h_a[h] / p.kt1;
    const double koutJ = ((double)p.J * p.kt1) / h_a[h];
    const double osc = h_oscale[h];
    const double u1 = U[1];
    double* out = p.uk + ((long long)z * p.nm + m0 + h) * (long long)p.ldk;
    for (int k = tid; k < p.nk; k += TT) {
      const double kk = p.ks[k];
      double v;
      if (kk < kout1) {
        v = u1;
      } else if (kk > koutJ) {
        v = 0.0;
      } else {
        const double t = kk * inv;
        int j = (int)t;
        j = max(1, min(j, p.J - 1));
        v = fma(t - (double)j, U[j + 1] - U[j], U[j]);
      }
      out[k] = v * osc;
    }
  }
}

template <int HB>
static int launch_transform(const TParams& p, cudaStream_t st) {
  const size_t smem = ((size_t)HB * p.JS + (size_t)NCH * HB + 32) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(profile_transform_kernel<HB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return fail(HMV_E_CUDA, "profile_transform smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  TParams q = p;
  q.nmg = cdiv(p.nm, HB);
  profile_transform_kernel<HB><<<q.nz * q.nmg, TT, smem, st>>>(q);
  return check_launch("profile_transform_kernel");
}

}  // namespace hmv
using namespace hmv;

extern "C" int hmv_profile_transform(int nz, int nm, int nk, int ldk, const double* zs_d, const double* ks_d,
                                     double kmax, const double* rs_d, const double* cmax_d, const double* xc_d,
                                     const double* alpha_d, const double* expo_d, const double* amp_d,
                                     const double* outscale_d, double gamma, double xmax, int nxs, int do_mass_norm,
                                     double* uk_d, void* stream) {
  HMV_REQUIRE(nz > 0 && nm > 0 && nk > 0 && ldk >= nk, "hmv_profile_transform: bad sizes");
  HMV_REQUIRE(nxs >= 4 && xmax > 0, "hmv_profile_transform: need nxs>=4 and xmax>0");
  HMV_REQUIRE(zs_d && ks_d && rs_d && cmax_d && xc_d && alpha_d && expo_d && amp_d && uk_d,
              "hmv_profile_transform: null pointer");
  TParams p;
  p.nz = nz; p.nm = nm; p.nk = nk; p.ldk = ldk; p.N = nxs; p.J = nxs / 2; p.JS = p.J + 2;
  p.do_mass_norm = do_mass_norm;
  p.gamma = gamma;
  p.dx = xmax / nxs;                           // xs = linspace(0,xmax,nxs+1)[1:]           (fft.py:73)
  p.step = (xmax - 1.0 * p.dx) / nxs;          // (x[-1]-x[0])/N                            (fft.py:44-46)
  p.kt1 = (1.0 * (1.0 / (nxs * p.step))) * 2.0 * M_PI;  // rfftfreq(N,step)[1]*2pi          (fft.py:50)
  p.kmax = kmax;
  p.zs = zs_d; p.ks = ks_d; p.rs = rs_d; p.cmax = cmax_d; p.xc = xc_d; p.alpha = alpha_d; p.expo = expo_d;
  p.amp = amp_d; p.outscale = outscale_d; p.uk = uk_d; p.nmg = 0;
  cudaStream_t st = (cudaStream_t)stream;
  // pick the widest halo batch whose bin table fits in 200 KB of shared memory
  const size_t budget = 200 * 1024;
  auto need = [&](int hb) { return ((size_t)hb * p.JS + (size_t)NCH * hb + 32) * sizeof(double); };
  if (need(8) <= budget) return launch_transform<8>(p, st);
  if (need(4) <= budget) return launch_transform<4>(p, st);
  if (need(2) <= budget) return launch_transform<2>(p, st);
  if (need(1) <= budget) return launch_transform<1>(p, st);
  return fail(HMV_E_LIMIT, "hmv_profile_transform: nxs=%d needs %zu B of shared memory per halo (limit %zu)", nxs,
              need(1), budget);
}
