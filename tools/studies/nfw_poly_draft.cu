// nfw_poly_draft.cu -- DRAFT of the next NFW cube kernel.  NOT part of the library and NOT yet
// run on a GPU: it is compile-checked only (nvcc -gencode arch=compute_100a,code=sm_100a -I../../hmvec_b200/csrc -c)
// and kept here so that the next round starts from code instead of a description.  Numerics are validated in float64
// by nfw_poly_emulation.py (worst 3.4e-12 of an interval's max|u| up to x c = 64).
//
// Idea: u_NFW(x c; c) on 0 < x c <= 64 as per-halo piecewise polynomials in y = (x c)^2 (21 intervals, degree 5..13)
// instead of a 5..39-term Maclaurin series up to x c = 16 and the Si/Ci closed form beyond.
//   pre-pass  nfw_poly_record_kernel : one thread per (halo, interval): u at the interval's Chebyshev nodes from the
//                                      closed form (nfw_bracket), monomial coefficients = M_n . node values
//   cube      uk_nfw_poly_kernel     : one CTA per halo row; a sorted k axis is cut at the interval edges by a
//                                      lane-parallel binary search, every warp takes 256-wide chunks of one interval
//                                      (warp-uniform coefficients from shared memory, 8 wavenumbers per lane), the
//                                      few wavenumbers beyond x c = 64 go through nfw_bracket as today.
#include "common.cuh"
#include "nfw_device.cuh"
#include "nfw_poly_tables.inc"

namespace hmv {

constexpr int NFWP_META = 6;                                     // c, a, a*c, ln(1+c), 1/m_c, pad
constexpr int NFWP_REC = NFWP_NI * NFWP_STRIDE + NFWP_META;      // doubles per halo record
#ifndef NFWP_EV
#define NFWP_EV 8
#endif
constexpr int NFWP_T = 256, NFWP_E = NFWP_EV, NFWP_CH = 32 * NFWP_E;

__global__ void __launch_bounds__(128) nfw_poly_record_kernel(int nz, int nm, const double* __restrict__ zs,
                                                              const double* __restrict__ cs,
                                                              const double* __restrict__ rvir,
                                                              double* __restrict__ rec) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nz * nm * NFWP_NI) return;
  const long long row = idx / NFWP_NI;
  const int iv = (int)(idx - row * NFWP_NI);
  const int z = (int)(row / nm);
  const double c = cs[row];
  const double ln1pc = log1p(c), mc = ln1pc - c / (1.0 + c), inv_mc = 1.0 / mc;       // hmvec.py:348
  double* r = rec + row * NFWP_REC;
  if (iv == 0) {
    const double a = rvir[row] / c * (1.0 + zs[z]);                                    // x = k rs (1+z), hmvec.py:342,349
    double* meta = r + NFWP_NI * NFWP_STRIDE;
    meta[0] = c; meta[1] = a; meta[2] = a * c; meta[3] = ln1pc; meta[4] = inv_mc; meta[5] = 0.0;
  }
  const int n = c_nfwp_deg[iv] + 1;
  int ni = 0;
  while (c_nfwp_nlist[ni] != n) ++ni;
  const double lo = c_nfwp_edge[iv], hi = c_nfwp_edge[iv + 1];
  const double ymid = 0.5 * (hi * hi + lo * lo), yhalf = 0.5 * (hi * hi - lo * lo);
  double f[NFWP_STRIDE];
#pragma unroll
  for (int i = 0; i < NFWP_STRIDE; ++i) {
    f[i] = 0.0;
    if (i < n) {
      const double xc = sqrt(fma(yhalf, c_nfwp_nodes[ni][i], ymid));
      f[i] = nfw_bracket(xc / c, c, ln1pc) * inv_mc;
    }
  }
  const double* M = c_nfwp_mat + c_nfwp_moff[ni];
  double* out = r + iv * NFWP_STRIDE;
#pragma unroll
  for (int j = 0; j < NFWP_STRIDE; ++j) {
    double m = 0.0;
    if (j < n) {
#pragma unroll
      for (int i = 0; i < NFWP_STRIDE; ++i)
        if (i < n) m = fma(M[j * n + i], f[i], m);
    }
    out[j] = m;
  }
}

// `sorted`: device flag (1 = ks non-decreasing), e.g. written by the chunk pre-pass; unsorted axes evaluate the closed
// form per element.
__global__ void __launch_bounds__(NFWP_T, 3) uk_nfw_poly_kernel(int nk, int ldk, const double* __restrict__ ks,
                                                                 const double* __restrict__ rec,
                                                                 const int* __restrict__ sorted,
                                                                 double* __restrict__ uk) {
  __shared__ __align__(16) double R[NFWP_REC];
  __shared__ int kb[NFWP_NI + 1];          // first k index with x c >= edge
  __shared__ int pre[NFWP_NI + 2];         // prefix of 256-wide chunk counts per interval, then the tail
  const long long row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NFWP_REC; i += NFWP_T) R[i] = __ldg(rec + row * NFWP_REC + i);
  __syncthreads();
  const double* meta = R + NFWP_NI * NFWP_STRIDE;
  const double c = meta[0], a = meta[1], ac = meta[2], ln1pc = meta[3], inv_mc = meta[4];
  double* out = uk + row * (long long)ldk;
  if (!*sorted) {
    for (int k = tid; k < nk; k += NFWP_T) out[k] = nfw_bracket(__ldg(ks + k) * a, c, ln1pc) * inv_mc;
    return;
  }
  if (warp == 0) {
    if (lane <= NFWP_NI) {
      const double edge = c_nfwp_edge[lane];
      int lo = 0, hi = nk;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ks + mid) * ac >= edge) hi = mid; else lo = mid + 1;
      }
      kb[lane] = lo;
    }
    __syncwarp();
    if (lane == 0) {
      int acc = 0;
      for (int i = 0; i < NFWP_NI; ++i) { pre[i] = acc; acc += (kb[i + 1] - kb[i] + NFWP_CH - 1) / NFWP_CH; }
      pre[NFWP_NI] = acc;
      pre[NFWP_NI + 1] = acc + (nk - kb[NFWP_NI] + 31) / 32;       // tail: 32-wide slices
    }
  }
  __syncthreads();
  const int npoly = pre[NFWP_NI], nall = pre[NFWP_NI + 1];
  for (int w = warp; w < nall; w += NFWP_T / 32) {
    if (w < npoly) {
      int iv = 0;
      while (pre[iv + 1] <= w) ++iv;                               // warp-uniform
      const int kbeg = kb[iv] + (w - pre[iv]) * NFWP_CH, kend = kb[iv + 1];
      const int deg = c_nfwp_deg[iv];
      const double ts = c_nfwp_tscale[iv], to = c_nfwp_toffs[iv];
      const double* m = R + iv * NFWP_STRIDE;
      double t[NFWP_E], u[NFWP_E];
#pragma unroll
      for (int e = 0; e < NFWP_E; ++e) {
        const double xc = __ldg(ks + min(kbeg + lane + 32 * e, nk - 1)) * ac;
        t[e] = fma(xc * xc, ts, to);
      }
      const double top = m[deg];
#pragma unroll
      for (int e = 0; e < NFWP_E; ++e) u[e] = top;
      for (int j = deg - 1; j >= 0; --j) {
        const double mj = m[j];
#pragma unroll
        for (int e = 0; e < NFWP_E; ++e) u[e] = fma(u[e], t[e], mj);
      }
#pragma unroll
      for (int e = 0; e < NFWP_E; ++e) {
        const int k = kbeg + lane + 32 * e;
        if (k < kend) out[k] = u[e];
      }
    } else {
      const int k = kb[NFWP_NI] + (w - npoly) * 32 + lane;          // beyond the last edge: closed form
      if (k < nk) out[k] = nfw_bracket(__ldg(ks + k) * a, c, ln1pc) * inv_mc;
    }
  }
}

}  // namespace hmv

// ---- study hook (tools/studies/run_nfw_poly_draft.py): the draft as a stand-alone shared library ---------------------
__global__ void nfwp_sorted_flag_kernel(int nk, const double* __restrict__ ks, int* __restrict__ flag) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  for (int k = threadIdx.x; k + 1 < nk; k += blockDim.x)
    if (!(ks[k] <= ks[k + 1])) ok = 0;
  __syncthreads();
  if (threadIdx.x == 0) *flag = ok;
}

extern "C" long long nfwp_ws_doubles(int nz, int nm) { return (long long)nz * nm * hmv::NFWP_REC + 2; }

extern "C" int nfwp_uk_nfw(int nz, int nm, int nk, int ldk, const double* zs_d, const double* ks_d, const double* cs_d,
                           const double* rvir_d, double* ws_d, double* uk_d, void* stream) {
  using namespace hmv;
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)nz * nm;
  int* flag = reinterpret_cast<int*>(ws_d + rows * NFWP_REC);
  nfwp_sorted_flag_kernel<<<1, 1024, 0, st>>>(nk < 0 ? -nk : nk, ks_d, flag);   // study hook: nk < 0 skips the pre-pass, ldk == 0 skips the cube
  if (nk > 0) nfw_poly_record_kernel<<<(unsigned)((rows * NFWP_NI + 127) / 128), 128, 0, st>>>(nz, nm, zs_d, cs_d, rvir_d, ws_d);
  if (ldk > 0) uk_nfw_poly_kernel<<<(unsigned)rows, NFWP_T, 0, st>>>(nk < 0 ? -nk : nk, ldk, ks_d, ws_d, flag, uk_d);
  return (int)cudaGetLastError();
}
