// gnfw_eval.cuh -- lean, lock-step exp / log(1+w) for the on-the-fly GNFW profile evaluation of the transform kernel
//      rho(x) = amp * t^gamma * (1 + t^alpha)^(-expo),  t = x/xc        (hmvec.py:856-860, 918-927)
// evaluated as exp(gamma lt - expo log(1 + exp(alpha lt))), lt = log t.
//
// Why not the CUDA math library: the producer warps share the FP64 pipe with their own DMMAs, exp + log1p + exp from
// libdevice cost ~100 FP64 instructions per sample, and -- decisive with only two producer warps per scheduler --
// the compiler does not interleave the eight halos a thread evaluates (each call is one long dependent Horner chain),
// so the evaluation ran latency-bound ("wait" stalls, ncu).  Here every step is written across V values at once, so
// the V chains are independent instructions back to back, with table-driven range reduction to shorten them:
//   exp(y):  y = (16 k + i) ln2/16 + r, |r| <= ln2/32  ->  2^k * T[i] * P6(r)          11 FP64 instructions
//   log(u):  u = 2^e m, m in [1,2), c_i ~ 1/m from the top 7 mantissa bits,
//            r = m c_i - 1, |r| <= 2^-8  ->  e ln2 - log(c_i) + P6(r)                   10 FP64 instructions
// Absolute error of log ~2e-16, relative error of exp ~4e-16: both enter rho at the 1e-15 level (parity bar: 1e-6).
// Arguments are not range-checked beyond saturation (|y| > ~690 clamps the exponent, u must be a positive normal).
#pragma once

namespace hmv {

constexpr int GNFW_EXP_TAB = 16, GNFW_LOG_TAB = 128;

struct GnfwTables {
  double ex[GNFW_EXP_TAB];          // 2^(i/16)
  double2 lg[GNFW_LOG_TAB];         // { c_i = fl(1/(1 + (i+0.5)/128)), -log(c_i) }
};

// filled once per CTA (any GNFW_LOG_TAB threads); callers synchronise afterwards
__device__ __forceinline__ void gnfw_tables_init(GnfwTables& t, int tid, int nthreads) {
  for (int i = tid; i < GNFW_LOG_TAB; i += nthreads) {
    const double c = 1.0 / (1.0 + ((double)i + 0.5) / (double)GNFW_LOG_TAB);
    t.lg[i] = make_double2(c, -log(c));
    if (i < GNFW_EXP_TAB) t.ex[i] = exp2((double)i / (double)GNFW_EXP_TAB);
  }
}

template <int V>
__device__ __forceinline__ void exp_lockstep(const GnfwTables& t, const double (&y)[V], double (&out)[V]) {
  const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: rint() in the low word
  const double INV = 23.083120654223414;                   // 16 / ln 2
  const double L_HI = 0x1.62e42fee00000p-5;                // ln2/16, low 21 mantissa bits zero (exact products for |k| < 2^20)
  const double L_LO = 1.1926343307941173e-11;              // ln2/16 - L_HI
  double kd[V], r[V], p[V];
  int ki[V];
#pragma unroll
  for (int v = 0; v < V; ++v) kd[v] = fma(y[v], INV, MAGIC);
#pragma unroll
  for (int v = 0; v < V; ++v) { ki[v] = __double2loint(kd[v]); kd[v] -= MAGIC; }
#pragma unroll
  for (int v = 0; v < V; ++v) r[v] = fma(-kd[v], L_HI, y[v]);
#pragma unroll
  for (int v = 0; v < V; ++v) r[v] = fma(-kd[v], L_LO, r[v]);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(r[v], 1.0 / 720.0, 1.0 / 120.0);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], 1.0 / 24.0);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], 1.0 / 6.0);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], 0.5);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], 1.0);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], 1.0);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double s = t.ex[ki[v] & (GNFW_EXP_TAB - 1)] * p[v];                // in [1, 2.1)
    const int e = min(max(ki[v] >> 4, -1000), 1000);                           // saturate instead of wrapping the exponent
    out[v] = __hiloint2double(__double2hiint(s) + (e << 20), __double2loint(s));
  }
}

// log(add + w) for positive normal add + w (absolute accuracy ~2e-16: it is used inside an exponent);
// add = 1 gives the log(1 + t^alpha) of the profile, add = 0 a plain logarithm
template <int V>
__device__ __forceinline__ void log_lockstep(const GnfwTables& t, const double (&w)[V], double add, double (&out)[V]) {
  const double MAGIC = 6755399441055744.0;
  const double LN2 = 0.6931471805599453;
  double r[V], p[V], base[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double u = add + w[v];
    const int hi = __double2hiint(u);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u));   // [1, 2)
    const double2 c = t.lg[(hi >> 13) & (GNFW_LOG_TAB - 1)];
    r[v] = fma(m, c.x, -1.0);
    const int e = (hi >> 20) - 1023;
    const double ed = __hiloint2double(0x43380000 + (e >> 31), e) - MAGIC;                  // (double)e
    base[v] = fma(ed, LN2, c.y);
  }
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(r[v], -1.0 / 6.0, 0.2);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], -0.25);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], 1.0 / 3.0);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], -0.5);
#pragma unroll
  for (int v = 0; v < V; ++v) p[v] = fma(p[v], r[v], 1.0);
#pragma unroll
  for (int v = 0; v < V; ++v) out[v] = fma(p[v], r[v], base[v]);
}

}  // namespace hmv
