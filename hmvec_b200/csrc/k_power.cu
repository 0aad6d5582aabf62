// k_power.cu -- K5: fused 1-halo + 2-halo mass integrals (reference hmvec.py:469-572).
//
// HBM-bound streaming reduction over the mass axis of the u(z,M,k) cubes.  A CTA owns one redshift and a
// 64-wide k tile (one 512-byte segment per cube row); its 8 warps split the mass axis, each lane carries two
// adjacent k (16-byte vector loads), partial sums meet in shared memory, and the epilogue applies the
// trapezoid-in-linear-M weights (folded into per-(z,M) coefficient rows by a small prep kernel), the 1-halo
// damping, the consistency terms and P_lin.  Algorithmic traffic: 8*nm bytes per (z,k) per distinct cube read.
#include "common.cuh"

namespace hmv {

constexpr int PT = 256, PW = PT / 32, KT = 64;   // threads, warps, k per tile

// ---------------------------------------------------------------------------------------------------------
// prep: per-(z,M) coefficient rows + per-z offsets
//   generic pair (form 0):  coef = {aA, bA, aB, bB, w1, w2, -}     1h += w1 (aA UC_A + bA US_A)(aB UC_B + bB US_B)
//   hod square   (form 1):  coef = {aA, bA, aB, bB, c1, w2, c2}    1h += US_A (c1 UC_A + c2 US_A)
//   w1 = trapzw*n, w2 = trapzw*n*b_h ;  zoff[z] = {b_A - C_A, b_B - C_B}
// ---------------------------------------------------------------------------------------------------------
struct TracerArgs {
  int kind;
  const double *Nc, *Ns, *NcNs, *NsNsm1, *ngal, *bias;
};

__device__ __forceinline__ void leg_coeffs(const TracerArgs& t, int z, long long i, double mu, double& a, double& b,
                                           double& t0) {
  if (t.kind == 0) {            // matter: M u / rho_m0 (hmvec.py:488-492)
    a = 0.0; b = mu; t0 = mu;
  } else if (t.kind == 1) {     // hod: (u_c Nc + u_s Ns)/ngal (hmvec.py:481-486)
    const double ig = 1.0 / t.ngal[z];
    a = t.Nc[i] * ig; b = t.Ns[i] * ig; t0 = (t.Nc[i] + t.Ns[i]) * ig;
  } else {                      // pressure (hmvec.py:494-497, 541-545)
    a = 0.0; b = 1.0; t0 = 0.0;
  }
}

__global__ void __launch_bounds__(256) power_prep_kernel(int nm, const double* __restrict__ ms,
                                                          const double* __restrict__ nzm,
                                                          const double* __restrict__ bh, double rho_m0,
                                                          TracerArgs A, TracerArgs B, int form,
                                                          double* __restrict__ coef, long long cstride,
                                                          double* __restrict__ zoff) {
  __shared__ double red[32];
  const int z = blockIdx.x;
  double cA = 0.0, cB = 0.0, gA = 0.0, gB = 0.0;
  for (int m = threadIdx.x; m < nm; m += blockDim.x) {
    const long long i = (long long)z * nm + m;
    const double wt = trapz_weight(ms, m, nm);
    const double w1 = wt * nzm[i], w2 = w1 * bh[i], mu = ms[m] / rho_m0;
    double aA, bA, tA0, aB, bB, tB0;
    leg_coeffs(A, z, i, mu, aA, bA, tA0);
    leg_coeffs(B, z, i, mu, aB, bB, tB0);
    coef[0 * cstride + i] = aA; coef[1 * cstride + i] = bA;
    coef[2 * cstride + i] = aB; coef[3 * cstride + i] = bB;
    coef[5 * cstride + i] = w2;
    if (form == 1) {            // hmvec.py:477-479
      const double ig = 1.0 / A.ngal[z], ig2 = ig * ig;
      coef[4 * cstride + i] = w1 * 2.0 * A.NcNs[i] * ig2;
      coef[6 * cstride + i] = w1 * A.NsNsm1[i] * ig2;
    } else {
      coef[4 * cstride + i] = w1;
      coef[6 * cstride + i] = 0.0;
    }
    cA = fma(w2, tA0, cA);      // consistency integrals (hmvec.py:567-568)
    cB = fma(w2, tB0, cB);
    if (A.kind == 1) gA = fma(w2, A.Nc[i] + A.Ns[i], gA);   // get_bg numerator (hmvec.py:465)
    if (B.kind == 1) gB = fma(w2, B.Nc[i] + B.Ns[i], gB);
  }
  cA = block_sum(cA, red); cB = block_sum(cB, red); gA = block_sum(gA, red); gB = block_sum(gB, red);
  if (threadIdx.x == 0) {
    double bA = (A.kind == 0) ? 1.0 : (A.kind == 1 ? gA / A.ngal[z] : 0.0);
    double bB = (B.kind == 0) ? 1.0 : (B.kind == 1 ? gB / B.ngal[z] : 0.0);
    if (A.bias) bA = A.bias[z];   // b1_in / b2_in (hmvec.py:558-561)
    if (B.bias) bB = B.bias[z];
    zoff[2 * z + 0] = bA - cA;
    zoff[2 * z + 1] = bB - cB;
  }
}

struct PairArgs {
  int nm, nk, ldk, form;
  const double *usA, *ucA, *usB, *ucB;   // cubes (uc may be null => 1)
  const double* coef;
  long long cstride;
  const double *zoff, *ks, *Pzk;
  double kstar;
  double *p1h, *p2h;
};

__device__ __forceinline__ double2 ld2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

__global__ void __launch_bounds__(PT) power_pair_kernel(const PairArgs a) {
  __shared__ double part[PW][3][KT];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int z = blockIdx.y, k0 = blockIdx.x * KT + 2 * lane;
  const bool active = k0 < a.ldk;      // ldk is even: a full double2 is always in-bounds of the padded row
  const long long zrow = (long long)z * a.nm;
  const double* cf = a.coef + zrow;
  const bool sameUS = (a.usB == a.usA), hasUCA = (a.ucA != nullptr), hasUCB = (a.ucB != nullptr);
  const bool sameUC = (a.ucB == a.ucA);
  double2 p1 = {0, 0}, iA = {0, 0}, iB = {0, 0};
  if (active) {
#pragma unroll 4
    for (int m = w; m < a.nm; m += PW) {
      const long long off = (zrow + m) * (long long)a.ldk + k0;
      const double2 usA = ld2(a.usA + off);
      double2 ucA = {1.0, 1.0}, usB = usA, ucB = {1.0, 1.0};
      if (hasUCA) ucA = ld2(a.ucA + off);
      if (!sameUS) usB = ld2(a.usB + off);
      if (hasUCB) ucB = sameUC ? ucA : ld2(a.ucB + off);
      const double aA = __ldg(cf + m), bA = __ldg(cf + a.cstride + m);
      const double aB = __ldg(cf + 2 * a.cstride + m), bB = __ldg(cf + 3 * a.cstride + m);
      const double c4 = __ldg(cf + 4 * a.cstride + m), w2 = __ldg(cf + 5 * a.cstride + m);
      double2 tA, tB;
      tA.x = fma(aA, ucA.x, bA * usA.x); tA.y = fma(aA, ucA.y, bA * usA.y);
      tB.x = fma(aB, ucB.x, bB * usB.x); tB.y = fma(aB, ucB.y, bB * usB.y);
      if (a.form == 0) {
        p1.x = fma(c4 * tA.x, tB.x, p1.x); p1.y = fma(c4 * tA.y, tB.y, p1.y);
      } else {
        const double c6 = __ldg(cf + 6 * a.cstride + m);
        p1.x = fma(usA.x, fma(c4, ucA.x, c6 * usA.x), p1.x);
        p1.y = fma(usA.y, fma(c4, ucA.y, c6 * usA.y), p1.y);
      }
      iA.x = fma(w2, tA.x, iA.x); iA.y = fma(w2, tA.y, iA.y);
      iB.x = fma(w2, tB.x, iB.x); iB.y = fma(w2, tB.y, iB.y);
    }
  }
  part[w][0][2 * lane] = p1.x; part[w][0][2 * lane + 1] = p1.y;
  part[w][1][2 * lane] = iA.x; part[w][1][2 * lane + 1] = iA.y;
  part[w][2][2 * lane] = iB.x; part[w][2][2 * lane + 1] = iB.y;
  __syncthreads();
  if (threadIdx.x < KT) {
    const int k = blockIdx.x * KT + threadIdx.x;
    if (k < a.nk) {
      double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
      for (int ww = 0; ww < PW; ++ww) { s0 += part[ww][0][threadIdx.x]; s1 += part[ww][1][threadIdx.x]; s2 += part[ww][2][threadIdx.x]; }
      const long long o = (long long)z * a.nk + k;
      if (a.p1h) {
        const double r = a.ks[k] / a.kstar;
        a.p1h[o] = s0 * (1.0 - exp(-r * r));                                   // hmvec.py:526
      }
      if (a.p2h) a.p2h[o] = a.Pzk[o] * (s1 + a.zoff[2 * z]) * (s2 + a.zoff[2 * z + 1]);   // hmvec.py:572
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// six spectra {mm, ee, me, gg, gm, ge} in one pass over (u_m, u_e); HOD satellites follow u_m, centrals u_c = 1
//   coef rows: A1 = w1 mu^2, c1 = w1 2 NcNs/ngal^2, c2 = w1 NsNsm1/ngal^2, B1 = w1 mu Nc/ngal, B2 = w1 mu Ns/ngal,
//              D1 = w2 mu, D2 = w2 Ns/ngal ;  zoff6[z] = {1 - C_m, bg - C_g + sum w2 Nc/ngal}
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) power_six_prep_kernel(int nm, const double* __restrict__ ms,
                                                              const double* __restrict__ nzm,
                                                              const double* __restrict__ bh, double rho_m0,
                                                              const double* __restrict__ Nc, const double* __restrict__ Ns,
                                                              const double* __restrict__ NcNs,
                                                              const double* __restrict__ NsNsm1,
                                                              const double* __restrict__ ngal,
                                                              double* __restrict__ coef, long long cstride,
                                                              double* __restrict__ zoff) {
  __shared__ double red[32];
  const int z = blockIdx.x;
  const double ig = 1.0 / ngal[z], ig2 = ig * ig;
  double cm = 0.0, cg = 0.0, gb = 0.0, gc = 0.0;
  for (int m = threadIdx.x; m < nm; m += blockDim.x) {
    const long long i = (long long)z * nm + m;
    const double wt = trapz_weight(ms, m, nm);
    const double w1 = wt * nzm[i], w2 = w1 * bh[i], mu = ms[m] / rho_m0;
    const double nc = Nc[i] * ig, ns = Ns[i] * ig;
    coef[0 * cstride + i] = w1 * mu * mu;
    coef[1 * cstride + i] = w1 * 2.0 * NcNs[i] * ig2;
    coef[2 * cstride + i] = w1 * NsNsm1[i] * ig2;
    coef[3 * cstride + i] = w1 * mu * nc;
    coef[4 * cstride + i] = w1 * mu * ns;
    coef[5 * cstride + i] = w2 * mu;
    coef[6 * cstride + i] = w2 * ns;
    cm = fma(w2, mu, cm);
    cg = fma(w2, (Nc[i] + Ns[i]) * ig, cg);
    gb = fma(w2, Nc[i] + Ns[i], gb);
    gc = fma(w2, nc, gc);
  }
  cm = block_sum(cm, red); cg = block_sum(cg, red); gb = block_sum(gb, red); gc = block_sum(gc, red);
  if (threadIdx.x == 0) {
    zoff[2 * z + 0] = 1.0 - cm;
    zoff[2 * z + 1] = gb * ig - cg + gc;   // bias - consistency + the k-independent central part of I_g
  }
}

struct SixArgs {
  int nz, nm, nk, ldk;
  const double *um, *ue, *coef;
  long long cstride;
  const double *zoff, *ks, *Pzk;
  double kstar;
  double *p1h, *p2h;
};

__global__ void __launch_bounds__(PT) power_six_kernel(const SixArgs a) {
  __shared__ double part[PW][9][KT];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int z = blockIdx.y, k0 = blockIdx.x * KT + 2 * lane;
  const bool active = k0 < a.ldk;
  const long long zrow = (long long)z * a.nm;
  const double* cf = a.coef + zrow;
  double2 acc[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) acc[q] = make_double2(0.0, 0.0);
  if (active) {
#pragma unroll 4
    for (int m = w; m < a.nm; m += PW) {
      const long long off = (zrow + m) * (long long)a.ldk + k0;
      const double2 um = ld2(a.um + off), ue = ld2(a.ue + off);
      const double A1 = __ldg(cf + m), c1 = __ldg(cf + a.cstride + m), c2 = __ldg(cf + 2 * a.cstride + m);
      const double B1 = __ldg(cf + 3 * a.cstride + m), B2 = __ldg(cf + 4 * a.cstride + m);
      const double D1 = __ldg(cf + 5 * a.cstride + m), D2 = __ldg(cf + 6 * a.cstride + m);
#define HMV_SIX(c)                                                          \
  {                                                                         \
    const double q1 = um.c * um.c, q2 = ue.c * ue.c, q3 = um.c * ue.c;      \
    acc[0].c = fma(A1, q1, acc[0].c);                                       \
    acc[1].c = fma(A1, q2, acc[1].c);                                       \
    acc[2].c = fma(A1, q3, acc[2].c);                                       \
    acc[3].c = fma(c1, um.c, fma(c2, q1, acc[3].c));                        \
    acc[4].c = fma(B1, um.c, fma(B2, q1, acc[4].c));                        \
    acc[5].c = fma(B1, ue.c, fma(B2, q3, acc[5].c));                        \
    acc[6].c = fma(D1, um.c, acc[6].c);                                     \
    acc[7].c = fma(D1, ue.c, acc[7].c);                                     \
    acc[8].c = fma(D2, um.c, acc[8].c);                                     \
  }
      HMV_SIX(x)
      HMV_SIX(y)
#undef HMV_SIX
    }
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) { part[w][q][2 * lane] = acc[q].x; part[w][q][2 * lane + 1] = acc[q].y; }
  __syncthreads();
  if (threadIdx.x < KT) {
    const int k = blockIdx.x * KT + threadIdx.x;
    if (k < a.nk) {
      double s[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        double t = 0;
#pragma unroll
        for (int ww = 0; ww < PW; ++ww) t += part[ww][q][threadIdx.x];
        s[q] = t;
      }
      const long long o = (long long)z * a.nk + k, S = (long long)a.nz * a.nk;
      const double r = a.ks[k] / a.kstar, damp = 1.0 - exp(-r * r);
      if (a.p1h) {
#pragma unroll
        for (int q = 0; q < 6; ++q) a.p1h[q * S + o] = s[q] * damp;
      }
      if (a.p2h) {
        const double P = a.Pzk[o];
        const double Lm = s[6] + a.zoff[2 * z], Le = s[7] + a.zoff[2 * z], Lg = s[8] + a.zoff[2 * z + 1];
        a.p2h[0 * S + o] = P * Lm * Lm; a.p2h[1 * S + o] = P * Le * Le; a.p2h[2 * S + o] = P * Lm * Le;
        a.p2h[3 * S + o] = P * Lg * Lg; a.p2h[4 * S + o] = P * Lg * Lm; a.p2h[5 * S + o] = P * Lg * Le;
      }
    }
  }
}

static int tracer_args(const hmv_tracer* t, const char* which, TracerArgs* out) {
  HMV_REQUIRE(t != nullptr, "hmv_power: tracer %s is null", which);
  HMV_REQUIRE(t->kind >= 0 && t->kind <= 2, "hmv_power: tracer %s has unknown kind %d", which, t->kind);
  HMV_REQUIRE(t->us_d != nullptr, "hmv_power: tracer %s has no profile cube", which);
  if (t->kind == 1)
    HMV_REQUIRE(t->Nc_d && t->Ns_d && t->NcNs_d && t->NsNsm1_d && t->ngal_d, "hmv_power: hod tracer %s lacks occupation arrays", which);
  out->kind = t->kind; out->Nc = t->Nc_d; out->Ns = t->Ns_d; out->NcNs = t->NcNs_d; out->NsNsm1 = t->NsNsm1_d;
  out->ngal = t->ngal_d; out->bias = t->bias_d;
  return HMV_OK;
}

}  // namespace hmv
using namespace hmv;

extern "C" long long hmv_power_ws_doubles(int nz, int nm) {
  if (nz <= 0 || nm <= 0) return 0;
  return 7LL * nz * nm + 2LL * nz;
}

extern "C" int hmv_power(int nz, int nm, int nk, int ldk, const double* ms_d, const double* ks_d,
                         const double* nzm_d, const double* bh_d, const double* Pzk_d, double rho_m0, double kstar,
                         const hmv_tracer* A, const hmv_tracer* B, double* ws_d, double* p1h_d, double* p2h_d,
                         void* stream) {
  HMV_REQUIRE(nz > 0 && nm >= 2 && nk > 0 && ldk >= nk, "hmv_power: bad sizes (nz=%d nm=%d nk=%d ldk=%d)", nz, nm, nk, ldk);
  HMV_REQUIRE((ldk & 1) == 0, "hmv_power: ldk must be even (16-byte vector loads); got %d", ldk);
  HMV_REQUIRE(nz <= 65535, "hmv_power: nz=%d exceeds grid.y limit 65535", nz);
  HMV_REQUIRE(ms_d && ks_d && nzm_d && bh_d && ws_d, "hmv_power: null pointer");
  HMV_REQUIRE(p2h_d == nullptr || Pzk_d != nullptr, "hmv_power: P2h requested without Pzk");
  TracerArgs ta, tb;
  int rc = tracer_args(A, "A", &ta);
  if (rc) return rc;
  rc = tracer_args(B, "B", &tb);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long cs = (long long)nz * nm;
  double* coef = ws_d;
  double* zoff = ws_d + 7 * cs;
  int form = 0;
  if (A->kind == 1 && B->kind == 1) form = 1;            // hmvec.py:510-511 (uses leg A's HOD only)
  power_prep_kernel<<<nz, 256, 0, st>>>(nm, ms_d, nzm_d, bh_d, rho_m0, ta, tb, form, coef, cs, zoff);
  rc = check_launch("power_prep_kernel");
  if (rc) return rc;
  PairArgs a;
  a.nm = nm; a.nk = nk; a.ldk = ldk; a.form = form;
  a.usA = A->us_d; a.ucA = A->uc_d; a.usB = B->us_d; a.ucB = B->uc_d;
  a.coef = coef; a.cstride = cs; a.zoff = zoff; a.ks = ks_d; a.Pzk = Pzk_d; a.kstar = kstar;
  a.p1h = p1h_d; a.p2h = p2h_d;
  dim3 grid(cdiv(nk, KT), nz);
  if (A->kind == 2 && B->kind == 2 && A->us_d != B->us_d && p1h_d) {
    // both pressure: the 1h integrand is pk_A^2 (hmvec.py:512-513) while the 2h legs stay A and B.
    // Leg-B coefficient rows equal leg A's for pressure (a=0, b=1), so one prep serves both passes.
    PairArgs a1 = a;
    a1.usB = A->us_d; a1.ucB = A->uc_d; a1.p2h = nullptr;
    power_pair_kernel<<<grid, PT, 0, st>>>(a1);
    rc = check_launch("power_pair_kernel(1h)");
    if (rc) return rc;
    a.p1h = nullptr;
    if (!p2h_d) return HMV_OK;
  }
  power_pair_kernel<<<grid, PT, 0, st>>>(a);
  return check_launch("power_pair_kernel");
}

extern "C" int hmv_power_six(int nz, int nm, int nk, int ldk, const double* ms_d, const double* ks_d,
                             const double* nzm_d, const double* bh_d, const double* Pzk_d, double rho_m0,
                             double kstar, const double* um_d, const double* ue_d, const double* Nc_d,
                             const double* Ns_d, const double* NcNs_d, const double* NsNsm1_d, const double* ngal_d,
                             double* ws_d, double* p1h_d, double* p2h_d, void* stream) {
  HMV_REQUIRE(nz > 0 && nm >= 2 && nk > 0 && ldk >= nk, "hmv_power_six: bad sizes");
  HMV_REQUIRE((ldk & 1) == 0, "hmv_power_six: ldk must be even; got %d", ldk);
  HMV_REQUIRE(nz <= 65535, "hmv_power_six: nz=%d exceeds grid.y limit 65535", nz);
  HMV_REQUIRE(ms_d && ks_d && nzm_d && bh_d && um_d && ue_d && Nc_d && Ns_d && NcNs_d && NsNsm1_d && ngal_d && ws_d,
              "hmv_power_six: null pointer");
  HMV_REQUIRE(p2h_d == nullptr || Pzk_d != nullptr, "hmv_power_six: P2h requested without Pzk");
  cudaStream_t st = (cudaStream_t)stream;
  const long long cs = (long long)nz * nm;
  double* coef = ws_d;
  double* zoff = ws_d + 7 * cs;
  power_six_prep_kernel<<<nz, 256, 0, st>>>(nm, ms_d, nzm_d, bh_d, rho_m0, Nc_d, Ns_d, NcNs_d, NsNsm1_d, ngal_d, coef,
                                            cs, zoff);
  int rc = check_launch("power_six_prep_kernel");
  if (rc) return rc;
  SixArgs a;
  a.nz = nz; a.nm = nm; a.nk = nk; a.ldk = ldk; a.um = um_d; a.ue = ue_d; a.coef = coef; a.cstride = cs;
  a.zoff = zoff; a.ks = ks_d; a.Pzk = Pzk_d; a.kstar = kstar; a.p1h = p1h_d; a.p2h = p2h_d;
  dim3 grid(cdiv(nk, KT), nz);
  power_six_kernel<<<grid, PT, 0, st>>>(a);
  return check_launch("power_six_kernel");
}
