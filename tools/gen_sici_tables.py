#!/usr/bin/env python
"""Generate hmvec_b200/csrc/sici_tables.inc : FP64 coefficient tables for the device Si/Ci routine.

The analytic NFW profile (reference hmvec.py:349-352) calls scipy.special.sici on [nz,nm,nk]
elements twice.  The CUDA kernel needs its own Si/Ci; rather than transcribing a library's tables
we derive ours from first principles with mpmath (40 digits):

  x <= 4 :  Maclaurin series in z = x^2
            Si(x) = x * sum_n a_n z^n ,            a_n = (-1)^n / ((2n+1)(2n+1)!)
            Ci(x) = gamma + ln x + z * sum_n b_n z^n,  b_n = (-1)^(n+1) / ((2n+2)(2n+2)!)
  x  > 4 :  auxiliary functions  Si = pi/2 - f cos x - g sin x ,  Ci = f sin x - g cos x  with
            F(s) = x f(x),  G(s) = x^2 g(x),  s = 16/x^2 in (0,1], each a degree-DEG polynomial in
            the local variable u in [-1,1] on NSEG UNIFORM sub-intervals of s (Chebyshev interpolants
            converted to monomials for Horner/FMA evaluation).  Uniform segments make the segment index
            one multiply + float->int; 128 segments of degree 6 reach 7e-16, and the (F,G) coefficient
            pairs are stored interleaved in global memory (14 KB, L1-resident, read as double2).

Run:  python tools/gen_sici_tables.py   (prints the max relative error of F, G and of Si, Ci against mpmath)
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 50
NMAC = 17
DEG = 6
NSEG = 128
EDGES = [i / float(NSEG) for i in range(NSEG + 1)]   # uniform in s = 16/x^2

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hmvec_b200", "csrc", "sici_tables.inc")


def fg(x):
    x = mp.mpf(x)
    si, ci = mp.si(x), mp.ci(x)
    s, c = mp.sin(x), mp.cos(x)
    return ci * s + (mp.pi / 2 - si) * c, -ci * c + (mp.pi / 2 - si) * s


def F(s):
    if s == 0:
        return mp.mpf(1)
    x = 4 / mp.sqrt(s)
    return x * fg(x)[0]


def G(s):
    if s == 0:
        return mp.mpf(1)
    x = 4 / mp.sqrt(s)
    return x * x * fg(x)[1]


def cheb_monomial(fun, a, b, deg):
    """Chebyshev interpolant of fun on [a,b] -> monomial coefficients in u=(s-mid)/half (mp precision)."""
    n = deg + 1
    a, b = mp.mpf(a), mp.mpf(b)
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    vals = [fun((a + b) / 2 + (b - a) / 2 * t) for t in nodes]
    c = [2 * sum(vals[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) / n for j in range(n)]
    c[0] /= 2
    # T_j(u) monomial expansion by recurrence
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        prev, prev2 = T[j - 1], T[j - 2]
        cur = [mp.mpf(0)] + [2 * v for v in prev]
        for i, v in enumerate(prev2):
            cur[i] -= v
        T.append(cur)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            mono[i] += c[j] * v
    return mono


def cheb_to_plain(fun, a, b, deg):
    """Chebyshev interpolant of fun on [a,b] as plain monomial coefficients in the ORIGINAL variable."""
    mono_u = cheb_monomial(fun, a, b, deg)            # in u = (w - mid)/half
    mid, half = (mp.mpf(a) + mp.mpf(b)) / 2, (mp.mpf(b) - mp.mpf(a)) / 2
    out = [mp.mpf(0)] * (deg + 1)
    # expand sum_i m_i ((w-mid)/half)^i
    for i, m in enumerate(mono_u):
        for k in range(i + 1):
            out[k] += m * mp.binomial(i, k) * (-mid) ** (i - k) / half ** i
    return out


def horner(co, u):
    r = np.zeros_like(u) + co[-1]
    for c in co[-2::-1]:
        r = r * u + c
    return r


def main():
    a = [mp.mpf((-1) ** n) / ((2 * n + 1) * mp.factorial(2 * n + 1)) for n in range(NMAC)]
    b = [mp.mpf((-1) ** (n + 1)) / ((2 * n + 2) * mp.factorial(2 * n + 2)) for n in range(NMAC)]
    nseg = len(EDGES) - 1
    Fc, Gc = [], []
    for lo, hi in zip(EDGES[:-1], EDGES[1:]):
        Fc.append([float(v) for v in cheb_monomial(F, lo, hi, DEG)])
        Gc.append([float(v) for v in cheb_monomial(G, lo, hi, DEG)])

    # ---- accuracy report (numpy double emulation of the device arithmetic) ----
    worst = dict(F=0.0, G=0.0, Si=0.0, Ci=0.0)
    for i, (lo, hi) in enumerate(zip(EDGES[:-1], EDGES[1:])):
        s = np.linspace(max(lo, 1e-12), hi, 24)
        u = (s - 0.5 * (lo + hi)) * (2.0 / (hi - lo))
        Fe = np.array([float(F(mp.mpf(v))) for v in s])
        Ge = np.array([float(G(mp.mpf(v))) for v in s])
        worst["F"] = max(worst["F"], np.max(np.abs(horner(Fc[i], u) / Fe - 1)))
        worst["G"] = max(worst["G"], np.max(np.abs(horner(Gc[i], u) / Ge - 1)))
    xs = np.geomspace(1e-6, 4.0, 600)
    z = xs * xs
    si = xs * horner([float(v) for v in a], z)
    ci = float(mp.euler) + np.log(xs) + z * horner([float(v) for v in b], z)
    sie = np.array([float(mp.si(mp.mpf(v))) for v in xs])
    cie = np.array([float(mp.ci(mp.mpf(v))) for v in xs])
    worst["Si"] = np.max(np.abs(si / sie - 1))
    worst["Ci"] = np.max(np.abs(ci - cie))  # absolute (Ci crosses zero near 0.6165)
    print("max rel err F %.2e  G %.2e ; Maclaurin Si rel %.2e  Ci abs %.2e" % (worst["F"], worst["G"], worst["Si"], worst["Ci"]))

    # ---- sin/cos kernels on |r| <= pi/4 for the device sincos used by the Si/Ci tail (arguments up to ~1e5):
    #      sin r = r * PS(r^2), cos r = PC(r^2); Chebyshev interpolants in w = r^2 on [0, (pi/4)^2]
    wmax = (mp.pi / 4) ** 2 * mp.mpf("1.02")
    ps = [float(v) for v in cheb_to_plain(lambda w: mp.mpf(1) if w == 0 else mp.sin(mp.sqrt(w)) / mp.sqrt(w), 0, wmax, 7)]
    pc = [float(v) for v in cheb_to_plain(lambda w: mp.cos(mp.sqrt(w)), 0, wmax, 8)]
    r = np.linspace(-np.pi / 4, np.pi / 4, 2001)
    w = r * r
    es = np.max(np.abs(r * horner(ps, w) - np.array([float(mp.sin(mp.mpf(float(v)))) for v in r])))
    ec = np.max(np.abs(horner(pc, w) - np.array([float(mp.cos(mp.mpf(float(v)))) for v in r])))
    print("sin/cos kernel abs err %.2e %.2e" % (es, ec))

    def arr(name, vals):
        return "static __device__ __constant__ double %s[%d] = {\n  %s\n};\n" % (
            name, len(vals), ",\n  ".join(repr(float(v)) for v in vals))

    with open(OUT, "w") as fh:
        fh.write("// GENERATED by tools/gen_sici_tables.py -- do not edit.  See that script for the derivation.\n")
        fh.write("#define HMV_SICI_NMAC %d\n#define HMV_SICI_DEG %d\n#define HMV_SICI_NSEG %d\n" % (NMAC, DEG, nseg))
        fh.write(arr("c_sin_k", ps))
        fh.write(arr("c_cos_k", pc))
        fh.write(arr("c_si_mac", a))
        fh.write(arr("c_ci_mac", b))
        fh.write("// (F,G) monomial coefficients in u = 2 (s NSEG - seg) - 1, interleaved: g_sici_FG[seg*(DEG+1)+i] = {F_i, G_i}\n")
        fh.write("static __device__ const double2 g_sici_FG[%d] = {\n" % (nseg * (DEG + 1)))
        fh.write(",\n".join("  {%r, %r}" % (Fc[sg][i], Gc[sg][i]) for sg in range(nseg) for i in range(DEG + 1)))
        fh.write("\n};\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
