#!/usr/bin/env python
"""Feasibility study for the next NFW kernel: how many polynomial terms does u_NFW(x c; c) need on
sub-intervals of 0 < x c <= 16 if each halo carries piecewise Chebyshev fits instead of one Maclaurin series?

u is evaluated with mpmath (30 digits) from the closed form of hmvec.py:349-352; for every concentration and interval
the script reports the lowest Chebyshev degree whose interpolant (in the variable y = (x c)^2, like the present series)
reaches 1e-12 relative to max|u| on the interval, next to the Maclaurin term count the kernel uses today.
CPU only; nothing here is imported by the product."""
import numpy as np
import mpmath as mp

mp.mp.dps = 30


def u_exact(xc, c):
    x = mp.mpf(xc) / c
    X = (1 + c) * x
    mc = mp.log(1 + c) - c / (1 + c)
    si = lambda t: mp.si(t)
    ci = lambda t: mp.ci(t)
    return (mp.sin(x) * (si(X) - si(x)) - mp.sin(c * x) / X + mp.cos(x) * (ci(X) - ci(x))) / mc


def cheb_degree(c, lo, hi, tol=1e-12, dmax=40):
    ylo, yhi = lo * lo, hi * hi
    test = np.linspace(ylo, yhi, 201)
    ref = np.array([float(u_exact(np.sqrt(y), c)) for y in test])
    scale = np.max(np.abs(ref))
    for d in range(2, dmax):
        k = np.arange(d + 1)
        nodes = 0.5 * (ylo + yhi) + 0.5 * (yhi - ylo) * np.cos(np.pi * (k + 0.5) / (d + 1))
        vals = np.array([float(u_exact(np.sqrt(y), c)) for y in nodes])
        coef = np.polynomial.chebyshev.chebfit(2 * (nodes - ylo) / (yhi - ylo) - 1, vals, d)
        fit = np.polynomial.chebyshev.chebval(2 * (test - ylo) / (yhi - ylo) - 1, coef)
        if np.max(np.abs(fit - ref)) <= tol * scale:
            return d
    return None


def maclaurin_terms(xc):
    n = 5 if xc < 0.03 else 7 if xc < 0.3 else 10 if xc < 1.0 else min(41, int(1.8 * xc + 10.5))
    return n | 1


if __name__ == "__main__":
    edges = [0.0, 1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0, 14.0, 16.0]
    print("interval (x c)      Maclaurin terms today   Chebyshev degree in y for 1e-12 (c = 3, 6, 10)")
    for lo, hi in zip(edges[:-1], edges[1:]):
        degs = [cheb_degree(mp.mpf(c), max(lo, 1e-3), hi) for c in (3, 6, 10)]
        print("[%4.1f, %4.1f]        %2d .. %2d               %s" % (lo, hi, maclaurin_terms(max(lo, 1e-3)), maclaurin_terms(hi), degs))
