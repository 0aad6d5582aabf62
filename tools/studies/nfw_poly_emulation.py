#!/usr/bin/env python
"""Float64 emulation of the planned piecewise-polynomial NFW kernel, against mpmath.

Per interval [lo, hi] of x c (edges below) and per halo: u is sampled at the n = degree+1 Chebyshev nodes of
y = (x c)^2, the n monomial coefficients in the local variable t in [-1, 1] come from ONE precomputed n x n matrix
M_n = (Chebyshev -> monomial) . (node values -> Chebyshev coefficients), and the kernel evaluates Horner in t.
The script builds M_n in 40-digit arithmetic, rounds it to float64, runs the float64 pipeline (node values rounded,
float64 matrix-vector product, float64 Horner) and reports the worst error relative to max|u| of the interval.
CPU only; nothing here is imported by the product."""
import numpy as np
import mpmath as mp

mp.mp.dps = 40
EDGES = [0.0, 1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0, 14.0, 16.0] + [20.0 + 4.0 * i for i in range(12)]   # ... 64
DEGREE = [5, 6, 8, 9, 9, 10, 10, 11, 11] + [13] * 12


def u_exact(xc, c):
    if xc == 0:
        return mp.mpf(1)
    x = mp.mpf(xc) / c
    X = (1 + c) * x
    mc = mp.log(1 + c) - c / (1 + c)
    return (mp.sin(x) * (mp.si(X) - mp.si(x)) - mp.sin(c * x) / X + mp.cos(x) * (mp.ci(X) - mp.ci(x))) / mc


def matrix(n):
    """M[j][i]: monomial coefficient j from node value i (nodes t_i = cos(pi (i + 1/2)/n))."""
    nodes = [mp.cos(mp.pi * (i + mp.mpf(1) / 2) / n) for i in range(n)]
    # Chebyshev coefficients a_k = (2 - [k==0])/n sum_i f_i T_k(t_i)
    A = mp.matrix(n, n)
    for k in range(n):
        for i in range(n):
            A[k, i] = (1 if k == 0 else 2) * mp.cos(k * mp.acos(nodes[i])) / n
    # T_k in monomials
    T = [[mp.mpf(0)] * n for _ in range(n)]
    T[0][0] = mp.mpf(1)
    if n > 1:
        T[1][1] = mp.mpf(1)
    for k in range(2, n):
        for j in range(n):
            T[k][j] = (2 * T[k - 1][j - 1] if j > 0 else 0) - T[k - 2][j]
    M = mp.matrix(n, n)
    for j in range(n):
        for i in range(n):
            M[j, i] = sum(T[k][j] * A[k, i] for k in range(n))
    return nodes, M


if __name__ == "__main__":
    mats = {}
    worst = 0.0
    print("interval      degree   max |error| / max|u|  for c = 2.5, 4, 6, 10")
    for (lo, hi), d in zip(zip(EDGES[:-1], EDGES[1:]), DEGREE):
        n = d + 1
        if n not in mats:
            nodes, M = matrix(n)
            mats[n] = ([float(t) for t in nodes], np.array([[float(M[j, i]) for i in range(n)] for j in range(n)]))
        tn, M64 = mats[n]
        ylo, yhi = lo * lo, hi * hi
        errs = []
        for c in (2.5, 4.0, 6.0, 10.0):
            f = np.array([float(u_exact(mp.sqrt(0.5 * (ylo + yhi) + 0.5 * (yhi - ylo) * mp.mpf(t)), mp.mpf(c))) for t in tn])
            m = M64 @ f                                   # float64, as the pre-pass kernel would
            ts = np.linspace(-1.0, 1.0, 41)
            ref = np.array([float(u_exact(mp.sqrt(0.5 * (ylo + yhi) + 0.5 * (yhi - ylo) * mp.mpf(t)), mp.mpf(c))) for t in ts])
            p = np.zeros_like(ts)
            for coef in m[::-1]:
                p = p * ts + coef                         # float64 Horner, as the cube kernel would
            errs.append(np.max(np.abs(p - ref)) / np.max(np.abs(ref)))
        worst = max(worst, max(errs))
        print("[%4.1f, %4.1f]   %2d       %s" % (lo, hi, d, "  ".join("%.1e" % e for e in errs)))
    print("worst: %.2e ; largest |M| entry per n: %s" % (worst, {n: float(np.max(np.abs(m[1]))) for n, m in mats.items()}))
