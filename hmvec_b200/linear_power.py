"""Host-side EH98 transfer function (the synthetic linear P(k) generator used with accuracy='low').

Plays the role of cosmology.py:404-504 (`Cosmology.Tk`), i.e. Eisenstein & Hu 1998 (ApJ 496, 605): the zero-baryon
shape fit, eqs. 28-31, and the full CDM+baryon fit with acoustic oscillations, eqs. 2-24.  Pure numpy, O(nk)."""
import numpy as np


def eisenstein_hu(k_mpc, h, omch2, ombh2, omm0, wiggles=True, tcmb=2.726):
    k = np.asarray(k_mpc, dtype=np.float64) / h            # h/Mpc
    wm, wb = omch2 + ombh2, ombh2
    fb, fc = wb / wm, omch2 / wm
    t2 = (tcmb / 2.7) ** 2
    k_eq = 7.46e-2 * wm / t2 / h                            # eq. 3
    z_eq = 2.50e4 * wm / t2 ** 2                            # eq. 2
    zb1 = 0.313 * wm ** -0.419 * (1.0 + 0.607 * wm ** 0.674)
    zb2 = 0.238 * wm ** 0.223
    z_d = 1291.0 * wm ** 0.251 / (1.0 + 0.659 * wm ** 0.828) * (1.0 + zb1 * wb ** zb2)   # eq. 4
    Rd = 31.5 * wb / t2 ** 2 * (1.0e3 / z_d)                # eq. 5
    Req = 31.5 * wb / t2 ** 2 * (1.0e3 / z_eq)
    s = 2.0 / (3.0 * k_eq) * np.sqrt(6.0 / Req) * np.log((np.sqrt(1.0 + Rd) + np.sqrt(Req + Rd)) / (1.0 + np.sqrt(Req)))  # eq. 6
    k_silk = 1.6 * wb ** 0.52 * wm ** 0.73 * (1.0 + (10.4 * wm) ** -0.95) / h              # eq. 7
    if not wiggles:
        ag = 1.0 - 0.328 * np.log(431.0 * wm) * fb + 0.38 * np.log(22.3 * wm) * fb ** 2    # eq. 31
        geff = omm0 * h * (ag + (1.0 - ag) / (1.0 + (0.43 * k * s) ** 4))                   # eq. 30
        q = k * t2 / geff
        L = np.log(2.0 * np.e + 1.8 * q)
        Cq = 14.2 + 731.0 / (1.0 + 62.5 * q)
        return L / (L + Cq * q * q)                                                         # eq. 29

    def T0(kk, alpha, beta):                                # eqs. 10, 19, 20
        q = kk / (13.41 * k_eq)
        L = np.log(np.e + 1.8 * beta * q)
        Cq = 14.2 / alpha + 386.0 / (1.0 + 69.9 * q ** 1.08)
        return L / (L + Cq * q * q)

    a1 = (46.9 * wm) ** 0.670 * (1.0 + (32.1 * wm) ** -0.532)   # eqs. 11, 12
    a2 = (12.0 * wm) ** 0.424 * (1.0 + (45.0 * wm) ** -0.582)
    alpha_c = a1 ** -fb * a2 ** (-fb ** 3)
    b1 = 0.944 / (1.0 + (458.0 * wm) ** -0.708)
    b2 = (0.395 * wm) ** -0.0266
    beta_c = 1.0 / (1.0 + b1 * (fc ** b2 - 1.0))
    f = 1.0 / (1.0 + (k * s / 5.4) ** 4)                    # eq. 18
    Tc = f * T0(k, 1.0, beta_c) + (1.0 - f) * T0(k, alpha_c, beta_c)   # eq. 17
    y = (1.0 + z_eq) / (1.0 + z_d)
    sq = np.sqrt(1.0 + y)
    G = y * (-6.0 * sq + (2.0 + 3.0 * y) * np.log((sq + 1.0) / (sq - 1.0)))   # eq. 15
    alpha_b = 2.07 * k_eq * s * (1.0 + Rd) ** -0.75 * G     # eq. 14
    beta_node = 8.41 * wm ** 0.435                          # eq. 23
    s_t = s / (1.0 + (beta_node / (k * s)) ** 3) ** (1.0 / 3.0)   # eq. 22
    beta_b = 0.5 + fb + (3.0 - 2.0 * fb) * np.sqrt((17.2 * wm) ** 2 + 1.0)   # eq. 24
    Tb = (T0(k, 1.0, 1.0) / (1.0 + (k * s / 5.2) ** 2)
          + alpha_b / (1.0 + (beta_b / (k * s)) ** 3) * np.exp(-(k / k_silk) ** 1.4)) * np.sinc(k * s_t / np.pi)  # eq. 21
    return fb * Tb + fc * Tc                                # eq. 16
