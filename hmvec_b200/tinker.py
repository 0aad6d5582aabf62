"""Tinker et al. 2010 mass function and bias (mirror of the reference's hmvec/tinker.py:26-90).

`nu = deltac / sigma`, arrays shaped (nz, nm).  On the hot path these formulas run on the device
(`hmv_mass_function_tinker`, csrc/k_sigma2.cu); this module holds the host-side pieces the device path needs --
the per-redshift parameters (`redshift_parameters`) with alpha(z) from the normalisation table -- and numpy mirrors of
the reference's free functions for callers that use them directly.

The table `data/alpha_consistency.txt` (1000 rows, z = linspace(0,3,1000)) is the reference's own data file
(hmvec/data/alpha_consistency.txt): alpha(z) such that int b(nu) f(nu) dnu = 1.  The reference looks for it one
directory too high (tinker.py:64), so `mass_function='tinker'` raises there; here the intended file is read.
"""
import os

import numpy as np

constants = {'deltac': 1.686}
default_params = {'tinker_f_nu_alpha_z0_delta_200': 0.368}        # Tinker et al. 2010, table 4

_ALPHA_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "alpha_consistency.txt")
_alpha_cache = None


def _alpha_of_z(z):
    """Linear interpolation in the normalisation table; outside [0,3] is an error, as with the reference's
    interp1d(bounds_error=True) (tinker.py:65-66)."""
    global _alpha_cache
    if _alpha_cache is None:
        _alpha_cache = np.loadtxt(_ALPHA_TABLE, unpack=True)
    tz, ta = _alpha_cache
    z = np.asarray(z, dtype=np.float64)
    if np.any(z < tz[0]) or np.any(z > tz[-1]):
        raise ValueError("A value in x_new is outside the interpolation range.")
    return np.interp(z, tz, ta)


def clamp_redshift(zs):
    """tinker.py:56: z*H(3-z) + 3*H(z-3) with H(0)=0 -- redshifts above 3 use z=3 (and z == 3 exactly maps to 0,
    a quirk of the two half-open step functions that is kept for parity)."""
    zs = np.asarray(zs, dtype=np.float64)
    return zs * np.heaviside(3 - zs, 0) + 3 * np.heaviside(zs - 3, 0)


def redshift_parameters(zs, norm_consistency=True, alpha=default_params['tinker_f_nu_alpha_z0_delta_200']):
    """[nz,5] array (alpha, beta, phi, eta, gamma) of f(nu) at each redshift (tinker.py:56-66); the device kernel's
    per-z input."""
    zc = clamp_redshift(np.asarray(zs, dtype=np.float64).reshape(-1))
    out = np.empty((zc.size, 5))
    out[:, 0] = _alpha_of_z(zc) if norm_consistency else alpha
    out[:, 1] = 0.589 * (1 + zc) ** 0.20
    out[:, 2] = -0.729 * (1 + zc) ** (-0.08)
    out[:, 3] = -0.243 * (1 + zc) ** 0.27
    out[:, 4] = 0.864 * (1 + zc) ** (-0.01)
    return out


def bias(nu, delta=200.):
    """Tinker 2010 eq. 6 (tinker.py:26-40)."""
    dc = constants['deltac']
    y = np.log10(delta)
    ey = np.exp(-(4. / y) ** 4.)
    A, a = 1. + 0.24 * y * ey, 0.44 * y - 0.88
    C = 0.019 + 0.107 * y + 0.19 * ey
    nua = nu ** a
    return 1 - A * nua / (nua + dc ** a) + 0.183 * nu ** 1.5 + C * nu ** 2.4


def f_nu(nu, zs, delta=200., norm_consistency=True, alpha=default_params['tinker_f_nu_alpha_z0_delta_200']):
    """f(nu) of Tinker 2010 (tinker.py:43-67); the multiplicity function is nu*f(nu) (hmvec.py:145)."""
    assert np.isclose(delta, 200.), "delta!=200 note implemented yet."
    zc = clamp_redshift(zs)
    beta = 0.589 * (1 + zc) ** 0.20
    phi = -0.729 * (1 + zc) ** (-0.08)
    eta = -0.243 * (1 + zc) ** 0.27
    gamma = 0.864 * (1 + zc) ** (-0.01)
    if norm_consistency:
        alpha = _alpha_of_z(zc)
    return alpha * (1. + (beta * nu) ** (-2. * phi)) * nu ** (2 * eta) * np.exp(-gamma * nu ** 2. / 2.)


def simple_f_nu(nu, delta=200.):
    """Tinker 2008 form (tinker.py:70-78)."""
    assert np.isclose(delta, 200.), "delta!=200 note implemented yet."
    sigma = constants['deltac'] / nu
    return 0.186 * (1. + (sigma / 2.57) ** (-1.47)) * np.exp(-1.19 / sigma ** 2.)


def NlnMsub(Msubs, Mhosts):
    """Subhalo mass function dN/dlnM_sub, Tinker & Wetzel 2010 eq. 12 (tinker.py:81-90): (Msubs, Mhosts) grid."""
    r = np.asarray(Msubs)[:, None] / np.asarray(Mhosts)[None, :]
    return 0.3 * r ** (-0.7) * np.exp(-9.9 * r ** 2.5)
