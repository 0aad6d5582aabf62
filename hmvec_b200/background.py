"""Analytic flat-LCDM background used when CAMB is not installed (accuracy='low' only).

Host-side producer of the per-redshift scalars the device path takes as input (H(z), chi(z)); it exposes the four
calls the reference makes on camb's results object (cosmology.py:83-130):  hubble_parameter [km/s/Mpc],
h_of_z [1/Mpc], comoving_radial_distance, angular_diameter_distance, get_Omega.
    H(z) = H0 sqrt(Om (1+z)^3 + 1 - Om),  Om = (ombh2 + omch2)/h^2,  chi(z) = c int_0^z dz'/H  (Gauss-Legendre).
"""
import numpy as np

_C_KMS = 299792.458
_NODES, _WEIGHTS = np.polynomial.legendre.leggauss(24)


class AnalyticBackground(object):
    def __init__(self, H0, ombh2, omch2):
        self.H0 = float(H0)
        self.omm = (ombh2 + omch2) / (self.H0 / 100.0) ** 2

    def hubble_parameter(self, z):
        z = np.asarray(z, dtype=np.float64)
        return self.H0 * np.sqrt(self.omm * (1.0 + z) ** 3 + (1.0 - self.omm))

    def h_of_z(self, z):
        return self.hubble_parameter(z) / _C_KMS

    def comoving_radial_distance(self, z):
        scalar = np.ndim(z) == 0
        zz = np.atleast_1d(np.asarray(z, dtype=np.float64)).reshape(-1)
        # chi = c int_0^{ln(1+z)} e^t dt / H(e^t - 1): 2 panels of 24-point Gauss-Legendre in t, vectorised over z
        # (4e-16 of 4 x 96 points up to z = 1100; an eighth of the evaluations)
        npan = 2
        tmax = np.log1p(zz)[:, None, None]
        lo = tmax * (np.arange(npan) / npan)[None, :, None]
        half = 0.5 * tmax / npan
        t = lo + half * (_NODES[None, None, :] + 1.0)
        f = np.exp(t) * _C_KMS / self.hubble_parameter(np.expm1(t))
        out = np.sum(half * _WEIGHTS[None, None, :] * f, axis=(1, 2))
        return float(out[0]) if scalar else out.reshape(np.shape(z))

    def angular_diameter_distance(self, z):
        return self.comoving_radial_distance(z) / (1.0 + np.asarray(z, dtype=np.float64))

    def angular_diameter_distance2(self, z1, z2):
        return (self.comoving_radial_distance(z2) - self.comoving_radial_distance(z1)) / (1.0 + np.asarray(z2))

    def get_Omega(self, what):
        return 0.0
