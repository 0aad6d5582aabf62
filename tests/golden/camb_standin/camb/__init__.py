"""Test-side stand-in for the (absent) `camb` package.

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference
(`/root/reference/hmvec`) can be imported in the build container to generate golden
vectors (tests/golden/make_golden.py).  It is never imported by the product package
`hmvec_b200` and must never shadow a real camb: make_golden.py only prepends this
directory to sys.path after checking that `import camb` fails.

It provides the handful of calls the reference makes at
cosmology.py:164-179 (set_params / get_background) and cosmology.py:83-130
(hubble_parameter, h_of_z, comoving_radial_distance, angular_diameter_distance,
get_Omega) with flat-LCDM closed forms:
    H(z)  = H0 sqrt(Om (1+z)^3 + 1 - Om),   Om = (ombh2+omch2)/h^2
    chi(z)= c * int_0^z dz'/H(z')           (Gauss-Legendre in ln(1+z), 4 panels x 128 nodes)
    D_A   = chi/(1+z)
"""
import numpy as np
from . import model  # noqa: F401  (reference does `from camb import model`)

_C_KMS = 299792.458
_GL_X, _GL_W = np.polynomial.legendre.leggauss(128)


class _Pars(object):
    def __init__(self, **kw):
        self.kw = dict(kw)
        self.YHe = kw.get("YHe", None)
        if self.YHe is None:
            self.YHe = 0.24
        self.WantTransfer = False
        self.WantTensors = False
        self.H0 = kw["H0"]
        self.ombh2 = kw["ombh2"]
        self.omch2 = kw["omch2"]
        self.ns = kw.get("ns", 0.965)
        self.As = kw.get("As", 2.2e-9)
        self.NonLinear = None

    def set_matter_power(self, redshifts=None, kmax=None, silent=True, **kw):
        self.pk_redshifts, self.pk_kmax = list(redshifts or []), kmax
        return self


class _Background(object):
    def __init__(self, pars):
        self.H0 = float(pars.H0)
        h = self.H0 / 100.0
        self.om = (pars.ombh2 + pars.omch2) / h ** 2

    def hubble_parameter(self, z):
        z = np.asarray(z, dtype=np.float64)
        return self.H0 * np.sqrt(self.om * (1.0 + z) ** 3 + (1.0 - self.om))

    def h_of_z(self, z):
        return self.hubble_parameter(z) / _C_KMS

    def comoving_radial_distance(self, z):
        zz = np.atleast_1d(np.asarray(z, dtype=np.float64)).reshape(-1)
        # substitute t = ln(1+z'): chi = c int_0^{ln(1+z)} e^t / H dt, 4 panels x 128 nodes (good to z ~ 1100)
        chi = np.zeros(zz.size)
        for p in range(4):
            lo = np.log1p(zz) * (p / 4.0)
            hi = np.log1p(zz) * ((p + 1) / 4.0)
            t = lo[:, None] + 0.5 * (hi - lo)[:, None] * (_GL_X[None, :] + 1.0)
            f = np.exp(t) * _C_KMS / self.hubble_parameter(np.expm1(t))
            chi += 0.5 * (hi - lo) * np.sum(f * _GL_W[None, :], axis=1)
        if np.ndim(z) == 0:
            return float(chi[0])
        return chi.reshape(np.shape(z))

    def angular_diameter_distance(self, z):
        return self.comoving_radial_distance(z) / (1.0 + np.asarray(z, dtype=np.float64))

    def angular_diameter_distance2(self, z1, z2):
        # flat background: D_A(z1 -> z2) = (chi(z2) - chi(z1)) / (1 + z2)
        return (self.comoving_radial_distance(z2) - self.comoving_radial_distance(z1)) / (1.0 + np.asarray(z2, dtype=np.float64))

    def get_Omega(self, what):
        return 0.0


class _MatterPower(object):
    """Stand-in for the object camb.get_matter_power_interpolator returns: P(z,k) in Mpc^3 with the call convention
    the reference uses, `PK.P(z, k, grid=True)` -> [nz,nk].  A smooth closed form (BBKS transfer function, Carroll-
    Press-Turner growth), NOT a Boltzmann solution: it only has to be one deterministic function that the reference
    and hmvec_b200 both receive when accuracy='medium'/'high' is exercised without CAMB."""

    def __init__(self, pars):
        h = float(pars.H0) / 100.0
        self.h, self.ns, self.As = h, float(pars.ns), float(pars.As)
        self.om = (pars.ombh2 + pars.omch2) / h ** 2
        self.gam = self.om * h * np.exp(-pars.ombh2 / h ** 2 * (1.0 + np.sqrt(2.0 * h) / self.om))

    def _growth(self, z):
        a3 = (1.0 + np.asarray(z, dtype=np.float64)) ** 3
        omz = self.om * a3 / (self.om * a3 + 1.0 - self.om)
        olz = 1.0 - omz
        g = 2.5 * omz / (omz ** (4.0 / 7.0) - olz + (1.0 + omz / 2.0) * (1.0 + olz / 70.0))
        g0 = 2.5 * self.om / (self.om ** (4.0 / 7.0) - (1.0 - self.om) + (1.0 + self.om / 2.0) * (1.0 + (1.0 - self.om) / 70.0))
        return g / g0 / (1.0 + np.asarray(z, dtype=np.float64))

    def P(self, z, k, grid=True):
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        k = np.atleast_1d(np.asarray(k, dtype=np.float64))
        q = k / (self.gam * self.h)
        T = np.log(1.0 + 2.34 * q) / (2.34 * q) * (1.0 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
        pk0 = 2.0e13 * self.As / 2.2e-9 * k * (k / 0.05) ** (self.ns - 1.0) * T ** 2
        return self._growth(z)[:, None] ** 2 * pk0[None, :]


def get_matter_power_interpolator(pars, nonlinear=False, hubble_units=False, k_hunit=False, kmax=None, zmax=None,
                                  var1=None, var2=None, **kw):
    return _MatterPower(pars)


def set_params(**kw):
    return _Pars(**kw)


def get_background(pars):
    return _Background(pars)
