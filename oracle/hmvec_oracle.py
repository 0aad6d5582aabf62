"""CPU oracle for the hmvec halo-model hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module.  The product package `hmvec_b200` never does: its path is
CUDA-only and fails loudly when the extension is missing.

What this is: a numpy/scipy float64 restatement of the algorithm the reference
(simonsobs/hmvec, mounted at /root/reference in the build container) runs on the path
    sigma^2(R,z) -> mass function / bias -> u(k|M,z) profiles -> HOD -> 1h/2h mass integrals -> Limber
written against plain arrays (no class state shared with the product), every function citing
the reference file:line whose arithmetic it follows.  The numerical *method* of each step is kept
identical to the reference (Simpson on the geomspace grid, rFFT sine transform + per-halo
np.interp, table-inverse SHMR, all-z bisection, trapezoid in linear M, bilinear Limber lookup)
because parity to rtol 1e-6 depends on those choices, not only on the maths.

Parity pin: `tests/golden/*.npz` were produced by running the UNMODIFIED reference in the build
container (numpy 2.3.5, scipy 1.18.1, camb replaced by tests/golden/camb_standin, accuracy='low');
`tests/test_oracle_golden.py` checks every function here against them.  Third-party arithmetic
(numpy pocketfft/interp/gradient, scipy sici/erf/simpson/newton/FITPACK) is un-pinned by the
reference itself (its pyproject lists no versions), so the effective pin is that environment.
The one function that cannot be executed from the reference under SciPy>=1.14 is
`limber_integral` (interp2d / dfitpack.bispeu were removed); it is restated here with
RectBivariateSpline(kx=ky=1).ev, SciPy's documented bug-for-bug replacement, and additionally
checked against a hand-written clamped bilinear interpolation.
"""
import numpy as np
from scipy.special import sici, erf, hyp2f1
from scipy.integrate import simpson
from scipy.optimize import newton
from scipy.interpolate import RectBivariateSpline
import scipy.constants as sc

try:  # numpy>=2 renamed trapz
    _trapz = np.trapezoid
except AttributeError:  # pragma: no cover
    _trapz = np.trapz

C_KMS = 299792.458  # cosmology.py:28

# ----------------------------------------------------------------------------------------------
# parameter tables (values are part of parity; reference: params.py:2-113)
# ----------------------------------------------------------------------------------------------
GAS_FAMILIES = {
    # (A0, alpha_m, alpha_z) for rho0, alpha, beta          params.py:3-24
    "AGN": dict(rho0=(4000.0, 0.29, -0.66), alpha=(0.88, -0.03, 0.19), beta=(3.83, 0.04, -0.025)),
    "SH": dict(rho0=(19000.0, 0.09, -0.95), alpha=(0.70, -0.017, 0.27), beta=(4.43, 0.005, 0.037)),
}
PRES_FAMILIES = {
    # (A0, alpha_m, alpha_z) for P0, xc, beta               params.py:27-37
    "pres": dict(P0=(18.1, 0.154, -0.758), xc=(0.497, -0.00865, 0.731), beta=(4.35, 0.0393, 0.415)),
}

DEFAULTS = dict(
    st_A=0.3222, st_a=0.707, st_p=0.3, st_deltac=1.686,                       # params.py:43-46
    sigma2_kmin=1e-4, sigma2_kmax=2000.0, sigma2_numks=10000, Wkr_taylor_switch=0.01,  # :47-50
    duffy_vir=(7.85, -0.081, -0.71), duffy_mean=(10.14, -0.081, -1.01),       # :53-58
    gas_gamma=-0.2, pres_gamma=-0.3, pres_alpha=1.0,                          # :65-69
    kstar_damping=0.01,                                                       # :72
    omch2=0.1198, ombh2=0.02225, H0=67.3, ns=0.9645, As=2.2e-9, pivot_scalar=0.05,  # :76-83
    parsec=3.08567758e16, mSun=1.989e30,                                      # :90-91
    hod_A_log10mthresh=1.0, hod_sig_log_mstellar=0.2, hod_alphasat=1.0,       # :97-99
    hod_Bsat=9.04, hod_betasat=0.74, hod_Bcut=1.65, hod_betacut=0.59,         # :100-103
    hod_bisect_lo=7.0, hod_bisect_hi=14.0, hod_bisect_rtol=1e-4,              # :104-106
)


# ----------------------------------------------------------------------------------------------
# background (stand-in for CAMB; mirrors tests/golden/camb_standin) and EH98 linear power
# ----------------------------------------------------------------------------------------------
class Background(object):
    """Flat-LCDM closed forms standing in for camb.get_background (cosmology.py:164-179)."""

    _x, _w = np.polynomial.legendre.leggauss(128)

    def __init__(self, p=None):
        q = dict(DEFAULTS)
        if p:
            q.update(p)
        self.p = q
        self.H0 = q["H0"]
        self.h = self.H0 / 100.0
        self.omm0 = (q["omch2"] + q["ombh2"]) / self.h ** 2          # cosmology.py:213-215
        self.oml0 = 1.0 - self.omm0                                   # cosmology.py:217 (omk=0)

    def hubble(self, z):  # km/s/Mpc
        z = np.asarray(z, dtype=np.float64)
        return self.H0 * np.sqrt(self.omm0 * (1 + z) ** 3 + 1.0 - self.omm0)

    def h_of_z(self, z):  # 1/Mpc
        return self.hubble(z) / C_KMS

    def chi(self, z):
        zz = np.atleast_1d(np.asarray(z, dtype=np.float64)).reshape(-1)
        out = np.zeros(zz.size)
        for p in range(4):  # t = ln(1+z'), 4 panels x 128 nodes
            lo, hi = np.log1p(zz) * (p / 4.0), np.log1p(zz) * ((p + 1) / 4.0)
            t = lo[:, None] + 0.5 * (hi - lo)[:, None] * (self._x[None, :] + 1.0)
            out += 0.5 * (hi - lo) * np.sum(self._w[None, :] * np.exp(t) * C_KMS / self.hubble(np.expm1(t)), axis=1)
        return out.reshape(np.shape(z)) if np.ndim(z) else float(out[0])

    # cosmology.py:239-243
    def rho_crit(self, z):
        Hsi = self.hubble(z) * 3.241e-20
        return 3.0 * Hsi ** 2 / 8.0 / np.pi / 6.67259e-11 * 1.477543e37

    # cosmology.py:232-234
    def rho_matter(self, z):
        return self.rho_crit(0.0) * self.omm0 * (1 + np.atleast_1d(z)) ** 3

    # hmvec.py:105-109 (Bryan & Norman)
    def deltav(self, z):
        x = self.rho_matter(z) / self.rho_crit(z) - 1.0
        return 18.0 * np.pi ** 2 + 82.0 * x - 39.0 * x ** 2


def growth_anorm(bg, a):
    """D(a) normalised to a in matter domination; cosmology.py:297-332 (type='anorm', exact=False)."""
    def d_over_a_times_a(aa):
        aa = np.asarray(aa, dtype=np.float64)
        x = (bg.oml0 / bg.omm0) ** (1.0 / 3.0) * aa
        return np.sqrt(1.0 + x ** 3) * hyp2f1(5.0 / 6.0, 1.5, 11.0 / 6.0, -x ** 3) * aa
    today = d_over_a_times_a(1.0)
    return d_over_a_times_a(a) / today * today  # val*mul with mul = D_approx(1), cosmology.py:321-331


def eh98_transfer(bg, ks):
    """Eisenstein & Hu 1998 transfer function with baryon wiggles; cosmology.py:404-504."""
    p = bg.p
    h = bg.h
    k = np.asarray(ks, dtype=np.float64) / h
    th2 = (2.726 / 2.7) ** 2
    wm = p["omch2"] + p["ombh2"]
    wb = p["ombh2"]
    fb = wb / wm
    fc = p["omch2"] / wm
    k_eq = 7.46e-2 * wm / th2 / h
    z_eq = 2.50e4 * wm / th2 ** 2
    bz1 = 0.313 * wm ** -0.419 * (1.0 + 0.607 * wm ** 0.674)
    bz2 = 0.238 * wm ** 0.223
    z_d = 1291.0 * wm ** 0.251 / (1.0 + 0.659 * wm ** 0.828) * (1.0 + bz1 * wb ** bz2)
    R_d = 31.5 * wb / th2 ** 2 * (1.0e3 / z_d)
    R_eq = 31.5 * wb / th2 ** 2 * (1.0e3 / z_eq)
    s = 2.0 / (3.0 * k_eq) * np.sqrt(6.0 / R_eq) * np.log(
        (np.sqrt(1.0 + R_d) + np.sqrt(R_eq + R_d)) / (1.0 + np.sqrt(R_eq)))
    k_silk = 1.6 * wb ** 0.52 * wm ** 0.73 * (1.0 + (10.4 * wm) ** -0.95) / h

    a1 = (46.9 * wm) ** 0.670 * (1.0 + (32.1 * wm) ** -0.532)
    a2 = (12.0 * wm) ** 0.424 * (1.0 + (45.0 * wm) ** -0.582)
    alpha_c = a1 ** -fb * a2 ** (-fb ** 3)
    c1 = 0.944 / (1.0 + (458.0 * wm) ** -0.708)
    c2 = (0.395 * wm) ** -0.0266
    beta_c = 1.0 / (1.0 + c1 * (fc ** c2 - 1.0))

    def t0(kk, alpha, beta):
        q = kk / (13.41 * k_eq)
        L = np.log(np.e + 1.8 * beta * q)
        C = 14.2 / alpha + 386.0 / (1.0 + 69.9 * q ** 1.08)
        return L / (L + C * q * q)

    f = 1.0 / (1.0 + (k * s / 5.4) ** 4)
    Tc = f * t0(k, 1.0, beta_c) + (1.0 - f) * t0(k, alpha_c, beta_c)
    y = (1.0 + z_eq) / (1.0 + z_d)
    x = np.sqrt(1.0 + y)
    G = y * (-6.0 * x + (2.0 + 3.0 * y) * np.log((x + 1.0) / (x - 1.0)))
    alpha_b = 2.07 * k_eq * s * (1.0 + R_d) ** -0.75 * G
    beta_node = 8.41 * wm ** 0.435
    s_tilde = s / (1.0 + (beta_node / (k * s)) ** 3) ** (1.0 / 3.0)
    beta_b = 0.5 + fb + (3.0 - 2.0 * fb) * np.sqrt((17.2 * wm) ** 2 + 1.0)
    Tb = (t0(k, 1.0, 1.0) / (1.0 + (k * s / 5.2) ** 2)
          + alpha_b / (1.0 + (beta_b / (k * s)) ** 3) * np.exp(-(k / k_silk) ** 1.4)) \
        * np.sinc(k * s_tilde / np.pi)
    return fb * Tb + fc * Tc


def plin_approx(bg, ks, zs):
    """Synthetic linear P(z,k) used with accuracy='low'; cosmology.py:391-402."""
    ks = np.asarray(ks, dtype=np.float64)
    zs = np.atleast_1d(np.asarray(zs, dtype=np.float64))
    p = bg.p
    tk = eh98_transfer(bg, ks)[None, :]
    D = growth_anorm(bg, 1.0 / (1.0 + zs))[:, None]
    omh2 = (p["omch2"] + p["ombh2"]) * 100.0 ** 2      # Omega_nu = 0
    kfac = (ks / p["pivot_scalar"]) ** (p["ns"] - 1.0) * ks
    pref = 8.0 * np.pi ** 2 * p["As"] / 25.0 / omh2 ** 2 * C_KMS ** 4
    return pref * kfac[None, :] * D ** 2 * tk ** 2


# ----------------------------------------------------------------------------------------------
# a1: sigma^2(R,z)
# ----------------------------------------------------------------------------------------------
def R_from_M(M, rho, delta):
    """hmvec.py:627-628"""
    return (3.0 * M / 4.0 / np.pi / delta / rho) ** (1.0 / 3.0)


def tophat_window(kR, taylor_switch=0.01):
    """cosmology.py:30-38: exact top-hat, 3-term Taylor below the switch."""
    kR = np.asarray(kR, dtype=np.float64)
    with np.errstate(all="ignore"):
        w = 3.0 * (np.sin(kR) - kR * np.cos(kR)) / kR ** 3
    small = kR < taylor_switch
    x2 = kR[small] ** 2
    w[small] = 1.0 - 0.1 * x2 + 0.00357142857143 * x2 * x2
    return w


def sigma2_grid(p=None):
    q = dict(DEFAULTS)
    if p:
        q.update(p)
    return np.geomspace(q["sigma2_kmin"], q["sigma2_kmax"], int(q["sigma2_numks"]))   # cosmology.py:254


def sigma2(Rs, ks_sig, sPzk, taylor_switch=0.01, zchunk=8):
    """cosmology.py:261-265: integrand P W^2 k^2/(2 pi^2), scipy Simpson on the geomspace grid.

    Rs [nm], ks_sig [nks], sPzk [nz,nks] -> [nz,nm].  Chunked over z only to bound memory."""
    W2 = tophat_window(ks_sig[None, :] * Rs[:, None], taylor_switch) ** 2       # [nm,nks]
    out = np.empty((sPzk.shape[0], Rs.size))
    for z0 in range(0, sPzk.shape[0], zchunk):
        integrand = sPzk[z0:z0 + zchunk, None, :] * W2[None] * ks_sig[None, None, :] ** 2 / 2.0 / np.pi ** 2
        out[z0:z0 + zchunk] = simpson(integrand, x=ks_sig, axis=-1)
    return out


# ----------------------------------------------------------------------------------------------
# a2: Sheth-Tormen mass function and bias
# ----------------------------------------------------------------------------------------------
def st_fsigma(s2, A=0.3222, a=0.707, p=0.3, dc=1.686):
    """hmvec.py:137-141"""
    sig = np.sqrt(s2)
    return A * np.sqrt(2.0 * a / np.pi) * (1.0 + (s2 / a / dc ** 2) ** p) * (dc / sig) * np.exp(-a * dc ** 2 / 2.0 / s2)


def st_bias(s2, a=0.707, p=0.3, dc=1.686):
    """hmvec.py:152-156"""
    nu2a = a * dc ** 2 / s2
    return 1.0 + (nu2a - 1.0) / dc + (2.0 * p / dc) / (1.0 + nu2a ** p)


def mass_function(s2, ms, rho_m0, fsigma=None, **st):
    """hmvec.py:178-185: n(M,z) = rho_m0 f(sigma) dln(sigma^-1)/dlnM / M^2 with np.gradient on ln M."""
    g = np.gradient(-0.5 * np.log(s2), np.log(ms), axis=-1)
    f = st_fsigma(s2, **st) if fsigma is None else fsigma
    return rho_m0 * f * g / ms[None, :] ** 2


def tinker_fsigma(s2, zs, alpha_table, dc=1.686):
    """Tinker et al. 2010 multiplicity nu f(nu) as hmvec.py:142-145 calls tinker.py:43-67: redshifts above 3 are
    evaluated at 3 (two half-open Heaviside steps: z == 3 exactly maps to 0), f's parameters scale with (1+z), and
    alpha(z) is linearly interpolated in the reference's table hmvec/data/alpha_consistency.txt
    (`alpha_table` = its two columns; out-of-range redshifts are an error there and here)."""
    tz, ta = alpha_table
    zs = np.asarray(zs, dtype=np.float64)
    zc = zs * np.heaviside(3 - zs, 0) + 3 * np.heaviside(zs - 3, 0)
    if np.any(zc < tz[0]) or np.any(zc > tz[-1]):
        raise ValueError("redshift outside the alpha(z) table")
    alpha = np.interp(zc, tz, ta)[:, None]
    beta = (0.589 * (1 + zc) ** 0.20)[:, None]
    phi = (-0.729 * (1 + zc) ** (-0.08))[:, None]
    eta = (-0.243 * (1 + zc) ** 0.27)[:, None]
    gamma = (0.864 * (1 + zc) ** (-0.01))[:, None]
    nu = dc / np.sqrt(s2)
    return nu * alpha * (1.0 + (beta * nu) ** (-2.0 * phi)) * nu ** (2 * eta) * np.exp(-gamma * nu ** 2 / 2.0)


def tinker_bias(s2, dc=1.686, delta=200.0):
    """tinker.py:26-40 (eq. 6 of Tinker et al. 2010) at nu = dc/sigma; the exponents use tinker.py's own 1.686."""
    nu = dc / np.sqrt(s2)
    y = np.log10(delta)
    ey = np.exp(-(4.0 / y) ** 4)
    A, a, C = 1.0 + 0.24 * y * ey, 0.44 * y - 0.88, 0.019 + 0.107 * y + 0.19 * ey
    return 1.0 - A * nu ** a / (nu ** a + 1.686 ** a) + 0.183 * nu ** 1.5 + C * nu ** 2.4


# ----------------------------------------------------------------------------------------------
# a3/a4: concentrations and the analytic NFW transform
# ----------------------------------------------------------------------------------------------
def duffy(ms, zs, h, A, alpha, beta):
    """hmvec.py:68-73"""
    return A * (h * ms[None, :] / 2.0e12) ** alpha * (1.0 + zs[:, None]) ** beta


def nfw_mc(c):
    """hmvec.py:737"""
    return np.log(1.0 + c) - c / (1.0 + c)


def uk_nfw_analytic(ks, zs, cs, rvirs):
    """hmvec.py:346-352: Si/Ci closed form of the truncated-NFW Fourier profile.  -> [nz,nm,nk]"""
    c = cs[..., None]
    rs = (rvirs / cs)[..., None]
    x = ks[None, None, :] * rs * (1.0 + zs[:, None, None])
    si1, ci1 = sici(x)
    si2, ci2 = sici((1.0 + c) * x)
    return (np.sin(x) * (si2 - si1) - np.sin(c * x) / ((1.0 + c) * x) + np.cos(x) * (ci2 - ci1)) / nfw_mc(c)


# ----------------------------------------------------------------------------------------------
# a5: mass-definition conversion
# ----------------------------------------------------------------------------------------------
def mdelta_convert(ms, cs, drho1, drho2):
    """hmvec.py:759-798: solve M1/mc(C1) = M2/mc(C2(M2)) in ln M2 by scipy's array secant."""
    M1 = ms[None, :] + 0.0 * cs
    r = (drho1 / drho2)[:, None]
    lhs = M1 / nfw_mc(cs)

    def resid(lnM2):
        c2 = cs * (np.exp(lnM2 - np.log(M1)) * r) ** (1.0 / 3.0)
        return lhs - np.exp(lnM2) / nfw_mc(c2)

    return np.exp(newton(resid, np.log(M1)))


# ----------------------------------------------------------------------------------------------
# a6: Battaglia GNFW profiles
# ----------------------------------------------------------------------------------------------
def _plaw(m200, z, triple):
    """hmvec.py:800-802"""
    A0, am, az = triple
    return A0 * (m200 / 1.0e14) ** am * (1.0 + z) ** az


def gas_density_x(x, m200, z, omb, omm, rhoc, gamma, fam):
    """hmvec.py:856-860 (note the sign convention of the second gamma)."""
    rho0 = _plaw(m200, z, fam["rho0"])
    al = _plaw(m200, z, fam["alpha"])
    be = _plaw(m200, z, fam["beta"])
    return (omb / omm) * rhoc * rho0 * x ** gamma * (1.0 + x ** al) ** (-(be + gamma) / al)


def gas_pressure_x(x, m200, r200, z, omb, omm, rhoc, alpha, gamma, fam, parsec, msun):
    """hmvec.py:918-927"""
    P0 = _plaw(m200, z, fam["P0"])
    xc = _plaw(m200, z, fam["xc"])
    be = _plaw(m200, z, fam["beta"])
    XH = 0.76
    efrac = 2.0 * (XH + 1.0) / (5.0 * XH + 3.0)
    G = sc.G / (parsec * 1e6) ** 3 * msun
    return efrac * (omb / omm) * 200.0 * m200 * G * rhoc / (2.0 * r200) * P0 * (x / xc) ** gamma \
        * (1.0 + (x / xc) ** alpha) ** (-be)


# ----------------------------------------------------------------------------------------------
# a7/a8: numerical profile transform (rFFT sine transform + per-halo interpolation)
# ----------------------------------------------------------------------------------------------
def sine_transform(xs, ys):
    """fft.py:44-51: int dx x sin(kx) y(x) via -Im rfft(x*y)*step with step=(x[-1]-x[0])/N."""
    N = xs.size
    step = (xs[-1] - xs[0]) / N
    U = -np.fft.rfft(xs * ys, axis=-1).imag * step
    kt = np.fft.rfftfreq(N, step) * 2.0 * np.pi
    return kt, U


def profile_transform(prof_of_x, cmaxs, rss, zs, ks, xmax, nxs, mass_norm=True):
    """fft.py:73-94 + the per-(z,M) interpolation loop fft.py:97-115.  -> [nz,nm,nk]"""
    xs = np.linspace(0.0, xmax, nxs + 1)[1:]
    rho = prof_of_x(xs)
    if rho.ndim == 1:
        rho = rho[None, None, :]
    rho = rho + 0.0 * cmaxs[..., None]
    inside = np.where(np.abs(xs)[None, None, :] > cmaxs[..., None], 0.0, 1.0)
    if mass_norm:
        mnorm = _trapz(inside * rho * xs ** 2, xs)
    else:
        mnorm = np.ones(cmaxs.shape)
    kt, U = sine_transform(xs, rho * inside)
    with np.errstate(all="ignore"):
        u = U / kt[None, None, :] / mnorm[..., None]
        kout = kt / rss[..., None] / (1.0 + zs[:, None, None])
    nz, nm = cmaxs.shape
    out = np.zeros((nz, nm, ks.size))
    for i in range(nz):
        for j in range(nm):
            good = kout[i, j] > 0
            kk = kout[i, j][good]
            uu = u[i, j][good]
            out[i, j] = np.interp(ks, kk, uu, left=uu[0], right=0)
    return out


# ----------------------------------------------------------------------------------------------
# a9/a10: HOD
# ----------------------------------------------------------------------------------------------
_SHMR_LO = (10.72, 0.55, 12.35, 0.28, 0.44, 0.18, 1.56, 2.51, 0.57, 0.17)     # z<=0.8, hmvec.py:668-677
_SHMR_HI = (11.09, 0.56, 12.27, -0.84, 0.65, 0.31, 1.12, -0.53, 0.56, -0.12)  # z>0.8,  hmvec.py:682-691


def shmr_log10mhalo(z, log10mstar):
    """hmvec.py:648-695.  z [nz,1] (or [nz]), log10mstar [1,n] or [nz,n] -> [nz,n]"""
    z = np.asarray(z, dtype=np.float64).reshape(-1, 1)
    L = np.asarray(log10mstar, dtype=np.float64) + 0.0 * z
    a1 = 1.0 / (1.0 + z) - 1.0
    out = np.empty(L.shape)
    for sel, P in ((z[:, 0] <= 0.8, _SHMR_LO), (z[:, 0] > 0.8, _SHMR_HI)):
        Ms0, Msa, M1, M1a, b0, ba, g0, ga, d0, da = P
        aa = a1[sel]
        d = L[sel] - (Ms0 + Msa * aa)
        out[sel] = -0.5 + (M1 + M1a * aa) + (b0 + ba * aa) * d \
            + 10.0 ** ((d0 + da * aa) * d) / (1.0 + 10.0 ** (-(g0 + ga * aa) * d))
    return out


def shmr_log10mstar(zs, log10mhalo):
    """hmvec.py:634-646: inverse SHMR through a 4000-point table and np.interp per z."""
    L = np.linspace(-18.0, 18.0, 4000)[None, :]
    mh = shmr_log10mhalo(zs, L)
    out = np.empty((np.size(zs), np.size(log10mhalo)))
    for i in range(out.shape[0]):
        out[i] = np.interp(np.ravel(log10mhalo), mh[i], L[0])
    return out


def hod_occupations(zs, ms, log10mthresh, hp, corr="max"):
    """hmvec.py:698-731.  log10mthresh [nz] -> Nc, Ns, NsNsm1, NcNs each [nz,nm]"""
    zs = np.asarray(zs, dtype=np.float64)
    lmh = np.log10(ms)[None, :]
    lth = np.asarray(log10mthresh, dtype=np.float64)[:, None]
    Nc = 0.5 * (1.0 - erf((lth - shmr_log10mstar(zs, lmh)) / (np.sqrt(2.0) * hp["hod_sig_log_mstellar"])))
    mth = shmr_log10mhalo(zs, lth)
    Msat = 1e12 * hp["hod_Bsat"] * 10.0 ** ((mth - 12.0) * hp["hod_betasat"])
    Mcut = 1e12 * hp["hod_Bcut"] * 10.0 ** ((mth - 12.0) * hp["hod_betacut"])
    masses = 10.0 ** lmh
    Ns = Nc * (masses / Msat) ** hp["hod_alphasat"] * np.exp(-Mcut / masses)
    if corr == "max":
        with np.errstate(all="ignore"):
            NsNsm1 = Ns ** 2 / Nc
        NsNsm1[np.isclose(Nc, 0.0)] = 0.0
        NcNs = Ns
    else:
        NsNsm1 = Ns ** 2
        NcNs = Ns * Nc
    return Nc, Ns, NsNsm1, NcNs


def hod_ngal(nzm, ms, Nc, Ns):
    """hmvec.py:956-957"""
    return _trapz(nzm * (Nc + Ns), ms, axis=-1)


def hod_bias(nzm, bh, ms, Nc, Ns, ngal):
    """hmvec.py:464-466"""
    return _trapz(nzm * (Nc + Ns) * bh, ms, axis=-1) / ngal


def bisect_all(target, x_of_y, lo, hi, rtol, decreasing=True, max_iter=200):
    """utils.py:19-42: every element keeps bisecting until *all* meet rtol; returns the last midpoint."""
    yl = target * 0 + lo
    yr = target * 0 + hi
    err = np.inf
    it = 0
    y = None
    while np.any(np.abs(err) > rtol):
        y = 0.5 * (yl + yr)
        err = (x_of_y(y) - target) / target
        up = err > 0
        if decreasing:
            yl[up] = y[up]
            yr[~up] = y[~up]
        else:
            yr[up] = y[up]
            yl[~up] = y[~up]
        it += 1
        if it > max_iter:
            raise RuntimeError("bisection did not converge")
    return y, it


def hod_solve_mthresh(ngal_target, zs, ms, nzm, hp):
    """hmvec.py:415-433"""
    f = lambda y: hod_ngal(nzm, ms, *hod_occupations(zs, ms, y, hp)[:2])
    y, it = bisect_all(np.asarray(ngal_target, dtype=np.float64), f, hp["hod_bisect_lo"], hp["hod_bisect_hi"],
                       hp["hod_bisect_rtol"], decreasing=True)
    return y * hp["hod_A_log10mthresh"], it


# ----------------------------------------------------------------------------------------------
# a12-a14: tracer terms and the 1-halo / 2-halo mass integrals
# ----------------------------------------------------------------------------------------------
class Tracer(object):
    """One leg of a spectrum.  kind in {'matter','hod','pressure'}."""

    def __init__(self, kind, u=None, uc=None, hod=None):
        self.kind, self.u, self.uc, self.hod = kind, u, uc, hod

    def term(self, ms, rho_m0, lowk=False):
        if self.kind == "matter":                                   # hmvec.py:488-492
            u = 1.0 if lowk else self.u
            return ms[None, :, None] * u / rho_m0
        if self.kind == "hod":                                      # hmvec.py:481-486
            uc = 1.0 if (lowk or self.uc is None) else self.uc
            us = 1.0 if lowk else self.u
            h = self.hod
            return (uc * h["Nc"][..., None] + us * h["Ns"][..., None]) / h["ngal"][..., None, None]
        if self.kind == "pressure":                                 # hmvec.py:494-497
            if lowk:
                return self.u[:, :, :1] + 0.0 * self.u
            return self.u
        raise ValueError(self.kind)

    def hod_square(self):                                           # hmvec.py:477-479
        uc = 1.0 if self.uc is None else self.uc
        h = self.hod
        return (2.0 * uc * self.u * h["NcNs"][..., None] + h["NsNsm1"][..., None] * self.u ** 2) \
            / h["ngal"][..., None, None] ** 2


def power_1halo(A, B, nzm, ms, ks, rho_m0, kstar=0.01):
    """hmvec.py:504-526"""
    if A.kind == "hod" and B.kind == "hod":
        sq = A.hod_square()
    elif A.kind == "pressure" and B.kind == "pressure":
        sq = A.term(ms, rho_m0) ** 2
    else:
        sq = A.term(ms, rho_m0) * B.term(ms, rho_m0)
    return _trapz(nzm[..., None] * sq, ms[:, None], axis=-2) * (1.0 - np.exp(-(ks / kstar) ** 2))


def power_2halo(A, B, nzm, bh, ms, Pzk, rho_m0, bA=None, bB=None):
    """hmvec.py:528-572"""
    def leg(T, b_in):
        I = _trapz(nzm[..., None] * T.term(ms, rho_m0) * bh[..., None], ms[:, None], axis=-2)
        if T.kind == "pressure":                                    # hmvec.py:541-545: b = C = 0
            b, C = 0.0, 0.0
        else:
            C = _trapz(nzm[..., None] * T.term(ms, rho_m0, lowk=True) * bh[..., None], ms[:, None], axis=-2)
            if T.kind == "matter":
                b = 1.0
            else:
                b = hod_bias(nzm, bh, ms, T.hod["Nc"], T.hod["Ns"], T.hod["ngal"])[:, None]
        if b_in is not None:
            b = np.asarray(b_in).reshape(-1, 1)
        return I + b - C
    return Pzk * leg(A, bA) * leg(B, bB)


# ----------------------------------------------------------------------------------------------
# a15/a16: lensing window and Limber
# ----------------------------------------------------------------------------------------------
def lensing_window(bg, ezs, zsrc, dndz=None):
    """cosmology.py:506-534"""
    ezs = np.asarray(ezs, dtype=np.float64)
    zsrc = np.array(zsrc, dtype=np.float64).reshape(-1)
    H0 = bg.h_of_z(0.0)
    H = bg.h_of_z(ezs)
    chis = bg.chi(ezs)
    chistar = bg.chi(zsrc)
    if zsrc.size == 1:
        integral = (chistar - chis) / chistar
        integral[ezs > zsrc] = 0
    else:
        nd = np.asarray(dndz, dtype=np.float64) / _trapz(dndz, zsrc)
        g = (chistar[None, :] - chis[:, None]) / chistar[None, :] * nd[None, :]
        g[zsrc[None, :] < ezs[:, None]] = 0
        integral = _trapz(g, zsrc, axis=-1)
    return 1.5 * bg.omm0 * H0 ** 2 * (1.0 + ezs) * chis / H * integral


def bilinear_clamped(ks, zs, Pzk, kq, zq):
    """Hand-written equivalent of FITPACK bispeu for a kx=ky=1 spline: clamp to the table, then bilinear."""
    kq = np.clip(kq, ks[0], ks[-1])
    zq = np.clip(zq, zs[0], zs[-1])
    ik = np.clip(np.searchsorted(ks, kq, side="right") - 1, 0, ks.size - 2)
    iz = np.clip(np.searchsorted(zs, zq, side="right") - 1, 0, zs.size - 2)
    tk = (kq - ks[ik]) / (ks[ik + 1] - ks[ik])
    tz = (zq - zs[iz]) / (zs[iz + 1] - zs[iz])
    p00, p01 = Pzk[iz, ik], Pzk[iz, ik + 1]
    p10, p11 = Pzk[iz + 1, ik], Pzk[iz + 1, ik + 1]
    return (1 - tz) * ((1 - tk) * p00 + tk * p01) + tz * ((1 - tk) * p10 + tk * p11)


def limber(ells, zs, ks, Pzk, gzs, W1, W2, hzs, chis, use_fitpack=True):
    """cosmology.py:882-904 with interp2d/bispeu -> RectBivariateSpline(kx=ky=1).ev (see module docstring)."""
    gzs = np.asarray(gzs, dtype=np.float64).reshape(-1)
    hzs = np.array(hzs, dtype=np.float64).reshape(-1)
    chis = np.array(chis, dtype=np.float64).reshape(-1)
    pref = hzs * np.array(W1, dtype=np.float64).reshape(-1) * np.array(W2, dtype=np.float64).reshape(-1) / chis ** 2
    if zs.size > 1:
        spl = RectBivariateSpline(ks, zs, Pzk.T, kx=1, ky=1) if use_fitpack else None
    out = np.zeros(len(ells))
    for i, ell in enumerate(ells):
        kev = (ell + 0.5) / chis
        if zs.size > 1:
            val = spl.ev(kev, gzs) if use_fitpack else bilinear_clamped(ks, zs, Pzk, kev, gzs)
        else:
            val = np.interp(kev, ks, Pzk[0])
        out[i] = (val * pref)[0] if gzs.size == 1 else _trapz(val * pref, gzs)
    return out


# ----------------------------------------------------------------------------------------------
# P(z,k) interpolator in front of the path: utils.py:53-182, cosmology.py:227-229,353-382
# ----------------------------------------------------------------------------------------------
class PKOracle(object):
    """RectBivariateSpline in (z, ln k) of log|P| (or of P when it changes sign), bicubic when the table allows,
    with the optional two-node power-law extension to extrap_kmax; P(z,k) = sign*exp(spline) (utils.py:95-103)."""

    def __init__(self, ks, zs, pk, log_interp=True, extrap_kmax=None):
        from scipy.interpolate import RectBivariateSpline
        ks, zs, pk = (np.asarray(a, dtype=np.float64) for a in (ks, zs, pk))
        self.sign = 1
        if log_interp and np.any(pk <= 0):                                               # utils.py:139-144
            if np.all(pk < 0):
                self.sign = -1
            else:
                log_interp = False
        vals = np.log(self.sign * pk) if log_interp else pk
        logk = np.log(ks)
        if extrap_kmax and extrap_kmax > ks[-1]:                                         # utils.py:150-169
            top = np.log(extrap_kmax)
            delta = top - logk[-1]
            ext = np.empty((vals.shape[0], vals.shape[1] + 2))
            ext[:, :-2] = vals
            dlog = (ext[:, -3] - ext[:, -4]) / (logk[-1] - logk[-2])
            ext[:, -1] = ext[:, -3] + dlog * delta
            ext[:, -2] = ext[:, -3] + dlog * delta * 0.9
            logk = np.hstack((logk, top - delta * 0.1, top))
            vals = ext
        self.islog = bool(log_interp)
        self.spl = RectBivariateSpline(zs, logk, vals, kx=min(len(zs) - 1, 3), ky=min(len(logk) - 1, 3))   # :171-172

    def P(self, z, k, grid=True):
        v = self.spl(z, np.log(k), grid=grid)
        return self.sign * np.exp(v) if self.islog else v


# ----------------------------------------------------------------------------------------------
# A small driver object so parity tests and the CPU baseline read like reference usage
# ----------------------------------------------------------------------------------------------
class OracleHaloModel(object):
    """Mirrors HaloModel(zs,ks,ms,accuracy='low') of hmvec.py:75-572 on top of the functions above."""

    def __init__(self, zs, ks, ms, params=None, mdef="vir", skip_nfw=False, Pzk=None, sPzk=None,
                 mass_function_mode="sheth-torman", alpha_table=None):
        self.zs = np.asarray(zs, dtype=np.float64)
        self.ks = np.asarray(ks, dtype=np.float64)
        self.ms = np.asarray(ms, dtype=np.float64)
        self.bg = Background(params)
        self.p = self.bg.p
        self.mdef = mdef
        self.h = self.bg.h
        self.omm0 = self.bg.omm0
        self.rho_m0 = float(self.bg.rho_matter(0.0)[0])
        self.Pzk = plin_approx(self.bg, self.ks, self.zs) if Pzk is None else Pzk       # hmvec.py:98-99
        self.ks_sig = sigma2_grid(self.p)
        self.sPzk = plin_approx(self.bg, self.ks_sig, self.zs) if sPzk is None else sPzk  # cosmology.py:259-260
        R = R_from_M(self.ms, self.rho_m0, 1.0)                                          # hmvec.py:117-118
        self.sigma2 = sigma2(R, self.ks_sig, self.sPzk, self.p["Wkr_taylor_switch"])
        st = dict(A=self.p["st_A"], a=self.p["st_a"], p=self.p["st_p"], dc=self.p["st_deltac"])
        if mass_function_mode == "tinker":                                               # hmvec.py:142-145,157-159
            f = tinker_fsigma(self.sigma2, self.zs, alpha_table, dc=st["dc"])
            self.nzm = mass_function(self.sigma2, self.ms, self.rho_m0, fsigma=f)
            self.bh = tinker_bias(self.sigma2, dc=st["dc"])
        else:
            self.nzm = mass_function(self.sigma2, self.ms, self.rho_m0, **st)
            self.bh = st_bias(self.sigma2, a=st["a"], p=st["p"], dc=st["dc"])
        self.uk_profiles, self.pk_profiles, self.hods = {}, {}, {}
        if not skip_nfw:
            self.add_nfw_profile("nfw")

    # -- geometry -----------------------------------------------------------------------------
    def concentration(self):
        A, al, be = self.p["duffy_vir"] if self.mdef == "vir" else self.p["duffy_mean"]
        return duffy(self.ms, self.zs, self.h, A, al, be)                                # hmvec.py:163-174

    def rvirs(self):
        if self.mdef == "vir":                                                           # hmvec.py:111-115
            return R_from_M(self.ms[None, :], self.bg.rho_crit(self.zs)[:, None], self.bg.deltav(self.zs)[:, None])
        return R_from_M(self.ms[None, :], self.bg.rho_matter(self.zs)[:, None], 200.0)

    def _m200c(self):
        rhoc = self.bg.rho_crit(self.zs)
        d1 = rhoc * self.bg.deltav(self.zs) if self.mdef == "vir" else self.bg.rho_matter(self.zs) * 200.0
        m200 = mdelta_convert(self.ms, self.concentration(), d1, 200.0 * rhoc)           # hmvec.py:216-224
        r200 = R_from_M(m200, rhoc[:, None], 200.0)
        return m200, r200, rhoc

    # -- profiles -----------------------------------------------------------------------------
    def add_nfw_profile(self, name, numeric=False, nxs=40000, xmax=200.0):
        cs = self.concentration()
        rv = self.rvirs()
        if numeric:                                                                      # hmvec.py:343-345
            self.uk_profiles[name] = profile_transform(lambda x: 1.0 / x / (1.0 + x) ** 2, cs, rv / cs,
                                                       self.zs, self.ks, xmax, nxs)
        else:
            self.uk_profiles[name] = uk_nfw_analytic(self.ks, self.zs, cs, rv)
        return self.ks, self.uk_profiles[name]

    def add_battaglia_profile(self, name, family="AGN", nxs=5000, xmax=20.0, overrides=None):
        fam = {k: tuple(v) for k, v in GAS_FAMILIES[family].items()}
        gamma = self.p["gas_gamma"]
        for key, val in (overrides or {}).items():                                       # hmvec.py:204-213
            if key == "battaglia_gas_gamma":
                gamma = val
            else:
                q, which = key.rsplit("_", 1)
                if q in fam:
                    t = list(fam[q]); t[("A0", "alpham", "alphaz").index(which)] = val; fam[q] = tuple(t)
        m200, r200, rhoc = self._m200c()
        omb = self.p["ombh2"] / self.h ** 2
        rgs = r200 / 2.0                                                                 # hmvec.py:247-248
        f = lambda x: gas_density_x(x, m200[..., None], self.zs[:, None, None], omb, self.omm0,
                                    rhoc[:, None, None], gamma, fam)
        self.uk_profiles[name] = profile_transform(f, self.rvirs() / rgs, rgs, self.zs, self.ks, xmax, nxs)

    def add_battaglia_pres_profile(self, name, family="pres", nxs=5000, xmax=20.0):
        fam = PRES_FAMILIES[family]
        m200, r200, rhoc = self._m200c()
        omb = self.p["ombh2"] / self.h ** 2
        f = lambda x: gas_pressure_x(x, m200[..., None], r200[..., None], self.zs[:, None, None], omb, self.omm0,
                                     rhoc[:, None, None], self.p["pres_alpha"], self.p["pres_gamma"], fam,
                                     self.p["parsec"], self.p["mSun"])
        pk = profile_transform(f, self.rvirs() / r200, r200, self.zs, self.ks, xmax, nxs, mass_norm=False)
        sigT = sc.physical_constants["Thomson cross section"][0]
        me = sc.physical_constants["electron mass"][0] / self.p["mSun"]
        scale = 4.0 * np.pi * (sigT / (me * sc.c ** 2)) * (r200 ** 3 * ((1 + self.zs) ** 2 / self.bg.h_of_z(self.zs))[:, None])
        self.pk_profiles[name] = pk * scale[..., None]                                   # hmvec.py:313-316

    # -- HOD ----------------------------------------------------------------------------------
    def add_hod(self, name, mthresh=None, ngal=None, corr="max", satellite_profile_name="nfw",
                central_profile_name=None):
        hp = self.p
        iters = 0
        if ngal is not None:
            l10, iters = hod_solve_mthresh(ngal, self.zs, self.ms, self.nzm, hp)
            mthresh = 10.0 ** l10
        l10 = np.log10(np.asarray(mthresh, dtype=np.float64))
        Nc, Ns, NsNsm1, NcNs = hod_occupations(self.zs, self.ms, l10, hp, corr)
        ng = hod_ngal(self.nzm, self.ms, Nc, Ns)
        self.hods[name] = dict(Nc=Nc, Ns=Ns, NsNsm1=NsNsm1, NcNs=NcNs, ngal=ng,
                               bg=hod_bias(self.nzm, self.bh, self.ms, Nc, Ns, ng),
                               satellite_profile=satellite_profile_name, central_profile=central_profile_name,
                               log10mthresh=l10[:, None], iterations=iters)

    # -- spectra ------------------------------------------------------------------------------
    def _tracer(self, name):
        if name in self.hods:
            h = self.hods[name]
            uc = None if h["central_profile"] is None else self.uk_profiles[h["central_profile"]]
            return Tracer("hod", u=self.uk_profiles[h["satellite_profile"]], uc=uc, hod=h)
        if name in self.uk_profiles:
            return Tracer("matter", u=self.uk_profiles[name])
        if name in self.pk_profiles:
            return Tracer("pressure", u=self.pk_profiles[name])
        raise ValueError(name)

    def get_power_1halo(self, name="nfw", name2=None):
        name2 = name if name2 is None else name2
        return power_1halo(self._tracer(name), self._tracer(name2), self.nzm, self.ms, self.ks, self.rho_m0,
                           self.p["kstar_damping"])

    def get_power_2halo(self, name="nfw", name2=None, b1_in=None, b2_in=None):
        name2 = name if name2 is None else name2
        return power_2halo(self._tracer(name), self._tracer(name2), self.nzm, self.bh, self.ms, self.Pzk,
                           self.rho_m0, b1_in, b2_in)

    def get_power(self, name, name2=None):
        return self.get_power_1halo(name, name2) + self.get_power_2halo(name, name2)

    # -- Limber -------------------------------------------------------------------------------
    def C_kk(self, ells, zs, ks, Pmm, lzs1=None, ldndz1=None, lzs2=None, ldndz2=None):
        """cosmology.py:563-568"""
        w1 = lensing_window(self.bg, zs, lzs1, ldndz1)
        w2 = lensing_window(self.bg, zs, lzs2, ldndz2)
        return limber(ells, zs, ks, Pmm, zs, w1, w2, self.bg.h_of_z(zs), self.bg.chi(zs))

    def C_kg(self, ells, zs, ks, Pgm, gzs, gdndz=None, lzs=None, ldndz=None):
        """cosmology.py:536-547"""
        gzs = np.array(gzs, dtype=np.float64).reshape(-1)
        w1 = lensing_window(self.bg, gzs, lzs, ldndz)
        w2 = np.asarray(gdndz) / _trapz(gdndz, gzs) if gzs.size > 1 else np.ones(1)
        return limber(ells, zs, ks, Pgm, gzs, w1, w2, self.bg.h_of_z(gzs), self.bg.chi(gzs))

    def C_gg(self, ells, zs, ks, Pgg, gzs, gdndz=None, zmin=None, zmax=None):
        """cosmology.py:549-561: dn/dz branch, or one effective redshift with a top-hat [zmin, zmax]"""
        gzs = np.asarray(gzs, dtype=np.float64).reshape(-1)
        hz, chis = self.bg.h_of_z(gzs), self.bg.chi(gzs)
        if gzs.size > 1:
            w1 = w2 = np.asarray(gdndz) / _trapz(gdndz, gzs)
        else:
            dchi = self.bg.chi(np.atleast_1d(zmax))[0] - self.bg.chi(np.atleast_1d(zmin))[0]
            w1, w2 = np.ones(1), 1.0 / dchi / hz
        return limber(ells, zs, ks, Pgg, gzs, w1, w2, hz, chis)

    def C_ky(self, ells, zs, ks, Pym, lzs1=None, ldndz1=None):
        """cosmology.py:585-589"""
        w1 = lensing_window(self.bg, zs, lzs1, ldndz1)
        return limber(ells, zs, ks, Pym, zs, w1, np.ones(np.size(zs)), self.bg.h_of_z(zs), self.bg.chi(zs))

    def C_yy(self, ells, zs, ks, Ppp):
        """cosmology.py:591-597"""
        one = np.ones(np.size(zs))
        return limber(ells, zs, ks, Ppp, zs, one, one, self.bg.h_of_z(zs), self.bg.chi(zs))
