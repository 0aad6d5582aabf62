"""Module-level helpers of the reference's `hmvec/hmvec.py` (lines 627-957), re-exported by `hmvec_b200` so that
`from hmvec_b200 import *` offers every public name `from hmvec import *` does.

These are the small closed-form functions user scripts call directly (the tSZ notebook uses `mdelta_from_mdelta`,
`R_from_M`, `P_e_generic_x`; `ksz.py` uses the HOD moments).  On the hot path the same formulas run inside the CUDA
kernels (k_halo.cu, k_hod.cu, gnfw_eval.cuh); the functions here act on whatever arrays the caller passes, with the
reference's argument order, defaults and broadcasting.  `mdelta_from_mdelta` runs on the device (hmv_mdelta).
"""
import numpy as np
import scipy.constants as constants
from scipy.special import erf

from .params import default_params, battaglia_defaults

_gas = battaglia_defaults[default_params['battaglia_gas_family']]
_pres = battaglia_defaults['pres']


# ---- mass / radius ---------------------------------------------------------------------------------------
def R_from_M(M, rho, delta):
    """hmvec.py:627-628"""
    return (3. * M / 4. / np.pi / delta / rho) ** (1. / 3.)


def duffy_concentration(m, z, A=None, alpha=None, beta=None, h=None):
    """Duffy et al. 2008 c(M,z) with the reference's defaults (hmvec.py:68-73)."""
    A = default_params['duffy_A_mean'] if A is None else A
    alpha = default_params['duffy_alpha_mean'] if alpha is None else alpha
    beta = default_params['duffy_beta_mean'] if beta is None else beta
    h = default_params['H0'] / 100. if h is None else h
    return A * ((h * m / 2.e12) ** alpha) * (1 + z) ** beta


def Fcon(c):
    """ln(1+c) - c/(1+c)  (hmvec.py:737)"""
    return np.log(1. + c) - c / (1. + c)


def rhoscale_nfw(mdelta, rdelta, cdelta):
    """NFW amplitude m/(4 pi rs^3 F(c)) (hmvec.py:739-742; the reference multiplies by an undefined name `pref` and
    raises NameError -- the amplitude itself is returned here)."""
    rs = rdelta / cdelta
    return mdelta / (4. * np.pi * rs ** 3.) / Fcon(cdelta)


def rho_nfw_x(x, rhoscale):
    return rhoscale / x / (1. + x) ** 2.


def rho_nfw(r, rhoscale, rs):
    return rho_nfw_x(r / rs, rhoscale)


def mdelta_from_mdelta(M1, C1, delta_rhos1, delta_rhos2, vectorized=True):
    """M1(m) -> M2(z,m) between two overdensity definitions for NFW halos (hmvec.py:748-768), solved on the device
    (hmv_mdelta: the reference's secant iteration in ln M).  C1: [nz,nm]; delta_rhos1/2: [nz]."""
    import torch
    from . import _capi as capi
    if not torch.cuda.is_available():
        raise RuntimeError("hmvec_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback.")
    C1 = np.ascontiguousarray(C1, dtype=np.float64)
    nz, nm = C1.shape
    dev = lambda a, n: torch.as_tensor(np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1), (n,))),
                                       device="cuda")
    out = torch.empty((nz, nm), dtype=torch.float64, device="cuda")
    ms_d, cs_d, d1, d2 = dev(M1, nm), torch.as_tensor(C1, device="cuda"), dev(delta_rhos1, nz), dev(delta_rhos2, nz)
    capi.check(capi.lib.hmv_mdelta(nz, nm, capi.ptr(ms_d), capi.ptr(cs_d), capi.ptr(d1), capi.ptr(d2), capi.ptr(out),
                                   capi.stream()), "hmv_mdelta")
    return out.cpu().numpy()


def mdelta_from_mdelta_unvectorized(M1, C1, delta_rhos1, delta_rhos2):
    """Scalar / broadcast form of the conversion (hmvec.py:770-798): same equation M1 F(C1) = M2 F(C2(M2)) with
    C2 = C1 ((M2/M1) drho1/drho2)^(1/3), solved by secant iteration in ln M2 (the reference calls
    scipy.optimize.newton without a derivative, i.e. the secant method)."""
    M1, C1, r = np.broadcast_arrays(np.asarray(M1, dtype=np.float64), np.asarray(C1, dtype=np.float64),
                                    np.asarray(delta_rhos1, dtype=np.float64) / np.asarray(delta_rhos2, dtype=np.float64))
    lnM1 = np.log(M1)
    lhs = M1 / Fcon(C1)

    def resid(x):
        return lhs - np.exp(x) / Fcon(C1 * (np.exp(x - lnM1) * r) ** (1. / 3.))

    x0 = lnM1
    x1 = x0 * (1 + 1e-4) + 1e-4
    f0, f1 = resid(x0), resid(x1)
    for _ in range(60):
        with np.errstate(all="ignore"):
            step = np.where(f1 != f0, f1 * (x1 - x0) / (f1 - f0), 0.)
        x0, f0 = x1, f1
        x1 = x1 - step
        f1 = resid(x1)
        if np.all(np.abs(step) < 1e-12):
            break
    return np.exp(x1)


# ---- HOD (Leauthaud et al. 2012 SHMR; hmvec.py:634-731) ----------------------------------------------------
_SHMR = {  # (Mstar00, Mstara, M1, M1a, beta0, beta_a, gamma0, gamma_a, delta0, delta_a): z <= 0.8 / z > 0.8
    'lo': (10.72, 0.55, 12.35, 0.28, 0.44, 0.18, 1.56, 2.51, 0.57, 0.17),
    'hi': (11.09, 0.56, 12.27, -0.84, 0.65, 0.31, 1.12, -0.53, 0.56, -0.12),
}


def Mhalo_stellar_core(log10mstellar, a, Mstar00, Mstara, M1, M1a, beta0, beta_a, gamma0, gamma_a, delta0, delta_a):
    """log10 M_halo(M_star) of arXiv:1001.0015 eq. 2 with parameters linear in (a-1) (hmvec.py:648-656)."""
    da = a - 1
    x = log10mstellar - (Mstar00 + Mstara * da)
    return (-0.5 + (M1 + M1a * da) + (beta0 + beta_a * da) * x
            + 10 ** ((delta0 + delta_a * da) * x) / (1. + 10 ** (-(gamma0 + gamma_a * da) * x)))


def Mhalo_stellar(z, log10mstellar):
    """hmvec.py:658-695: the two parameter sets switch at z = 0.8.  z: [nz,1] (or [nz]); log10mstellar: [.., n]."""
    z = np.asarray(z, dtype=np.float64)
    ls = log10mstellar + z * 0
    a = 1. / (1 + z)
    out = np.zeros((z.size, np.shape(log10mstellar)[-1]))
    zf = z.reshape(-1)
    for key, sel in (('lo', np.where(zf <= 0.8)), ('hi', np.where(zf > 0.8))):
        out[sel] = Mhalo_stellar_core(ls[sel], a[sel], *_SHMR[key])
    return out


def Mstellar_halo(z, log10mhalo):
    """Inverse SHMR by table lookup: 4000 points of log10 M_star in [-18,18], linear interpolation per redshift
    (hmvec.py:634-646)."""
    grid = np.linspace(-18, 18, 4000)[None, :]
    mh = Mhalo_stellar(z, grid)
    out = np.zeros((z.shape[0], log10mhalo.shape[-1]))
    for i in range(z.size):
        out[i] = np.interp(log10mhalo[0], mh[i], grid[0])
    return out


def avg_Nc(log10mhalo, z, log10mstellar_thresh, sig_log_mstellar):
    """<N_c(m)> (hmvec.py:698-703)"""
    d = log10mstellar_thresh - Mstellar_halo(z, log10mhalo)
    return 0.5 * (1. - erf(d / (np.sqrt(2.) * sig_log_mstellar)))


def hod_default_mfunc(mthresh, Bamp, Bind):
    return (10. ** (12.)) * Bamp * 10 ** ((mthresh - 12) * Bind)


def avg_Ns(log10mhalo, z, log10mstellar_thresh, Nc=None, sig_log_mstellar=None, alphasat=None, Bsat=None, betasat=None,
           Bcut=None, betacut=None, Msat_override=None, Mcut_override=None):
    """<N_s(m)> (hmvec.py:708-716)"""
    mth = Mhalo_stellar(z, log10mstellar_thresh)
    Msat = Msat_override if Msat_override is not None else hod_default_mfunc(mth, Bsat, betasat)
    Mcut = Mcut_override if Mcut_override is not None else hod_default_mfunc(mth, Bcut, betacut)
    if Nc is None:
        Nc = avg_Nc(log10mhalo, z, log10mstellar_thresh, sig_log_mstellar=sig_log_mstellar)
    m = 10 ** log10mhalo
    return Nc * ((m / Msat) ** alphasat) * np.exp(-Mcut / m)


def avg_NsNsm1(Nc, Ns, corr="max"):
    """<N_s(N_s-1)> (hmvec.py:719-725)"""
    if corr == 'max':
        with np.errstate(all="ignore"):
            ret = Ns ** 2. / Nc
        ret[np.isclose(Nc, 0.)] = 0
        return ret
    elif corr == 'min':
        return Ns ** 2.


def avg_NcNs(Nc, Ns, corr="max"):
    """<N_c N_s> (hmvec.py:727-731)"""
    if corr == 'max':
        return Ns
    elif corr == 'min':
        return Ns * Nc


def ngal_from_mthresh(log10mthresh=None, zs=None, nzm=None, ms=None, sig_log_mstellar=None, Ncs=None, Nss=None,
                      alphasat=None, Bsat=None, betasat=None, Bcut=None, betacut=None, Msat_override=None,
                      Mcut_override=None):
    """n_gal(z) = int dM n(M,z) (N_c + N_s)  (hmvec.py:936-957)"""
    if (Ncs is None) and (Nss is None):
        thr = log10mthresh[:, None]
        lmh = np.log10(ms[None, :])
        Ncs = avg_Nc(lmh, zs[:, None], thr, sig_log_mstellar)
        Nss = avg_Ns(lmh, zs[:, None], thr, Ncs, sig_log_mstellar, alphasat, Bsat, betasat, Bcut, betacut,
                     Msat_override=Msat_override, Mcut_override=Mcut_override)
    else:
        assert log10mthresh is None
        assert zs is None
        assert sig_log_mstellar is None
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return trapz(nzm * (Ncs + Nss), ms, axis=-1)


# ---- Battaglia gas density / pressure (hmvec.py:800-927) -----------------------------------------------------
def battaglia_gas_fit(m200critz, z, A0x, alphamx, alphazx):
    return A0x * (m200critz / 1.e14) ** alphamx * (1. + z) ** alphazx


def rho_gas_generic_x(x, m200critz, z, omb, omm, rhocritz, gamma=default_params['battaglia_gas_gamma'],
                      rho0_A0=_gas['rho0_A0'], rho0_alpham=_gas['rho0_alpham'], rho0_alphaz=_gas['rho0_alphaz'],
                      alpha_A0=_gas['alpha_A0'], alpha_alpham=_gas['alpha_alpham'], alpha_alphaz=_gas['alpha_alphaz'],
                      beta_A0=_gas['beta_A0'], beta_alpham=_gas['beta_alpham'], beta_alphaz=_gas['beta_alphaz']):
    """GNFW electron density in x = 2r/R200c (hmvec.py:844-860; note the exponent -(beta+gamma)/alpha)."""
    rho0 = battaglia_gas_fit(m200critz, z, rho0_A0, rho0_alpham, rho0_alphaz)
    alpha = battaglia_gas_fit(m200critz, z, alpha_A0, alpha_alpham, alpha_alphaz)
    beta = battaglia_gas_fit(m200critz, z, beta_A0, beta_alpham, beta_alphaz)
    return (omb / omm) * rhocritz * rho0 * (x ** gamma) * (1. + x ** alpha) ** (-(beta + gamma) / alpha)


def rho_gas_generic(r, m200critz, z, omb, omm, rhocritz, gamma=default_params['battaglia_gas_gamma'], **fit):
    """hmvec.py:819-842: the density at physical radius r."""
    R200 = R_from_M(m200critz, rhocritz, delta=200)
    return rho_gas_generic_x(2 * r / R200, m200critz, z, omb, omm, rhocritz, gamma, **fit)


def rho_gas(r, m200critz, z, omb, omm, rhocritz, gamma=default_params['battaglia_gas_gamma'], profile="AGN"):
    """hmvec.py:804-817"""
    fam = battaglia_defaults[profile]
    return rho_gas_generic(r, m200critz, z, omb, omm, rhocritz, gamma=gamma,
                           **{k: fam[k] for k in fam if k.split('_')[0] in ('rho0', 'alpha', 'beta')})


def P_e_generic_x(x, m200critz, R200critz, z, omb, omm, rhocritz, alpha=default_params['battaglia_pres_alpha'],
                  gamma=default_params['battaglia_pres_gamma'], P0_A0=_pres['P0_A0'], P0_alpham=_pres['P0_alpham'],
                  P0_alphaz=_pres['P0_alphaz'], xc_A0=_pres['xc_A0'], xc_alpham=_pres['xc_alpham'],
                  xc_alphaz=_pres['xc_alphaz'], beta_A0=_pres['beta_A0'], beta_alpham=_pres['beta_alpham'],
                  beta_alphaz=_pres['beta_alphaz']):
    """GNFW electron pressure in x = r/R200c (hmvec.py:906-927)."""
    P0 = battaglia_gas_fit(m200critz, z, P0_A0, P0_alpham, P0_alphaz)
    xc = battaglia_gas_fit(m200critz, z, xc_A0, xc_alpham, xc_alphaz)
    beta = battaglia_gas_fit(m200critz, z, beta_A0, beta_alpham, beta_alphaz)
    XH = .76
    eFrac = 2.0 * (XH + 1.0) / (5.0 * XH + 3.0)
    G_newt = constants.G / (default_params['parsec'] * 1e6) ** 3 * default_params['mSun']
    t = x / xc
    return eFrac * (omb / omm) * 200 * m200critz * G_newt * rhocritz / (2 * R200critz) * P0 * t ** gamma * (1. + t ** alpha) ** (-beta)


def P_e_generic(r, m200critz, z, omb, omm, rhocritz, alpha=default_params['battaglia_pres_alpha'],
                gamma=default_params['battaglia_pres_gamma'], **fit):
    """hmvec.py:881-904"""
    R200 = R_from_M(m200critz, rhocritz, delta=200)
    return P_e_generic_x(r / R200, m200critz, R200, z, omb, omm, rhocritz, alpha, gamma, **fit)


def P_e(r, m200critz, z, omb, omm, rhocritz, alpha=default_params['battaglia_pres_alpha'],
        gamma=default_params['battaglia_pres_gamma'], profile="pres"):
    """hmvec.py:864-879"""
    return P_e_generic(r, m200critz, z, omb, omm, rhocritz, alpha=alpha, gamma=gamma, **battaglia_defaults[profile])


def a2z(a):
    return (1.0 / a) - 1.0


__all__ = [n for n in dir() if not n.startswith('_') and n not in ('np', 'constants', 'erf')]
