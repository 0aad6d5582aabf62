"""kSZ tomography forecasts on top of the halo model: mirror of the reference's `hmvec/ksz.py` core (lines 1-336,
420-433, 435-468, 876-934) -- the main in-tree consumer of P_gg, P_ge and hods['g']['bg'].

`class kSZ(HaloModel)` keeps the reference's constructor signature and attributes (kLs, krs, mu, kS, sPggs, sPges,
Pmms, fs, d2vs, adotf, kstars, chistars, vrec, sPggtot, sPge, bgs, ngals_mpc3) so that code written against it runs
on the device-backed `HaloModel`: the small-scale spectra come from the cached six-spectra pass, and the
short-wavelength integral of the reconstruction noise (`Nvv_core_integral`, ksz.py:299-336) runs on the device
(hmv_ksz_nvv_integral).  Everything else here is O(n_mu n_kL) host arithmetic on [nz]-sized inputs.

Two things differ from the reference, both because its code path needs packages that are not part of the hot path:
  * the reference defaults to engine='class' and takes f(z) from CLASS (`get_growth_rate_f` raises for engine='camb',
    cosmology.py:345-350).  Here the default engine is 'camb' and f = dlnD/dlna comes from the closed-form growing
    mode (cosmology.py:297-313), documented in `Cosmology.get_growth_rate_f`.
  * the linear P(k_L) of the large-scale spectra uses `P_lin_slow` (CAMB) in the reference; with accuracy='low' (no
    CAMB) this module uses the EH98 `P_lin_approx`, the same substitution HaloModel itself makes (hmvec.py:98-99).
The survey-level helpers of ksz.py:340-418, 471-875, 936-988 (template C_l, Ma & Fry, squeezed limit, astropy-based
survey SNR) are forecast scripts outside the path and are not mirrored.
"""
import warnings

import numpy as np
import torch

from . import _capi as capi
from .cosmology import Cosmology, _trapz
from .hmvec import HaloModel
from .params import default_params

defaults = {'min_mass': 1e6, 'max_mass': 1e16, 'num_mass': 1000}
constants = {
    'thompson_SI': 6.6524e-29,
    'meter_to_megaparsec': 3.241e-23,
    'G_SI': 6.674e-11,
    'mProton_SI': 1.673e-27,
    'H100_SI': 3.241e-18,
}


def Ngg(ngalMpc3):
    """Shot noise 1/n_gal (ksz.py:31-32)."""
    return 1. / ngalMpc3


def get_survey_volume(zmin, zmax, fsky):
    """Comoving volume of a shell in Gpc^3 (ksz.py:35-39)."""
    c = Cosmology(engine='camb', accuracy='low')
    chimin, chimax = c.comoving_radial_distance(zmin), c.comoving_radial_distance(zmax)
    return fsky * (4. / 3.) * np.pi * (chimax ** 3. - chimin ** 3.) / 1e9


def get_kmin(volume_gpc3):
    """pi / V^(1/3) (ksz.py:66-68)."""
    return np.pi / (volume_gpc3 * 1e9) ** (1. / 3.)


def chi(Yp, NHe):
    return (1 - Yp * (1 - NHe / 4.)) / (1 - Yp / 2.)


def ne0_shaw(ombh2, Yp, NHe=0, me=1.14, gasfrac=0.9):
    """Mean electron density today in 1/m^3, eq. 3 of arXiv:1109.0553 (ksz.py:75-84)."""
    mu_e = 1.14
    return chi(Yp, NHe) * gasfrac * ombh2 * 3. * (constants['H100_SI'] ** 2.) / constants['mProton_SI'] / 8. / np.pi \
        / constants['G_SI'] / mu_e


def ksz_radial_function(z, ombh2, Yp, gasfrac=0.9, xe=1, tau=0, params=None):
    """K(z) = T_CMB sigma_T n_e0 x_e e^-tau (1+z)^2, eq. 4 of arXiv:1810.13423 (ksz.py:86-96; `gasfrac` is accepted
    and, as in the reference, not forwarded to ne0_shaw)."""
    if params is None:
        params = default_params
    return params['T_CMB'] * constants['thompson_SI'] * ne0_shaw(ombh2, Yp) * (1. + z) ** 2. \
        / constants['meter_to_megaparsec'] * xe * np.exp(-tau)


def _sanitize(inp):
    inp[~np.isfinite(inp)] = 0
    return inp


def get_interpolated_cls(Cls, chistar, kss):
    """C_tot at l = floor(chi* k); zero below l = 2 (set in place, as the reference does), inf beyond the table
    (ksz.py:422-433)."""
    Cls[:2] = 0
    ell = np.asarray(chistar * np.asarray(kss, dtype=np.float64))
    out = np.full(ell.shape, np.inf)
    ok = ell <= Cls.size - 1
    out[ok] = Cls[ell[ok].astype(np.int64)]
    return out


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("hmvec_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def _ks_integral(kSs, Pge, Pgg_tot, clk, Pgg_photo_tot=None):
    """trapz over kS of kS Pge^2/(Pgg_tot C_l) on the device; arrays are [nk] or [..., nk] (leading axes = (mu,kL))."""
    dev = _device()
    kSs = np.ascontiguousarray(kSs, dtype=np.float64)
    nk = kSs.size
    lead = np.broadcast_shapes(np.shape(Pgg_tot)[:-1], np.shape(Pge)[:-1] if Pge is not None and np.ndim(Pge) else (),
                               np.shape(Pgg_photo_tot)[:-1] if Pgg_photo_tot is not None else ())
    nb = int(np.prod(lead)) if lead else 1

    def up(a):
        if a is None:
            return None, 0
        a = np.asarray(a, dtype=np.float64)
        if a.ndim <= 1:
            a = np.ascontiguousarray(np.broadcast_to(a, (nk,)))
            return torch.as_tensor(a, device=dev), 0
        a = np.ascontiguousarray(np.broadcast_to(a, lead + (nk,))).reshape(nb, nk)
        return torch.as_tensor(a, device=dev), nk

    ks_d, clk_d = torch.as_tensor(kSs, device=dev), torch.as_tensor(np.ascontiguousarray(clk, dtype=np.float64), device=dev)
    e_d, es = up(Pge)
    g_d, gs = up(Pgg_tot)
    p_d, ps = up(Pgg_photo_tot)
    out = torch.empty(nb, dtype=torch.float64, device=dev)
    capi.check(capi.lib.hmv_ksz_nvv_integral(nb, nk, capi.ptr(ks_d), capi.ptr(e_d), es, capi.ptr(g_d), gs, capi.ptr(p_d),
                                             ps, capi.ptr(clk_d), capi.ptr(out), capi.stream()), "hmv_ksz_nvv_integral")
    res = out.cpu().numpy()
    return res.reshape(lead) if lead else res[0]


def Nvv_core_integral(chi_star, Fstar, mu, kL, kSs, Cls, Pge, Pgg_tot, Pgg_photo_tot=None, errs=False,
                      robust_term=False, photo=True):
    """Velocity-reconstruction noise N_vv(mu, kL) (ksz.py:299-336): prefactor mu^-2 2 pi chi*^2/K*^2 over the kS
    integral of kS Pge^2/(Pgg_tot C_tot).  errs=True sets Pge = 1 and also returns the input Pge."""
    if robust_term:
        if photo:
            print("WARNING: photo_zs were True for an Nvv(robust_term=True) call. Overriding to False.")
        photo = False
        assert Pgg_photo_tot is not None
    ret_Pge = None
    if errs:
        ret_Pge = np.array(Pge, copy=True)
        Pge = None
    amu = np.resize(mu, (kL.size, mu.size)).T
    with np.errstate(all="ignore"):
        prefact = amu ** (-2.) * 2. * np.pi * chi_star ** 2. / Fstar ** 2.
    clk = get_interpolated_cls(Cls, chi_star, kSs)
    integral = _ks_integral(kSs, Pge, Pgg_tot, clk, Pgg_photo_tot if robust_term else None)
    Nvv = prefact / integral
    assert np.all(np.isfinite(Nvv))
    return (Nvv, ret_Pge) if errs else Nvv


def pge_err_core(pgv_int, kstar, chistar, volume_gpc3, kss, ks_bin_edges, pggtot, Cls):
    """Band-power errors on P_ge (ksz.py:43-63)."""
    volume = volume_gpc3 * 1e9
    cltot = get_interpolated_cls(Cls, chistar, kss)
    with np.errstate(all="ignore"):
        integrand = kss / (pggtot * cltot)
    ints = []
    for kleft, kright in zip(ks_bin_edges[:-1], ks_bin_edges[1:]):
        sel = np.logical_and(kss > kleft, kss <= kright)
        ints.append(_trapz(_sanitize(integrand[sel]), kss[sel]))
    return (volume * kstar ** 2 / 12 / np.pi ** 3 / chistar ** 2. * pgv_int * np.asarray(ints)) ** (-0.5)


class kSZ(HaloModel):
    def __init__(self, zs, volumes_gpc3, ngals_mpc3, kL_max=0.1, num_kL_bins=100, kS_min=0.1, kS_max=10.0,
                 num_kS_bins=101, num_mu_bins=102, ms=None, params=None, mass_function="sheth-torman", halofit=None,
                 mdef='vir', nfw_numeric=False, skip_nfw=False, electron_profile_name='e',
                 electron_profile_family='AGN', skip_electron_profile=False, electron_profile_param_override=None,
                 electron_profile_nxs=None, electron_profile_xmax=None, skip_hod=False, hod_name="g", hod_corr="max",
                 hod_param_override=None, mthreshs_override=None, verbose=False, b1=None, b2=None, sigz=None,
                 engine='camb', accuracy='medium', device=None):
        """ksz.py:102-235.  Extra keywords: accuracy (the reference always runs HaloModel's default 'medium', which
        needs CAMB; 'low' uses the EH98 power throughout) and device."""
        if ms is None:
            ms = np.geomspace(defaults['min_mass'], defaults['max_mass'], defaults['num_mass'])
        volumes_gpc3 = np.atleast_1d(volumes_gpc3)
        assert len(zs) == len(volumes_gpc3) == len(ngals_mpc3)
        ngals_mpc3 = np.asarray(ngals_mpc3)
        ks = np.geomspace(kS_min, kS_max, num_kS_bins)
        self.mu = np.linspace(-1., 1., num_mu_bins)
        HaloModel.__init__(self, zs, ks, ms=ms, params=params if params is not None else {}, mass_function=mass_function,
                           halofit=halofit, mdef=mdef, nfw_numeric=nfw_numeric, skip_nfw=skip_nfw, engine=engine,
                           accuracy=accuracy, device=device)
        self.kS = self.ks
        if not skip_electron_profile:
            self.add_battaglia_profile(name=electron_profile_name, family=electron_profile_family,
                                       param_override=electron_profile_param_override, nxs=electron_profile_nxs,
                                       xmax=electron_profile_xmax, ignore_existing=False)
        if not skip_hod:
            self.add_hod(hod_name, mthresh=mthreshs_override, ngal=ngals_mpc3, corr=hod_corr,
                         satellite_profile_name='nfw', central_profile_name=None, ignore_existing=False,
                         param_override=hod_param_override)
        self.Pmms, self.fs, self.adotf, self.d2vs = [], [], [], []
        self.sigma_z_func = lambda z: sigz * (1. + z)
        self.Hphotozs = self.h_of_z(self.zs)                          # 1/Mpc
        self.kLs = np.geomspace(get_kmin(np.max(volumes_gpc3)), kL_max, num_kL_bins)
        self.krs = self.mu.reshape((self.mu.size, 1)) * self.kLs.reshape((1, self.kLs.size))   # (mu, kL)
        self.sigz = sigz
        win = lambda zi, power: self.Wphoto(zi).reshape((self.mu.size, self.kLs.size, 1)) ** power
        if not skip_hod:
            self.sPggs = self.get_power(hod_name, name2=hod_name, verbose=verbose, b1=b1, b2=b1)
            self.sPges = self.get_power(hod_name, name2=electron_profile_name, verbose=verbose, b1=b1)
            if sigz is not None:
                oPggs, oPges = self.sPggs.copy(), self.sPges.copy()
                self.sPggs = np.asarray([oPggs[zi] * win(zi, 2.) for zi in range(oPggs.shape[0])])
                self.sPges = np.asarray([oPges[zi] * win(zi, 1.) for zi in range(oPges.shape[0])])
        if np.max(volumes_gpc3) != np.min(volumes_gpc3):
            warnings.warn('Using equal k_min at each z, despite different volumes at each z')
        p = self._plin_large_scale(self.kLs, self.zs)
        growth = self.get_growth_rate_f(self.zs)[None, ...]
        self.kstars, self.chistars, self.Vs = [], [], volumes_gpc3
        self.vrec, self.sPggtot, self.sPge, self.bgs = [], [], [], []
        # the reference looks the spectra up under the literal names 'g' and 'e' here (ksz.py:182-183)
        aPgg = self.get_power('g', 'g', verbose=verbose)
        aPge = self.get_power('g', 'e', verbose=verbose)
        bgs_all = self.hods['g']['bg']
        for zindex, volume_gpc3 in enumerate(volumes_gpc3):
            self.Pmms.append(np.resize(p[zindex].copy(), (self.mu.size, self.kLs.size)))
            self.fs.append(growth[:, zindex].copy())
            z = self.zs[zindex]
            a = 1. / (1. + z)
            H = self.h_of_z(z)
            self.kstars.append(self.ksz_radial_function(zindex))
            self.d2vs.append(self.fs[zindex] * a * H / self.kLs)
            self.adotf.append(self.fs[zindex] * a * H)
            self.chistars.append(self.comoving_radial_distance(z))
            bg = bgs_all[zindex]
            self.bgs.append(bg)
            ngg = Ngg(ngals_mpc3[zindex])
            flPgg = self.lPgg(zindex, bg1=bg, bg2=bg)[0, :] + ngg
            flPgv = self.lPgv(zindex, bg=bg)[0, :]
            kls = self.kLs
            with np.errstate(all="ignore"):
                integrand = _sanitize((kls ** 2.) * (flPgv * flPgv) / flPgg)
            self.vrec.append(_trapz(integrand, kls).copy())
            Pgg = aPgg[zindex].copy()
            if sigz is not None:
                Pgg = Pgg[None, None] * win(zindex, 2.)
            self.sPggtot.append((Pgg + ngg).copy())
            Pge = aPge[zindex].copy()
            if sigz is not None:
                Pge = Pge[None, None] * win(zindex, 1.)
            self.sPge.append(Pge.copy())
        self.ngals_mpc3 = ngals_mpc3

    def _plin_large_scale(self, kLs, zs):
        """Linear P(k_L, z) of the large-scale spectra: CAMB's (`P_lin_slow`, ksz.py:167) when available, the EH98
        power with accuracy='low'."""
        if self.accuracy == 'low':
            return self.P_lin_approx(kLs, zs)
        return self.P_lin_slow(kLs, zs)

    def Pge_err(self, zindex, ks_bin_edges, Cls):
        return pge_err_core(self.vrec[zindex], self.kstars[zindex], self.chistars[zindex], self.Vs[zindex], self._ks64,
                            ks_bin_edges, self.sPggtot[zindex][0], Cls)

    def lPvv(self, zindex, bv1=1, bv2=1):
        """Long-wavelength P_vv = (f a H/k_L)^2 P_mm b_v1 b_v2 on the (mu, kL) grid (ksz.py:246-257)."""
        return (self.d2vs[zindex]) ** 2. * self.Pmms[zindex] * bv1 * bv2

    def lPgg(self, zindex, bg1, bg2):
        Pgg = self.Pmms[zindex] * bg1 * bg2
        if self.sigz is not None:
            Pgg = Pgg[..., None] * (self.Wphoto(zindex).reshape((self.mu.size, self.kLs.size, 1)) ** 2.)
        return Pgg

    def lPgv(self, zindex, bg, bv=1):
        Pgv = self.Pmms[zindex] * bg * bv * (self.d2vs[zindex])
        if self.sigz is not None:
            Pgv = Pgv[..., None] * (self.Wphoto(zindex).reshape((self.mu.size, self.kLs.size, 1)))
        return Pgv

    def ksz_radial_function(self, zindex, gasfrac=0.9, xe=1, tau=0, params=None):
        return ksz_radial_function(self.zs[zindex], self.ombh2, self.YHe, gasfrac=gasfrac, xe=xe, tau=tau, params=params)

    def Wphoto(self, zindex):
        """Photo-z damping exp(-sigma_z^2 k_r^2 / 2H^2) on the (mu, kL) grid (ksz.py:283-287)."""
        H = self.Hphotozs[zindex]
        return np.exp(-self.sigma_z_func(self.zs[zindex]) ** 2. * self.krs ** 2. / 2. / H ** 2.)

    def Nvv(self, zindex, Cls):
        return Nvv_core_integral(self.chistars[zindex], self.ksz_radial_function(zindex), self.mu, self.kLs, self.kS,
                                 Cls, self.sPge[zindex], self.sPggtot[zindex], Pgg_photo_tot=None, errs=False,
                                 robust_term=False, photo=True)


def get_ksz_snr(volume_gpc3, z, ngal_mpc3, Cls, bg=None, params=None, kL_max=0.1, num_kL_bins=100, kS_min=0.1,
                kS_max=10.0, num_kS_bins=101, num_mu_bins=102, ms=None, mass_function="sheth-torman", mdef='vir',
                nfw_numeric=False, electron_profile_family='AGN', electron_profile_nxs=None, electron_profile_xmax=None,
                sigz=None, **kw):
    """SNR^2 = V int dmu dkL kL^2/(2 pi)^2 Pgv^2/(Pgg_tot Nvv)  (ksz.py:435-468).  Returns (snr, kSZ object)."""
    fksz = kSZ([z], [volume_gpc3], [ngal_mpc3], kL_max=kL_max, num_kL_bins=num_kL_bins, kS_min=kS_min, kS_max=kS_max,
               num_kS_bins=num_kS_bins, num_mu_bins=num_mu_bins, ms=ms, params=params, mass_function=mass_function,
               halofit=None, mdef=mdef, nfw_numeric=nfw_numeric, skip_nfw=False, electron_profile_name='e',
               electron_profile_family=electron_profile_family, skip_electron_profile=False,
               electron_profile_param_override=params, electron_profile_nxs=electron_profile_nxs,
               electron_profile_xmax=electron_profile_xmax, skip_hod=False, hod_name="g", hod_corr="max",
               hod_param_override=None, sigz=sigz, **kw)
    V = volume_gpc3 * 1e9
    ngg = Ngg(ngal_mpc3)
    Nvv_ = fksz.Nvv(0, Cls)
    if bg is None:
        bg = fksz.bgs[0]
    lPgg, lPgv = fksz.lPgg(zindex=0, bg1=bg, bg2=bg), fksz.lPgv(zindex=0, bg=bg)
    if sigz is not None:
        lPgg, lPgv = lPgg[..., 0], lPgv[..., 0]
    kls = fksz.kLs
    with np.errstate(all="ignore"):
        integrand = _sanitize((kls ** 2.) * (lPgv ** 2) / (lPgg + ngg) / Nvv_)
    snr2 = _trapz(_trapz(integrand, kls), fksz.mu) / (2. * np.pi) ** 2.
    return np.sqrt(V * snr2), fksz


def Nvv(z, vol_gpc3, ngals_mpc3, Cl_total, sigz=None, kL_max=0.1, num_kL_bins=100, kS_min=0.1, kS_max=10.0,
        num_kS_bins=101, num_mu_bins=102, **kw):
    """Convenience wrapper (ksz.py:876-934): returns (mus, kLs, N_vv[mu, kL]) for one redshift box."""
    hksz = kSZ([z], [vol_gpc3], [ngals_mpc3], kL_max=kL_max, num_kL_bins=num_kL_bins, kS_min=kS_min, kS_max=kS_max,
               num_kS_bins=num_kS_bins, num_mu_bins=num_mu_bins, sigz=sigz, **kw)
    return hksz.mu, hksz.kLs, hksz.Nvv(0, Cl_total)
