#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; never at test/bench time):

    python tests/golden/make_golden.py [case ...]

The reference (simonsobs/hmvec) is imported from /root/reference with `camb` replaced by the
test-side stand-in in tests/golden/camb_standin (flat-LCDM closed forms) and accuracy='low', which
routes both Pzk (hmvec.py:98-99) and the sigma^2 spectrum (cosmology.py:259-260) through the
reference's own EH98 P_lin_approx (cosmology.py:391-402).  With that the reference's own files
hmvec.py / fft.py / utils.py / params.py / cosmology.py execute the whole hot path.
`limber_integral` cannot run under SciPy>=1.14 (interp2d / dfitpack.bispeu were removed), so
`scipy.interpolate.interp2d` and `dfitpack.bispeu` are shimmed for the duration of this script
with RectBivariateSpline(kx=ky=1) -- SciPy's documented bug-for-bug replacement; the reference's
own limber_integral body (cosmology.py:867-904) is what executes.

Cases: readme, mini, mini_mean, largeslab (the path on four grids), kat (the reference's two known-answer ideas),
cky (C_ky: tSZ x lensing), pkspline (utils.get_matter_power_interpolator_generic on an EH98 table: the P(z,k)
ingestion in front of the path).
"""
import os
import sys
import time
import types
import warnings

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

warnings.filterwarnings("ignore")


def _import_reference():
    """The unmodified reference at /root/reference with the camb stand-in and the SciPy / data-path shims
    (oracle/ref_loader.py documents each one)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import ref_loader
    return ref_loader.load(REF)


def _meta():
    return dict(numpy_version=np.__version__, scipy_version=scipy.__version__,
                generated=time.strftime("%Y-%m-%d"), reference="simonsobs/hmvec @ /root/reference")


SPECTRA = [("mm", "nfw", "nfw"), ("ee", "electron", "electron"), ("me", "nfw", "electron"),
           ("gg", "g", "g"), ("gm", "g", "nfw"), ("ge", "g", "electron")]


def _common(h, out):
    out.update(zs=h.zs, ks=h.ks, ms=h.ms, Pzk=h.Pzk, sigma2=h.sigma2, nzm=h.nzm, bh=h.bh,
               hubble=h.hubble_parameter(h.zs), h_of_z=h.h_of_z(h.zs), chi=h.comoving_radial_distance(h.zs),
               rho_crit=h.rho_critical_z(h.zs), rho_m0=h.rho_matter_z(0.), deltav=h.deltav(h.zs),
               cs=h.concentration(), rvirs=h.rvir(h.ms[None, :], h.zs[:, None]))


def _hod(h, name, out, tag=None):
    tag = tag or name
    for k in ("Nc", "Ns", "NsNsm1", "NcNs", "ngal", "bg", "log10mthresh"):
        out["hod_%s_%s" % (tag, k)] = np.asarray(h.hods[name][k])


def case_readme(hm):
    """C1-C3 of SURVEY 8(d) on the README grid (README.rst:55-84)."""
    zs = np.linspace(0., 3., 20)
    ms = np.geomspace(2e10, 1e17, 200)
    ks = np.geomspace(1e-4, 100, 1001)
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low')
    out = {}
    _common(h, out)
    out["sPzk_sub"] = h.sPzk[:, ::10]
    h.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
    # cubes are 32 MB each: keep a strided sample (every 8th mass, every 4th k) + two full halos
    out["uk_nfw_sub"] = h.uk_profiles["nfw"][:, ::8, ::4]
    out["uk_e_sub"] = h.uk_profiles["electron"][:, ::8, ::4]
    out["uk_nfw_rows"] = h.uk_profiles["nfw"][[0, 7, 19]][:, [0, 100, 199]]
    out["uk_e_rows"] = h.uk_profiles["electron"][[0, 7, 19]][:, [0, 100, 199]]
    h.add_hod("g", mthresh=10 ** 10.5 + zs * 0.)
    h.add_hod("g2", ngal=np.geomspace(1e-3, 1e-5, zs.size))
    out["g2_ngal_target"] = np.geomspace(1e-3, 1e-5, zs.size)
    _hod(h, "g", out)
    _hod(h, "g2", out)
    for tag, a, b in SPECTRA + [("g2g2", "g2", "g2"), ("g2e", "g2", "electron")]:
        out["P1h_" + tag] = h.get_power_1halo(a, b)
        out["P2h_" + tag] = h.get_power_2halo(a, b)
    return out


MINI_ZS = np.array([0.01, 0.4, 0.8, 0.81, 1.6, 3.0])
MINI_MS = np.geomspace(1e11, 1e16, 64)
MINI_KS = np.geomspace(1e-3, 50, 257)


def case_mini(hm):
    """Full-resolution small grid exercising every branch of the path incl. pressure and Limber."""
    zs, ms, ks = MINI_ZS, MINI_MS, MINI_KS
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low')
    out = {}
    _common(h, out)
    out["sPzk_sub"] = h.sPzk[:, ::10]
    # mass conversion as used inside add_battaglia_profile (hmvec.py:216-225)
    rhoc = h.rho_critical_z(zs)
    m200 = hm.mdelta_from_mdelta(ms, h.concentration(), rhoc * h.deltav(zs), 200. * rhoc)
    out["m200c"] = m200
    h.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
    h.add_battaglia_profile("electron_sh", family="SH", xmax=10, nxs=2000)
    h.add_battaglia_profile("electron_ov", family="AGN", xmax=20, nxs=5000,
                            param_override={"battaglia_gas_gamma": -0.3, "rho0_A0": 3000., "alpha_alphaz": 0.25,
                                            "not_a_key": 1.0})
    h.add_battaglia_pres_profile("y", family="pres", xmax=20, nxs=5000)
    h.add_nfw_profile("nfwnum", numeric=True, nxs=8000, xmax=100)
    out["uk_nfw"] = h.uk_profiles["nfw"]
    out["uk_e"] = h.uk_profiles["electron"]
    out["uk_e_sh"] = h.uk_profiles["electron_sh"]
    out["uk_e_ov"] = h.uk_profiles["electron_ov"]
    out["uk_nfwnum"] = h.uk_profiles["nfwnum"]
    out["pk_y"] = h.pk_profiles["y"]
    h.add_hod("g", mthresh=10 ** 10.5 + zs * 0.)
    ngt = np.geomspace(2e-3, 3e-5, zs.size)
    h.add_hod("g2", ngal=ngt)
    h.add_hod("gmin", mthresh=10 ** (10.2 + 0.1 * zs), corr="min")
    h.add_hod("gcen", mthresh=10 ** 10.8 + zs * 0., central_profile_name="electron", satellite_profile_name="nfwnum")
    h.add_hod("gov", mthresh=10 ** 10.5 + zs * 0.,
              param_override={"hod_sig_log_mstellar": 0.3, "hod_alphasat": 1.1, "hod_Bsat": 8.0, "hod_betacut": 0.5})
    out["g2_ngal_target"] = ngt
    for n in ("g", "g2", "gmin", "gcen", "gov"):
        _hod(h, n, out)
    pairs = SPECTRA + [("g2g2", "g2", "g2"), ("g2e", "g2", "electron"), ("gming", "gmin", "gmin"),
                       ("gminm", "gmin", "nfw"), ("gcengcen", "gcen", "gcen"), ("gcene", "gcen", "electron"),
                       ("gg2", "g", "g2"), ("govgov", "gov", "gov"), ("yy", "y", "y"), ("ym", "y", "nfw"),
                       ("yg", "y", "g"), ("nn", "nfwnum", "nfwnum"), ("shsh", "electron_sh", "electron_sh"),
                       ("ovm", "electron_ov", "nfw"), ("eg", "electron", "g")]
    for tag, a, b in pairs:
        out["P1h_" + tag] = h.get_power_1halo(a, b)
        out["P2h_" + tag] = h.get_power_2halo(a, b)
    b1 = np.linspace(1.2, 2.0, zs.size)
    b2 = np.linspace(0.9, 1.4, zs.size)
    out["b1_in"], out["b2_in"] = b1, b2
    out["P2h_ge_bin"] = h.get_power_2halo("g", "electron", b1_in=b1, b2_in=b2)
    # Limber (cosmology.py:536-597, 867-904)
    ells = np.geomspace(10, 1e4, 40)
    out["ells"] = ells
    Pmm = out["P1h_mm"] + out["P2h_mm"]
    Pgm = out["P1h_gm"] + out["P2h_gm"]
    Pgg = out["P1h_gg"] + out["P2h_gg"]
    Pyy = out["P1h_yy"] + out["P2h_yy"]
    out["lens_window_25"] = h.lensing_window(zs, 2.5)
    out["C_kk"] = h.C_kk(ells, zs, ks, Pmm, lzs1=2.5, lzs2=2.5)
    out["C_kg"] = h.C_kg(ells, zs, ks, Pgm, gzs=0.8, lzs=2.5)
    out["C_yy"] = h.C_yy(ells, zs, ks, Pyy)
    lz = np.linspace(0.2, 2.8, 30)
    ldn = lz ** 2 * np.exp(-(lz / 0.9) ** 1.5)
    out["lz"], out["ldndz"] = lz, ldn
    out["lens_window_dndz"] = h.lensing_window(zs, lz, ldn)
    out["C_kk_dndz"] = h.C_kk(ells, zs, ks, Pmm, lzs1=lz, ldndz1=ldn, lzs2=1.1)
    gz = np.linspace(0.05, 2.5, 25)
    gdn = gz * np.exp(-gz / 0.5)
    out["gz"], out["gdndz"] = gz, gdn
    out["C_kg_dndz"] = h.C_kg(ells, zs, ks, Pgm, gzs=gz, gdndz=gdn, lzs=1100.)
    out["C_gg_dndz"] = h.C_gg(ells, zs, ks, Pgg, gzs=gz, gdndz=gdn)
    return out


def case_mini_mean(hm):
    """mdef='mean' branch (hmvec.py:111-115,166-169,219-220)."""
    zs, ms, ks = MINI_ZS, MINI_MS, MINI_KS
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low', mdef='mean',
                     params={"omch2": 0.125, "H0": 70.0, "ns": 0.97, "st_a": 0.75, "kstar_damping": 0.02})
    out = {}
    _common(h, out)
    h.add_battaglia_profile("electron", xmax=20, nxs=5000)
    out["uk_nfw"] = h.uk_profiles["nfw"]
    out["uk_e"] = h.uk_profiles["electron"]
    h.add_hod("g", mthresh=10 ** 10.5 + zs * 0.)
    _hod(h, "g", out)
    for tag, a, b in SPECTRA:
        out["P1h_" + tag] = h.get_power_1halo(a, b)
        out["P2h_" + tag] = h.get_power_2halo(a, b)
    return out


def case_largeslab(hm):
    """Two redshifts at the LARGE grid's M and k resolution (2000 M x 10000 k): only [nz,nk] outputs kept."""
    zs = np.array([0.01, 3.0])
    ms = np.geomspace(2e10, 1e17, 2000)
    ks = np.geomspace(1e-4, 100, 10000)
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low')
    out = dict(zs=zs, ms=ms, ks=ks, sigma2=h.sigma2, nzm=h.nzm, bh=h.bh)
    h.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
    h.add_hod("g", mthresh=10 ** 10.5 + zs * 0.)
    out["uk_nfw_rows"] = h.uk_profiles["nfw"][:, ::400]
    out["uk_e_rows"] = h.uk_profiles["electron"][:, ::400]
    for tag, a, b in SPECTRA:
        out["P1h_" + tag] = h.get_power_1halo(a, b)
        out["P2h_" + tag] = h.get_power_2halo(a, b)
    return out


def case_kat(hm):
    """Known-answer pieces the reference's own code names."""
    from hmvec import fft as rfft_mod, utils as rutils
    out = {}
    xs = np.linspace(0., 30., 6001)[1:]
    kt, U = rfft_mod.fft_integral(xs, np.exp(-xs ** 2 / 2.))
    out["gauss_xs"], out["gauss_kt"], out["gauss_U"] = xs, kt, U
    out["gauss_analytic"] = rfft_mod.analytic_fft_integral(kt)
    x = np.array([2., 4., 6.])
    out["bisect_x"] = x
    out["bisect_y"] = rutils.vectorized_bisection_search(x, lambda y: np.sqrt(y), (1, 40), 'increasing',
                                                         rtol=1e-4, verbose=False)
    return out


def case_cky(hm):
    """The one Limber wrapper the mini fixture does not hold: C_ky = tSZ x CMB lensing (cosmology.py:585-589) from the
    pressure x matter spectrum, delta-function and dn/dz lensing sources."""
    zs, ms, ks = MINI_ZS, MINI_MS, MINI_KS
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low')
    h.add_battaglia_pres_profile("y", family="pres", xmax=20, nxs=5000)
    out = dict(zs=zs, ms=ms, ks=ks)
    out["P1h_ym"], out["P2h_ym"] = h.get_power_1halo("y", "nfw"), h.get_power_2halo("y", "nfw")
    ells = np.geomspace(10, 1e4, 40)
    Pym = out["P1h_ym"] + out["P2h_ym"]
    out["ells"] = ells
    out["C_ky"] = h.C_ky(ells, zs, ks, Pym, lzs1=2.5)
    lz = np.linspace(0.2, 2.8, 30)
    ldn = lz ** 2 * np.exp(-(lz / 0.9) ** 1.5)
    out["lz"], out["ldndz"] = lz, ldn
    out["C_ky_dndz"] = h.C_ky(ells, zs, ks, Pym, lzs1=lz, ldndz1=ldn)
    # C_gg with a single effective redshift and a top-hat [zmin, zmax] (cosmology.py:556-560)
    h.add_hod("g", mthresh=10 ** 10.5 + zs * 0.)
    out["Pgg"] = h.get_power("g", "g")
    out["C_gg_tophat"] = h.C_gg(ells, zs, ks, out["Pgg"], gzs=np.array([0.8]), zmin=0.7, zmax=0.9)
    return out


def case_pkspline(hm):
    """P(z,k) through the reference's own interpolator builder (utils.py:53-182) on a synthetic CLASS-like table:
    the reference's EH98 P_lin_approx on a coarse (z,k) grid.  Variants: bicubic log-interpolation, the power-law
    extension beyond kmax, a sign-changing table (linear interpolation of P itself) and a 3-redshift table (kx=2)."""
    from hmvec import utils as rutils
    zs_t = np.linspace(0., 4., 25)
    ks_t = np.geomspace(5e-5, 30., 140)
    h = hm.HaloModel(np.array([0.5]), np.geomspace(1e-3, 1., 8), ms=np.geomspace(1e11, 1e15, 8), accuracy='low')
    pk = h.P_lin_approx(ks_t, zs_t)
    zq = np.linspace(0., 4., 57)
    kq = np.geomspace(5e-5, 30., 411)
    out = dict(zs_tab=zs_t, ks_tab=ks_t, pk_tab=pk, zq=zq, kq=kq)
    PK = rutils.get_matter_power_interpolator_generic(ks_t, zs_t, pk, silent=True)
    out["P_log"] = PK.P(zq, kq, grid=True)
    out["P_log_scalar_z"] = PK.P(1.2345, kq)
    kq_x = np.geomspace(1e-4, 200., 300)
    PKx = rutils.get_matter_power_interpolator_generic(ks_t, zs_t, pk, extrap_kmax=200., silent=True)
    out["kq_x"], out["P_extrap"] = kq_x, PKx.P(zq, kq_x, grid=True)
    pk_sc = pk * np.cos(3.0 * np.log(ks_t))[None, :]                 # crosses zero: log_interp is dropped
    PKs = rutils.get_matter_power_interpolator_generic(ks_t, zs_t, pk_sc, silent=True)
    out["pk_tab_sc"], out["P_signchange"] = pk_sc, PKs.P(zq, kq, grid=True)
    PKn = rutils.get_matter_power_interpolator_generic(ks_t, zs_t, -pk, silent=True)
    out["P_negative"] = PKn.P(zq, kq, grid=True)
    z3 = np.array([0., 1., 2.5])
    PK3 = rutils.get_matter_power_interpolator_generic(ks_t, z3, h.P_lin_approx(ks_t, z3), silent=True)
    zq3 = np.linspace(0., 2.5, 11)
    out["zs_tab3"], out["zq3"], out["P_kx2"] = z3, zq3, PK3.P(zq3, kq, grid=True)
    return out


def case_tinker(hm):
    """mass_function='tinker' with mdef='mean' (the tSZ notebook's configuration, examples/tSZ example.ipynb cell 4;
    hmvec.py:142-145,157-159 -> tinker.py:26-67), plus the free functions of tinker.py on fixed arguments."""
    from hmvec import tinker as rt
    zs, ms, ks = MINI_ZS, MINI_MS, MINI_KS
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low', mass_function='tinker', mdef='mean')
    out = {}
    _common(h, out)
    h.add_battaglia_pres_profile("y", family="pres", xmax=20, nxs=5000)
    h.add_hod("g", mthresh=10 ** 10.5 + zs * 0.)
    _hod(h, "g", out)
    for tag, a, b in (("mm", "nfw", "nfw"), ("gg", "g", "g"), ("gm", "g", "nfw"), ("yy", "y", "y"), ("ym", "y", "nfw")):
        out["P1h_" + tag] = h.get_power_1halo(a, b)
        out["P2h_" + tag] = h.get_power_2halo(a, b)
    nu = np.geomspace(0.2, 6., 23)[None, :] * np.ones((zs.size, 1))
    out["kat_nu"] = nu
    out["kat_bias"] = rt.bias(nu)
    out["kat_f_nu"] = rt.f_nu(nu, zs[:, None])
    out["kat_f_nu_nonorm"] = rt.f_nu(nu, zs[:, None], norm_consistency=False)
    out["kat_simple_f_nu"] = rt.simple_f_nu(nu)
    out["kat_NlnMsub"] = rt.NlnMsub(np.geomspace(1e9, 1e13, 7), np.geomspace(1e12, 1e15, 5))
    return out


def case_hostfuncs(hm):
    """The free functions of hmvec.py:627-957 on fixed arguments (what `from hmvec import *` hands to user scripts),
    and kappa_2h_profiles (hmvec.py:598-622) for a single lens redshift, the case its broadcasting supports."""
    out = {}
    z = np.array([0.1, 0.8, 0.81, 2.5])
    lmh = np.log10(np.geomspace(1e10, 1e16, 31))[None, :]
    thr = np.array([10.2, 10.5, 10.5, 11.0])[:, None]
    out["z"], out["log10mhalo"], out["thresh"] = z, lmh, thr
    out["Mhalo_stellar"] = hm.Mhalo_stellar(z[:, None], np.linspace(8., 12., 9)[None, :])
    out["Mstellar_halo"] = hm.Mstellar_halo(z[:, None], lmh)
    Nc = hm.avg_Nc(lmh, z[:, None], thr, 0.2)
    Ns = hm.avg_Ns(lmh, z[:, None], thr, Nc, 0.2, 1.0, 9.04, 0.74, 1.65, 0.59)
    out["avg_Nc"], out["avg_Ns"] = Nc, Ns
    out["avg_NsNsm1_max"], out["avg_NsNsm1_min"] = hm.avg_NsNsm1(Nc, Ns, "max"), hm.avg_NsNsm1(Nc, Ns, "min")
    out["avg_NcNs_max"], out["avg_NcNs_min"] = hm.avg_NcNs(Nc, Ns, "max"), hm.avg_NcNs(Nc, Ns, "min")
    out["hod_default_mfunc"] = hm.hod_default_mfunc(np.array([11.5, 12.0, 13.0]), 9.04, 0.74)
    ms = np.geomspace(1e11, 1e15, 9)
    C1 = hm.duffy_concentration(ms[None, :], z[:, None], 7.85, -0.081, -0.71, 0.673)
    d1 = np.array([1.2e13, 3e13, 3.1e13, 2e14]); d2 = np.array([2.6e13, 6e13, 6.1e13, 4e14])
    out["ms"], out["C1"], out["d1"], out["d2"] = ms, C1, d1, d2
    out["duffy_default"] = hm.duffy_concentration(ms, 0.5)
    out["mdelta"] = hm.mdelta_from_mdelta(ms, C1, d1, d2)
    out["mdelta_unvec"] = hm.mdelta_from_mdelta_unvectorized(ms[3], C1[1, 3], d1[1], d2[1])
    out["R_from_M"] = hm.R_from_M(ms, 3e10, 200.)
    out["Fcon"] = hm.Fcon(np.array([0.5, 4., 11.]))
    out["rho_nfw"] = hm.rho_nfw(np.array([0.1, 1., 3.]), 2e15, 0.4)
    x = np.geomspace(1e-2, 20., 17)
    m200 = np.array([3e12, 1e14, 2e15])[:, None]
    out["x"], out["m200"] = x, m200
    out["battaglia_gas_fit"] = hm.battaglia_gas_fit(m200, 0.7, 4000., 0.29, -0.66)
    out["rho_gas_generic_x"] = hm.rho_gas_generic_x(x[None, :], m200, 0.7, 0.049, 0.31, 1.5e11)
    out["rho_gas_SH"] = hm.rho_gas(x[None, :], m200, 0.7, 0.049, 0.31, 1.5e11, profile="SH")
    out["P_e_generic_x"] = hm.P_e_generic_x(x[None, :], m200, 1.3, 0.7, 0.049, 0.31, 1.5e11)
    out["P_e"] = hm.P_e(x[None, :], m200, 0.7, 0.049, 0.31, 1.5e11)
    out["a2z"] = hm.a2z(np.array([0.2, 0.5, 1.0]))
    # kappa_2h_profiles: one lens redshift
    zl = np.array([0.45])
    ks = np.geomspace(1e-3, 50, 400)
    h = hm.HaloModel(zl, ks, ms=MINI_MS, accuracy='low', skip_nfw=True)
    thetas = np.geomspace(1e-4, 1e-2, 9)
    out["k2h_zl"], out["k2h_ks"], out["k2h_thetas"] = zl, ks, thetas
    out["k2h"] = h.kappa_2h_profiles(thetas, 3e14, 1.1, verbose=False)
    return out


def case_ksz(hm):
    """kSZ consumer: the free functions of ksz.py on fixed arguments, and the attributes + Nvv of the
    reference's own `kSZ` class.  The class needs three things this image lacks, all shimmed test-side:
    accuracy='medium' power (HaloModel's default; the camb stand-in's closed-form PK, see camb_standin), engine='camb'
    (no classy), and a growth rate for that engine -- `Cosmology.get_growth_rate_f` raises for 'camb'
    (cosmology.py:345-350) and is replaced by dlnD/dlna of the reference's own D_growth_approx (cosmology.py:297-313)."""
    from hmvec import ksz as rk
    import hmvec.cosmology as hcosm

    def growth_rate_f(self, zs):
        a = 1. / (1. + np.atleast_1d(np.asarray(zs, dtype=np.float64)))
        e = 1e-4
        return (np.log(self.D_growth_approx(a * np.exp(e))) - np.log(self.D_growth_approx(a * np.exp(-e)))) / (2. * e)

    hcosm.Cosmology.get_growth_rate_f = growth_rate_f
    out = {}
    # free functions
    out["ne0_shaw"] = rk.ne0_shaw(0.02225, 0.24)
    zz = np.array([0.3, 1.0, 2.0])
    out["kat_z"], out["ksz_radial_function"] = zz, rk.ksz_radial_function(zz, 0.02225, 0.24)
    out["get_kmin"] = rk.get_kmin(np.array([1.0, 50.0]))
    out["chi"] = rk.chi(0.24, 1)
    Cls = 1e-6 / (1.0 + np.arange(6000.) / 300.) ** 2 + 3e-8
    out["Cls"] = Cls.copy()
    kS = np.geomspace(0.1, 10., 101)
    out["kS"] = kS
    out["interp_cls"] = rk.get_interpolated_cls(Cls.copy(), 1800., kS)
    mu, kL = np.linspace(-1., 1., 12), np.geomspace(2e-3, 0.1, 9)
    Pge1, Pgg1 = 3e2 * kS ** -1.5, 4e3 * kS ** -1.2 + 1e4
    out["mu"], out["kL"], out["Pge1"], out["Pgg1"] = mu, kL, Pge1, Pgg1
    out["Nvv_1d"] = rk.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), Pge1, Pgg1)
    W = np.exp(-(mu[:, None] * kL[None, :] * 40.) ** 2)[..., None]
    Pge3, Pgg3 = Pge1[None, None] * W, Pgg1[None, None] * W ** 2 + 50.
    out["Nvv_3d"] = rk.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), Pge3, Pgg3)
    nv, pe = rk.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), Pge3, Pgg3, errs=True)
    out["Nvv_errs"] = nv
    out["Nvv_robust"] = rk.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), Pge3, Pgg3,
                                             Pgg_photo_tot=Pgg3 * 1.3, robust_term=True, photo=False)
    edges = np.geomspace(0.1, 10., 6)
    out["ks_bin_edges"] = edges
    out["pge_err_core"] = rk.pge_err_core(3.5, 2.2e-7, 1800., 4.0, kS, edges, Pgg1, Cls.copy())
    # the class, two redshift boxes, with and without photo-z scatter
    zs = [0.4, 1.1]
    vols, ngals = [5., 12.], [3e-4, 1e-4]
    ms = np.geomspace(1e10, 1e16, 96)
    out["c_zs"], out["c_vols"], out["c_ngals"], out["c_ms"] = np.array(zs), np.array(vols), np.array(ngals), ms
    for tag, sigz in (("", None), ("_pz", 0.03)):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            k = rk.kSZ(np.array(zs), vols, ngals, num_kL_bins=24, num_kS_bins=41, num_mu_bins=14, ms=ms, sigz=sigz,
                       engine='camb', electron_profile_nxs=5000, electron_profile_xmax=20)
        for a in ("kLs", "krs", "mu", "kS", "Hphotozs", "sPggs", "sPges", "Pmms", "fs", "adotf", "d2vs", "kstars",
                  "chistars", "vrec", "sPggtot", "sPge", "bgs"):
            out["c%s_%s" % (tag, a)] = np.asarray(getattr(k, a))
        out["c%s_Nvv0" % tag] = k.Nvv(0, Cls.copy())
        out["c%s_Nvv1" % tag] = k.Nvv(1, Cls.copy())
        out["c%s_lPgg" % tag] = k.lPgg(0, 1.5, 1.7)
        out["c%s_lPgv" % tag] = k.lPgv(1, 1.4)
        out["c%s_lPvv" % tag] = k.lPvv(1)
        if sigz is None:       # with the photo-z window the reference's Pge_err indexes a (kL, kS) slice and fails
            out["c%s_Pge_err" % tag] = k.Pge_err(0, edges, Cls.copy())
        out["c%s_sigma2" % tag], out["c%s_Pzk" % tag] = k.sigma2, k.Pzk
    return out


CASES = dict(ksz=case_ksz, tinker=case_tinker, hostfuncs=case_hostfuncs, cky=case_cky, pkspline=case_pkspline, readme=case_readme, mini=case_mini, mini_mean=case_mini_mean, largeslab=case_largeslab, kat=case_kat)

if __name__ == "__main__":
    hm = _import_reference()
    which = sys.argv[1:] or list(CASES)
    for name in which:
        t = time.time()
        out = CASES[name](hm)
        out = {k: np.asarray(v) for k, v in out.items()}
        for k, v in _meta().items():
            out["_meta_" + k] = np.asarray(v)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-10s %6.1fs  %7.2f MB  %d arrays" % (name, time.time() - t, os.path.getsize(path) / 1e6, len(out)))
