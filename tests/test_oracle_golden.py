"""Pin the CPU oracle (oracle/hmvec_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU-only; runs in the driver's `-m "not gpu"` pass."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import hmvec_oracle as orc

# Intermediates that cross zero (u(k) of a truncated profile oscillates) get an absolute floor of
# 1e-12*max|array|; every [nz,nk] spectrum and every [nz,nm] weight is compared with pure rtol=1e-6.
OSC = 1e-12


def _build(g, mdef="vir", params=None):
    return orc.OracleHaloModel(g["zs"], g["ks"], g["ms"], params=params, mdef=mdef)


@pytest.fixture(scope="module")
def mini(golden_mini):
    g = golden_mini
    o = _build(g)
    o.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
    o.add_hod("g", mthresh=10 ** 10.5 + g["zs"] * 0.0)
    return o


def test_background_and_plin(golden_mini, mini):
    g, o = golden_mini, mini
    assert_close(o.bg.hubble(g["zs"]), g["hubble"], 1e-12)
    assert_close(o.bg.chi(g["zs"]), g["chi"], 1e-12)
    assert_close(o.bg.rho_crit(g["zs"]), g["rho_crit"], 1e-12)
    assert_close(o.bg.deltav(g["zs"]), g["deltav"], 1e-12)
    assert_close(o.Pzk, g["Pzk"], 1e-10, name="Pzk")
    assert_close(o.sPzk[:, ::10], g["sPzk_sub"], 1e-10, name="sPzk")


def test_mass_function(golden_mini, mini):
    g, o = golden_mini, mini
    assert_close(o.sigma2, g["sigma2"], 1e-10, name="sigma2")
    assert_close(o.nzm, g["nzm"], 1e-9, name="nzm")
    assert_close(o.bh, g["bh"], 1e-10, name="bh")
    assert_close(o.concentration(), g["cs"], 1e-12)
    assert_close(o.rvirs(), g["rvirs"], 1e-12)
    assert_close(o._m200c()[0], g["m200c"], 1e-10, name="m200c")


def test_profiles(golden_mini, mini):
    g, o = golden_mini, mini
    assert_close(o.uk_profiles["nfw"], g["uk_nfw"], 1e-6, OSC, "uk_nfw")
    assert_close(o.uk_profiles["electron"], g["uk_e"], 1e-6, OSC, "uk_e")
    o.add_battaglia_profile("esh", family="SH", xmax=10, nxs=2000)
    assert_close(o.uk_profiles["esh"], g["uk_e_sh"], 1e-6, OSC, "uk_e_sh")
    o.add_battaglia_profile("eov", family="AGN", xmax=20, nxs=5000,
                            overrides={"battaglia_gas_gamma": -0.3, "rho0_A0": 3000., "alpha_alphaz": 0.25,
                                       "not_a_key": 1.0})
    assert_close(o.uk_profiles["eov"], g["uk_e_ov"], 1e-6, OSC, "uk_e_ov")
    o.add_nfw_profile("nfwnum", numeric=True, nxs=8000, xmax=100)
    assert_close(o.uk_profiles["nfwnum"], g["uk_nfwnum"], 1e-6, OSC, "uk_nfwnum")
    o.add_battaglia_pres_profile("y")
    assert_close(o.pk_profiles["y"], g["pk_y"], 1e-6, OSC, "pk_y")


def test_hod_and_spectra(golden_mini, mini):
    g, o = golden_mini, mini
    zs = g["zs"]
    o.add_hod("g2", ngal=g["g2_ngal_target"])
    o.add_hod("gmin", mthresh=10 ** (10.2 + 0.1 * zs), corr="min")
    for n in ("g", "g2", "gmin"):
        for k in ("Nc", "Ns", "NsNsm1", "NcNs", "ngal", "bg", "log10mthresh"):
            assert_close(o.hods[n][k], g["hod_%s_%s" % (n, k)], 1e-8, name="%s.%s" % (n, k))
    pairs = [("mm", "nfw", "nfw"), ("ee", "electron", "electron"), ("me", "nfw", "electron"), ("gg", "g", "g"),
             ("gm", "g", "nfw"), ("ge", "g", "electron"), ("g2g2", "g2", "g2"), ("g2e", "g2", "electron"),
             ("gming", "gmin", "gmin"), ("gminm", "gmin", "nfw"), ("gg2", "g", "g2"), ("eg", "electron", "g")]
    for tag, a, b in pairs:
        assert_close(o.get_power_1halo(a, b), g["P1h_" + tag], 1e-6, name="P1h_" + tag)
        assert_close(o.get_power_2halo(a, b), g["P2h_" + tag], 1e-6, name="P2h_" + tag)
    assert_close(o.get_power_2halo("g", "electron", b1_in=g["b1_in"], b2_in=g["b2_in"]), g["P2h_ge_bin"], 1e-6)


def test_central_profile_override_and_pressure_spectra(golden_mini, mini):
    g, o = golden_mini, mini
    zs = g["zs"]
    if "nfwnum" not in o.uk_profiles:
        o.add_nfw_profile("nfwnum", numeric=True, nxs=8000, xmax=100)
    if "y" not in o.pk_profiles:
        o.add_battaglia_pres_profile("y")
    o.add_hod("gcen", mthresh=10 ** 10.8 + zs * 0., central_profile_name="electron", satellite_profile_name="nfwnum")
    hp = dict(o.p)
    hp.update(hod_sig_log_mstellar=0.3, hod_alphasat=1.1, hod_Bsat=8.0, hod_betacut=0.5)
    keep = o.p
    o.p = hp
    o.add_hod("gov", mthresh=10 ** 10.5 + zs * 0.)
    o.p = keep
    for n in ("gcen", "gov"):
        for k in ("Nc", "Ns", "NsNsm1", "ngal", "bg"):
            assert_close(o.hods[n][k], g["hod_%s_%s" % (n, k)], 1e-8, name="%s.%s" % (n, k))
    for tag, a, b in [("gcengcen", "gcen", "gcen"), ("gcene", "gcen", "electron"), ("govgov", "gov", "gov"),
                      ("yy", "y", "y"), ("ym", "y", "nfw"), ("yg", "y", "g"), ("nn", "nfwnum", "nfwnum")]:
        assert_close(o.get_power_1halo(a, b), g["P1h_" + tag], 1e-6, name="P1h_" + tag)
        assert_close(o.get_power_2halo(a, b), g["P2h_" + tag], 1e-6, name="P2h_" + tag)


def test_limber(golden_mini, mini):
    g, o = golden_mini, mini
    zs, ks, ells = g["zs"], g["ks"], g["ells"]
    Pmm = g["P1h_mm"] + g["P2h_mm"]
    Pgm = g["P1h_gm"] + g["P2h_gm"]
    Pgg = g["P1h_gg"] + g["P2h_gg"]
    Pyy = g["P1h_yy"] + g["P2h_yy"]
    assert_close(orc.lensing_window(o.bg, zs, 2.5), g["lens_window_25"], 1e-10)
    assert_close(orc.lensing_window(o.bg, zs, g["lz"], g["ldndz"]), g["lens_window_dndz"], 1e-10)
    assert_close(o.C_kk(ells, zs, ks, Pmm, lzs1=2.5, lzs2=2.5), g["C_kk"], 1e-9)
    assert_close(o.C_kg(ells, zs, ks, Pgm, gzs=0.8, lzs=2.5), g["C_kg"], 1e-9)
    assert_close(o.C_yy(ells, zs, ks, Pyy), g["C_yy"], 1e-9)
    assert_close(o.C_kk(ells, zs, ks, Pmm, lzs1=g["lz"], ldndz1=g["ldndz"], lzs2=1.1), g["C_kk_dndz"], 1e-9)
    assert_close(o.C_kg(ells, zs, ks, Pgm, gzs=g["gz"], gdndz=g["gdndz"], lzs=1100.), g["C_kg_dndz"], 1e-9)
    assert_close(o.C_gg(ells, zs, ks, Pgg, g["gz"], g["gdndz"]), g["C_gg_dndz"], 1e-9)
    # hand-written clamped bilinear == FITPACK kx=ky=1 (the restated piece of limber_integral)
    chis, hz = o.bg.chi(zs), o.bg.h_of_z(zs)
    w = g["lens_window_25"]
    a = orc.limber(ells, zs, ks, Pmm, zs, w, w, hz, chis, use_fitpack=True)
    b = orc.limber(ells, zs, ks, Pmm, zs, w, w, hz, chis, use_fitpack=False)
    assert_close(a, b, 1e-12)


def test_mean_mdef(golden_mini_mean):
    g = golden_mini_mean
    o = _build(g, mdef="mean", params={"omch2": 0.125, "H0": 70.0, "ns": 0.97, "st_a": 0.75, "kstar_damping": 0.02})
    o.add_battaglia_profile("electron", xmax=20, nxs=5000)
    o.add_hod("g", mthresh=10 ** 10.5 + g["zs"] * 0.0)
    assert_close(o.sigma2, g["sigma2"], 1e-10)
    assert_close(o.nzm, g["nzm"], 1e-9)
    assert_close(o.concentration(), g["cs"], 1e-12)
    assert_close(o.rvirs(), g["rvirs"], 1e-12)
    assert_close(o.uk_profiles["nfw"], g["uk_nfw"], 1e-6, OSC)
    assert_close(o.uk_profiles["electron"], g["uk_e"], 1e-6, OSC)
    for tag, a, b in [("mm", "nfw", "nfw"), ("ee", "electron", "electron"), ("me", "nfw", "electron"),
                      ("gg", "g", "g"), ("gm", "g", "nfw"), ("ge", "g", "electron")]:
        assert_close(o.get_power_1halo(a, b), g["P1h_" + tag], 1e-6, name="P1h_" + tag)
        assert_close(o.get_power_2halo(a, b), g["P2h_" + tag], 1e-6, name="P2h_" + tag)


def test_readme_grid(golden_readme):
    """C1-C3 (README.rst:55-84) incl. the 19-iteration ngal bisection."""
    g = golden_readme
    o = _build(g)
    assert_close(o.sigma2, g["sigma2"], 1e-10)
    assert_close(o.nzm, g["nzm"], 1e-9)
    assert_close(o.bh, g["bh"], 1e-10)
    o.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
    assert_close(o.uk_profiles["nfw"][:, ::8, ::4], g["uk_nfw_sub"], 1e-6, OSC)
    assert_close(o.uk_profiles["electron"][:, ::8, ::4], g["uk_e_sub"], 1e-6, OSC)
    o.add_hod("g", mthresh=10 ** 10.5 + g["zs"] * 0.0)
    o.add_hod("g2", ngal=g["g2_ngal_target"])
    assert o.hods["g2"]["iterations"] == 19
    for n in ("g", "g2"):
        for k in ("Nc", "Ns", "NsNsm1", "NcNs", "ngal", "bg", "log10mthresh"):
            assert_close(o.hods[n][k], g["hod_%s_%s" % (n, k)], 1e-8, name="%s.%s" % (n, k))
    for tag, a, b in [("mm", "nfw", "nfw"), ("ee", "electron", "electron"), ("me", "nfw", "electron"),
                      ("gg", "g", "g"), ("gm", "g", "nfw"), ("ge", "g", "electron"), ("g2g2", "g2", "g2"),
                      ("g2e", "g2", "electron")]:
        assert_close(o.get_power_1halo(a, b), g["P1h_" + tag], 1e-6, name="P1h_" + tag)
        assert_close(o.get_power_2halo(a, b), g["P2h_" + tag], 1e-6, name="P2h_" + tag)


def test_known_answers(golden_kat):
    """The reference's own known-answer ideas: Gaussian sine transform (fft.py:36-42,53) and
    utils.test_bisection_search (utils.py:45-51)."""
    g = golden_kat
    kt, U = orc.sine_transform(g["gauss_xs"], np.exp(-g["gauss_xs"] ** 2 / 2.0))
    assert_close(kt, g["gauss_kt"], 1e-14)
    assert_close(U, g["gauss_U"], 1e-9, 1e-13)
    # analytic sqrt(pi/2) k exp(-k^2/2): limited by the rectangle rule, not by parity
    sel = kt < 4
    assert np.max(np.abs(U[sel] - g["gauss_analytic"][sel])) < 5e-3
    y, _ = orc.bisect_all(g["bisect_x"], np.sqrt, 1.0, 40.0, 1e-4, decreasing=False)
    assert_close(y, g["bisect_y"], 1e-14)
    assert np.all(np.isclose(y, [4.0, 16.0, 36.0], rtol=1e-3))


def test_pk_interpolator(golden_pkspline):
    """The P(z,k) interpolator in front of the path (utils.py:53-182) against the reference's own builder."""
    from conftest import PKSPLINE_CASES
    g = golden_pkspline
    for key, tab, zq, kq, zt, kw in PKSPLINE_CASES:
        pk = -g[tab[1:]] if tab.startswith("-") else g[tab]
        PK = orc.PKOracle(g["ks_tab"], g[zt], pk, **kw)
        assert_close(PK.P(g[zq], g[kq]), g[key], 1e-12, name=key)
    z3 = g["zs_tab3"]
    bg = orc.Background()
    PK3 = orc.PKOracle(g["ks_tab"], z3, orc.plin_approx(bg, g["ks_tab"], z3))
    assert_close(PK3.P(g["zq3"], g["kq"]), g["P_kx2"], 1e-9, name="P_kx2")
    assert_close(orc.PKOracle(g["ks_tab"], g["zs_tab"], g["pk_tab"]).P(1.2345, g["kq"])[0], g["P_log_scalar_z"], 1e-12,
                 name="P_log_scalar_z")


def test_c_ky(golden_cky):
    """tSZ x lensing (cosmology.py:585-589): pressure x matter spectra and both lensing-source forms."""
    g = golden_cky
    o = orc.OracleHaloModel(g["zs"], g["ks"], g["ms"])
    o.add_battaglia_pres_profile("y", family="pres", xmax=20, nxs=5000)
    assert_close(o.get_power_1halo("y", "nfw"), g["P1h_ym"], 1e-6, name="P1h_ym")
    assert_close(o.get_power_2halo("y", "nfw"), g["P2h_ym"], 1e-6, name="P2h_ym")
    Pym = g["P1h_ym"] + g["P2h_ym"]
    assert_close(o.C_ky(g["ells"], g["zs"], g["ks"], Pym, lzs1=2.5), g["C_ky"], 1e-9)
    assert_close(o.C_ky(g["ells"], g["zs"], g["ks"], Pym, lzs1=g["lz"], ldndz1=g["ldndz"]), g["C_ky_dndz"], 1e-9)
    assert_close(o.C_gg(g["ells"], g["zs"], g["ks"], g["Pgg"], 0.8, zmin=0.7, zmax=0.9), g["C_gg_tophat"], 1e-9)
