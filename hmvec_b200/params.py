"""Parameter tables.  Keys and values mirror the reference's flat dicts (hmvec/params.py:2-113) because user code
indexes them by name (`hm.default_params['H0']`, `params={'st_a': ...}`); they are data, part of the parity contract."""


def _fit(names, rows):
    out = {}
    for q, (A0, am, az) in zip(names, rows):
        out[q + "_A0"], out[q + "_alpham"], out[q + "_alphaz"] = A0, am, az
    return out


# Battaglia 2016 power-law fits  X = A0 (M200c/1e14)^alpham (1+z)^alphaz          (params.py:2-37)
battaglia_defaults = {
    "AGN": _fit(("rho0", "alpha", "beta"), ((4000., 0.29, -0.66), (0.88, -0.03, 0.19), (3.83, 0.04, -0.025))),
    "SH": _fit(("rho0", "alpha", "beta"), ((19000., 0.09, -0.95), (0.70, -0.017, 0.27), (4.43, 0.005, 0.037))),
    "pres": _fit(("P0", "xc", "beta"), ((18.1, 0.154, -0.758), (0.497, -0.00865, 0.731), (4.35, 0.0393, 0.415))),
}

default_params = {}
# mass function / sigma^2 grid                                                    (params.py:43-50)
default_params.update(st_A=0.3222, st_a=0.707, st_p=0.3, st_deltac=1.686,
                      sigma2_kmin=1e-4, sigma2_kmax=2000, sigma2_numks=10000, Wkr_taylor_switch=0.01)
# profiles                                                                        (params.py:53-69)
default_params.update(duffy_A_vir=7.85, duffy_alpha_vir=-0.081, duffy_beta_vir=-0.71,
                      duffy_A_mean=10.14, duffy_alpha_mean=-0.081, duffy_beta_mean=-1.01,
                      nfw_integral_numxs=40000, nfw_integral_xmax=200,
                      electron_density_profile_integral_numxs=5000, electron_density_profile_integral_xmax=20,
                      electron_pressure_profile_integral_numxs=5000, electron_pressure_profile_integral_xmax=20,
                      battaglia_gas_gamma=-0.2, battaglia_gas_family='AGN',
                      battaglia_pres_gamma=-0.3, battaglia_pres_alpha=1., battaglia_pres_family='pres')
# power spectra                                                                   (params.py:72-73)
default_params.update(kstar_damping=0.01, default_halofit='mead')
# cosmology and constants                                                         (params.py:76-94)
default_params.update(omch2=0.1198, ombh2=0.02225, H0=67.3, ns=0.9645, As=2.2e-9, mnu=0.0, omk=0.0,
                      pivot_scalar=0.05, w0=-1.0, tau=0.06, nnu=3.046, wa=0., num_massive_neutrinos=3,
                      T_CMB=2.7255e6, parsec=3.08567758e16, mSun=1.989e30, thompson_SI=6.6524e-29,
                      meterToMegaparsec=3.241e-23, Yp=0.24)
# HOD                                                                             (params.py:97-107)
default_params.update(hod_A_log10mthresh=1.0, hod_sig_log_mstellar=0.2, hod_alphasat=1.0, hod_Bsat=9.04,
                      hod_betasat=0.74, hod_Bcut=1.65, hod_betacut=0.59,
                      hod_bisection_search_min_log10mthresh=7., hod_bisection_search_max_log10mthresh=14.,
                      hod_bisection_search_rtol=1e-4, hod_bisection_search_warn_iter=20)
default_params['class_output'] = ''                                             # params.py:110
