"""kSZ consumer against values produced by the reference's own ksz.py (tests/golden/ksz.npz,
make_golden.py::case_ksz): the device Nvv integral, hmvec_b200.ksz.kSZ, and -- the attribute contract -- the
reference's UNMODIFIED kSZ class body executed on top of hmvec_b200.HaloModel."""
import os
import sys
import types
import warnings

import numpy as np
import pytest

from conftest import assert_close, load_golden, ROOT, GOLDEN

ATTRS = ("kLs", "krs", "mu", "kS", "Hphotozs", "sPggs", "sPges", "Pmms", "fs", "adotf", "d2vs", "kstars", "chistars",
         "vrec", "sPggtot", "sPge", "bgs")


def test_ksz_free_functions_cpu():
    """Closed-form helpers (no device): ksz.py:31-96, 422-433, 43-63."""
    warnings.filterwarnings("ignore")
    from hmvec_b200 import ksz
    g = load_golden("ksz")
    assert_close(ksz.ne0_shaw(0.02225, 0.24), g["ne0_shaw"], 1e-14)
    assert_close(ksz.ksz_radial_function(g["kat_z"], 0.02225, 0.24), g["ksz_radial_function"], 1e-14)
    assert_close(ksz.get_kmin(np.array([1.0, 50.0])), g["get_kmin"], 1e-14)
    assert_close(ksz.chi(0.24, 1), g["chi"], 1e-15)
    assert_close(ksz.get_interpolated_cls(g["Cls"].copy(), 1800., g["kS"]), g["interp_cls"], 0)
    assert_close(ksz.pge_err_core(3.5, 2.2e-7, 1800., 4.0, g["kS"], g["ks_bin_edges"], g["Pgg1"], g["Cls"].copy()),
                 g["pge_err_core"], 1e-12)


@pytest.fixture()
def camb_standin():
    """accuracy='medium' without CAMB: put the test-side stand-in on the path for the duration of one test (its
    closed-form PK is what the reference received when the golden file was made)."""
    try:
        import camb  # noqa: F401
        pytest.skip("a real camb is installed")
    except ImportError:
        pass
    path = os.path.join(GOLDEN, "camb_standin")
    sys.path.insert(0, path)
    yield
    sys.path.remove(path)
    for m in [m for m in sys.modules if m == "camb" or m.startswith("camb.")]:
        del sys.modules[m]


@pytest.mark.gpu
def test_nvv_core_integral_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    from hmvec_b200 import ksz
    g = load_golden("ksz")
    mu, kL, kS, Cls = g["mu"], g["kL"], g["kS"], g["Cls"]
    assert_close(ksz.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), g["Pge1"], g["Pgg1"]), g["Nvv_1d"], 1e-10)
    W = np.exp(-(mu[:, None] * kL[None, :] * 40.) ** 2)[..., None]
    Pge3, Pgg3 = g["Pge1"][None, None] * W, g["Pgg1"][None, None] * W ** 2 + 50.
    assert_close(ksz.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), Pge3, Pgg3), g["Nvv_3d"], 1e-10)
    nv, pe = ksz.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), Pge3, Pgg3, errs=True)
    assert_close(nv, g["Nvv_errs"], 1e-10)
    assert_close(pe, Pge3, 0)
    assert_close(ksz.Nvv_core_integral(1800., 2.2e-7, mu, kL, kS, Cls.copy(), Pge3, Pgg3, Pgg_photo_tot=Pgg3 * 1.3,
                                       robust_term=True, photo=False), g["Nvv_robust"], 1e-10)


def _check_class(k, g, tag, edges):
    for a in ATTRS:
        assert_close(np.asarray(getattr(k, a)), g["c%s_%s" % (tag, a)], 1e-6, name="kSZ%s.%s" % (tag, a))
    Cls = g["Cls"]
    assert_close(k.Nvv(0, Cls.copy()), g["c%s_Nvv0" % tag], 1e-6, name="Nvv0" + tag)
    assert_close(k.Nvv(1, Cls.copy()), g["c%s_Nvv1" % tag], 1e-6, name="Nvv1" + tag)
    assert_close(k.lPgg(0, 1.5, 1.7), g["c%s_lPgg" % tag], 1e-6)
    assert_close(k.lPgv(1, 1.4), g["c%s_lPgv" % tag], 1e-6)
    assert_close(k.lPvv(1), g["c%s_lPvv" % tag], 1e-6)
    if not tag:
        assert_close(k.Pge_err(0, edges, Cls.copy()), g["c_Pge_err"], 1e-6)
    assert_close(k.sigma2, g["c%s_sigma2" % tag], 1e-9, name="sigma2 (accuracy='medium')")
    assert_close(k.Pzk, g["c%s_Pzk" % tag], 1e-12, name="Pzk (accuracy='medium')")


@pytest.mark.gpu
@pytest.mark.parametrize("sigz", [None, 0.03])
def test_ksz_class_matches_reference(camb_standin, sigz):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    warnings.filterwarnings("ignore")
    from hmvec_b200 import ksz
    g = load_golden("ksz")
    tag = "" if sigz is None else "_pz"
    k = ksz.kSZ(g["c_zs"], list(g["c_vols"]), list(g["c_ngals"]), num_kL_bins=24, num_kS_bins=41, num_mu_bins=14,
                ms=g["c_ms"], sigz=sigz, electron_profile_nxs=5000, electron_profile_xmax=20)
    _check_class(k, g, tag, g["ks_bin_edges"])


@pytest.mark.gpu
def test_reference_ksz_class_runs_on_device_halomodel(camb_standin):
    """The attribute contract: the reference's own `class kSZ(HaloModel)` source (oracle/_ref/hmvec/ksz.py, unmodified),
    with its relative imports resolved inside hmvec_b200, constructs and evaluates on the device-backed HaloModel and
    reproduces the numbers the reference produced on its own HaloModel."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    src = os.path.join(ROOT, "oracle", "_ref", "hmvec", "ksz.py")
    if not os.path.exists(src):
        pytest.skip("oracle/_ref not built (oracle/build_ref.sh)")
    warnings.filterwarnings("ignore")
    import hmvec_b200  # noqa: F401
    mod = types.ModuleType("hmvec_b200._reference_ksz")
    mod.__package__ = "hmvec_b200"
    mod.__file__ = src
    exec(compile(open(src).read(), src, "exec"), mod.__dict__)
    g = load_golden("ksz")
    k = mod.kSZ(g["c_zs"], list(g["c_vols"]), list(g["c_ngals"]), num_kL_bins=24, num_kS_bins=41, num_mu_bins=14,
                ms=g["c_ms"], sigz=None, engine='camb', electron_profile_nxs=5000, electron_profile_xmax=20)
    assert isinstance(k, hmvec_b200.HaloModel)
    for a in ATTRS:
        assert_close(np.asarray(getattr(k, a)), g["c_%s" % a], 1e-6, name="reference kSZ on hmvec_b200: " + a)
    assert_close(k.Nvv(0, g["Cls"].copy()), g["c_Nvv0"], 1e-6)
