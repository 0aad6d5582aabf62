"""Import the UNMODIFIED reference package (simonsobs/hmvec) for test fixtures and CPU baselines -- TEST INFRASTRUCTURE.

`load(root)` imports `hmvec` from `root` (the installed copy `oracle/_ref`, built by oracle/build_ref.sh, or
/root/reference in the build container) with the three shims it needs to run in this image:
  * `camb` is not installed: tests/golden/camb_standin (flat-LCDM closed forms) stands in; with accuracy='low' both
    Pzk (hmvec.py:98-99) and the sigma^2 spectrum (cosmology.py:259-260) come from the reference's own EH98
    P_lin_approx, so no CAMB product is ever needed;
  * SciPy >= 1.14 removed `interp2d` and `dfitpack.bispeu`, which limber_integral calls (cosmology.py:890,899):
    shimmed with RectBivariateSpline(kx=ky=1) / `_fitpack.bispeu`, SciPy's documented bug-for-bug replacement, so the
    reference's own function body executes;
  * tinker.py:64 looks for its table one directory too high: the one dirname() call it makes is redirected to the
    package's own data/ directory.
Only tests/golden/make_golden.py and the CPU legs of bench.py call this."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STANDIN = os.path.join(os.path.dirname(HERE), "tests", "golden", "camb_standin")
DEFAULT_ROOT = os.path.join(HERE, "_ref")


def available(root=DEFAULT_ROOT):
    return os.path.exists(os.path.join(root, "hmvec", "hmvec.py"))


def load(root=DEFAULT_ROOT):
    try:
        import camb  # noqa: F401
        if "camb_standin" not in getattr(camb, "__file__", ""):
            raise SystemExit("a real camb is installed; the stand-in must not shadow it")
    except ImportError:
        pass
    if STANDIN not in sys.path:
        sys.path.insert(0, STANDIN)
    if root not in sys.path:
        sys.path.insert(0, root)
    import hmvec  # noqa
    import hmvec.cosmology as hcosm
    import hmvec.tinker as htinker
    import scipy.interpolate._fitpack as _fp
    from scipy.interpolate import RectBivariateSpline

    class _Interp2d(object):
        """interp2d(ks, zs, Pzks[nz,nk]) for a regular grid == FITPACK regrid with kx=ky=1, s=0."""

        def __init__(self, x, y, z, bounds_error=False, **kw):
            tx, ty, c = RectBivariateSpline(np.asarray(x), np.asarray(y), np.asarray(z).T, kx=1, ky=1, s=0).tck
            self.tck = (tx, ty, c, 1, 1)

    hcosm.interp2d = _Interp2d
    hcosm.si = types.SimpleNamespace(dfitpack=types.SimpleNamespace(bispeu=_fp.bispeu))
    data_dir = os.path.join(os.path.dirname(os.path.abspath(hmvec.__file__)), "data")
    htinker.os = types.SimpleNamespace(path=types.SimpleNamespace(dirname=lambda f: data_dir))
    return hmvec
