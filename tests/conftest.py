import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture %s missing" % path)
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_mini():
    return load_golden("mini")


@pytest.fixture(scope="session")
def golden_readme():
    return load_golden("readme")


@pytest.fixture(scope="session")
def golden_mini_mean():
    return load_golden("mini_mean")


@pytest.fixture(scope="session")
def golden_largeslab():
    return load_golden("largeslab")


@pytest.fixture(scope="session")
def golden_kat():
    return load_golden("kat")


PKSPLINE_CASES = [  # (golden key, table key, query-z key, query-k key, redshift-table key, builder kwargs)
    ("P_log", "pk_tab", "zq", "kq", "zs_tab", {}),
    ("P_extrap", "pk_tab", "zq", "kq_x", "zs_tab", {"extrap_kmax": 200.0}),
    ("P_signchange", "pk_tab_sc", "zq", "kq", "zs_tab", {}),
    ("P_negative", "-pk_tab", "zq", "kq", "zs_tab", {}),
]


def assert_close(got, want, rtol=1e-6, atol_frac=0.0, name=""):
    """rtol parity with an optional absolute floor expressed as a fraction of max|want|.

    The floor is only used for oscillatory intermediates (u(k) crosses zero); final spectra use 0."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (name, got.shape, want.shape)
    finite = np.isfinite(want)
    scale = np.max(np.abs(want[finite])) if finite.any() else 0.0
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol_frac * scale, equal_nan=True, err_msg=name)


@pytest.fixture(scope="session")
def golden_pkspline():
    return load_golden("pkspline")


@pytest.fixture(scope="session")
def golden_cky():
    return load_golden("cky")
