"""The C-ABI library loads and exports every symbol include/hmvec_b200.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hmvec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hmv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "hmvec_b200", "libhmvec_b200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/hmvec_b200.h but not exported: %s" % missing


def test_binding_covers_header_and_reports_errors():
    from hmvec_b200 import _capi as capi
    assert set(_declared()) == set(capi.EXPORTS)
    assert capi.lib.hmv_abi_version() == 1
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    assert capi.lib.hmv_power_ws_doubles(0, 5) == 0
    assert capi.lib.hmv_sigma2_ws_doubles(4, 64, 1000) >= 64 * 1000
    rc = capi.lib.hmv_limber(0, None, 1, 2, 2, None, None, None, None, 1, None, None, None, None, None)
    assert rc == -1 and "bad sizes" in capi.last_error()
    rc = capi.lib.hmv_hod(3, 10, None, None, None, None, 0, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and "null pointer" in capi.last_error()


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle (or the reference) anywhere."""
    pkg = os.path.join(ROOT, "hmvec_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in txt.replace("test-side oracle", ""), fn
            assert "/root/reference" not in txt, fn
