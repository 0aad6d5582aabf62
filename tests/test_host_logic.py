"""Host-side producers of the product package (background, EH98 P_lin, Simpson weights, windows, slabs, bisection)
against the CPU oracle / scipy.  CPU-only."""
import warnings

import numpy as np
import pytest
from scipy.integrate import simpson

from conftest import assert_close
from oracle import hmvec_oracle as orc


@pytest.fixture(scope="module")
def cosmo():
    from hmvec_b200.cosmology import Cosmology
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return Cosmology({}, accuracy='low')


def test_simpson_weights_match_scipy():
    from hmvec_b200.cosmology import simpson_weights
    rng = np.random.default_rng(0)
    for n in (2, 3, 4, 7, 10, 101, 1000, 10000):
        x = np.geomspace(1e-4, 2000, n) if n > 2 else np.array([0.3, 1.7])
        y = rng.standard_normal((3, n))
        assert_close(y @ simpson_weights(x), simpson(y, x=x, axis=-1), 1e-12, 1e-15)


def test_background_and_linear_power(cosmo, golden_mini):
    g = golden_mini
    zs = g["zs"]
    assert_close(cosmo.hubble_parameter(zs), g["hubble"], 1e-12)
    assert_close(cosmo.comoving_radial_distance(zs), g["chi"], 1e-11)
    assert_close(cosmo.rho_critical_z(zs), g["rho_crit"], 1e-12)
    assert_close(cosmo.P_lin_approx(g["ks"], zs), g["Pzk"], 1e-9)
    assert_close(cosmo.lensing_window(zs, 2.5), g["lens_window_25"], 1e-9)
    assert_close(cosmo.lensing_window(zs, g["lz"], g["ldndz"]), g["lens_window_dndz"], 1e-9)
    assert float(np.ravel(cosmo.comoving_radial_distance(1100.))[0]) == pytest.approx(orc.Background().chi(1100.), rel=1e-9)


def test_make_inputs_and_slabs(golden_mini):
    from hmvec_b200 import pipeline, zshard
    g = golden_mini
    inp = pipeline.make_inputs(g["zs"], g["ms"], g["ks"], ells=g["ells"])
    assert_close(inp["Pzk"], g["Pzk"], 1e-9)
    assert_close(inp["sPzk"][:, ::10], g["sPzk_sub"], 1e-9)
    assert inp["kw"].shape == inp["ks_sig"].shape == (10000,)
    b = zshard.slab_bounds(g["zs"].size, 4)
    assert b[0] == 0 and b[-1] == g["zs"].size and np.all(np.diff(b) >= 1)
    parts = [pipeline.slab_inputs(inp, zshard.slab(g["zs"].size, r, 4)) for r in range(4)]
    assert_close(np.concatenate([p["Pzk"] for p in parts]), inp["Pzk"], 0)
    assert_close(np.concatenate([p["zs"] for p in parts]), g["zs"], 0)
    assert parts[0]["chis"].size == g["zs"].size          # Limber inputs stay replicated
    for nz, w in ((200, 8), (7, 3), (5, 8)):
        bb = zshard.slab_bounds(nz, w)
        assert bb[-1] == nz and np.diff(bb).max() - np.diff(bb).min() <= 1


def test_utils_bisection(golden_kat):
    from hmvec_b200 import utils
    utils.test_bisection_search()
    y = utils.vectorized_bisection_search(golden_kat["bisect_x"], np.sqrt, (1, 40), 'increasing', verbose=False)
    assert_close(y, golden_kat["bisect_y"], 1e-14)
    with pytest.raises(AssertionError):
        utils.vectorized_bisection_search(np.ones(2), np.sqrt, (1, 40), 'sideways')


def test_device_ops_fail_loudly_without_gpu(cosmo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cosmo.get_sigma2_R(np.array([1.0, 2.0]), np.array([0.5]))
    from hmvec_b200 import HaloModel
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            HaloModel(np.array([0.1]), np.geomspace(1e-3, 1, 8), ms=np.geomspace(1e12, 1e14, 4), accuracy='low')


def test_profile_dict_membership_does_not_download():
    """`name in hm.uk_profiles` / `.keys()` are what add_hod and get_power use to classify tracers; they must not
    trigger DeviceCubes.__getitem__ (a 32 GB device->host copy per test on the LARGE grid)."""
    from hmvec_b200.hmvec import DeviceCubes

    class Boom(DeviceCubes):
        def __getitem__(self, name):
            raise AssertionError("membership test downloaded a cube")

    dc = Boom(owner=None)
    dc._t["nfw"] = object()
    assert "nfw" in dc and "electron" not in dc
    assert "nfw" in dc.keys() and "electron" not in dc.keys()
    assert list(dc) == ["nfw"] and len(dc) == 1


def test_pk_table_preparation_matches_oracle():
    """Host half of the P(z,k) interpolator (utils.py:139-170): log|P| / sign / extension nodes handed to the spline
    fit are the oracle's, for a positive, a negative and a sign-changing table, with and without extrap_kmax."""
    import numpy as np
    from hmvec_b200.utils import _table_for_spline
    from oracle import hmvec_oracle as orc
    zs = np.linspace(0.0, 3.0, 9)
    ks = np.geomspace(1e-4, 20.0, 60)
    pk = orc.plin_approx(orc.Background(), ks, zs)
    for tab, kw in ((pk, {}), (pk, {"extrap_kmax": 150.0}), (-pk, {}), (pk * np.cos(2.0 * np.log(ks))[None, :], {})):
        z_t, logk, vals, islog, sign = _table_for_spline(ks, zs, tab, True, kw.get("extrap_kmax"))
        o = orc.PKOracle(ks, zs, tab, **kw)
        tx, ty, _ = o.spl.tck
        assert islog == o.islog and sign == o.sign
        assert logk.size == ty.size - 4 and np.isclose(logk[-1], ty[-1]) and np.isclose(logk[0], ty[0])
        # the oracle's spline interpolates exactly the table the product prepares
        np.testing.assert_allclose(o.spl(z_t, logk), vals, rtol=1e-10, atol=1e-12)
    import pytest
    with pytest.raises(ValueError):
        _table_for_spline(ks, zs, pk * np.cos(2.0 * np.log(ks))[None, :], True, 150.0)


def test_launch_orders_are_permutations():
    """The work orders the kernels pick for load balance (host-side introspection entry points, no device work): the
    profile transform's queue (a mixed head, then jointly heavy-first) and the tile-slow order of the table reduction
    must visit every (z, mass group) / (z, k tile) exactly once for any grid, incl. ragged ones and slabs shorter than
    the head; the six-spectra kernel's tile width stays a multiple of 16 within [448, 512]."""
    import ctypes as C
    import os
    from hmvec_b200 import _capi as capi
    os.environ.pop("HMV_K1_TAIL", None)
    os.environ.pop("HMV_TILE", None)
    for nz, nm, grid in ((1, 16, 0), (3, 48, 0), (6, 200, 0), (25, 2000, 0), (64, 2000, 148), (200, 2000, 148),
                         (7, 37, 4), (40, 1000, 16)):
        nmg = -(-nm // 16)
        z = (C.c_int * (nz * nmg))()
        q = (C.c_int * (nz * nmg))()
        assert capi.lib.hmv_debug_k1_order(nz, nm, grid, z, q) == nmg
        seen = {(z[i], q[i]) for i in range(nz * nmg)}
        assert len(seen) == nz * nmg and all(0 <= a < nz and 0 <= b < nmg for a, b in seen), (nz, nm, grid)
        if nz >= 25:
            # the launch ends on the lightest mass groups (q counts from the heavy end) of the heavy-first redshifts
            assert q[nz * nmg - 1] == nmg - 1 and q[nz * nmg - 2] == nmg - 1
    os.environ["HMV_K1_TAIL"] = "-1"          # measurement variant: heaviest and lightest items alternate
    try:
        z, q = (C.c_int * (6 * 13))(), (C.c_int * (6 * 13))()
        capi.lib.hmv_debug_k1_order(6, 200, 0, z, q)
        assert len({(z[i], q[i]) for i in range(6 * 13)}) == 6 * 13
    finally:
        del os.environ["HMV_K1_TAIL"]
    for nz, nk in ((1, 10), (6, 1001), (25, 10000), (100, 10000), (130, 3000), (200, 10000), (201, 513)):
        n = nz * (-(-nk // 512))
        z, k0 = (C.c_int * n)(), (C.c_int * n)()
        tile = capi.lib.hmv_debug_tab_order(nz, nk, z, k0)
        assert tile == 512
        seen = {(z[i], k0[i]) for i in range(n)}
        assert len(seen) == n and all(0 <= a < nz and 0 <= b < nk and b % tile == 0 for a, b in seen), (nz, nk)
        assert k0[0] == (-(-nk // 512) - 1) * 512 and k0[n - 1] == 0      # highest wavenumbers first, lowest last
    tiles = {nz: capi.lib.hmv_debug_wave_tile(nz, 10000) for nz in (12, 13, 25, 50, 100, 200)}
    assert all(t % 16 == 0 and 448 <= t <= 512 for t in tiles.values()), tiles
    assert tiles[25] == 448 and tiles[50] == tiles[100] == tiles[200] == 512, tiles
