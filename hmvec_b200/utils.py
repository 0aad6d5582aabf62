"""Host-side helpers of the reference's `hmvec.utils` that user code calls directly (utils.py:6-51).

`vectorized_bisection_search` here is the generic, arbitrary-callable version (the callable is Python, so it runs on
the host); the HOD mthresh<->ngal solve on the hot path does NOT use it -- that is hmv_hod_bisect/hmv_hod_pick."""
import numpy as np


def interp(x, y, bounds_error=False, fill_value=0., **kwargs):
    from scipy.interpolate import interp1d
    return interp1d(x, y, bounds_error=bounds_error, fill_value=fill_value, **kwargs)


def vectorized_bisection_search(x, inv_func, ybounds, monotonicity, rtol=1e-4, verbose=True, hang_check_num_iter=20):
    """Find y with inv_func(y) == x element-wise by bisection on ybounds; every element keeps bisecting until
    ALL elements satisfy |inv_func(y)/x - 1| <= rtol, and the last midpoint is returned (utils.py:9-42)."""
    if monotonicity not in ('increasing', 'decreasing'):
        raise AssertionError("monotonicity must be 'increasing' or 'decreasing'")
    x = np.asarray(x, dtype=np.float64)
    lo = np.full(x.shape, float(ybounds[0]))
    hi = np.full(x.shape, float(ybounds[1]))
    err = np.full(x.shape, np.inf)
    n, warned, mid = 0, False, 0.5 * (lo + hi)
    while np.any(np.abs(err) > rtol):
        mid = 0.5 * (lo + hi)
        err = (inv_func(mid) - x) / x
        above, below = err > 0, err <= 0
        if monotonicity == 'decreasing':
            lo = np.where(above, mid, lo)
            hi = np.where(below, mid, hi)
        else:
            hi = np.where(above, mid, hi)
            lo = np.where(below, mid, lo)
        n += 1
        if n > hang_check_num_iter and not warned:
            print("WARNING: Bisection search has done more than ", hang_check_num_iter, " loops. Still searching...")
            warned = True
    if verbose:
        print("Bisection search converged in ", n, " iterations.")
    return mid


def test_bisection_search():
    """The reference's only executable assertion (utils.py:45-51)."""
    xs = np.array([2., 4., 6.])
    d = vectorized_bisection_search(xs, np.sqrt, (1, 40), 'increasing', rtol=1e-4, verbose=False)
    assert np.all(np.isclose(d, xs ** 2, rtol=1e-3))
