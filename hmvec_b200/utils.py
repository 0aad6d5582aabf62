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


# ---- P(z,k) interpolators (utils.py:53-182): host fit, device evaluation -------------------------------------------
def _table_for_spline(ks, zs, pk, log_interp=True, extrap_kmax=None):
    """The (zs, ln k, values) table the reference hands to RectBivariateSpline (utils.py:139-170): log of |P| unless
    P changes sign, plus the two power-law extension nodes beyond kmax when extrap_kmax is given."""
    ks, zs, pk = (np.asarray(a, dtype=np.float64) for a in (ks, zs, pk))
    sign = 1
    if log_interp and np.any(pk <= 0):
        if np.all(pk < 0):
            sign = -1
        else:
            log_interp = False
    vals = np.log(sign * pk) if log_interp else pk
    logk = np.log(ks)
    if extrap_kmax and extrap_kmax > ks[-1]:
        if not log_interp:
            raise ValueError("Cannot use extrap_kmax with log_interp=False (P(k) crosses zero)")
        top = np.log(extrap_kmax)
        delta = top - logk[-1]
        slope = (vals[:, -1] - vals[:, -2]) / (logk[-1] - logk[-2])
        vals = np.concatenate([vals, (vals[:, -1] + slope * delta * 0.9)[:, None], (vals[:, -1] + slope * delta)[:, None]], axis=1)
        logk = np.concatenate([logk, [top - delta * 0.1, top]])
    return zs, logk, vals, bool(log_interp), sign


class PKInterpolatorDevice(object):
    """`PK.P(z, k, grid=True)` evaluated on the GPU (hmv_pk_spline) from a host-side FITPACK fit.

    Built either from a table (`get_matter_power_interpolator_generic`, the CLASS route of the reference) or from any
    scipy RectBivariateSpline-derived interpolator carrying `islog` / `logsign` (CAMB's
    get_matter_power_interpolator result) with `from_spline`.  Attributes kmin, kmax, zmin, zmax, islog, logsign as
    in the reference.  There is no CPU evaluation path here: `P` needs a CUDA device."""

    def __init__(self, tck, degrees, islog, logsign, device=None):
        import torch
        from . import _capi as capi
        if not torch.cuda.is_available():
            raise RuntimeError("hmvec_b200: PKInterpolatorDevice needs a CUDA device (no CPU fallback)")
        self._capi, self._torch = capi, torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        tx, ty, c = (np.ascontiguousarray(a, dtype=np.float64) for a in tck)
        self.kx, self.ky = int(degrees[0]), int(degrees[1])
        self.nx, self.ny = tx.size, ty.size
        if c.size != (self.nx - self.kx - 1) * (self.ny - self.ky - 1):
            raise ValueError("spline coefficients do not match the knot vectors")
        self._tx, self._ty, self._c = (torch.as_tensor(a, device=self.device) for a in (tx, ty, c))
        self.islog, self.logsign = bool(islog), int(logsign)

    @classmethod
    def from_spline(cls, spl, device=None):
        return cls(spl.tck, spl.degrees, getattr(spl, "islog", False), getattr(spl, "logsign", 1), device=device)

    def P_device(self, z, k, scale=1.0):
        """[nz,nk] float64 CUDA tensor of scale * P(z,k) on the outer-product grid of z and k."""
        torch, capi = self._torch, self._capi
        zs = torch.as_tensor(np.ascontiguousarray(np.atleast_1d(z), dtype=np.float64), device=self.device)
        ks = torch.as_tensor(np.ascontiguousarray(np.atleast_1d(k), dtype=np.float64), device=self.device)
        out = torch.empty((zs.numel(), ks.numel()), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            capi.check(capi.lib.hmv_pk_spline(zs.numel(), ks.numel(), capi.ptr(zs), capi.ptr(ks), self.nx, self.ny, self.kx,
                                              self.ky, capi.ptr(self._tx), capi.ptr(self._ty), capi.ptr(self._c),
                                              int(self.islog), float(scale) * (self.logsign if self.islog else 1.0),
                                              capi.ptr(out), capi.stream()), "hmv_pk_spline")
        return out

    def P(self, z, k, grid=None):
        """numpy result with the reference's shape conventions (utils.py:95-103)."""
        if grid is None:
            grid = not np.isscalar(z) and not np.isscalar(k)
        if not grid and not (np.isscalar(z) or np.isscalar(k)):
            zz, kk = np.broadcast_arrays(np.asarray(z, dtype=np.float64), np.asarray(k, dtype=np.float64))
            flat = np.array([self.P_device(a, b).cpu().numpy()[0, 0] for a, b in zip(zz.ravel(), kk.ravel())])
            return flat.reshape(zz.shape)
        out = self.P_device(z, k).cpu().numpy()
        if np.isscalar(z) and np.isscalar(k):
            return out[0, 0]
        if grid:
            return out
        return out[0] if np.isscalar(z) else out[:, 0]


def get_matter_power_interpolator_generic(ks, zs, pk, return_z_k=False, log_interp=True, extrap_kmax=None,
                                          silent=False, device=None):
    """utils.py:53-182 with the evaluation moved to the GPU: pk[z,k] on increasing zs, ks -> PKInterpolatorDevice."""
    from scipy.interpolate import RectBivariateSpline
    ks = np.asarray(ks, dtype=np.float64)
    zs_t, logk, vals, islog, sign = _table_for_spline(ks, zs, pk, log_interp, extrap_kmax)
    if zs_t.size < 2:
        raise NotImplementedError("single-redshift interpolators (interp1d in the reference) stay on the host")
    spl = RectBivariateSpline(zs_t, logk, vals, kx=min(zs_t.size - 1, 3), ky=min(logk.size - 1, 3))   # host FITPACK fit
    res = PKInterpolatorDevice(spl.tck, spl.degrees, islog, sign, device=device)
    res.kmin, res.kmax, res.zmin, res.zmax = float(np.min(ks)), float(ks[-1]), float(np.min(zs_t)), float(np.max(zs_t))
    return (res, zs_t, ks) if return_z_k else res
