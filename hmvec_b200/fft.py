"""Profile transforms: drop-in for the reference's `hmvec/fft.py`.

`generic_profile_fft` -- the numerical Fourier transform of ANY radial profile onto the target wavenumbers
(fft.py:56-115) -- evaluates the caller's profile function on the host (it is an arbitrary Python callable, exactly
as in the reference) and hands the samples to the device kernel that `add_battaglia_profile` uses
(hmv_profile_transform_samples: theta-cut, mass norm, tensor-core sine sums, interpolation, one fused pass).  The
built-in Battaglia / NFW profiles never go through here: their samples are evaluated inside the kernel.

`fft_integral`, `uk_fft`, `uk_brute_force` are the small 1-D helpers of fft.py:8-53 (host numpy; not on the path).
"""
import numpy as np
import torch

from . import _capi as capi


def _trapz(y, x, axis=-1):
    f = getattr(np, "trapezoid", None) or np.trapz
    return f(y, x, axis=axis)


def fft_integral(x, y, axis=-1):
    """int_0^inf dx x sin(kx) y(x) by an FFT (fft.py:35-51): returns (ks, uk) with the reference's step and
    frequency conventions (step = (x[-1]-x[0])/N)."""
    assert x.ndim == 1
    N = x.size
    step = (x[-1] - x[0]) / N
    uk = -np.fft.rfft(x * y, axis=axis).imag * step
    return np.fft.rfftfreq(N, step) * 2 * np.pi, uk


def analytic_fft_integral(ks):
    """fft_integral of y = exp(-x^2/2) in closed form (fft.py:53)."""
    return np.sqrt(np.pi / 2.) * np.exp(-ks ** 2. / 2.) * ks


def uk_fft(rhofunc, rvir, dr=0.001, rmax=100):
    """fft.py:8-19"""
    rvir = np.asarray(rvir)
    rs = np.arange(dr, rmax, dr)
    rhos = rhofunc(np.abs(rs))
    integrand = rhos * (np.abs(rs) <= rvir[..., None])
    m = _trapz(integrand * rs ** 2., rs, axis=-1) * 4. * np.pi
    ks, ukt = fft_integral(rs, integrand)
    return ks, 4. * np.pi * ukt / ks / m[..., None]


def uk_brute_force(r, rho, rvir, ks):
    """Direct quadrature of the same transform (fft.py:22-33)."""
    sel = np.where(r < rvir)
    rs, rhos = r[sel], rho[sel]
    m = _trapz(rhos * rs ** 2., rs) * 4. * np.pi
    integrand = 4. * np.pi * rs[:, None] * np.sin(rs[:, None] * ks[None, :]) * rhos[:, None] / ks[None, :]
    return _trapz(integrand, rs, axis=0) / m


def generic_profile_fft(rhofunc_x, cmaxs, rss, zs, ks, xmax, nxs, do_mass_norm=True, device=None):
    """u(k|M,z) of the profile rhofunc_x(x), x = r/r_s, truncated at cmaxs (fft.py:56-115).

    rhofunc_x: callable on xs = linspace(0,xmax,nxs+1)[1:] returning [nz,nm,nxs]; cmaxs, rss: [nz,nm]; zs: [nz];
    ks: [nk].  Returns (ks, ukouts [nz,nm,nk]) as numpy, like the reference."""
    if not torch.cuda.is_available():
        raise RuntimeError("hmvec_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback.")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    xs = np.linspace(0., xmax, nxs + 1)[1:]
    rhos = np.asarray(rhofunc_x(xs), dtype=np.float64)
    if rhos.ndim == 1:
        rhos = rhos[None, None, :]
    cmaxs = np.asarray(cmaxs, dtype=np.float64)
    nz, nm = cmaxs.shape
    rhos = np.ascontiguousarray(np.broadcast_to(rhos, (nz, nm, nxs)))
    ks64 = np.asarray(ks, dtype=np.float64).reshape(-1)
    nk = ks64.size
    ldk = ((nk + 15) // 16) * 16
    up = lambda a: torch.as_tensor(np.array(a, dtype=np.float64, order='C'), device=dev)
    rho_d, cm_d, rs_d = up(rhos), up(cmaxs), up(np.broadcast_to(np.asarray(rss, dtype=np.float64), (nz, nm)))
    zs_d, ks_d = up(np.asarray(zs, dtype=np.float64).reshape(-1)), up(ks64)
    out = torch.empty((nz, nm, ldk), dtype=torch.float64, device=dev)
    ws = torch.empty(int(capi.lib.hmv_profile_transform_ws_doubles(nz, nm, int(nxs))), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        capi.check(capi.lib.hmv_profile_transform_samples(nz, nm, nk, ldk, capi.ptr(zs_d), capi.ptr(ks_d),
                                                          float(np.max(ks64)), capi.ptr(rs_d), capi.ptr(cm_d),
                                                          capi.ptr(rho_d), None, float(xmax), int(nxs),
                                                          int(bool(do_mass_norm)), capi.ptr(ws), capi.ptr(out),
                                                          capi.stream()), "hmv_profile_transform_samples")
        res = out[..., :nk].cpu().numpy()
    return ks, res
