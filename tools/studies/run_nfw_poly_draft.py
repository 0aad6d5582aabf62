#!/usr/bin/env python
"""One-shot GPU check of the polynomial NFW draft against the library's hmv_uk_nfw: max difference on a small and on
the nz=25 slab of the LARGE grid, and CUDA-event timings of both.  Run on a GPU box from the repo root."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hmvec_b200 as hm  # noqa: E402
from hmvec_b200 import _capi as capi  # noqa: E402

lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnfwpoly_draft.so"))
lib.nfwp_ws_doubles.restype = C.c_longlong
lib.nfwp_ws_doubles.argtypes = [C.c_int, C.c_int]
lib.nfwp_uk_nfw.restype = C.c_int
lib.nfwp_uk_nfw.argtypes = [C.c_int] * 4 + [C.c_void_p] * 7


def run(nz, nm, nk, reps):
    zs = np.linspace(0.01, 3.0, nz); ms = np.geomspace(2e10, 1e17, nm); ks = np.geomspace(1e-4, 100, nk)
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low', skip_nfw=True)
    dev = h.device
    f64 = dict(dtype=torch.float64, device=dev)
    ldk = (nk + 15) // 16 * 16
    zs_d, ks_d = torch.as_tensor(zs, device=dev), torch.as_tensor(ks, device=dev)
    cs_d, rv_d = h._cs_d, h._rvir_d
    A = torch.zeros((nz, nm, ldk), **f64); B = torch.zeros((nz, nm, ldk), **f64)
    wsA = torch.empty(int(capi.lib.hmv_uk_nfw_ws_doubles(nz, nm, nk)), **f64)
    wsB = torch.empty(int(lib.nfwp_ws_doubles(nz, nm)), **f64)
    st = capi.stream()
    fa = lambda: capi.check(capi.lib.hmv_uk_nfw(nz, nm, nk, ldk, capi.ptr(zs_d), capi.ptr(ks_d), float(ks.max()),
                                                capi.ptr(cs_d), capi.ptr(rv_d), capi.ptr(wsA), capi.ptr(A), st), "hmv_uk_nfw")
    fb = lambda: lib.nfwp_uk_nfw(nz, nm, nk, ldk, zs_d.data_ptr(), ks_d.data_ptr(), cs_d.data_ptr(), rv_d.data_ptr(),
                                 wsB.data_ptr(), B.data_ptr(), st)
    out = {}
    for name, f in (("library", fa), ("poly_draft", fb)):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / reps
    d = (A - B).abs()
    print("nz=%d nm=%d nk=%d: library %.3f ms, poly draft %.3f ms; max|diff| %.3e (max|u| %.3e), rel-to-max per row worst %.3e"
          % (nz, nm, nk, out["library"], out["poly_draft"], d.max().item(), A.abs().max().item(),
             (d.amax(dim=-1) / A.abs().amax(dim=-1)).max().item()))


def split(nz, nm, nk, reps, so):
    """pre-pass and cube kernel timed apart (study hook of the draft library: nk < 0 skips the pre-pass, ldk == 0 the cube)"""
    L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), so))
    L.nfwp_ws_doubles.restype = C.c_longlong; L.nfwp_ws_doubles.argtypes = [C.c_int, C.c_int]
    L.nfwp_uk_nfw.restype = C.c_int; L.nfwp_uk_nfw.argtypes = [C.c_int] * 4 + [C.c_void_p] * 7
    zs = np.linspace(0.01, 3.0, nz); ms = np.geomspace(2e10, 1e17, nm); ks = np.geomspace(1e-4, 100, nk)
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low', skip_nfw=True)
    dev = h.device; f64 = dict(dtype=torch.float64, device=dev); ldk = (nk + 15) // 16 * 16
    zs_d, ks_d = torch.as_tensor(zs, device=dev), torch.as_tensor(ks, device=dev)
    B = torch.zeros((nz, nm, ldk), **f64); ws = torch.empty(int(L.nfwp_ws_doubles(nz, nm)), **f64); st = capi.stream()
    call = lambda a, b: L.nfwp_uk_nfw(nz, nm, a, b, zs_d.data_ptr(), ks_d.data_ptr(), h._cs_d.data_ptr(), h._rvir_d.data_ptr(),
                                      ws.data_ptr(), B.data_ptr(), st)
    res = {}
    for name, (a, b) in (("prepass", (nk, 0)), ("cube", (-nk, ldk))):
        call(nk, ldk); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call(a, b)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps
    print("%s: pre-pass %.3f ms, cube %.3f ms" % (so, res["prepass"], res["cube"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "split":
        for so in ("libnfwpoly_draft.so", "libnfwpoly_draft_e2.so", "libnfwpoly_draft_e4.so"):
            split(25, 2000, 10000, 3, so)
    else:
        run(3, 64, 1001, 2)
        run(25, 2000, 10000, 3)
