#!/usr/bin/env python
"""Time hmv_power (generic tracer pair) on the LARGE grid: per-spectrum HBM GB/s.  Run on a GPU box."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmvec_b200 import pipeline, _capi as capi  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 200
zs = np.linspace(0.01, 3., nz); ms = np.geomspace(2e10, 1e17, 2000); ks = np.geomspace(1e-4, 100, 10000)
g = pipeline.GridSix(pipeline.make_inputs(zs, ms, ks))
g.upload(); g.run(); torch.cuda.synchronize()
d = g.d


def tracer(kind):
    t = capi.Tracer()
    if kind == "g":
        t.kind = 1; t.us_d = g.um.data_ptr()
        t.Nc_d, t.Ns_d = d["Nc"].data_ptr(), d["Ns"].data_ptr()
        t.NcNs_d, t.NsNsm1_d, t.ngal_d = d["NcNs"].data_ptr(), d["NsNsm1"].data_ptr(), d["ngal"].data_ptr()
    else:
        t.kind = 0; t.us_d = (g.um if kind == "m" else g.ue).data_ptr()
    return t


ws = torch.empty(int(capi.lib.hmv_power_ws_doubles(g.nz, g.nm)), dtype=torch.float64, device=g.device)
o1 = torch.empty((g.nz, g.nk), dtype=torch.float64, device=g.device); o2 = torch.empty_like(o1)
out = {}
for tag, (a, b), ncube in (("mm", "mm", 1), ("me", "me", 2), ("gg", "gg", 1), ("ge", "ge", 2)):
    A, B = tracer(a), tracer(b)
    call = lambda: capi.check(capi.lib.hmv_power(g.nz, g.nm, g.nk, g.ldk, capi.ptr(d["ms"]), capi.ptr(d["ks"]),
                                                 capi.ptr(d["nzm"]), capi.ptr(d["bh"]), capi.ptr(d["Pzk"]), g.rho_m0,
                                                 float(g.p['kstar_damping']), C.byref(A), C.byref(B), capi.ptr(ws),
                                                 capi.ptr(o1), capi.ptr(o2), capi.stream()), "hmv_power")
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        call()
    e1.record(); torch.cuda.synchronize()
    ms_ = e0.elapsed_time(e1) / 5
    by = g.nz * g.nk * (8.0 * g.nm * ncube + 24.0)
    out[tag] = {"ms": ms_, "alg_bytes": by, "gbs": by / ms_ / 1e6, "pts_per_s": g.nz * g.nk / (ms_ * 1e-3)}
print(json.dumps(out))
