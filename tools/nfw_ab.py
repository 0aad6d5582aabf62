#!/usr/bin/env python
"""A/B of the two NFW evaluations (hmv_set_nfw_mode 0: piecewise polynomials, 1: series + Si/Ci) on a z-slab of the
LARGE grid: CUDA-event times of hmv_uk_nfw and the largest difference between the two cubes.  Run on a GPU box."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmvec_b200 import _capi as capi, pipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nz", type=int, default=64)
ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
zs_all = np.linspace(0.01, 3., 200)
pick = np.linspace(0, 199, a.nz).round().astype(int)
zs = zs_all[pick]; ms = np.geomspace(2e10, 1e17, 2000); ks = np.geomspace(1e-4, 100, 10000)
import warnings; warnings.filterwarnings("ignore")
g = pipeline.GridSix(pipeline.make_inputs(zs, ms, ks, ngal=np.geomspace(1e-3, 1e-5, 200)[pick]))
g.upload(); g.run(); torch.cuda.synchronize()
L, d, ptr, st = capi.lib, g.d, capi.ptr, capi.stream()
nz, nm, nk, ldk = g.nz, g.nm, g.nk, g.ldk
out = {}
cubes = {}
for mode in (1, 0):
    capi.check(L.hmv_set_nfw_mode(mode), "mode")
    cube = torch.zeros_like(g.um)
    f = lambda: capi.check(L.hmv_uk_nfw(nz, nm, nk, ldk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["cs"]), ptr(d["rvir"]),
                                        ptr(d["nfw_ws"]), ptr(cube), st), "hmv_uk_nfw")
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out["mode%d_ms" % mode] = float(np.median(ts))
    cubes[mode] = cube
diff = (cubes[0] - cubes[1]).abs()
out["max_abs_diff"] = float(diff.max())
i = int(diff.argmax()); zi, r = divmod(i, nm * ldk); mi, ki = divmod(r, ldk)
out["argmax"] = [zi, mi, ki, float(cubes[1].reshape(-1)[i]), float(cubes[0].reshape(-1)[i])]
rel = diff / cubes[1].abs().clamp_min(1e-300)
big = cubes[1].abs() > 1e-6
out["max_rel_diff_where_|u|>1e-6"] = float(rel[big].max())
print("NFWAB " + json.dumps(out))
