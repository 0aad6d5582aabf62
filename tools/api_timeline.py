#!/usr/bin/env python
"""GPU timeline of ONE step of the drop-in API (torch.profiler / CUPTI): every kernel and memcpy with its start and
duration relative to the first host call, and the idle gaps between them -- where a short slab's step is host-bound.

    python tools/api_timeline.py 25
"""
import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hmvec_b200 as hm  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 25
zs = np.linspace(0.01, 3., nz); ms = np.geomspace(2e10, 1e17, 2000); ks = np.geomspace(1e-4, 100, 10000)
ells = np.geomspace(10, 1e4, 1000)
ngal = np.geomspace(1e-3, 1e-5, nz)
PAIRS = (("nfw", "nfw"), ("electron", "electron"), ("nfw", "electron"), ("g", "g"), ("g", "nfw"), ("g", "electron"), ("y", "y"))


def step():
    with contextlib.redirect_stdout(io.StringIO()):
        h = hm.HaloModel(zs, ks, ms=ms, accuracy='low')
        h.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
        h.add_battaglia_pres_profile("y", family="pres", xmax=20, nxs=5000)
        h.add_hod("g", ngal=ngal)
        P = {p: h.get_power(*p) for p in PAIRS}
    h.C_kk(ells, zs, ks, P[("nfw", "nfw")], lzs1=2.5, lzs2=2.5)
    h.C_kg(ells, zs, ks, P[("g", "nfw")], gzs=0.8, lzs=2.5)
    h.C_yy(ells, zs, ks, P[("y", "y")])
    torch.cuda.synchronize()


for _ in range(3):
    step()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    step()
    wall = (time.perf_counter() - t0) * 1e3
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
first = min(e.time_range.start for e in prof.events())
print("step wall %.2f ms under the profiler; GPU activities:" % wall)
end_prev, busy = None, 0.0
for e in evs:
    s, d = (e.time_range.start - first) / 1e3, (e.time_range.end - e.time_range.start) / 1e3
    gap = "" if end_prev is None or s - end_prev < 0.02 else "   <-- idle %.3f ms" % (s - end_prev)
    if d >= 0.02 or gap:
        print("%8.3f ms  +%7.3f  %s%s" % (s, d, e.name[:70], gap))
    end_prev = max(end_prev or 0.0, s + d)
    busy += d
print("sum of GPU activity %.2f ms, last activity ends at %.2f ms" % (busy, end_prev))
