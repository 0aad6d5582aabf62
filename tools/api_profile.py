#!/usr/bin/env python
"""Where the drop-in API's end-to-end step spends its time (LARGE grid, the bench's api_step workflow):
host-side return time of each call without synchronising (a call that returns late is host-bound), the same with a
synchronise after each call (device time per call), and a cProfile of one step.  Run on a GPU box."""
import cProfile
import contextlib
import io
import os
import pstats
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hmvec_b200 as hm  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 200
zs = np.linspace(0.01, 3., nz); ms = np.geomspace(2e10, 1e17, 2000); ks = np.geomspace(1e-4, 100, 10000)
ells = np.geomspace(10, 1e4, 1000)
ngal = np.geomspace(1e-3, 1e-5, nz)
PAIRS = (("nfw", "nfw"), ("electron", "electron"), ("nfw", "electron"), ("g", "g"), ("g", "nfw"), ("g", "electron"), ("y", "y"))


def step(sync):
    marks = []
    def mark(name):
        if sync:
            torch.cuda.synchronize()
        marks.append((name, time.perf_counter()))
    mark("start")
    with contextlib.redirect_stdout(io.StringIO()):
        h = hm.HaloModel(zs, ks, ms=ms, accuracy='low'); mark("ctor")
        h.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000); mark("electron")
        h.add_battaglia_pres_profile("y", family="pres", xmax=20, nxs=5000); mark("pressure")
        h.add_hod("g", ngal=ngal); mark("hod")
        P = {}
        for p in PAIRS:
            P[p] = h.get_power(*p); mark("P" + p[0][0] + p[1][0])
    ckk = h.C_kk(ells, zs, ks, P[("nfw", "nfw")], lzs1=2.5, lzs2=2.5); mark("C_kk")
    ckg = h.C_kg(ells, zs, ks, P[("g", "nfw")], gzs=0.8, lzs=2.5); mark("C_kg")
    cyy = h.C_yy(ells, zs, ks, P[("y", "y")]); mark("C_yy")
    torch.cuda.synchronize(); marks.append(("sync", time.perf_counter()))
    return marks


for _ in range(2):
    step(False)
for sync in (False, True):
    for rep in range(2):
        m = step(sync)
        print(("sync-after-each " if sync else "free-running    ") +
              " | ".join("%s %.1f" % (m[i][0], (m[i][1] - m[i - 1][1]) * 1e3) for i in range(1, len(m))) +
              " | total %.1f ms" % ((m[-1][1] - m[0][1]) * 1e3), flush=True)
pr = cProfile.Profile()
pr.enable(); step(False); pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
