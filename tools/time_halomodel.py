#!/usr/bin/env python
"""Wall-clock of the drop-in HaloModel API on the LARGE grid (README workflow): constructor, electron profile,
ngal-HOD, six get_power calls, C_kk/C_kg.  Run on a GPU box."""
import os
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
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h = hm.HaloModel(zs, ks, ms=ms, accuracy='low')
    torch.cuda.synchronize(); t1 = time.perf_counter()
    h.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    h.add_hod("g", ngal=ngal)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    P = {}
    for a, b in (("nfw", "nfw"), ("electron", "electron"), ("nfw", "electron"), ("g", "g"), ("g", "nfw"), ("g", "electron")):
        P[(a, b)] = h.get_power(a, b)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    p1, p2 = h.get_power_six("nfw", "electron", "g")
    torch.cuda.synchronize(); t5 = time.perf_counter()
    ckk = h.C_kk(ells, zs, ks, P[("nfw", "nfw")], lzs1=2.5, lzs2=2.5)
    ckg = h.C_kg(ells, zs, ks, P[("g", "nfw")], gzs=0.8, lzs=2.5)
    torch.cuda.synchronize(); t6 = time.perf_counter()
    print("rep %d: ctor %.3f s | electron %.3f | hod %.3f | 6x get_power %.3f | get_power_six %.3f | C_kk+C_kg %.3f | total %.3f s"
          % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t6 - t0))
    np.testing.assert_allclose(P[("g", "electron")], p1["ge"] + p2["ge"], rtol=1e-10)
    del h, P, p1, p2
    torch.cuda.empty_cache()
