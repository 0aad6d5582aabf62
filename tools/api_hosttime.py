#!/usr/bin/env python
"""Inclusive host time per method of the drop-in API's step (perf_counter wrappers around every method of HaloModel,
Cosmology and ZComm-free helpers; no profiler overhead), averaged over `reps` free-running steps of the bench's
workflow on an nz-redshift slab.  Shows what is left on the host once the device work of a small slab is short.

    python tools/api_hosttime.py 25 [reps]
"""
import contextlib
import functools
import io
import os
import sys
import time
import types
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hmvec_b200 as hm  # noqa: E402
from hmvec_b200 import cosmology, hmvec, _capi  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
zs = np.linspace(0.01, 3., nz); ms = np.geomspace(2e10, 1e17, 2000); ks = np.geomspace(1e-4, 100, 10000)
ells = np.geomspace(10, 1e4, 1000)
ngal = np.geomspace(1e-3, 1e-5, nz)
PAIRS = (("nfw", "nfw"), ("electron", "electron"), ("nfw", "electron"), ("g", "g"), ("g", "nfw"), ("g", "electron"), ("y", "y"))
acc = {}


def wrap(owner, name, fn, label):
    @functools.wraps(fn)
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            e = acc.setdefault(label, [0.0, 0])
            e[0] += time.perf_counter() - t0
            e[1] += 1
    setattr(owner, name, w)


for cls in (hmvec.HaloModel, cosmology.Cosmology, hmvec.DeviceCubes):
    for name, fn in list(vars(cls).items()):
        if isinstance(fn, types.FunctionType):
            wrap(cls, name, fn, cls.__name__ + "." + name)
        elif isinstance(fn, staticmethod):
            f = fn.__func__
            wrap(cls, name, f, cls.__name__ + "." + name)
            setattr(cls, name, staticmethod(getattr(cls, name)))
for name in ("check", "ptr", "stream"):
    wrap(_capi, name, getattr(_capi, name), "_capi." + name)


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
acc.clear()
t0 = time.perf_counter()
for _ in range(reps):
    step()
tot = (time.perf_counter() - t0) / reps * 1e3
print("step %.2f ms wall (nz=%d, %d reps); inclusive host ms per step, calls per step" % (tot, nz, reps))
for k, (t, n) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%8.3f  %5.1f  %s" % (t / reps * 1e3, n / reps, k))
