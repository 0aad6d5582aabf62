#!/usr/bin/env python
"""Table mode of the transform on a z-slab of the LARGE grid: (1) hmv_profile_tables + hmv_profile_expand against
hmv_profile_transform (must be bit-identical), (2) hmv_power_six_tab against hmv_power_six on the materialised cube,
(3) CUDA-event times of the pieces.  Run on a GPU box."""
import argparse
import json
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmvec_b200 import _capi as capi, pipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nz", type=int, default=64)
ap.add_argument("--nm", type=int, default=2000)
ap.add_argument("--nk", type=int, default=10000)
ap.add_argument("--reps", type=int, default=6)
a = ap.parse_args()
warnings.filterwarnings("ignore")
zs_all = np.linspace(0.01, 3., 200)
pick = np.linspace(0, 199, a.nz).round().astype(int)
zs = zs_all[pick]; ms = np.geomspace(2e10, 1e17, a.nm); ks = np.geomspace(1e-4, 100, a.nk)
g = pipeline.GridSix(pipeline.make_inputs(zs, ms, ks, ngal=np.geomspace(1e-3, 1e-5, 200)[pick]), tsz_tables=False)
g.upload(); g.run(); torch.cuda.synchronize()
L, d, ptr, st = capi.lib, g.d, capi.ptr, capi.stream()
nz, nm, nk, ldk = g.nz, g.nm, g.nk, g.ldk
E = lambda n: torch.empty(int(n), dtype=torch.float64, device=g.ue.device)
tab = E(L.hmv_profile_table_doubles(nz, nm, g.nxs))
ws6 = E(L.hmv_power_six_tab_ws_doubles(nz, nm, nk))


def timed(f, reps=a.reps):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def transform():
    capi.check(L.hmv_profile_transform(nz, nm, nk, ldk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["rs"]), ptr(d["cmax"]),
                                       ptr(d["xc"]), ptr(d["alpha"]), ptr(d["expo"]), ptr(d["amp"]), ptr(d["oscale"]),
                                       g.gamma, g.xmax, g.nxs, 1, ptr(d["tr_ws"]), ptr(g.ue), st), "transform")


def tables():
    capi.check(L.hmv_profile_tables(nz, nm, nk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["rs"]), ptr(d["cmax"]),
                                    ptr(d["xc"]), ptr(d["alpha"]), ptr(d["expo"]), ptr(d["amp"]), ptr(d["oscale"]),
                                    g.gamma, g.xmax, g.nxs, 1, ptr(d["tr_ws"]), ptr(tab), st), "tables")


cube2 = torch.zeros_like(g.ue)


def expand():
    capi.check(L.hmv_profile_expand(nz, nm, nk, ldk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["rs"]), g.xmax, g.nxs,
                                    ptr(d["tr_ws"]), ptr(tab), ptr(cube2), st), "expand")


p1a, p2a = torch.zeros(6, nz, nk, dtype=torch.float64, device=g.ue.device), torch.zeros(6, nz, nk, dtype=torch.float64, device=g.ue.device)
p1b, p2b = torch.zeros_like(p1a), torch.zeros_like(p2a)


def six(p1, p2):
    capi.check(L.hmv_power_six(nz, nm, nk, ldk, ptr(d["ms"]), ptr(d["ks"]), ptr(d["nzm"]), ptr(d["bh"]), ptr(d["Pzk"]),
                               g.rho_m0, float(g.p['kstar_damping']), ptr(g.um), ptr(g.ue), ptr(d["Nc"]), ptr(d["Ns"]),
                               ptr(d["NcNs"]), ptr(d["NsNsm1"]), ptr(d["ngal"]), ptr(d["pow_ws"]), 0, ptr(p1), ptr(p2), st),
               "six")


def six_tab(p1, p2):
    capi.check(L.hmv_power_six_tab(nz, nm, nk, ldk, ptr(d["ms"]), ptr(d["ks"]), ptr(d["nzm"]), ptr(d["bh"]), ptr(d["Pzk"]),
                                   g.rho_m0, float(g.p['kstar_damping']), ptr(g.um), ptr(tab), g.nxs, ptr(d["Nc"]),
                                   ptr(d["Ns"]), ptr(d["NcNs"]), ptr(d["NsNsm1"]), ptr(d["ngal"]), ptr(ws6), 0, ptr(p1),
                                   ptr(p2), st), "six_tab")


out = {"nz": nz}
transform(); tables(); expand(); torch.cuda.synchronize()
out["expand_vs_transform_max_abs"] = float((cube2[..., :nk] - g.ue[..., :nk]).abs().max())
six(p1a, p2a); six_tab(p1b, p2b); torch.cuda.synchronize()
rel = lambda x, y: float(((x - y).abs() / y.abs().clamp_min(1e-300)).max())
out["six_tab_vs_six_rel_p1"] = rel(p1b, p1a)
out["six_tab_vs_six_rel_p2"] = rel(p2b, p2a)
out["ms_transform"] = timed(transform)
out["ms_tables"] = timed(tables)
out["ms_expand"] = timed(expand)
out["ms_six"] = timed(lambda: six(p1a, p2a))
out["ms_six_tab"] = timed(lambda: six_tab(p1b, p2b))
# ---- tSZ leg: pressure tables + table-only auto spectrum against transform + cube + hmv_power ----
if g.tsz:
    import ctypes as C
    ytab = E(L.hmv_profile_table_doubles(nz, nm, g.p_nxs))
    wsy = E(L.hmv_power_ws_doubles(nz, nm))
    pgam = float(g.p['battaglia_pres_gamma'])

    def ytransform():
        capi.check(L.hmv_profile_transform(nz, nm, nk, ldk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["y_rs"]),
                                           ptr(d["y_cmax"]), ptr(d["y_xc"]), ptr(d["y_alpha"]), ptr(d["y_expo"]),
                                           ptr(d["y_amp"]), ptr(d["y_oscale"]), pgam, g.p_xmax, g.p_nxs, 0,
                                           ptr(d["tr_ws"]), ptr(g.uy), st), "ytransform")

    def ytables():
        capi.check(L.hmv_profile_tables(nz, nm, nk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["y_rs"]), ptr(d["y_cmax"]),
                                        ptr(d["y_xc"]), ptr(d["y_alpha"]), ptr(d["y_expo"]), ptr(d["y_amp"]),
                                        ptr(d["y_oscale"]), pgam, g.p_xmax, g.p_nxs, 0, ptr(d["tr_ws"]), ptr(ytab), st),
                   "ytables")

    py1, py2 = torch.zeros(nz, nk, dtype=torch.float64, device=g.ue.device), torch.zeros(nz, nk, dtype=torch.float64, device=g.ue.device)
    qy1, qy2 = torch.zeros_like(py1), torch.zeros_like(py2)

    def yy_cube():
        capi.check(L.hmv_power(nz, nm, nk, ldk, ptr(d["ms"]), ptr(d["ks"]), ptr(d["nzm"]), ptr(d["bh"]), ptr(d["Pzk"]),
                               g.rho_m0, float(g.p['kstar_damping']), C.byref(g.ty), C.byref(g.ty), ptr(d["pair_ws"]),
                               ptr(py1), ptr(py2), st), "yy_cube")

    def yy_tab():
        capi.check(L.hmv_power_tab(nz, nm, nk, ptr(d["ms"]), ptr(d["ks"]), ptr(d["nzm"]), ptr(d["bh"]), ptr(d["Pzk"]),
                                   g.rho_m0, float(g.p['kstar_damping']), 2, ptr(ytab), g.p_nxs, ptr(wsy), ptr(qy1),
                                   ptr(qy2), st), "yy_tab")

    ytransform(); ytables(); yy_cube(); yy_tab(); torch.cuda.synchronize()
    out["yy_tab_vs_cube_rel_p1"] = rel(qy1, py1)
    out["yy_tab_vs_cube_rel_p2"] = rel(qy2, py2)
    out["ms_ytransform"] = timed(ytransform)
    out["ms_ytables"] = timed(ytables)
    out["ms_yy_cube"] = timed(yy_cube)
    out["ms_yy_tab"] = timed(yy_tab)
print("TABAB " + json.dumps(out))
