#!/usr/bin/env python
"""Kernel-level A/B bench on a z-slab of the LARGE grid: times hmv_profile_transform (K1), hmv_uk_nfw (K2) and
hmv_power_six (K5) alone with CUDA events, for one or several builds of the library (same C ABI), and compares the
cubes each build writes with those of the first one.

    python tools/kbench.py --nz 25 --libs hmvec_b200/libhmvec_b200.so,gpurun_out/lib_v2.so [--reps 10] [--pres]

Each build runs in its own process (HMV_LIB selects the library).  Run on a GPU box."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(a):
    import torch
    from hmvec_b200 import _capi as capi, pipeline
    zs_all = np.linspace(0.01, 3., 200)
    pick = np.linspace(0, 199, a.nz).round().astype(int)          # spread over the whole redshift range
    zs = zs_all[pick]
    ms = np.geomspace(2e10, 1e17, a.nm)
    ks = np.geomspace(1e-4, 100, a.nk)
    inp = pipeline.make_inputs(zs, ms, ks, ngal=np.geomspace(1e-3, 1e-5, 200)[pick])
    g = pipeline.GridSix(inp)
    g.upload()
    g.run()
    torch.cuda.synchronize()
    L, d, ptr, st = capi.lib, g.d, capi.ptr, capi.stream()
    nz, nm, nk, ldk = g.nz, g.nm, g.nk, g.ldk

    def k1():
        capi.check(L.hmv_profile_transform(nz, nm, nk, ldk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["rs"]),
                                           ptr(d["cmax"]), ptr(d["xc"]), ptr(d["alpha"]), ptr(d["expo"]), ptr(d["amp"]),
                                           ptr(d["oscale"]), g.gamma, g.xmax, g.nxs, 1, ptr(d["tr_ws"]), ptr(g.ue), st),
                   "hmv_profile_transform")

    def k2():
        capi.check(L.hmv_uk_nfw(nz, nm, nk, ldk, ptr(d["zs"]), ptr(d["ks"]), g.kmax, ptr(d["cs"]), ptr(d["rvir"]),
                                ptr(d["nfw_ws"]), ptr(g.um), st), "hmv_uk_nfw")

    def k5():
        capi.check(L.hmv_power_six(nz, nm, nk, ldk, ptr(d["ms"]), ptr(d["ks"]), ptr(d["nzm"]), ptr(d["bh"]),
                                   ptr(d["Pzk"]), g.rho_m0, float(g.p['kstar_damping']), ptr(g.um), ptr(g.ue),
                                   ptr(d["Nc"]), ptr(d["Ns"]), ptr(d["NcNs"]), ptr(d["NsNsm1"]), ptr(d["ngal"]),
                                   ptr(d["pow_ws"]), nz * nk, ptr(g.p1), ptr(g.p2), st), "hmv_power_six")

    out = {"lib": os.environ.get("HMV_LIB", "default"), "nz": nz}
    for name, fn in (("k1", k1), ("k2", k2), ("k5", k5)):
        if name not in a.kernels:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name + "_ms"] = float(np.median(ts))
        out[name + "_min"] = float(np.min(ts))
    # samples of the cubes for the parity comparison between builds
    rows = slice(None, None, 37)
    np.savez(a.dump, ue=g.ue[:, rows, :nk].cpu().numpy(), um=g.um[:, rows, :nk].cpu().numpy(),
             p1=g.p1.cpu().numpy(), p2=g.p2.cpu().numpy())
    print("KBENCH " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nz", type=int, default=25)
    ap.add_argument("--nm", type=int, default=2000)
    ap.add_argument("--nk", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--kernels", default="k1,k2,k5")
    ap.add_argument("--libs", default="")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--dump", default="/tmp/kbench_dump.npz")
    a = ap.parse_args()
    a.kernels = a.kernels.split(",")
    if a.child:
        return child(a)
    libs = [l for l in a.libs.split(",") if l] or [os.path.join(ROOT, "hmvec_b200", "libhmvec_b200.so")]
    ref = None
    for i, lib in enumerate(libs):
        env = dict(os.environ, HMV_LIB=os.path.abspath(lib))
        dump = "/tmp/kbench_dump_%d.npz" % i
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--nz", str(a.nz), "--nm", str(a.nm), "--nk", str(a.nk),
               "--reps", str(a.reps), "--kernels", ",".join(a.kernels), "--dump", dump]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("KBENCH ")]
        if r.returncode != 0 or not line:
            print("KBENCH-FAIL %s rc=%d\n%s" % (lib, r.returncode, (r.stdout + r.stderr)[-2000:]), flush=True)
            continue
        res = json.loads(line[0][7:])
        cur = dict(np.load(dump))
        if ref is None:
            ref = cur
        else:
            for k in ("ue", "um", "p1", "p2"):
                sc = np.max(np.abs(ref[k]))
                res["dmax_" + k] = float(np.max(np.abs(cur[k] - ref[k])) / sc)
            with np.errstate(all="ignore"):
                for k in ("p1", "p2"):
                    rel = np.abs(cur[k] - ref[k]) / np.abs(ref[k])
                    res["rel_" + k] = float(np.nanmax(np.where(np.isfinite(rel), rel, 0.0)))
        print("KBENCH " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
