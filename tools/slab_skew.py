#!/usr/bin/env python
"""How uneven are the z-slabs of an N-way sharded run?  Times every rank's slab of the LARGE grid alone on ONE GPU
(the launch sequence without the gather + Limber stage) for the contiguous partition (zshard.slab_bounds) and for a
round-robin one (rank r owns z_r, z_{r+N}, ...).  The step of an N-GPU run is the slowest rank's.

    python tools/slab_skew.py --world 8 [--steps 10]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--nz", type=int, default=200)
    ap.add_argument("--ranks", default="", help="comma-separated ranks to time (default: all)")
    ap.add_argument("--partitions", default="contiguous,round_robin")
    a = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    import torch
    from hmvec_b200 import pipeline, zshard
    zs = np.linspace(0.01, 3., a.nz)
    ms = np.geomspace(2e10, 1e17, 2000)
    ks = np.geomspace(1e-4, 100, 10000)
    ngal = np.geomspace(1e-3, 1e-5, a.nz)
    inp = pipeline.make_inputs(zs, ms, ks, ngal=ngal)
    b = zshard.slab_bounds(a.nz, a.world)
    out = {}
    ranks = [int(x) for x in a.ranks.split(",") if x] or list(range(a.world))
    for name in a.partitions.split(","):
        rows = []
        for r in ranks:
            idx = np.arange(b[r], b[r + 1]) if name == "contiguous" else np.arange(r, a.nz, a.world)
            g = pipeline.GridSix(pipeline.slab_inputs(inp, idx))
            g.upload()
            for _ in range(3):
                g.run()
            torch.cuda.synchronize()
            nst = len(g.STAGES) + 1
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(nst)] for _ in range(a.steps)]
            for s in range(a.steps):
                g.run(events=evs[s])
            torch.cuda.synchronize()
            st = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(nst - 1)] for e in evs]).mean(axis=0)
            tot = float(np.mean([e[0].elapsed_time(e[-1]) for e in evs]))
            # the launch sequence with the side streams: level 1 = sigma^2/n(M)/HOD leg, 2 = + electron and tSZ legs;
            # interleaved repetitions so that clock / power drift hits every level alike
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lv = {0: [], 1: [], 2: []}
            for rep in range(3):
                for level in (0, 1, 2):
                    g.overlap = level
                    for _ in range(2):
                        g.run()
                    torch.cuda.synchronize()
                    e0.record()
                    for s in range(a.steps):
                        g.run()
                    e1.record()
                    torch.cuda.synchronize()
                    lv[level].append(e0.elapsed_time(e1) / a.steps)
            rows.append({"rank": r, "ms": tot, **{"ms_level%d" % k: round(float(np.median(v)), 4) for k, v in lv.items()},
                         **{n: round(float(v), 3) for n, v in zip(g.STAGES, st)}})
            print(name, json.dumps(rows[-1]), flush=True)
            del g
            torch.cuda.empty_cache()
        t = np.array([x["ms"] for x in rows])
        out[name] = {"max_ms": float(t.max()), "mean_ms": float(t.mean()), "min_ms": float(t.min())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
