#!/usr/bin/env python
"""bench.py -- P(k,z) grid points per second for the hmvec hot path on B200 (BASELINE.json metric).

A step = one pass of the whole path over the LARGE synthetic grid (BASELINE.json configs[3]+[4]):
    zs=linspace(0.01,3,200), ms=geomspace(2e10,1e17,2000), ks=geomspace(1e-4,100,10000), EH98 linear power;
    sigma^2 -> n(M,z), b -> u_NFW cube -> Battaglia-AGN electron cube (xmax=20, nxs=5000) -> ngal-solved HOD ->
    {mm,ee,me,gg,gm,ge} 1h+2h spectra -> [all-gather over z] -> Limber C_kk, C_kg at 1000 ells.
One "grid point" = one (z,k) of one spectrum with both its 1-halo and 2-halo terms: 6*nz*nk points per step.
With N GPUs the z axis is sharded (nz/N redshifts per rank, total work fixed -> "strong" scaling).

  value : device-resident inputs, CUDA events around K steps, max over ranks
  e2e   : the same step through pinned HOST buffers (H2D of the linear power + background, D2H of the 12 spectra and
          the two C_ell), copies inside the timed region
  --impl reference : the reference algorithm on the host cores (oracle port of the numpy path, one z-slab per worker)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "P(k,z) grid pts/s, Battaglia+HOD 1h+2h"
UNIT = "pts/s"


def grids(a):
    zs = np.linspace(0.01, 3., a.nz)
    ms = np.geomspace(2e10, 1e17, a.nm)
    ks = np.geomspace(1e-4, 100, a.nk)
    ells = np.geomspace(10, 1e4, a.nl)
    return zs, ms, ks, ells


def workload_name(a):
    return "C4+C5: zs=%d, ms=%d (2e10-1e17), ks=%d (1e-4-100), Battaglia AGN electron (xmax=20,nxs=5000) + ngal-HOD, " \
           "six spectra 1h+2h, Limber C_kk/C_kg at %d ells" % (a.nz, a.nm, a.nk, a.nl)


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy path, one z-slab per worker process
# ---------------------------------------------------------------------------------------------------------
def _oracle_slab(job):
    zs, ms, ks, ngal = job
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import hmvec_oracle as orc
    t = time.perf_counter()
    o = orc.OracleHaloModel(zs, ks, ms)
    o.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
    o.add_hod("g", ngal=ngal)
    chk = 0.0
    for a, b in (("nfw", "nfw"), ("electron", "electron"), ("nfw", "electron"), ("g", "g"), ("g", "nfw"),
                 ("g", "electron")):
        chk += float(np.sum(o.get_power_1halo(a, b)[:, ::97])) + float(np.sum(o.get_power_2halo(a, b)[:, ::97]))
    return time.perf_counter() - t, chk


def cpu_workers():
    n = os.cpu_count() or 1
    try:
        import psutil
        n = min(n, max(1, int(psutil.virtual_memory().available / 4e9)))   # ~3 GB peak RSS per 1-z slab
    except Exception:
        pass
    return max(1, min(n, 64))


def cpu_sample(a, nworkers, zper=1):
    """Time `nworkers` concurrent slabs of `zper` redshifts each at full M,k resolution; returns (pts/s, seconds)."""
    zs, ms, ks, _ = grids(a)
    ngal = np.geomspace(1e-3, 1e-5, zs.size)
    pick = np.linspace(0, zs.size - 1, nworkers * zper).round().astype(int)
    jobs = [(zs[pick[i * zper:(i + 1) * zper]], ms, ks, ngal[pick[i * zper:(i + 1) * zper]]) for i in range(nworkers)]
    t = time.perf_counter()
    if nworkers == 1:
        _oracle_slab(jobs[0])
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(nworkers) as pool:
            pool.map(_oracle_slab, jobs)
    dt = time.perf_counter() - t
    return 6.0 * nworkers * zper * ks.size / dt, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = cpu_workers()
    for _ in range(a.warmup if a.warmup < 1 else 1):      # one warm-up step is enough to page numpy/scipy in
        cpu_sample(a, P)
    vals, secs = [], []
    for _ in range(a.steps):
        v, dt = cpu_sample(a, P)
        vals.append(v)
        secs.append(dt)
    total_pts = 6.0 * P * a.nk * a.steps
    value = total_pts / sum(secs)
    sample = "%d concurrent 1-redshift slabs (of %d z) at full %d M x %d k resolution per step, whole path except " \
             "Limber" % (P, a.nz, a.nm, a.nk)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * sum(secs) / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "parallelism": "host processes x%d" % P},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """Polls SM clock and throttle reasons through NVML from a thread while the timed region runs."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, index, period=0.01):
        import threading
        self.sm, self.bits, self.mx, self.power = [], 0, None, []
        self._stop = threading.Event()
        self._ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._ok = True
        except Exception:
            return
        self.period = period
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0}
        if not self._ok:
            return out
        self._stop.set()
        self.t.join(timeout=2)
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), samples=len(self.sm),
                       power_w_max=float(np.max(self.power)) if self.power else None)
        out["reasons"] = [n for n, b in self.REASONS if self.bits & b]
        return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from hmvec_b200 import _capi as capi, pipeline, zshard

    zs, ms, ks, ells = grids(a)
    inp = pipeline.make_inputs(zs, ms, ks, ells=ells)
    zc = zshard.ZComm(a.nz, None) if world > 1 else None
    sl = zc.slab if zc is not None else slice(0, a.nz)
    g = pipeline.GridSix(pipeline.slab_inputs(inp, sl), device=dev, zcomm=zc, nz_total_zs=zs,
                         fused_nfw=a.fused_nfw)
    capi.check(capi.lib.hmv_set_transform_mode(a.transform_mode), "hmv_set_transform_mode")
    g.transform_mode = a.transform_mode
    g.upload()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms_):
        if world == 1:
            return ms_
        t = torch.tensor([ms_], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(a.warmup, 3)):
        g.run()
    barrier()

    # ---- device-resident timing, stage boundaries marked with events on the launching stream -------------------
    nst = len(g.STAGES) + 1
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(nst)] for _ in range(a.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0.record()
    for s in range(a.steps):
        g.run(events=evs[s])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler is not None else None
    stage_ms = np.array([[evs[s][i].elapsed_time(evs[s][i + 1]) for i in range(nst - 1)] for s in range(a.steps)])
    stage_ms = stage_ms.mean(axis=0)

    # ---- end to end through pinned host buffers -----------------------------------------------------------------
    for _ in range(2):
        g.upload(); g.run(overlap_d2h=True); g.finish_e2e()
    barrier()
    e0.record()
    for s in range(a.steps):
        g.upload()                   # pinned host -> HBM: linear power, background, ngal targets
        g.run(overlap_d2h=True)      # spectra go back chunk by chunk while later z-chunks are still being reduced
        g.finish_e2e()               # the step's results are on the host before the next one starts
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pts_per_step = 6.0 * a.nz * a.nk
    value = pts_per_step * a.steps / (ms_total * 1e-3)
    e2e = pts_per_step * a.steps / (ms_e2e * 1e-3)

    # ---- roofline per stage (algorithmic bytes per launch of the stage's main kernel; this rank's slab) ------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    nzl, nm, nk = g.nz, g.nm, g.nk
    alg_bytes = {
        "uk_nfw": 8.0 * nzl * nm * nk,                                  # store of the cube (K2; FP64-pipe bound)
        "uk_electron": 8.0 * nzl * nm * nk,                             # store of the cube (K1; FP64-pipe bound)
        # two-cube kernel: read 2 cubes once, write 12 spectra, read Pzk.  Fused kernel: one cube + 448 B of per-halo
        # coefficient/NFW records per (z,M) for each of the ceil(nk/512) k tiles
        "power_six": (nzl * nk * (16.0 * nm + 96.0 + 8.0) if not g.fused_nfw else
                      nzl * nk * (8.0 * nm + 96.0 + 8.0) + 448.0 * nzl * nm * ((g.ldk + 511) // 512)),
        "sigma2": 8.0 * (nzl * g.nks + 2.0 * g.nks * nm + nzl * nm),    # sPzk + W2 table write/read + sigma2 out
    }
    if g.fused_nfw:
        del alg_bytes["uk_nfw"]          # no NFW cube in the spectra-only fusion
    kernels = {}
    for name, ms_ in zip(g.STAGES, stage_ms):
        k = {"ms": float(ms_)}
        if name in alg_bytes:
            gbs = alg_bytes[name] / (ms_ * 1e-3) / 1e9
            k.update(alg_bytes=alg_bytes[name], gbs=gbs, frac_hbm=gbs / hbm_peak)
        kernels[name] = k
    # K1 runs on the FP64 tensor cores: algorithmic flops = 2 * sum over halos of (samples inside the theta-cut) x
    # (bins the k-range needs), from the per-halo parameters the step just produced
    fp64_tf = float(capi.lib.hmv_bench_dfma(20000, capi.stream()))
    dmma_tf = float(capi.lib.hmv_bench_dmma(2000, capi.stream()))
    dx = g.xmax / g.nxs
    kt1 = 2.0 * np.pi / (g.nxs * ((g.xmax - dx) / g.nxs))
    rs_h, cmax_h = g.d["rs"].cpu().numpy(), g.d["cmax"].cpu().numpy()
    zs_h = g.d["zs"].cpu().numpy()
    ncut = np.minimum(g.nxs, np.floor(cmax_h / dx))
    jneed = np.minimum(g.nxs // 2, np.floor(g.kmax * rs_h * (1.0 + zs_h[:, None]) / kt1) + 1.0)
    k1_flop = float(2.0 * np.sum(ncut * jneed))
    k1 = kernels["uk_electron"]
    k1.update(alg_flop=k1_flop, tflops=k1_flop / (k1["ms"] * 1e-3) / 1e12)
    k1["frac_fp64_tensor"] = k1["tflops"] / dmma_tf
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)     # ncu dram__bytes_read+write per launch of each stage's main kernel (full grid)
    except Exception:
        pass
    dom = max(alg_bytes, key=lambda n: kernels[n]["ms"])
    if dom == "uk_electron":
        roof = {"kernel": "%s (K1, FP64 mma.sync m8n8k4)" % ("profile_transform_ws_kernel" if g.transform_mode == 0 else "profile_transform_kernel"), "bound": "tensor",
                "achieved": k1["tflops"], "peak": dmma_tf, "unit": "TFLOP/s", "frac": k1["frac_fp64_tensor"],
                "peak_source": "FP64 DMMA peak measured in this run (hmv_bench_dmma); MEASURED_PEAKS.json has no FP64 figure",
                "alg_flop_per_launch": k1_flop, "hbm_store_frac": k1["frac_hbm"]}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_hbm"], "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes[dom]}
    roof.update(ms_per_launch=kernels[dom]["ms"], traffic=traffic.get(dom) if world == 1 else None,
                fp64_dfma_peak_tflops_measured=fp64_tf, fp64_dmma_peak_tflops_measured=dmma_tf)
    # the HBM-bound kernel of the path, for reference beside the dominant one
    roof["hbm_kernel"] = {"kernel": "power_six_nfw_kernel (K5+K2 fused)" if g.fused_nfw else "power_six_kernel (K5)", "achieved": kernels["power_six"]["gbs"], "peak": hbm_peak,
                          "unit": "GB/s", "frac": kernels["power_six"]["frac_hbm"],
                          "traffic": traffic.get("power_six") if world == 1 else None}

    wl_bytes = 8.0 * a.nz * a.nm * a.nk * 4 + 96.0 * a.nz * a.nk      # whole C4 workload, 1.28e11 B
    roof["workload"] = {"alg_bytes_per_step": wl_bytes, "achieved": wl_bytes / (ms_total / a.steps * 1e-3) / 1e9 / world,
                        "peak": hbm_peak, "unit": "GB/s per GPU",
                        "frac": wl_bytes / (ms_total / a.steps * 1e-3) / 1e9 / world / hbm_peak,
                        "note": "SURVEY 8(d) algorithmic traffic of the whole step (2 cubes written + read once, 12 "
                                "spectra) over the step time"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": "z-sharded x%d" % world,
                       "nfw": "evaluated inside the mass reduction (spectra-only fusion)" if g.fused_nfw else
                              "cube materialised in HBM",
                       "l2": "inputs exceed L2 (two %.1f GB cubes per rank)" % (8e-9 * nzl * nm * g.ldk)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": g.h2d_bytes() * world,
                    "d2h_bytes_per_step": g.d2h_bytes() * world, "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(g.launches_per_run * a.steps * world), "clocks": clocks, "roofline": roof, "kernels": kernels}

    if world == 1 and not a.no_cpu:
        v, dt = cpu_sample(a, 1, zper=a.cpu_nz)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
                                "sample": "%d of %d redshifts at full %d M x %d k resolution, whole path except "
                                          "Limber, single numpy thread" % (a.cpu_nz, a.nz, a.nm, a.nk)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nz", type=int, default=200)
    ap.add_argument("--nm", type=int, default=2000)
    ap.add_argument("--nk", type=int, default=10000)
    ap.add_argument("--nl", type=int, default=1000)
    ap.add_argument("--cpu-nz", type=int, default=2, help="redshifts in the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--transform-mode", type=int, default=0, choices=[0, 1],
                    help="K1 launch plan: 0 = persistent warp-specialised kernel (default), 1 = bin-count-class kernels")
    ap.add_argument("--fused-nfw", action="store_true",
                    help="evaluate the NFW profile inside the mass reduction (hmv_power_six_nfw) instead of writing "
                         "its cube to HBM and reading it back (hmv_uk_nfw + hmv_power_six, the faster default)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
