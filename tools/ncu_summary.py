#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof_r1a.ncu-rep > profiles/r1_full_top3.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(path):
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    agg = collections.OrderedDict()
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    byts = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows:
        n = r['Kernel Name'].split('(')[0][:80]
        a = agg.setdefault(n, [0, 0.0, 0.0])
        if r['Metric Name'] == 'gpu__time_duration.sum':
            a[0] += 1
            a[1] += float(r['Metric Value']) * scale.get(r['Metric Unit'], 1e-6)
        elif r['Metric Name'] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            a[2] += float(r['Metric Value']) * byts.get(r['Metric Unit'], 1.0)
    tot = sum(v[1] for v in agg.values())
    print("# per-launch device time from: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised)")
    print("# source: %s ; %d launches, %.3f ms total" % (path, sum(v[0] for v in agg.values()), tot))
    print("%-82s %5s %11s %10s %7s %12s" % ("kernel", "n", "total_ms", "avg_ms", "share", "dram_GB/launch"))
    for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-82s %5d %11.3f %10.3f %6.1f%% %12.3f" % (n, c, t, t / max(c, 1), 100 * t / tot, b / max(c, 1) / 1e9))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none --import-source on ; source: %s" % path)
    for r in rows[2:]:
        print("-" * 100)
        print("kernel: %s   grid %s block %s" % (r[hdr.index('Kernel Name')], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-72s %s %s" % (k, r[i], units[i]))
        rd, wr = float(r[hdr.index('dram__bytes_read.sum')]), float(r[hdr.index('dram__bytes_write.sum')])
        print("  %-72s %.6f %s" % ("traffic = dram read + write", rd + wr, units[hdr.index('dram__bytes_read.sum')]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
