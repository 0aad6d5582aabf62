#!/usr/bin/env python
"""Per-source-line warp-stall samples from an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep [kernel-regex] [top]
Prints, per captured kernel, the source lines with the most stall samples and their executed-instruction counts."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if rx:
        cmd += ["--kernel-name", "regex:" + rx]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernels, hdr, name = [], None, "?"
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            name = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            kernels.append((name, hdr, {}))
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit():
            continue        # SASS rows have an empty line number; the CUDA-line rows carry the per-line totals
        try:
            smp = int(r[hdr.index("# Samples")] or 0)
            ins = int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError:
            continue
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and r[i].strip().isdigit() and int(r[i]):
                st[h[6:]] = int(r[i])
        kernels[-1][2][(int(r[0]), r[1].strip()[:90])] = (smp, ins, st)
    for name, hdr, agg in kernels:
        if not agg:
            continue
        tot = sum(v[0] for v in agg.values()) or 1
        toti = sum(v[1] for v in agg.values()) or 1
        print("=== %s   (samples %d, warp-instructions %d)" % (name[:90], tot, toti))
        for (ln, src), (smp, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            why = " ".join("%s:%d%%" % (k, round(100.0 * v / max(1, sum(st.values()))))
                           for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print("  %5.1f%% smp %5.1f%% inst  L%-4d %-90s | %s" % (100.0 * smp / tot, 100.0 * ins / toti, ln, src, why))


if __name__ == "__main__":
    main()
