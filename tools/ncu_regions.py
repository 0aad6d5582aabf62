#!/usr/bin/env python
"""Stall samples of a captured kernel summed over source-line regions (needs -lineinfo + --import-source on).

    python tools/ncu_regions.py rep.ncu-rep kernel-regex nsections section:lo-hi=name [...]   (section = index of the source file in the report; the first nsections sections = one kernel instance)
Lines outside every region go to 'other'.  Only the first matching kernel instance is used."""
import collections, csv, io, subprocess, sys

def main():
    rep, rx = sys.argv[1], sys.argv[2]
    regions = []
    nsec = int(sys.argv.pop(3))
    for a in sys.argv[3:]:
        spec, name = a.split("=")
        f, rng = spec.split(":")
        lo, hi = rng.split("-")
        regions.append((f, int(lo), int(hi), name))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, sec, first, names = None, -1, False, {}
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r in rows:
        if not r: continue
        if r[0] == "Line No":
            hdr = r; sec += 1; first = True; continue
        if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit(): continue
        try:
            smp = int(r[hdr.index("# Samples")] or 0); ins = int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError: continue
        ln = int(r[0]); name = "other"
        if first:
            first = False; names[sec] = r[1].strip()[:60]
            if sec >= nsec: break
        for f, lo, hi, n in regions:
            if int(f) == sec and lo <= ln <= hi: name = n; break
        a = agg[name]; a[0] += smp; a[1] += ins
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and r[i].strip().isdigit(): a[2][h[6:]] += int(r[i])
    print("sections:", names)
    tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
    for n, (s, i, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        ss = sum(st.values()) or 1
        print("%-14s %5.1f%% smp %5.1f%% inst | %s" % (n, 100.0*s/tot, 100.0*i/toti,
              " ".join("%s:%d%%" % (k, round(100.0*v/ss)) for k, v in st.most_common(5))))
if __name__ == "__main__": main()
