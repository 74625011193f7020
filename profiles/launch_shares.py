#!/usr/bin/env python
"""Per-kernel launch counts, total time and share of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
    python profiles/launch_shares.py gpurun_out/r02_launches.csv > profiles/r02_launch_shares.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = d.get("Metric Unit", "")
        v = v / 1e6 if u in ("nsecond", "ns") else (v / 1e3 if u in ("usecond", "us") else v)
        a = agg.setdefault(d["Kernel Name"].split("(")[0][-70:], [0, 0.0])
        a[0] += 1
        a[1] += v
tot = sum(ms for _, ms in agg.values())
print("ncu launch list (cold caches, serialised: shares, not absolute times): %d kernels, %.2f ms" % (sum(n for n, _ in agg.values()), tot))
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %5d launches %10.3f ms %6.2f %%" % (k, n, ms, 100 * ms / tot))
