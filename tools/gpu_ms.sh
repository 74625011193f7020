#!/bin/bash
mkdir -p gpurun_out
python tools/prof_tc.py 512 5 2 cont 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ms_launches.csv python tools/prof_tc.py 128 5 1 cont > /dev/null 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_ms_launches.csv")))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); k = d["Kernel Name"][:60]
        try: v = float(d["Metric Value"].replace(",", ""))
        except ValueError: continue
        u = d.get("Metric Unit", "")
        v = v / 1e6 if u in ("nsecond", "ns") else (v / 1e3 if u in ("usecond", "us") else v)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, ms) in agg.items(): print("%-60s %4d launches %10.3f ms" % (k, n, ms))
P
ncu --set full --clock-control none --import-source on -k regex:ms_dist_tile -c 1 -o gpurun_out/r02_ms_tile python tools/prof_tc.py 128 5 1 cont > /dev/null 2>&1
