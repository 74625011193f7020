#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tc.sh
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_tc_launches.csv python tools/prof_tc.py 200 10 1 > /dev/null 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_tc_launches.csv")))
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
ncu --set full --clock-control none --import-source on -k regex:ptm_tc5_kernel -c 1 -o gpurun_out/r02_tc5_kernel python tools/prof_tc.py 100 10 1 > gpurun_out/r02_ncu_full.log 2>&1; tail -2 gpurun_out/r02_ncu_full.log
