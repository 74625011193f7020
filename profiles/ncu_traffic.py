#!/usr/bin/env python
"""Record the DRAM traffic of a kernel from an `ncu --set full` capture for bench.py's `roofline.traffic`:
    python profiles/ncu_traffic.py <report.ncu-rep> <kernel substring> <bench kernel name> <workload string>
adds  "<bench kernel name>|<workload>": dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch)  to
profiles/ncu_traffic.json.  Read here (no GPU needed); the capture itself comes from the GPU box."""
import csv
import io
import json
import os
import subprocess
import sys

rep, sub, name, workload = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tot = None
for row in rows[2:]:
    if sub in row[ik]:
        tot = float(row[ir]) * scale[units[ir]] + float(row[iw]) * scale[units[iw]]
        break
assert tot is not None, "kernel not in the report"
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[name + "|" + workload] = tot
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
print(name + "|" + workload, "%.4g bytes per launch" % tot)
