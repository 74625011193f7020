#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 2 --warmup 3 --cpu-budget 1 > gpurun_out/r02_bench_chk.json 2> gpurun_out/r02_bench_chk.err; echo "bench exit $?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_chk.json").read().strip().splitlines()[-1])
    print("value %.4g ms %.2f" % (d["value"], d["ms_per_step"]), json.dumps(d.get("search_viterbi_beam")))
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_chk.err").read()[-800:])
P
