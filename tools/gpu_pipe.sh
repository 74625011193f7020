#!/bin/bash
mkdir -p gpurun_out
for p in 2 3; do
PSB_PIPELINE=$p timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_pipe$p.json 2> gpurun_out/r02_bench_pipe$p.err
python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_pipe$p.json").read().strip().splitlines()[-1])
    print("pipe $p", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_pipe$p.err").read()[-400:])
P
done
