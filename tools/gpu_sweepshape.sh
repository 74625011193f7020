#!/bin/bash
mkdir -p gpurun_out
for sh in 1 2 3; do
PSB_SWEEP_SHAPE=$sh timeout 600 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/r02_bench_shape$sh.json 2> gpurun_out/r02_bench_shape$sh.err
python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_shape$sh.json").read().strip().splitlines()[-1])
    print("shape $sh", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "sweep", d["search_viterbi"]["ms"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_shape$sh.err").read()[-400:])
P
done
