#!/bin/bash
# device-resident step with the batch cut into n utterance ranges on their own streams (PSB_PIPELINE)
mkdir -p gpurun_out
for n in ${PIPES:-2 4 8}; do
PSB_PIPELINE=$n timeout 600 python bench.py --steps 5 --warmup 3 --cpu-budget 1 > gpurun_out/r02_bench_pipe$n.json 2> gpurun_out/r02_bench_pipe$n.err; echo "bench pipe $n exit $?"
python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_pipe$n.json").read().strip().splitlines()[-1])
    print("pipe $n", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["kernel_ms_unpipelined"], d.get("gpu_launches"))
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_pipe$n.err").read()[-600:])
P
done
