#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "hmmset or hmm" > gpurun_out/r02_hs.log 2>&1; echo "tests exit $?: $(tail -n 3 gpurun_out/r02_hs.log)"
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-budget 1 > gpurun_out/r02_bench_hs.json 2> gpurun_out/r02_bench_hs.err
python - <<P
import json
d = json.loads(open("gpurun_out/r02_bench_hs.json").read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], json.dumps(d["viterbi_stage"])[:400])
P
