#!/bin/bash
# two threads per frame in the tcgen05 filter: parity tests, then the headline bench with both variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tensor_core or topn_kernel_variants or ptm_batch or full_size" > gpurun_out/r02_split.log 2>&1; echo "tests exit $?: $(tail -n 3 gpurun_out/r02_split.log)"
for sp in 2 1; do
PSB_TC_SPLIT=$sp timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_split$sp.json 2> gpurun_out/r02_bench_split$sp.err; echo "bench split $sp exit $?"
python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_split$sp.json").read().strip().splitlines()[-1])
    print("split $sp", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["kernel_ms_unpipelined"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_split$sp.err").read()[-600:])
P
done
