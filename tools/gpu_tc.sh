#!/bin/bash
# tensor-core filter path: parity tests, then the bench with the old (5) and new (6) top-N paths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tensor_core or topn_kernel_variants or ptm_batch" -s > gpurun_out/r02_tc_test.log 2>&1
echo "tc tests exit $?: $(tail -n 15 gpurun_out/r02_tc_test.log)"
for v in 6; do
    PSB_TOPN_VARIANT=$v timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_tc$v.json 2> gpurun_out/r02_bench_tc$v.err
    echo "bench variant $v exit $?"; tail -c 400 gpurun_out/r02_bench_tc$v.err
    python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_tc$v.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["kernel_ms_unpipelined"])
except Exception as e:
    print("unreadable", e)
P
done
