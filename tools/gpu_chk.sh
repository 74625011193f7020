#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tensor_core or topn_kernel_variants or ptm_batch or hmmset or full_size" > gpurun_out/r02_chk.log 2>&1; echo "tests exit $?: $(tail -n 3 gpurun_out/r02_chk.log)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; echo "bench exit $?"
timeout 600 python bench.py --model semi --utts 512 --secs 5 --steps 3 --warmup 3 --search fwdtree > gpurun_out/r02_bench_semi.json 2> gpurun_out/r02_bench_semi.err; echo "bench semi exit $?"
for f in final2 semi; do python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["kernel_ms_unpipelined"], (d.get("search_viterbi") or {}).get("ms"), json.dumps(d.get("search_coupled"))[:300])
except Exception as e:
    print("$f unreadable", e); print(open("gpurun_out/r02_bench_$f.err").read()[-600:])
P
done
