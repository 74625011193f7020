#!/bin/bash
# continuous models: parity tests, then config-4 shape with 4 (default) and 8 frames per thread
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ms_" > gpurun_out/r02_ms2.log 2>&1; echo "tests exit $?: $(tail -n 3 gpurun_out/r02_ms2.log)"
for fu in 1 0; do
echo "PSB_MS_FUSE=$fu: $(PSB_MS_FUSE=$fu timeout 600 python tools/prof_tc.py 512 5 2 cont 2>&1 | tail -1)"
done
[ -n "$SKIP_BENCH" ] || timeout 900 python bench.py --model cont --utts 512 --secs 5 --steps 3 --warmup 3 --search fsg > gpurun_out/r02_bench_cont2.json 2> gpurun_out/r02_bench_cont2.err; echo "bench cont exit $?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_cont2.json").read().strip().splitlines()[-1])
    print("cont", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["kernel_ms_unpipelined"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_cont2.err").read()[-600:])
P
[ -n "$SKIP_NCU" ] || ncu --set full --clock-control none --import-source on -k regex:ms_dist_tile -c 1 -f -o gpurun_out/r02_ms_tile4 python tools/prof_tc.py 128 5 1 cont > /dev/null 2>&1; ls -la gpurun_out/r02_ms_tile4.ncu-rep
