#!/bin/bash
# round-2 measurement pass: every number DESIGN.md section 6 quotes comes from this script's outputs
mkdir -p gpurun_out
python tools/prof_tc.py 512 5 2 cont 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench default exit $?"
timeout 600 python bench.py --model en-us --steps 3 --warmup 3 > gpurun_out/r02_bench_enus.json 2> gpurun_out/r02_bench_enus.err; echo "bench en-us exit $?"
timeout 600 python bench.py --model semi --utts 512 --secs 5 --steps 3 --warmup 3 --search fwdtree > gpurun_out/r02_bench_semi.json 2> gpurun_out/r02_bench_semi.err; echo "bench semi exit $?"
timeout 600 python bench.py --model cont --utts 512 --secs 5 --steps 3 --warmup 3 --search fsg > gpurun_out/r02_bench_cont.json 2> gpurun_out/r02_bench_cont.err; echo "bench cont exit $?"
timeout 900 python bench.py --utts 1 --secs 3600 --steps 3 --warmup 3 > gpurun_out/r02_bench_longstream.json 2> gpurun_out/r02_bench_longstream.err; echo "bench long stream exit $?"
for f in final enus semi cont longstream; do python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["kernel_ms_unpipelined"], (d.get("search_viterbi") or {}).get("ms"), (d.get("cpu_baseline") or {}).get("value"), json.dumps(d.get("search_coupled"))[:300])
except Exception as e:
    print("$f unreadable", e); print(open("gpurun_out/r02_bench_$f.err").read()[-600:])
P
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hmmset_sweep_kernel -c 1 -o gpurun_out/r02_sweep_kernel python bench.py --steps 1 --warmup 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ptm_tc5_kernel -c 1 -o gpurun_out/r02_tc5_kernel_1000 python tools/prof_tc.py 1000 10 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hmmset_eval_kernel -c 1 -s 20 -o gpurun_out/r02_hmmset_eval_1000 python bench.py --steps 1 --warmup 0 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:ptm_senone4_kernel -c 1 -o gpurun_out/r02_senone4_1000 python tools/prof_tc.py 1000 10 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
