#!/bin/bash
# First GPU call of the next round: run the search kernels that have only been verified in host
# emulation, each test file under its own timeout so that a hang cannot take the
# box with it, in both bindings of the phase code.  Logs go to gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/run_unverified.sh'
set -u
mkdir -p gpurun_out
export PSB_RUN_UNVERIFIED=1
rc=0
for mode in 0 1; do
    export PSB_SEARCH_WARP=$mode
    for f in tests/test_gpu_zz_fsg.py tests/test_gpu_zz_ngram.py; do
        log=gpurun_out/unverified_$(basename $f .py)_warp$mode.log
        timeout 300 python -m pytest $f -x -q -m gpu > $log 2>&1
        r=$?
        echo "$f warp=$mode -> exit $r: $(tail -n 1 $log)"
        [ $r -ne 0 ] && rc=1
    done
done
unset PSB_SEARCH_WARP
timeout 300 python bench.py --steps 2 --warmup 1 > gpurun_out/unverified_bench.log 2>&1
echo "bench exit $?: $(python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/unverified_bench.log").read().strip().splitlines()[-1])
    print(json.dumps(d.get("search_stage", "no search_stage")))
except Exception as e:
    print("unreadable:", e)
P
)"
exit $rc
