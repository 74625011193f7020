#!/bin/bash
# The search kernels in both bindings of the phase code (CTA / warp per
# utterance), each test file under its own timeout so that a hang cannot take the box with it, then
# one small case per kernel under compute-sanitizer.  Logs go to gpurun_out/.  This was the first GPU
# call of round 2 (all green: profiles/r02_first_hw_run/).
#   gpurun --timeout 1500 -- 'bash tools/run_search_hw.sh'
set -u
mkdir -p gpurun_out
rc=0
for mode in 0 1; do
    export PSB_SEARCH_WARP=$mode
    for f in tests/test_gpu_zz_fsg.py tests/test_gpu_zz_ngram.py tests/test_gpu_zz_decoder.py; do
        log=gpurun_out/unverified_$(basename $f .py)_warp$mode.log
        timeout 300 python -m pytest $f -x -q -m gpu > $log 2>&1
        r=$?
        echo "$f warp=$mode -> exit $r: $(tail -n 1 $log)"
        [ $r -ne 0 ] && rc=1
    done
done
unset PSB_SEARCH_WARP
# One small case of each kernel under the sanitizer (memcheck: out-of-bounds in the carved work
# areas; racecheck: shared-memory hazards in the block scans).  Slow, so only when the plain runs
# passed, and never fatal for the exit code: the logs are what is wanted.
if [ $rc -eq 0 ]; then
    for tool in memcheck racecheck; do
        log=gpurun_out/unverified_sanitizer_$tool.log
        timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -x -q -m gpu \
            "tests/test_gpu_zz_fsg.py::test_block_scan_selftest" \
            "tests/test_gpu_zz_fsg.py::test_fsg_batch_matches_reference_and_oracle[go]" \
            "tests/test_gpu_zz_ngram.py::test_fwdtree_batch_matches_reference_and_oracle[default]" \
            "tests/test_gpu_zz_ngram.py::test_fwdflat_batch_matches_reference_and_oracle[flat_default]" \
            > $log 2>&1
        echo "sanitizer $tool -> exit $?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -n 1)"
    done
fi
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
