#!/bin/bash
# the beam-pruning cluster sweep: parity tests, memcheck + racecheck of one case, timing at the headline's shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sweep" > gpurun_out/r02_beam.log 2>&1; echo "tests exit $?: $(tail -n 5 gpurun_out/r02_beam.log)"
[ -n "$SKIP_MEMCHECK" ] || timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sweep_beam and 2500" > gpurun_out/r02_beam_memcheck.log 2>&1; echo "memcheck exit $?: $(tail -n 4 gpurun_out/r02_beam_memcheck.log)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sweep_beam and 900" > gpurun_out/r02_beam_racecheck.log 2>&1; echo "racecheck exit $?: $(tail -n 4 gpurun_out/r02_beam_racecheck.log)"
timeout 600 python tools/beam_sweep_time.py > gpurun_out/r02_beam_time.json 2> gpurun_out/r02_beam_time.err; echo "time exit $?"; cat gpurun_out/r02_beam_time.json; tail -n 5 gpurun_out/r02_beam_time.err
