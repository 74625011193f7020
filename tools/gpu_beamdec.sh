#!/bin/bash
mkdir -p gpurun_out
N=${OUT:-r02_beam_sweep_decode}
timeout 900 python tools/beam_sweep_decode.py ${REPS:-60} ${MODE:-} > gpurun_out/$N.json 2> gpurun_out/$N.err; echo "beam sweep exit $?"; cat gpurun_out/$N.json; tail -n 6 gpurun_out/$N.err
