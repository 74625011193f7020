#!/bin/bash
mkdir -p gpurun_out
ONLY=${ONLY:-cluster_sweep_beam_3000} REPS=1 UTTS=1000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hmmset_sweep_kernel -c 1 -f -o gpurun_out/r02_beam_sweep python tools/beam_sweep_time.py > gpurun_out/r02_beam_prof.log 2>&1; tail -3 gpurun_out/r02_beam_prof.log; ls -la gpurun_out/r02_beam_sweep.ncu-rep
