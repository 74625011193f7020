#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ptm_tc_kernel -c 1 -o gpurun_out/r02_tc2_kernel python tools/prof_tc.py 100 10 1 > gpurun_out/r02_ncu_full.log 2>&1; tail -2 gpurun_out/r02_ncu_full.log
