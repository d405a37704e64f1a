#!/bin/bash
# ncu --set full capture of one resident launch (20 steps at 1e6 walkers) + launch list
mkdir -p gpurun_out
AB_STEPS=${AB_STEPS:-20} timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_discrete -s 1 -c 1 -f -o gpurun_out/${PROF_NAME:-r02_run} python tools/prof_run.py > gpurun_out/${PROF_NAME:-r02_run}.log 2>&1
tail -3 gpurun_out/${PROF_NAME:-r02_run}.log
