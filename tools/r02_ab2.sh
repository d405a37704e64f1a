#!/bin/bash
# round-2 (re-entry): occupancy variants of the run kernel in single-launch-per-step mode at 1e6, + source-level ncu of the step kernel
mkdir -p gpurun_out
{
echo "== per-step kernel"; PVD_NO_RESIDENT=1 python tools/step_ab.py --one
for v in 2562 2563 2564 3842; do
  echo "-- run kernel variant $v single"; PVD_RUN_MAX_WALKERS=100000000 PVD_RUN_SINGLE=1 PVD_RUN_VARIANT=$v timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
  echo "-- run kernel variant $v resident dynamic"; PVD_RUN_MAX_WALKERS=100000000 PVD_RUN_VARIANT=$v timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
done
} > gpurun_out/r02_ab2.txt 2>&1
cat gpurun_out/r02_ab2.txt
PVD_NO_RESIDENT=1 AB_STEPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_discrete -s 101 -c 1 -f -o gpurun_out/r02_step python tools/prof_run.py > gpurun_out/r02_step.log 2>&1
tail -3 gpurun_out/r02_step.log
