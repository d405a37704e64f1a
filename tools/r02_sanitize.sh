#!/bin/bash
# compute-sanitizer over small runs of every kernel family (memcheck for all three discrete step kernels, racecheck for the
# default selection).  Output: gpurun_out/r02_sanitize.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
{
echo "== memcheck, default kernel selection (resident k_run_discrete)"
timeout 600 $CS --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 2>&1 | tail -18; echo "rc=$?"
echo "== memcheck, deferred-compaction step (k_step_gather)"
PVD_RUN_MAX_WALKERS=0 timeout 600 $CS --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 2>&1 | tail -6; echo "rc=$?"
echo "== memcheck, compaction inside the step (k_step_discrete)"
PVD_NO_RESIDENT=1 PVD_NO_GATHER=1 timeout 600 $CS --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 2>&1 | tail -6; echo "rc=$?"
echo "== racecheck, deferred-compaction step"
PVD_RUN_MAX_WALKERS=0 timeout 900 $CS --tool racecheck --error-exitcode 9 python tools/sanitize_small.py 2>&1 | tail -8; echo "rc=$?"
} > gpurun_out/r02_sanitize.txt 2>&1
cat gpurun_out/r02_sanitize.txt
