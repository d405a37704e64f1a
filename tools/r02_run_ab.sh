#!/bin/bash
# round-2: resident multi-step kernel -- parity tests, then A/B timings (run under gpurun).
# VARIANTS="2563 3842" selects occupancy variants, MODES="static dynamic single" the tile assignment / launch mode.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_run.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_run_tests.txt
cat gpurun_out/r02_run_tests.txt
{
for n in ${SIZES:-1000000 20000}; do
  echo "== walkers $n"
  PVD_NO_RESIDENT=1 AB_WALKERS=$n python tools/step_ab.py --one
  for v in ${VARIANTS:-2563}; do
    for m in ${MODES:-dynamic static}; do
      echo "-- variant $v mode $m"
      case $m in
        static) PVD_RUN_STATIC=1 PVD_RUN_VERBOSE=1 PVD_RUN_VARIANT=$v AB_WALKERS=$n timeout 300 python tools/step_ab.py --one 2>&1 | tail -2;;
        dynamic) PVD_RUN_VARIANT=$v AB_WALKERS=$n timeout 300 python tools/step_ab.py --one 2>&1 | tail -1;;
        single) PVD_RUN_SINGLE=1 PVD_RUN_VARIANT=$v AB_WALKERS=$n timeout 300 python tools/step_ab.py --one 2>&1 | tail -1;;
      esac
    done
  done
done
} > gpurun_out/r02_run_ab.txt 2>&1
cat gpurun_out/r02_run_ab.txt
if [ -n "$PROF_NAME" ]; then bash tools/r02_prof.sh; fi
