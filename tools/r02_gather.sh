#!/bin/bash
# round-2: deferred-compaction step -- parity tests, then A/B timings (run under gpurun)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gather.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gather_tests.txt
cat gpurun_out/r02_gather_tests.txt
{
for n in ${SIZES:-1000000 600000 200000 20000}; do
  echo "== walkers $n"
  echo "-- per-step"; AB_MODE=0 AB_WALKERS=$n python tools/step_ab.py --one 2>&1 | tail -1
  echo "-- auto"; AB_WALKERS=$n python tools/step_ab.py --one 2>&1 | tail -1
  for b in ${MINBS:-2 3}; do
    echo "-- gather minb $b"; PVD_GATHER_MINB=$b AB_MODE=3 AB_WALKERS=$n timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
  done
done
} > gpurun_out/r02_gather_ab.txt 2>&1
cat gpurun_out/r02_gather_ab.txt
