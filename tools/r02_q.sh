#!/bin/bash
mkdir -p gpurun_out
for n in 20000 100000; do echo "-- auto $n"; AB_WALKERS=$n python tools/step_ab.py --one 2>&1 | tail -1; done
echo "-- gather 1e6"; AB_MODE=3 python tools/step_ab.py --one 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_gather.py tests/test_gpu_run.py -m gpu -x -q 2>&1 | tail -3
