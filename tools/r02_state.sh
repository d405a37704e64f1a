#!/bin/bash
# round-2 state pass (run under gpurun): full GPU suite, then the driver's own bench invocation and a long one
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_state_tests.txt
cat gpurun_out/r02_state_tests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_state_bench20.json 2> gpurun_out/r02_state_bench20.err; tail -3 gpurun_out/r02_state_bench20.err
python bench.py --steps 2000 --warmup 50 --no-other-configs --no-cpu-baseline --no-e2e-run > gpurun_out/r02_state_bench2000.json 2>> gpurun_out/r02_state_bench20.err
cat gpurun_out/r02_state_bench20.json | head -c 3000; echo; cat gpurun_out/r02_state_bench2000.json | head -c 1500
