#!/bin/bash
# round-2 state pass (run under gpurun): full GPU suite, then the driver's own bench invocation
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_state_tests.txt
cat gpurun_out/r02_state_tests.txt
python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e-run > gpurun_out/r02_state_bench20.json 2> gpurun_out/r02_state_bench20.err; tail -3 gpurun_out/r02_state_bench20.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_state_bench20.json').read().strip().splitlines()[-1]); print('bench20 value %.4g ms/step %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']), d['clocks'], d['tutorial_20k']['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['gpu_launches'])"
AB_MODE=3 python tools/seg_time.py
