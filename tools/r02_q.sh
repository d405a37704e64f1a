#!/bin/bash
mkdir -p gpurun_out
echo "-- gather 1e6"; AB_MODE=3 python tools/step_ab.py --one 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_gather.py tests/test_gpu_run.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e-run > gpurun_out/r02_q20.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_q20.json').read().strip().splitlines()[-1]); print('bench20 value %.4g ms/step %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']), 'e2e %.4g'%d['e2e']['value'], d['tutorial_20k']['ms_per_step'])"
