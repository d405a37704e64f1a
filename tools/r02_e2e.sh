#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dmc_sim.py -m gpu -x -q -k "zpe or restart or device" 2>&1 | tail -6
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e-run > gpurun_out/r02_e2e_$i.json 2> gpurun_out/r02_e2e_$i.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_e2e_$i.json').read().strip().splitlines()[-1]); print('bench20 value %.4g ms/step %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']), 'e2e %.4g'%d['e2e']['value'], d['tutorial_20k']['ms_per_step'])"
done
