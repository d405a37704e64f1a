#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gather.py -m gpu -x -q 2>&1 | tail -3
{ echo "== gather"; AB_MODE=3 python tools/seg_time.py; echo "== gather warm 5"; AB_MODE=3 python tools/seg_time2.py; } > gpurun_out/r02_seg.txt 2>&1
cat gpurun_out/r02_seg.txt
python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e-run > gpurun_out/r02_b20.json 2> gpurun_out/r02_b20.err; tail -2 gpurun_out/r02_b20.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_b20.json').read().strip().splitlines()[-1]); print('bench20 value %.4g ms/step %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']), d['clocks'], d['tutorial_20k']['ms_per_step'], d['e2e']['value'])"
