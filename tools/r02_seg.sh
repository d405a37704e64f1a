#!/bin/bash
mkdir -p gpurun_out
export PVD_MBOX_TIMEOUT_S=2
timeout 600 python -m pytest tests/test_gpu_gather.py -m gpu -x -q 2>&1 | tail -12
unset PVD_MBOX_TIMEOUT_S
{ echo "== gather"; AB_MODE=3 timeout 300 python tools/seg_time.py; echo "== gather warm 5"; AB_MODE=3 timeout 300 python tools/seg_time2.py; 
for n in 600000 200000 20000; do echo "-- gather $n"; AB_WALKERS=$n AB_MODE=3 timeout 300 python tools/step_ab.py --one 2>&1 | tail -1; done; } > gpurun_out/r02_seg.txt 2>&1
cat gpurun_out/r02_seg.txt
