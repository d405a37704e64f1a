#!/bin/bash
mkdir -p gpurun_out
{
for st in 0 1000 2000 3500 5000; do
  echo "-- gather stagger $st"
  PVD_GATHER_STAGGER_NS=$st AB_MODE=3 timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
done
} > gpurun_out/r02_exp2.txt 2>&1
cat gpurun_out/r02_exp2.txt
