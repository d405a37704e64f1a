#!/bin/bash
mkdir -p gpurun_out
{
for b in 2 3; do for c in default 100 50 25; do
  echo "-- gather minb $b carveout $c"
  if [ $c = default ]; then unset PVD_GATHER_CARVEOUT; else export PVD_GATHER_CARVEOUT=$c; fi
  PVD_GATHER_MINB=$b AB_MODE=3 timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
done; done
} > gpurun_out/r02_exp2.txt 2>&1
cat gpurun_out/r02_exp2.txt
