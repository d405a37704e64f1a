#!/bin/bash
mkdir -p gpurun_out
{
for b in 1 2 3 4; do echo "== per-step kernel batch $b"; PVD_TICKET_BATCH_HEAVY=$b PVD_NO_RESIDENT=1 python tools/step_ab.py --one; done
for v in 2562 2563 2564; do
  echo "-- run kernel variant $v single static"; PVD_RUN_STATIC=1 PVD_RUN_MAX_WALKERS=100000000 PVD_RUN_SINGLE=1 PVD_RUN_VARIANT=$v timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
  echo "-- run kernel variant $v resident static"; PVD_RUN_STATIC=1 PVD_RUN_MAX_WALKERS=100000000 PVD_RUN_VARIANT=$v timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
done
} > gpurun_out/r02_ab3.txt 2>&1
cat gpurun_out/r02_ab3.txt
