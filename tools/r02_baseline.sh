#!/bin/bash
# round-2 baseline pass (run under gpurun): the driver's own bench invocation + ticket-batch A/B of the round-1 step kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_base_smi.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-other-configs > gpurun_out/r02_base_bench20.json 2> gpurun_out/r02_base_bench20.err
python bench.py --steps 2000 --warmup 50 --no-other-configs --no-cpu-baseline > gpurun_out/r02_base_bench2000.json 2>> gpurun_out/r02_base_bench20.err
for b in 1 2; do PVD_TICKET_BATCH_HEAVY=$b AB_RNG=zig python tools/step_ab.py --one; done > gpurun_out/r02_base_ab.txt 2>&1
AB_WALKERS=20000 AB_RNG=zig python tools/step_ab.py --one >> gpurun_out/r02_base_ab.txt 2>&1
PVD_NO_PDL=1 AB_RNG=zig python tools/step_ab.py --one >> gpurun_out/r02_base_ab.txt 2>&1
cat gpurun_out/r02_base_ab.txt
cat gpurun_out/r02_base_bench20.json | head -c 1500
