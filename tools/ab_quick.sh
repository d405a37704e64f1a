#!/bin/bash
# quick correctness + timing pass for step-kernel experiments (run under gpurun)
timeout 600 python -m pytest tests/test_gpu_sim.py tests/test_gpu_extended.py -m gpu -x -q 2>&1 | tail -2
for r in zig fp64 fast; do AB_RNG=$r python tools/step_ab.py --one; done
AB_WALKERS=20000 AB_RNG=zig python tools/step_ab.py --one
