#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest ${TESTS:-tests/test_gpu_impext.py tests/test_gpu_dmc_sim.py} -m gpu -x -q 2>&1 | tail -${TAILN:-30}
