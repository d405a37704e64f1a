#!/bin/bash
# round-2 multi-GPU pass (gpurun --gpus N): 2-GPU tests, then bench at N GPUs with the per-step and the resident kernel
N=${NGPU:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_multi_tests.txt
cat gpurun_out/r02_multi_tests.txt
run() { # name, extra env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/r02_multi_${name}.json 2> gpurun_out/r02_multi_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_multi_${name}.json").read().strip().splitlines()[-1])
    print("${name}", d["n_gpus"], d["steps"], "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d.get("parity_check"))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r02_multi_${name}.err").read()[-1500:])
PY
}
run perstep20 PVD_NO_RESIDENT=1
run resident20 PVD_RUN_MAX_WALKERS=100000000
STEPS=2000 WARMUP=50 run perstep2000 PVD_NO_RESIDENT=1
STEPS=2000 WARMUP=50 run resident2000 PVD_RUN_MAX_WALKERS=100000000
