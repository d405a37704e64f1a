#!/bin/bash
# round-2 multi-GPU pass (gpurun --gpus N): 2-GPU tests, then the bench line at N GPUs (driver's invocation)
N=${NGPU:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_multi_tests.txt
cat gpurun_out/r02_multi_tests.txt
for NN in ${NLIST:-$N}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NN --master-addr 127.0.0.1 --master-port 2965$NN bench.py --gpus $NN --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/r02_multi_$NN.json 2> gpurun_out/r02_multi_$NN.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_multi_$NN.json").read().strip().splitlines()[-1])
    print("N=$NN", d["steps"], "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d.get("parity_check"), d["clocks"])
except Exception as e:
    print("N=$NN FAILED", e); print(open("gpurun_out/r02_multi_$NN.err").read()[-2500:])
PY
done
python bench.py --steps ${STEPS:-20} --warmup ${WARMUP:-5} --no-cpu-baseline --no-other-configs --no-e2e-run > gpurun_out/r02_multi_1.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_multi_1.json').read().strip().splitlines()[-1]); print('N=1 value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']), d['clocks'])"
