#!/bin/bash
N=${NGPU:-2}
mkdir -p gpurun_out
for mode in overlap nooverlap; do
  for steps in 20 400; do
  if [ $mode = nooverlap ]; then export PVD_NO_GATHER_OVERLAP=1; else unset PVD_NO_GATHER_OVERLAP; fi
  PVD_MBOX_TIMEOUT_S=5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps $steps --warmup 5 > gpurun_out/r02_mab_${mode}_$steps.json 2> gpurun_out/r02_mab_${mode}_$steps.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_mab_${mode}_$steps.json").read().strip().splitlines()[-1])
    print("$mode steps=$steps N=$N value %.4g ms/step %.4f"%(d["value"],d["ms_per_step"]), d.get("parity_check",{}).get("mailbox_equals_nccl"))
except Exception as e:
    print("$mode FAILED", e); print(open("gpurun_out/r02_mab_${mode}_$steps.err").read()[-1500:])
PY
  done
done
unset PVD_NO_GATHER_OVERLAP
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e-run > gpurun_out/r02_mab_1.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_mab_1.json').read().strip().splitlines()[-1]); print('N=1 value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']))"
