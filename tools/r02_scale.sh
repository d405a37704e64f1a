#!/bin/bash
# scaling probe (gpurun --gpus 8): headline bench line at N = 8, 4, 2, 1 with the driver's --steps 20 --warmup 5, then the other
# BASELINE workloads (c3 continuous, c4 importance sampling, c5 NN) at N = 8
mkdir -p gpurun_out
for N in ${NLIST:-8 4 2}; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2966$N bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} ${EXTRA} > gpurun_out/r02_scale_$N.json 2> gpurun_out/r02_scale_$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms/step %.4f e2e %.4g"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d.get("parity_check"), d["clocks"])
except Exception as e:
    print("N=$N FAILED", e); print(open("gpurun_out/r02_scale_$N.err").read()[-1500:])
PY
done
python bench.py --steps ${STEPS:-20} --warmup ${WARMUP:-5} --no-cpu-baseline --no-other-configs --no-e2e-run > gpurun_out/r02_scale_1.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_scale_1.json').read().strip().splitlines()[-1]); print('N=1 value %.4g ms/step %.4f'%(d['value'],d['ms_per_step']))"
for W in ${WORKLOADS:-c3 c4 c5}; do
  NW=${WLN:-8}
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NW --master-addr 127.0.0.1 --master-port 2967$NW bench.py --gpus $NW --workload $W --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/r02_wl_${W}_$NW.json 2> gpurun_out/r02_wl_${W}_$NW.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_wl_${W}_$NW.json").read().strip().splitlines()[-1])
    print("$W N=$NW value %.4g ms/step %.4f"%(d["value"],d["ms_per_step"]), d.get("roofline",{}).get("frac"), d["config"].get("workload"))
except Exception as e:
    print("$W N=$NW FAILED", e); print(open("gpurun_out/r02_wl_${W}_$NW.err").read()[-1500:])
PY
done
