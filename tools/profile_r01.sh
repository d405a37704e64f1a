#!/bin/bash
# Round-1 profiling pass (run under gpurun on ONE B200).  part = a | b1 | b2 | b3  (separate calls keep what each writes to gpurun_out/ under the 64 MiB merge limit)
set -u
part=${1:-a}
NCU="ncu --clock-control none"
mkdir -p gpurun_out
if [ "$part" = "a" ]; then
  # launch list of the headline bench command (shares, not absolutes)
  $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/launches_r01.csv \
      python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/launches_r01.log 2>&1
  # launch list of the other configurations (per-GPU sizes)
  $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_cfg_r01.csv \
      python tools/config_bench.py --large-only --steps 6 --equil 30 > gpurun_out/launches_cfg_r01.log 2>&1
  # fused discrete step (the headline kernel, default ziggurat normals), with source
  $NCU --set full --import-source on -k regex:k_step_discrete -s 100 -c 1 -f -o gpurun_out/r01_step_discrete \
      python tools/step_ab.py --one > gpurun_out/prof_step.log 2>&1
  # PES alone
  $NCU --set full -k regex:k_pot_aos -s 1 -c 1 -f -o gpurun_out/r01_pot_aos \
      python -c "
import numpy as np, sys
sys.path.insert(0, '.')
from pyvibdmc_b200 import kernels as K
eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
x = eq[None] * 1.01 + np.random.default_rng(0).normal(0, 0.05, (1000000, 3, 3))
for _ in range(3): K.pes_h2o(x)
" > gpurun_out/prof_pot.log 2>&1
  # the bench line itself (not under a profiler) and the ZPE validation
  python bench.py > gpurun_out/BENCH_local.json 2> gpurun_out/bench_local.err
  python tools/zpe_validation.py gpurun_out/zpe_validation.json > gpurun_out/zpe_validation.log 2>&1
elif [ "$part" = "b1" ]; then
  $NCU --set full -k regex:k_cont_update -s 40 -c 1 -f -o gpurun_out/r01_cont_update \
      python tools/config_bench.py --only c3 --large-only --steps 6 --equil 60 > gpurun_out/prof_c3.log 2>&1
  $NCU --set full -k regex:k_imp_move -s 20 -c 1 -f -o gpurun_out/r01_imp_move \
      python tools/config_bench.py --only c4 --large-only --steps 6 --equil 30 > gpurun_out/prof_c4.log 2>&1
elif [ "$part" = "b2" ]; then
  $NCU --set full -k regex:k_nn_h4o2_tc2 -s 1 -c 1 -f -o gpurun_out/r01_nn_tc2 \
      python tools/nn_bench.py 4000000 > gpurun_out/prof_nn.log 2>&1
  $NCU --set full -k regex:k_branch_discrete -s 10 -c 1 -f -o gpurun_out/r01_branch_discrete \
      python tools/config_bench.py --only c5 --large-only --steps 8 --equil 10 > gpurun_out/prof_c5.log 2>&1
else
  $NCU --set full -k regex:k_displace_soa -s 10 -c 1 -f -o gpurun_out/r01_displace_soa \
      python tools/config_bench.py --only c5 --large-only --steps 8 --equil 10 > gpurun_out/prof_c5d.log 2>&1
fi
ls -la gpurun_out | tail -20
