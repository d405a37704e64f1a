#!/bin/bash
# k_branch_discrete occupancy A/B (configs 4 and 5 branch with it): 2 / 3 / 4 CTAs per SM
mkdir -p gpurun_out
L=pyvibdmc_b200/_lib
{
for lib in libpvd_prev.so libpvd_ab_br3.so libpvd_ab_br4.so libpvd_prev.so; do
  echo "== c5 (NN, 1.25e7 walkers, 20 steps) $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c5 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['mean_population'])"
done
for lib in libpvd_prev.so libpvd_ab_br3.so libpvd_ab_br4.so; do
  echo "== c4a (analytic importance sampling, equilibrated) $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c4a --steps 100 --warmup 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['mean_population'])"
done
} > gpurun_out/r02_s9.txt 2>&1
cat gpurun_out/r02_s9.txt
