#!/bin/bash
# A/B of the partial finite-difference stencil at EQUILIBRIUM (200 warm-up steps) and in the driver's 20-step shape
mkdir -p gpurun_out
L=pyvibdmc_b200/_lib
{
for rep in 1 2; do
for lib in libpvd_ab_nopartial.so libpvd_b200.so; do
  echo "== c4 equilibrated (warm-up 200, 100 steps) $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c4 --steps 100 --warmup 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['mean_population'])"
done
done
for lib in libpvd_ab_nopartial.so libpvd_b200.so; do
  echo "== c4 warm-up 5, 20 steps $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['mean_population'])"
done
} > gpurun_out/r02_s6.txt 2>&1
cat gpurun_out/r02_s6.txt
