#!/bin/bash
# finite-difference importance sampling: code size vs instruction fetch.  prev = seven inlined copies of the trial function,
# ab_inline = one copy per stencil pair (four), b200 = ONE out-of-line function
mkdir -p gpurun_out
L=pyvibdmc_b200/_lib
{
echo "== gpu tests (new library)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for rep in 1 2; do
for lib in libpvd_prev.so libpvd_ab_inline.so libpvd_b200.so; do
  echo "== c4 equilibrated (warm-up 200, 100 steps) $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c4 --steps 100 --warmup 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['mean_population'])"
done
done
for lib in libpvd_prev.so libpvd_ab_inline.so libpvd_b200.so; do
  echo "== c4 warm-up 5, 20 steps $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['mean_population'])"
done
} > gpurun_out/r02_s8.txt 2>&1
cat gpurun_out/r02_s8.txt
