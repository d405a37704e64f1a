#!/bin/bash
# round 2, session 3: A/B of the partial finite-difference stencil (config 4) and the paired swish (config 5)
mkdir -p gpurun_out
L=pyvibdmc_b200/_lib
{
echo "== gpu tests (new library)"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for lib in libpvd_prev.so libpvd_b200.so; do
  echo "== c4 (FD importance sampling) $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'))"
done
for lib in libpvd_prev.so libpvd_ab_nopair.so libpvd_b200.so; do
  echo "== c5 (NN) $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c5 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'))"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python tools/nn_bench.py 2>&1 | grep -A3 '"tcgen05": {' | head -4
done
echo "== NN accuracy (new library)"; timeout 300 python tools/nn_accuracy.py 2>&1 | tail -12
} > gpurun_out/r02_s3.txt 2>&1
cat gpurun_out/r02_s3.txt
