#!/bin/bash
# round 2, session 3: continuous weighting without the chained scan (kill flags as ballot words) -- tests and A/B on config 3
mkdir -p gpurun_out
L=pyvibdmc_b200/_lib
{
echo "== gpu tests (new library)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for lib in libpvd_prev.so libpvd_b200.so libpvd_prev.so libpvd_b200.so; do
  echo "== c3 (H2O continuous, 1e6 walkers) $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c3 --steps 200 --warmup 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['mean_population'])"
done
for lib in libpvd_prev.so libpvd_b200.so; do
  echo "== c3 at 20 000 walkers $lib"
  PVD_B200_LIB=$PWD/$L/$lib timeout 300 python bench.py --workload c3 --walkers 20000 --steps 500 --warmup 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['mean_population'])"
done
} > gpurun_out/r02_s5.txt 2>&1
cat gpurun_out/r02_s5.txt
# per-kernel durations of the continuous step (serialised, cold caches: shares only)
for lib in libpvd_prev.so libpvd_b200.so; do
PVD_B200_LIB=$PWD/$L/$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_s5_launches_$lib.csv python bench.py --workload c3 --steps 12 --warmup 3 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_s5_launches_$lib.csv 2>&1 | grep -i "cont" | head -12 >> gpurun_out/r02_s5.txt
done
cat gpurun_out/r02_s5.txt
