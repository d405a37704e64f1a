#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE B200).  part = a | b  (what a call writes to gpurun_out/ must stay under 64 MiB:
# the ncu reports are exported to CSV on the box and only the headline kernel's report travels back)
set -u
part=${1:-a}
NCU="ncu --clock-control none"
mkdir -p gpurun_out
if [ "$part" = "a" ]; then
  # launch list of the driver's own bench invocation (shares, not absolutes)
  $NCU --metrics gpu__time_duration.sum -c 4200 --csv --log-file gpurun_out/launches_r02.csv \
      python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e-run --hw-warmup-ms 20 > gpurun_out/launches_r02.log 2>&1
  # launch list of the other configurations (per-GPU sizes)
  $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_cfg_r02.csv \
      python tools/config_bench.py --large-only --steps 6 --equil 30 > gpurun_out/launches_cfg_r02.log 2>&1
  # the bench lines themselves (not under a profiler)
  python bench.py --steps 20 --warmup 5 > gpurun_out/r02_BENCH_20steps.json 2> gpurun_out/r02_bench20.err
  python bench.py --no-cpu-baseline --no-other-configs --no-e2e-run > gpurun_out/r02_BENCH_default.json 2> gpurun_out/r02_benchdef.err
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_BENCH_reference.json 2> gpurun_out/r02_benchref.err
else
  # the deferred-compaction step (headline kernel) with source; the materialisation; the self-compacting kernel it replaces
  AB_MODE=3 AB_STEPS=4 $NCU --set full --import-source on -k regex:k_step_gather -s 102 -c 1 -f -o gpurun_out/r02_step_gather python tools/prof_run.py > gpurun_out/prof_gather.log 2>&1
  AB_MODE=3 AB_STEPS=4 $NCU --set full -k regex:k_gather_materialise -s 1 -c 1 -f -o /tmp/r02_materialise python tools/prof_run.py > gpurun_out/prof_mat.log 2>&1
  AB_MODE=0 AB_STEPS=4 $NCU --set full -k regex:k_step_discrete -s 101 -c 1 -f -o /tmp/r02_step_discrete python tools/prof_run.py > gpurun_out/prof_disc.log 2>&1
  ncu -i /tmp/r02_materialise.ncu-rep --page raw --csv > gpurun_out/r02_materialise_raw.csv 2>/dev/null
  ncu -i /tmp/r02_step_discrete.ncu-rep --page raw --csv > gpurun_out/r02_step_discrete_raw.csv 2>/dev/null
fi
ls -la gpurun_out | tail -20
