#!/bin/bash
# round-2 ablation of the deferred-compaction step at 1e6 walkers (libraries built with -DPVD_EXP_*; timings only, results are wrong by design)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gather.py -m gpu -x -q 2>&1 | tail -5
{
for lib in "" ${LIBS:-exp_NO_RNG exp_NO_PES exp_NO_RNG_PES}; do
  for b in 2 3; do
    echo "-- lib ${lib:-full} gather minb $b"
    if [ -n "$lib" ]; then export PVD_B200_LIB=$PWD/pyvibdmc_b200/_lib/$lib.so; else unset PVD_B200_LIB; fi
    PVD_GATHER_MINB=$b AB_MODE=3 timeout 300 python tools/step_ab.py --one 2>&1 | tail -1
  done
done
unset PVD_B200_LIB
echo "-- per-step"; AB_MODE=0 python tools/step_ab.py --one 2>&1 | tail -1
for n in 600000 200000; do echo "-- gather $n"; AB_WALKERS=$n AB_MODE=3 python tools/step_ab.py --one 2>&1 | tail -1; echo "-- auto $n";  AB_WALKERS=$n python tools/step_ab.py --one 2>&1 | tail -1; done
} > gpurun_out/r02_exp.txt 2>&1
cat gpurun_out/r02_exp.txt
if [ -n "$PROF" ]; then
AB_MODE=3 AB_STEPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_gather -s 102 -c 1 -f -o gpurun_out/r02_gather python tools/prof_run.py > gpurun_out/r02_gather_prof.log 2>&1
tail -2 gpurun_out/r02_gather_prof.log
fi
