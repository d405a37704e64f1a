#!/bin/bash
# ncu --set full of k_imp_move<TrialH2O> on an EQUILIBRATED ensemble (launch 230 of bench.py --workload c4 with 200 warm-up steps)
mkdir -p gpurun_out
for lib in libpvd_b200.so libpvd_ab_nopartial.so; do
PVD_B200_LIB=$PWD/pyvibdmc_b200/_lib/$lib timeout 600 ncu --clock-control none --set full -k regex:k_imp_move -s 230 -c 1 -f -o /tmp/r02_imp_eq_$lib python bench.py --workload c4 --steps 40 --warmup 50 > gpurun_out/prof_imp_eq.log 2>&1
ncu -i /tmp/r02_imp_eq_$lib.ncu-rep --page raw --csv > gpurun_out/r02_imp_move_fd_eq_raw_$lib.csv 2>/dev/null
done
ls -la gpurun_out/r02_imp_move_fd_eq_raw*
