#!/bin/bash
# end of round 2 (session 3): tests, smoke, launch lists + bench lines (profile_r02.sh a), --set full captures of the two kernels
# this session changed (k_cont_update: config 3, k_imp_move<TrialH2O>: config 4)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_final_tests.txt
cat gpurun_out/r02_final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/profile_r02.sh a > /dev/null 2>&1
NCU="ncu --clock-control none"
timeout 600 $NCU --set full -k regex:k_cont_update -s 6 -c 1 -f -o /tmp/r02_cont_update python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/prof_cont.log 2>&1
timeout 600 $NCU --set full -k regex:k_imp_move -s 6 -c 1 -f -o /tmp/r02_imp_move_fd python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/prof_imp.log 2>&1
ncu -i /tmp/r02_cont_update.ncu-rep --page raw --csv > gpurun_out/r02_cont_update_raw.csv 2>/dev/null
ncu -i /tmp/r02_imp_move_fd.ncu-rep --page raw --csv > gpurun_out/r02_imp_move_fd_raw.csv 2>/dev/null
for wl in c3 c4 c4a c5; do timeout 300 python bench.py --workload $wl --steps 200 --warmup 50 > gpurun_out/r02_wl_${wl}_1.json 2> gpurun_out/r02_wl_${wl}_1.err; done
python - <<'PY'
import json
for f in ["r02_BENCH_20steps","r02_BENCH_default","r02_BENCH_reference","r02_wl_c3_1","r02_wl_c4_1","r02_wl_c4a_1","r02_wl_c5_1"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f"%(d["value"],d.get("ms_per_step",0)), "frac", (d.get("roofline") or {}).get("frac"), "e2e %.4g"%d["e2e"]["value"], (d.get("tutorial_20k") or {}).get("ms_per_step"), (d.get("e2e_run") or {}).get("seconds"))
        oc = d.get("other_configs") or {}
        for k, v in oc.items(): print("   ", k, v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"))
    except Exception as e: print(f,"FAILED",e)
PY
