#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err; tail -3 gpurun_out/r02_bench20.err
for w in c3 c4 c5; do python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -2 gpurun_out/r02_bench_$w.err; done
python - <<'PY'
import json
for f in ["r02_bench20","r02_bench_c3","r02_bench_c4","r02_bench_c5"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f frac %.3f e2e %.4g"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"]), d.get("parity_check"), (d.get("e2e_run") or {}).get("seconds"), (d.get("tutorial_20k") or {}).get("ms_per_step"))
    except Exception as e: print(f,"FAILED",e)
PY
