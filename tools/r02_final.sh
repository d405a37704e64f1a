#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_final_tests.txt
cat gpurun_out/r02_final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/profile_r02.sh a > /dev/null 2>&1
python - <<'PY'
import json
for f in ["r02_BENCH_20steps","r02_BENCH_default","r02_BENCH_reference"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f"%(d["value"],d.get("ms_per_step",0)), "frac", (d.get("roofline") or {}).get("frac"), "e2e %.4g"%d["e2e"]["value"], (d.get("tutorial_20k") or {}).get("ms_per_step"), (d.get("e2e_run") or {}).get("seconds"))
    except Exception as e: print(f,"FAILED",e)
PY
