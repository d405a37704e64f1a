#!/usr/bin/env python
"""Round 2: turn the ncu artefacts of tools/profile_r02.sh (gpurun_out/) into the small, tracked summaries under profiles/."""
import csv
import json
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_profiles as S1  # noqa: E402

G, OUT = S1.G, S1.OUT


def raw_csv(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d["Kernel Name"]}
        for k in S1.KEYS + ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]:
            if k in d:
                rec[k] = d[k] + " " + units[hdr.index(k)]
        rec["stalls_per_issue"] = {k.split("stalled_")[1].split("_per")[0]: float(d[k]) for k in hdr
                                   if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k] or 0) >= 0.05}
        out.append(rec)
    return out


def main():
    md = ["# Round 2 profiles (B200, `ncu --clock-control none`, tools/profile_r02.sh)", "",
          "Per-launch times under ncu are cold-cache and serialised (no programmatic dependent launch overlap): compare SHARES with the bench, not absolutes.", ""]
    lp = os.path.join(G, "launches_r02.csv")
    if os.path.exists(lp):
        md += ["## Launch list of `python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e-run --hw-warmup-ms 20`", "",
               "(the driver's invocation; the 296-CTA `k_step_gather` launches are the 1e6-walker steps -- warm-up, timed region, the three end-to-end "
               "repetitions -- each segment of steps ends with one `k_gather_materialise`; `k_run_discrete` is the resident kernel of the 20 000-walker "
               "tutorial leg (2 200 steps in two launches); `k_fp64_peak` is the roofline micro-benchmark.)", "", S1.launches(lp), ""]
    lp = os.path.join(G, "launches_cfg_r02.csv")
    if os.path.exists(lp):
        md += ["## Launch list of `python tools/config_bench.py --large-only --steps 6 --equil 30`", "",
               "(BASELINE configs 1, 3, 4, 5 at their per-GPU sizes.)", "", S1.launches(lp), ""]
    summ = {}
    rep = os.path.join(G, "r02_step_gather.ncu-rep")
    if os.path.exists(rep):
        import subprocess
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open("/tmp/_r02_gather_raw.csv", "w").write(txt)
        summ["r02_step_gather"] = raw_csv("/tmp/_r02_gather_raw.csv")
    for name in ("r02_materialise", "r02_step_discrete", "r02_cont_update", "r02_imp_move_fd"):
        p = os.path.join(G, name + "_raw.csv")
        if os.path.exists(p):
            summ[name] = raw_csv(p)
    for name, recs in summ.items():
        md += [f"## `ncu --set full` : {name}", "", "```json", json.dumps(recs, indent=1), "```", ""]
    json.dump(summ, open(os.path.join(OUT, "r02_ncu_summary.json"), "w"), indent=1)
    sg = summ.get("r02_step_gather")
    if sg:
        rd = float(sg[0]["dram__bytes_read.sum"].split()[0]) * 1e6
        wr = float(sg[0]["dram__bytes_write.sum"].split()[0]) * 1e6
        json.dump({"kernel": "k_step_gather<PotH2O, ziggurat normals>", "walkers_per_launch": 1_000_000, "dram_bytes_per_launch": rd + wr,
                   "dram_read": rd, "dram_write": wr, "algorithmic_bytes_per_launch": 164e6,
                   "note": "algorithmic: read 72 B coordinates + 4 B copy count of the source slot, write 72 B moved coordinates + 8 B V + 4 B copy count "
                           "per walker (+ 4 B tile totals per 32); the buffer a step writes is partly still in the 126 MB L2 when the next step gathers from it"},
                  open(os.path.join(OUT, "r02_step_kernel_traffic.json"), "w"), indent=1)
    for f in ("r02_BENCH_20steps.json", "r02_BENCH_default.json", "r02_BENCH_reference.json", "r02_exp.txt", "r02_seg.txt", "r02_ab2.txt", "r02_ab3.txt", "r02_gather_ab.txt"):
        src = os.path.join(G, f)
        if os.path.exists(src):
            open(os.path.join(OUT, f), "w").write(open(src).read())
    open(os.path.join(OUT, "r02_profiles.md"), "w").write("\n".join(md))
    print("\n".join(md)[:5000])


if __name__ == "__main__":
    main()
