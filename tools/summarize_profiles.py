#!/usr/bin/env python
"""Turn the ncu artefacts brought back in gpurun_out/ into the small, tracked summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_per_warp.ratio", "launch__occupancy_limit_registers",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d["Kernel Name"]}
        for k in KEYS:
            if k in d:
                rec[k] = d[k] + " " + units[hdr.index(k)]
        rec["stalls_per_issue"] = {k.split("stalled_")[1].split("_per")[0]: float(d[k]) for k in hdr
                                   if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k] or 0) >= 0.05}
        out.append(rec)
    return out


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            agg[(d["Kernel Name"].split("(")[0][:70], d["Grid Size"], d["Block Size"])].append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines = ["| kernel | grid | block | launches | mean us | share of GPU time |", "|---|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k[0]}` | {k[1]} | {k[2]} | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f} % |")
    return "\n".join(lines)


def main():
    os.makedirs(OUT, exist_ok=True)
    md = ["# Round 1 profiles (B200, ncu 2025, `--clock-control none`)", "",
          "Per-launch times under ncu are cold-cache and serialised: compare SHARES with the bench, not absolutes.", ""]
    lp = os.path.join(G, "launches_r01.csv")
    if os.path.exists(lp):
        md += ["## Launch list of `python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs`", "",
               "(`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200`, see `tools/profile_r01.sh`; the 296-CTA launches are the "
               "1e6-walker steps, the 122-CTA launches the 20 000-walker tutorial steps, `k_fp64_peak` is the roofline micro-benchmark.)", "",
               launches(lp), ""]
    lp = os.path.join(G, "launches_cfg_r01.csv")
    if os.path.exists(lp):
        md += ["## Launch list of `python tools/config_bench.py --large-only --steps 6 --equil 30`", "",
               "(BASELINE configs 1, 3, 4, 5 at their per-GPU sizes: 1e6 HO, 1e6 H2O continuous, 1.25e6 H2O importance sampling, 1.25e7 "
               "water dimers on the NN surface; kernels shared between configurations are pooled per grid size.)", "", launches(lp), ""]
    summ = {}
    for name in ("r01_step_discrete", "r01_pot_aos", "r01_cont_update", "r01_imp_move", "r01_nn_tc2", "r01_branch_discrete", "r01_displace_soa"):
        rep = os.path.join(G, name + ".ncu-rep")
        if os.path.exists(rep):
            summ[name] = raw(rep)
            md += [f"## `ncu --set full` : {name}", "", "```json", json.dumps(summ[name], indent=1), "```", ""]
    json.dump(summ, open(os.path.join(OUT, "r01_ncu_summary.json"), "w"), indent=1)
    sd = summ.get("r01_step_discrete")
    if sd:
        rd = float(sd[0]["dram__bytes_read.sum"].split()[0]) * 1e6
        wr = float(sd[0]["dram__bytes_write.sum"].split()[0]) * 1e6
        json.dump({"kernel": "k_step_discrete<PotH2O, ziggurat normals>", "walkers_per_launch": 1_000_000, "dram_bytes_per_launch": rd + wr,
                   "dram_read": rd, "dram_write": wr, "algorithmic_bytes_per_launch": 152e6,
                   "note": "writes of the compacted ensemble mostly stay in the 126 MB L2 until the next step reads them"},
                  open(os.path.join(OUT, "r01_step_kernel_traffic.json"), "w"), indent=1)
    for f in ("zpe_validation.json", "BENCH_local.json"):
        src = os.path.join(G, f)
        if os.path.exists(src):
            open(os.path.join(OUT, "r01_" + f), "w").write(open(src).read())
    open(os.path.join(OUT, "r01_profiles.md"), "w").write("\n".join(md))
    print("\n".join(md)[:6000])


if __name__ == "__main__":
    main()
