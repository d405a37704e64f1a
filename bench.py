#!/usr/bin/env python
"""bench.py -- walker-steps/s of the DMC propagation loop (H2O, shipped Partridge-Schwenke PES,
discrete weighting: BASELINE.json configs[1] physics) on N B200s.

    python bench.py --gpus 1 --steps 10000 --warmup 100      (the defaults: a timed region of about one second)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     (the reference's CPU path on the host cores)

One bench "step" = one DMC time step (move -> V -> birth/death -> Vref) of the resident
ensemble.  Default ensemble: 1,000,000 walkers per GPU (weak scaling), i.e. 144 MB of walker
state per GPU in two ping-pong buffers -- larger than the 126 MB L2, so every step streams from
HBM; the tutorial size (20,000 walkers) is timed too and reported under "tutorial_20k".
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
AMU = 1.0 / 6.02213670000e23 / 9.10938970000e-28
MASSES = np.array([1.00782503, 1.00782503, 15.99491462]) * AMU
DT = 5.0
# algorithmic work per walker-step of the fused step kernel
FLOP_PER_WS = 2000.0           # 1600 (PS PES) + ~400 (nine Box-Muller normals)
BYTES_PER_WS = 72 + 72 + 8     # read coords, write compacted coords + V (who_from only inside DW windows)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


def start_ensemble(n):
    return np.ascontiguousarray(np.broadcast_to(EQ[None] * 1.01, (n, 3, 3)))


def time_cpu_reference(budget_s=15.0):
    """The reference's multiprocessing CPU path (oracle restatement) on a bounded sample."""
    from oracle import cpu_reference_loop as R
    cores = os.cpu_count() or 1
    n = 200_000
    probe = R.time_h2o_discrete(n, 1, 1, cores=cores)
    steps = int(max(2, min(40, budget_s / max(probe["seconds"], 1e-3))))
    res = R.time_h2o_discrete(n, steps, 1, cores=cores)
    return {"value": res["value"], "unit": "walker-steps/s", "cores": cores, "kind": "port",
            "sample": f"H2O PS discrete, {n} walkers x {steps} time steps, Pool({cores}) potential like Potential.getpot",
            "seconds": res["seconds"]}, res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_reference_loop as R
    cores = os.cpu_count() or 1
    # bounded sample: as many walkers per step as let the K timed steps finish in about 90 s on this host (at most the
    # 200 000 of the cpu_baseline leg, at least 20 000 = the tutorial's population, where Pool.map is still efficient)
    probe = R.time_h2o_discrete(100_000, 2, 1, cores=cores)
    n = int(min(200_000, max(20_000, probe["value"] * 90.0 / max(args.steps, 1))))
    res = R.time_h2o_discrete(n, args.steps, min(args.warmup, 5), cores=cores)
    ms = 1e3 * res["seconds"] / max(args.steps, 1)
    line = {"impl": "reference", "metric": "walker-steps/s", "value": res["value"], "unit": "walker-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "h2o_ps_discrete", "walkers_per_step": n, "delta_t": DT,
                       "cpu_warmup_steps": min(args.warmup, 5),
                       "note": "reference CPU path (oracle port: NumPy loop + C PES behind Pool.map), bounded sample"},
            "cpu_baseline": {"value": res["value"], "unit": "walker-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{n} walkers x {args.steps} time steps"},
            "e2e": {"value": res["value"], "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    args.restore_stdout()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--walkers", type=int, default=1_000_000, help="walkers per GPU")
    ap.add_argument("--rng", default="ziggurat", choices=["ziggurat", "fp64", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="mailbox", choices=["mailbox", "nccl"],
                    help="per-step exchange for N > 1: NVLink peer-memory mailbox fused into the step kernel, or a NCCL all-reduce")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the per-GPU timings of BASELINE configs 1, 3, 4, 5")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library prints there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def restore_stdout():
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    args.restore_stdout = restore_stdout
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    from pyvibdmc_b200 import kernels as K, _capi
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if K.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    rng_mode = _capi.RNG_MODES[args.rng]
    n_loc = args.walkers
    n0 = n_loc * world
    stream = torch.cuda.Stream(device=dev)        # all kernels, NCCL calls and timing events share this stream
    torch.cuda.set_stream(stream)

    def make_sim(nw, nw_global, seed):
        s = K.DeviceSim(3, 3, MASSES, nw_global, DT, _capi.POT_H2O_PS, seed=seed, rng_mode=rng_mode, device=local_rank,
                        rank=rank, world_size=world, capacity=int(1.5 * nw) + 1024, stats_ring=max(1 << 14, args.warmup + args.steps + 8))
        s.set_stream(stream.cuda_stream)
        return s

    def connect(s):
        if world > 1 and args.collective == "mailbox":
            handles = [None] * world
            dist.all_gather_object(handles, s.mailbox_handle())
            s.mailbox_connect(handles)

    sim = make_sim(n_loc, n0, 1234 + rank)
    sums_t = torch.zeros(_capi.NSUMS, dtype=torch.float64, device=dev)
    if world > 1:
        sim.set_sums_ptr(sums_t.data_ptr())
        connect(sim)
    start = start_ensemble(n_loc)
    sim.upload(start)
    if world > 1:
        dist.all_reduce(sums_t)
        sim.init_finalize()

    def run_steps(s, k):
        if world == 1:
            s.run(k)
        elif args.collective == "mailbox":
            s.run_mailbox(k)
        else:
            for _ in range(k):
                s.step_local(1)
                dist.all_reduce(sums_t)
                s.step_finalize()

    def steps(k):
        run_steps(sim, k)

    launches0 = K.launch_count()
    steps(args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = K.launch_count()
    ev0.record(stream)
    steps(args.steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = K.launch_count() - launches1
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.summary() if sampler else None
    st = sim.state()
    stats = sim.stats(args.warmup, args.steps)
    assert st["step"] == args.warmup + args.steps, st
    walker_steps = float(stats["pop"].sum())          # sum_t global population (SURVEY 8d metric)
    value = walker_steps / (ms * 1e-3)

    # ---- multi-GPU end to end (every rank: pinned host shard -> device, K steps with the all-reduce, results back to the host)
    e2e_multi = None
    if world > 1:
        host_in = torch.from_numpy(start).pin_memory()
        cap_out = int(1.5 * n_loc) + 1024
        host_out = {"coords": torch.empty((cap_out, 3, 3), dtype=torch.float64).pin_memory().numpy(),
                    "pots": torch.empty(cap_out, dtype=torch.float64).pin_memory().numpy()}
        tot_ws, tot_s, h2d, d2h = 0.0, 0.0, 0, 0
        s2 = make_sim(n_loc, n0, 99 + 17 * rank)
        s2.set_sums_ptr(sums_t.data_ptr())
        connect(s2)
        for rep in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            s2.upload(host_in.numpy())
            dist.all_reduce(sums_t)
            s2.init_finalize()
            run_steps(s2, args.steps)
            out = s2.download(out=host_out)
            stt = s2.stats(0, args.steps)
            dt_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
            if rep > 0:
                tot_ws += float(stt["pop"].sum())
                tot_s += float(dt_t.item())
                h2d = host_in.numel() * 8 * world
                d2h = (out["coords"].nbytes + out["pots"].nbytes) * world + stt.nbytes
        s2.close()
        sim.set_sums_ptr(sums_t.data_ptr())
        e2e_multi = {"value": tot_ws / tot_s, "unit": "walker-steps/s", "h2d_bytes_per_step": h2d / args.steps,
                     "d2h_bytes_per_step": d2h / args.steps,
                     "what": f"per rank: upload shard (pinned host) + {args.steps} time steps with the per-step exchange + download; max over ranks"}

    if rank != 0:
        sim.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel of the step: k_step_discrete<PotH2O>
    hbm_peak, peak_kind = load_peaks()
    fp64_peak = K.fp64_peak()
    per_gpu_ws_per_s = value / world
    kernel_ms = ms / args.steps                       # one launch per step; events bracket the launches on their stream
    roofline = {"bound": "fp64", "kernel": "k_step_discrete<PotH2O>", "achieved": per_gpu_ws_per_s * FLOP_PER_WS / 1e12,
                "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": per_gpu_ws_per_s * FLOP_PER_WS / fp64_peak,
                "peak_kind": "measured here: register-resident DFMA chain (pvd_measure_fp64_peak)",
                "avg_launch_ms": kernel_ms, "flop_per_walker_step": FLOP_PER_WS, "traffic": None,
                "hbm": {"achieved": per_gpu_ws_per_s * BYTES_PER_WS / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": per_gpu_ws_per_s * BYTES_PER_WS / 1e9 / hbm_peak, "peak_kind": peak_kind,
                        "bytes_per_walker_step": BYTES_PER_WS}}
    prof = os.path.join(ROOT, "profiles", "r01_step_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- end to end through the public call: host start structures -> K time steps -> host results
    e2e = e2e_multi
    if world == 1:
        host_in = torch.from_numpy(start).pin_memory()
        cap_out = int(1.5 * n_loc) + 1024
        host_out = {"coords": torch.empty((cap_out, 3, 3), dtype=torch.float64).pin_memory().numpy(),
                    "pots": torch.empty(cap_out, dtype=torch.float64).pin_memory().numpy()}
        reps, tot_ws, tot_s, h2d, d2h = 3, 0.0, 0.0, 0, 0
        for rep in range(reps + 1):
            s2 = make_sim(n_loc, n0, 99 + rep)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s2.upload(host_in.numpy())
            s2.run(args.steps)
            out = s2.download(out=host_out)
            stt = s2.stats(0, args.steps)
            dt_s = time.perf_counter() - t0
            if rep > 0:
                tot_ws += float(stt["pop"].sum())
                tot_s += dt_s
                h2d = host_in.numel() * 8
                d2h = out["coords"].nbytes + out["pots"].nbytes + stt.nbytes
            s2.close()
        e2e = {"value": tot_ws / tot_s, "unit": "walker-steps/s", "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps,
               "what": f"upload start structures (pinned host) + {args.steps} time steps + download walkers and V (pinned host) and per-step Vref/pop"}

    # ---- tutorial size (configs[1] as shipped: 20,000 walkers)
    tut = None
    if world == 1:
        s3 = make_sim(20000, 20000, 5)
        s3.upload(start_ensemble(20000))
        s3.run(200)
        s3.sync()
        s3.run(2000)
        t_ms = s3.last_run_ms()
        tst = s3.stats(200, 2000)
        tut = {"walkers": 20000, "steps": 2000, "ms_per_step": t_ms / 2000, "value": float(tst["pop"].sum()) / (t_ms * 1e-3),
               "unit": "walker-steps/s"}
        s3.close()

    # ---- the other BASELINE configurations at their per-GPU sizes (parity-test cases; reported, not the headline)
    others = None
    if world == 1 and not args.no_other_configs:
        import importlib.util
        spec = importlib.util.spec_from_file_location("config_bench", os.path.join(ROOT, "tools", "config_bench.py"))
        cb = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cb)
        sim.close()
        try:
            others = cb.collect(("c1", "c3", "c4", "c4a", "c5"), large_only=False, steps=100)
        except Exception as e:             # never lose the headline line to a side measurement
            others = {"error": repr(e)}
        others["note"] = ("steady-state device-resident loop, CUDA events inside pvd_sim_run; c1 = 1-D HO discrete, c3 = H2O continuous "
                          "(1e6/GPU), c4 = H2O importance sampling with finite-difference drift (1.25e6/GPU), c4a = the same with the reference's analytic derivatives, c5 = (H2O)2 NN PES on tcgen05 "
                          "(1.25e7/GPU)")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _ = time_cpu_reference()
        except Exception as e:
            cpu = {"value": None, "unit": "walker-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + repr(e)}

    line = {"metric": "walker-steps/s", "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "h2o_ps_discrete", "potential": "Partridge-Schwenke H2O (shipped Fortran PES, CUDA fp64)",
                       "weighting": "discrete", "walkers_per_gpu": n_loc, "global_walkers": n0, "delta_t": DT,
                       "rng": "philox4x32-10 + " + {"ziggurat": "fp64 ziggurat (1024 layers)", "fp64": "fp64 Box-Muller",
                                                       "fast": "SFU Box-Muller"}[args.rng],
                       "l2": "walker state (2 x 80 MB ping-pong at 1e6 walkers) exceeds the 126 MB L2; no flush needed",
                       "parallelism": (f"walkers sharded over {world} GPU(s); per step {_capi.NSUMS} doubles per shard are exchanged "
                                       + ("by peer stores over NVLink from the step kernel's last CTA (mailbox), no collective kernel"
                                          if args.collective == "mailbox" else "by one NCCL all-reduce"))
                       if world > 1 else "single GPU, one kernel launch per time step"},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
            "tutorial_20k": tut, "other_configs": others, "final_population": int(st["n"]),
            "zpe_cm1_last_half": float(stats["vref"][args.steps // 2:].mean() / 4.556335281212229e-6)}
    args.restore_stdout()
    print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
