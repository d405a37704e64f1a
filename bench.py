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
FP64_NOMINAL = 148 * 64 * 2 * 1.965e9      # B200: 148 SMs x 64 FP64 lanes x 2 flop x 1.965 GHz = 37.2 TFLOP/s (SURVEY 8d fallback)
# SURVEY 8(d) per-unit figures of the other BASELINE configurations (dominant kernel of each)
OTHER_ROOFLINES = {
    "c1": ("hbm", "k_run_discrete<PotHarm<1>> / k_step_discrete<PotHarm<1>>", 8 + 8 + 8, "B per walker-step: read x, write compacted x + V (1-D HO)"),
    "c3": ("fp64", "k_cont_update<ContFused<PotH2O>>", 2000.0, "flop per walker-step: 1600 PS PES + ~400 for nine normals"),
    "c4": ("fp64", "k_imp_move<TrialH2O, PotH2O>", 2800.0, "flop per walker-step: 1600 PS PES + 20 trial-wfn evaluations x ~60"),
    "c4a": ("fp64", "k_imp_move<TrialH2OAn, PotH2O>", 1900.0, "flop per walker-step: 1600 PS PES + one analytic trial-wfn evaluation with derivatives (~300)"),
    "c5": ("tensor", "k_nn_h4o2_tc2", 61440.0, "flop per walker: 2 (15*120 + 120*120 + 120*120 + 120)"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def load_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])) * 1e12, "measured (sustained bf16 cuBLAS)"
    return 1.4e15, "fallback"


def other_roofline(label, rate, hbm_peak_gbs, fp64_self):
    """roofline object of one entry of other_configs (rate = walker-steps/s on one GPU)."""
    key = label.split("_")[0]
    if key not in OTHER_ROOFLINES:
        return None
    bound, kernel, per_unit, what = OTHER_ROOFLINES[key]
    if bound == "hbm":
        peak, unit, kind, scale = hbm_peak_gbs * 1e9, "GB/s", "measured copy bandwidth (MEASURED_PEAKS.json)", 1e9
    elif bound == "tensor":
        peak, kind = load_tensor_peak()
        unit, scale = "TFLOP/s", 1e12
    else:
        peak, unit, kind, scale = FP64_NOMINAL, "TFLOP/s", "nominal FP64 (37.2 TFLOP/s; MEASURED_PEAKS.json has no FP64 entry)", 1e12
    r = {"bound": bound, "kernel": kernel, "achieved": rate * per_unit / scale, "peak": peak / scale, "unit": unit,
         "frac": rate * per_unit / peak, "peak_kind": kind, "per_unit": per_unit, "per_unit_what": what, "traffic": None}
    if bound == "fp64":
        r["frac_of_self_measured_peak"] = rate * per_unit / fp64_self
    return r


class ClockSampler(threading.Thread):
    """SM clock, power and clock-event (throttle) reasons sampled DURING the timed region.  NVML in-process (one query
    costs tens of microseconds, so even a 2 ms timed region gets samples and no nvidia-smi process is forked next to the
    measurement); falls back to nvidia-smi every 200 ms when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.0005):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.period = index, [], threading.Event(), period_s
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self.nvml, self.source = None, "nvidia-smi"

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        nv, h = self.nvml, self.handle
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        flag = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))), str(self.max_mhz),
                str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0), flag(nv.nvmlClocksEventReasonHwSlowdown),
                flag(nv.nvmlClocksEventReasonHwThermalSlowdown), flag(nv.nvmlClocksEventReasonSwThermalSlowdown),
                flag(nv.nvmlClocksEventReasonSwPowerCap)]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(self.period if self.nvml else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "source": self.source}


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


def time_dmc_sim_run(cpu_rate=None):
    """End to end through the user API: BASELINE config 2 AS WRITTEN (the tutorial run, docs/getting_started.rst:98-114 of the
    reference): DMC_Sim(...).run() with 20 000 walkers x 20 000 time steps, dt = 5, checkpoints every 500 steps, wave-function
    dumps every 1000 steps with 300 steps of descendant weighting, log, sim_info -- wall clock from the constructor to the
    return of run(), files on disk included (what the reference's own logs time)."""
    import shutil
    import tempfile
    import pyvibdmc_b200 as pv
    folder = tempfile.mkdtemp(prefix="pvd_bench_run_")
    try:
        pdir = os.path.join(os.path.dirname(pv.__file__), "sample_potentials", "FortPots", "Partridge_Schwenke_H2O")
        pot = pv.Potential(potential_function="water_pot", python_file="h2o_potential.py", potential_directory=pdir, num_cores=1)
        n, T = 20000, 20000
        t0 = time.perf_counter()
        sim = pv.DMC_Sim(sim_name="tutorial_water", output_folder=folder, weighting="discrete", num_walkers=n, num_timesteps=T,
                         equil_steps=500, chkpt_every=500, wfn_every=1000, desc_wt_steps=300, atoms=["H", "H", "O"], delta_t=DT,
                         potential=pot, start_structures=EQ[None] * 1.01, seed=2024)
        sim.run()
        secs = time.perf_counter() - t0
        pops = np.asarray(sim._pop_vs_tau, dtype=np.float64)
        vref = np.asarray(sim._vref_vs_tau, dtype=np.float64)
        files = sum(len(f) for _, _, f in os.walk(folder))
        nbytes = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(folder) for f in fs)
        out = {"what": "pv.DMC_Sim(...).run(): BASELINE config 2 as written (20 000 walkers x 20 000 steps, dt 5, chkpt 500, wfn 1000, desc 300) incl. log / checkpoint / HDF5 files",
               "seconds": secs, "walker_steps": float(pops.sum()), "value": float(pops.sum()) / secs, "unit": "walker-steps/s",
               "files_written": files, "bytes_on_disk": nbytes,
               "zpe_cm1": float(vref[len(vref) // 4:].mean() / 4.556335281212229e-6)}
        if cpu_rate:
            out["cpu_port_same_run_estimate_s"] = float(pops.sum()) / cpu_rate
            out["cpu_port_note"] = "walker-steps of this run / the cpu_baseline rate (bounded sample of the same loop on all host cores); the reference's own tutorial logs show 6.6e5 walker-steps/s on 8 cores"
        return out
    finally:
        shutil.rmtree(folder, ignore_errors=True)


DIMER = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                  [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]]) / 0.529177


def workload_spec(name):
    """BASELINE configurations other than the headline, at their per-GPU sizes."""
    from pyvibdmc_b200 import _capi
    import importlib.util
    sp = os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials")
    mH, mO = MASSES[0], MASSES[2]
    if name == "c1":
        mu = mH * mO / (mH + mO)
        om = 3700 * 4.556335281212229e-6
        return dict(label="c1_ho_discrete", natoms=1, ndim=1, masses=[mu], n_loc=1_000_000, dt=10.0, pot=_capi.POT_HARMONIC,
                    pot_params=[(0.5 * mu) * om ** 2], weighting="discrete", start=np.zeros((1, 1)), equil=200,
                    what="1-D harmonic oscillator, discrete weighting (BASELINE config 1 at 1e6 walkers per GPU)")
    if name in ("c3", "c4", "c4a"):
        d = dict(natoms=3, ndim=3, masses=list(MASSES), pot=_capi.POT_H2O_PS, start=EQ * 1.01)
        if name == "c3":
            d.update(label="c3_h2o_continuous", n_loc=1_000_000, dt=5.0, weighting="continuous", equil=300,
                     what="H2O, PS surface, continuous weighting, 1e6 walkers per GPU (BASELINE config 3)")
            return d
        spec = importlib.util.spec_from_file_location("call_trl_h2o_b200", os.path.join(sp, "FortPots", "Partridge_Schwenke_H2O", "call_trl_h2o.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        d.update(n_loc=1_250_000, dt=1.0, weighting="discrete", equil=200)
        if name == "c4":
            d.update(label="c4_h2o_impsamp_fd", trial=_capi.TRIAL_H2O_FD, trial_table=mod.packed_table(),
                     what="H2O importance sampling, product trial wave function, finite-difference drift and local energy, 1.25e6 walkers per GPU (BASELINE config 4)")
        else:
            d.update(label="c4a_h2o_impsamp_analytic", trial=_capi.TRIAL_H2O_AN, trial_table=mod.packed_table_analytic(),
                     what="config 4 with the reference's analytic derivatives (dpsi_dx + chain rule)")
        return d
    if name == "c5":
        return dict(label="c5_dimer_nn", natoms=6, ndim=3, masses=[mO, mH, mH] * 2, n_loc=12_500_000, dt=5.0, pot=_capi.POT_NN_H4O2,
                    weighting="discrete", start=DIMER, equil=50, nn=np.load(os.path.join(sp, "TensorflowPots", "sample_h4o2_nn_packed.npy")),
                    what="(H2O)2 on the shipped NN surface (Coulomb descriptor + 15-120-120-120-1 MLP on tcgen05), 1.25e7 walkers per GPU (BASELINE config 5)")
    raise SystemExit("unknown workload " + name)


def run_other_workload(args):
    """--workload c1|c3|c4|c4a|c5 at N GPUs: same contract as the headline line (value = whole-job walker-steps/s)."""
    import torch
    from pyvibdmc_b200 import kernels as K, _capi
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if K.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = workload_spec(args.workload)
    n_loc = w["n_loc"] if args.walkers == 1_000_000 else args.walkers
    n0 = n_loc * world
    trial = w.get("trial", _capi.TRIAL_NONE)
    start = np.ascontiguousarray(np.broadcast_to(w["start"][None], (n_loc,) + w["start"].shape)).reshape(n_loc, w["natoms"], w["ndim"])
    host_in = torch.from_numpy(start).pin_memory().numpy()
    steps_k, warm = args.steps, max(args.warmup, w["equil"] if args.warmup >= 50 else args.warmup)
    ring = max(1 << 14, warm + steps_k + 8)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from pyvibdmc_b200.distributed import ShardedSim
        dist.init_process_group("nccl", device_id=dev)

        def make(seed):
            ss = ShardedSim(w["natoms"], w["ndim"], w["masses"], n0, w["dt"], w["pot"], weighting=w["weighting"], seed=seed,
                            pot_params=w.get("pot_params"), rebalance_every=0, trial=trial, trial_table=w.get("trial_table"),
                            stats_ring=ring, collective=args.collective)
            if "nn" in w:
                ss.sim.set_nn_weights(w["nn"])
            return ss
        sim = make(11)
        stream = sim.stream
        sim.upload(host_in)
        sim.run(warm)
        torch.cuda.synchronize()
        dist.barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gate = torch.zeros(1, dtype=torch.float64, device=dev)
        l0 = K.launch_count()
        with torch.cuda.stream(stream):
            dist.all_reduce(gate)
        ev0.record(stream)
        sim.run(steps_k)
        ev1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        launches = K.launch_count() - l0
        stats = sim.stats(warm, steps_k)
        coll = sim.collective
    else:
        def make(seed):
            s_ = K.DeviceSim(w["natoms"], w["ndim"], w["masses"], n0, w["dt"], w["pot"], weighting=w["weighting"], seed=seed,
                             pot_params=w.get("pot_params"), trial=trial, stats_ring=ring, device=local_rank)
            if trial != _capi.TRIAL_NONE:
                s_.set_trial_table(w["trial_table"])
            if "nn" in w:
                s_.set_nn_weights(w["nn"])
            return s_
        sim = make(11)
        sim.upload(host_in)
        sim.run(warm)
        sim.sync()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = K.launch_count()
        sim.run(steps_k)
        sim.sync()
        ms = sim.last_run_ms()
        launches = K.launch_count() - l0
        stats = sim.stats(warm, steps_k)
        coll = None
    clocks = sampler.summary() if sampler else None
    value = float(stats["pop"].astype(np.float64).sum()) / (ms * 1e-3)
    # end to end: pinned host start structures -> K steps -> walkers and energies back on the host
    tot_ws = tot_s = 0.0
    h2d = d2h = 0
    for rep in range(3):
        s2 = make(99 + rep)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        s2.upload(host_in)
        s2.run(steps_k)
        out = (s2.sim if dist else s2).download()
        stt = s2.stats(0, steps_k)
        dt_s = time.perf_counter() - t0
        if dist:
            tt = torch.tensor([dt_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_s = float(tt.item())
        if rep > 0:
            tot_ws += float(stt["pop"].astype(np.float64).sum())
            tot_s += dt_s
            h2d = host_in.nbytes * world
            d2h = (out["coords"].nbytes + out["pots"].nbytes) * world + stt.nbytes
        s2.close()
    sim.close()
    if dist:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
            return
    hbm_peak, _ = load_peaks()
    roofline = other_roofline(w["label"], value / world, hbm_peak, K.fp64_peak())
    line = {"metric": "walker-steps/s", "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": steps_k, "warmup": warm,
            "ms_per_step": ms / steps_k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.workload != "c5" else "f64 walkers, fp16-split tcgen05 MLP with fp32 accumulation", "data": "synthetic",
            "config": {"workload": w["label"], "what": w["what"], "walkers_per_gpu": n_loc, "global_walkers": n0, "delta_t": w["dt"],
                       "l2": "walker state exceeds the 126 MB L2; no flush needed",
                       "parallelism": (f"walkers sharded over {world} GPUs; per-step exchange: {coll}" if world > 1 else "single GPU")},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
            "e2e": {"value": tot_ws / tot_s, "unit": "walker-steps/s", "h2d_bytes_per_step": h2d / steps_k, "d2h_bytes_per_step": d2h / steps_k,
                    "what": f"upload start structures (pinned host) + {steps_k} time steps + download walkers and V"},
            "cpu_baseline": {"value": None, "unit": "walker-steps/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "not timed for this workload: the CPU arm (--impl reference) and the cpu_baseline leg run the headline workload"},
            "mean_population": float(stats["pop"].mean())}
    args.restore_stdout()
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


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
    ap.add_argument("--workload", default="h2o_ps_discrete", choices=["h2o_ps_discrete", "c1", "c3", "c4", "c4a", "c5"],
                    help="h2o_ps_discrete = the headline (BASELINE config 2 physics at 1e6 walkers per GPU); c1/c3/c4/c4a/c5 = the other "
                         "BASELINE configurations at their per-GPU sizes, also under torchrun (N GPUs)")
    ap.add_argument("--hw-warmup-ms", type=float, default=400.0,
                    help="before the W warm-up steps of the measured ensemble, a scratch ensemble of the same shape is stepped for about this "
                         "long (untimed) so that SM clocks and NVLink links have left their idle states")
    ap.add_argument("--no-e2e-run", action="store_true", help="skip the DMC_Sim(...).run() leg (BASELINE config 2 as written, with its files)")
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
    if args.workload != "h2o_ps_discrete":
        return run_other_workload(args)

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

    sums_t = torch.zeros(_capi.NSUMS, dtype=torch.float64, device=dev)
    start = start_ensemble(n_loc)

    def new_sim(seed, collective):
        s = make_sim(n_loc, n0, seed)
        if world > 1:
            s.set_sums_ptr(sums_t.data_ptr())
            if collective == "mailbox":
                connect(s)
        s.upload(start)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(sums_t)
            s.init_finalize()
        return s

    def run_steps(s, k, collective=None):
        collective = collective or args.collective
        if world == 1:
            s.run(k)
        elif collective == "mailbox":
            s.run_mailbox(k)
        else:
            for _ in range(k):
                s.step_local(1)
                dist.all_reduce(sums_t)
                s.step_finalize()

    # ---- hardware warm-up on a scratch ensemble (untimed; the measured ensemble still gets exactly W warm-up steps)
    hw_steps = int(max(0.0, args.hw_warmup_ms) / 0.11)
    if hw_steps > 0:
        hw = new_sim(4321 + rank, args.collective)
        run_steps(hw, hw_steps)
        torch.cuda.synchronize()
        hw.close()
        if world > 1:
            dist.barrier()

    sim = new_sim(1234 + rank, args.collective)

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
    if world > 1:
        # device-side start gate: a tiny all-reduce ON THE TIMING STREAM completes on every GPU at (nearly) the same time,
        # so the host-side skew with which the ranks leave dist.barrier() is not billed to the first timed step
        gate = torch.zeros(1, dtype=torch.float64, device=dev)
        with torch.cuda.stream(stream):
            dist.all_reduce(gate)
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

    # ---- correctness carried by the bench line itself (N > 1): the NVLink mailbox exchange and a NCCL all-reduce, started
    # from the same seeds, must give bit-identical populations and Vref on every rank, and births - deaths must be the change
    # of the global population in every timed step
    parity_check = None
    if world > 1:
        pc_steps = 50
        hist = {}
        for coll in ("mailbox", "nccl"):
            sp = new_sim(555 + rank, coll)
            run_steps(sp, pc_steps, coll)
            torch.cuda.synchronize()
            stp = sp.stats(0, pc_steps)
            hist[coll] = (stp["vref"].copy(), stp["pop"].copy(), sp.state()["n"])
            sp.close()
        sim.set_sums_ptr(sums_t.data_ptr())
        same_local = bool(np.array_equal(hist["mailbox"][0], hist["nccl"][0]) and np.array_equal(hist["mailbox"][1], hist["nccl"][1])
                          and hist["mailbox"][2] == hist["nccl"][2])
        h = torch.tensor(np.concatenate([hist["mailbox"][0], hist["mailbox"][1]]), device=dev)
        hmax, hmin = h.clone(), h.clone()
        dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
        bd_ok = bool(np.array_equal(np.diff(stats["pop"]), (stats["births"] - stats["deaths"])[1:]))
        flags = torch.tensor([float(same_local), float(torch.equal(hmax, hmin)), float(bd_ok)], dtype=torch.float64, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        parity_check = {"mailbox_equals_nccl": bool(flags[0].item() == 1.0), "same_history_on_all_ranks": bool(flags[1].item() == 1.0),
                        "births_minus_deaths_ok": bool(flags[2].item() == 1.0), "steps": pc_steps,
                        "what": "two extra ensembles per rank, same seeds, 50 steps through each exchange: Vref and population histories compared bit for bit"}
    elif world == 1:
        parity_check = {"births_minus_deaths_ok": bool(np.array_equal(np.diff(stats["pop"]), (stats["births"] - stats["deaths"])[1:]))}

    if rank != 0:
        sim.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel of the step (one launch per time step)
    step_kernel = ("k_step_gather<PotH2O>" if n_loc > 300000 and not os.environ.get("PVD_NO_GATHER") else
                   "k_run_discrete<PotH2O>") if world == 1 or args.collective == "mailbox" else "k_step_discrete<PotH2O>"
    hbm_peak, peak_kind = load_peaks()
    fp64_peak = K.fp64_peak()
    per_gpu_ws_per_s = value / world
    kernel_ms = ms / args.steps                       # one launch per step; events bracket the launches on their stream
    roofline = {"bound": "fp64", "kernel": step_kernel, "achieved": per_gpu_ws_per_s * FLOP_PER_WS / 1e12,
                "peak": FP64_NOMINAL / 1e12, "unit": "TFLOP/s", "frac": per_gpu_ws_per_s * FLOP_PER_WS / FP64_NOMINAL,
                "peak_kind": "nominal FP64 pipe, 148 SM x 64 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no FP64 entry: SURVEY 8d fallback)",
                "self_measured_peak": fp64_peak / 1e12, "frac_of_self_measured_peak": per_gpu_ws_per_s * FLOP_PER_WS / fp64_peak,
                "self_measured_peak_kind": "builder-side micro-benchmark in this process: register-resident DFMA chains (pvd_measure_fp64_peak)",
                "avg_launch_ms": kernel_ms, "flop_per_walker_step": FLOP_PER_WS, "traffic": None,
                "hbm": {"achieved": per_gpu_ws_per_s * BYTES_PER_WS / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": per_gpu_ws_per_s * BYTES_PER_WS / 1e9 / hbm_peak, "peak_kind": peak_kind,
                        "bytes_per_walker_step": BYTES_PER_WS}}
    # dram__bytes_read + dram__bytes_write of ONE launch of this kernel from the committed `ncu --set full` capture (not this run)
    for name in ("r02_step_kernel_traffic.json", "r01_step_kernel_traffic.json"):
        prof = os.path.join(ROOT, "profiles", name)
        if os.path.exists(prof):
            try:
                roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
                roofline["traffic_source"] = "profiles/" + name + " (ncu --set full capture of the same kernel, per launch)"
                break
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
               "unit": "walker-steps/s", "kernel": "k_run_discrete<PotH2O> (resident: the 2000 steps are one launch)"}
        tut["roofline"] = {"bound": "latency", "note": "625 tiles of 32 walkers on 148 SMs: one tile per warp and step; the step time is the "
                           "dependency chain tile -> Vref -> next tile, not a pipe", "achieved": tut["value"] * FLOP_PER_WS / 1e12,
                           "peak": FP64_NOMINAL / 1e12, "unit": "TFLOP/s", "frac": tut["value"] * FLOP_PER_WS / FP64_NOMINAL}
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
        for lab, ent in list(others.items()):
            if isinstance(ent, dict) and "walker_steps_per_s" in ent:
                ent["roofline"] = other_roofline(lab, ent["walker_steps_per_s"], hbm_peak, fp64_peak)
        others["note"] = ("steady-state device-resident loop, CUDA events inside pvd_sim_run; c1 = 1-D HO discrete, c3 = H2O continuous "
                          "(1e6/GPU), c4 = H2O importance sampling with finite-difference drift (1.25e6/GPU), c4a = the same with the reference's analytic derivatives, c5 = (H2O)2 NN PES on tcgen05 "
                          "(1.25e7/GPU)")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _ = time_cpu_reference()
        except Exception as e:
            cpu = {"value": None, "unit": "walker-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + repr(e)}

    e2e_run = None
    if world == 1 and not args.no_e2e_run:
        try:
            e2e_run = time_dmc_sim_run(cpu["value"] if cpu and cpu.get("value") else None)
        except Exception as e:             # never lose the headline line to a side measurement
            e2e_run = {"error": repr(e)}

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
                       if world > 1 else "single GPU, one kernel launch per time step (deferred compaction: the NEXT step's loads do the np.repeat gather; the output of the last step is compacted when the ensemble is next read -- pvd_sim_download in the e2e leg, one 0.05 ms kernel per read, not per step)",
                       "hardware_warmup": (f"{hw_steps} untimed steps of a scratch ensemble of the same shape before the W warm-up steps "
                                           "(SM clocks and NVLink links out of their idle states)") if hw_steps else "none"},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
            "parity_check": parity_check, "e2e_run": e2e_run,
            "tutorial_20k": tut, "other_configs": others, "final_population": int(st["n"])}
    args.restore_stdout()
    print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
