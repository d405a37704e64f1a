#!/usr/bin/env python
"""Steady-state walker-steps/s of the device-resident loop for the BASELINE configurations other than the
headline one (bench.py measures config 2's kernel): 1-D HO discrete (C1), H2O continuous (C3), H2O importance
sampling with finite-difference drift (C4), water dimer on the NN surface (C5).  One GPU; sizes are per-GPU shards.

Timing: CUDA events inside pvd_sim_run (pvd_sim_last_run_ms), after warm-up; every run is long enough that the
walker arrays (>= 150 MB at 1e6 walkers) do not stay in L2 between steps.
usage: config_bench.py [--quick] [--only c1,c3,c4,c5] [--steps K]
"""
import argparse
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from pyvibdmc_b200 import _capi, kernels as K  # noqa: E402
from pyvibdmc_b200.simulation_utilities import Constants  # noqa: E402

EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
DIMER = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                  [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]]) / 0.529177
WN = 4.556335281212229e-6
SP = os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials")


def water_table():
    spec = importlib.util.spec_from_file_location("call_trl_h2o_b200", os.path.join(SP, "FortPots", "Partridge_Schwenke_H2O", "call_trl_h2o.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.packed_table()


EQUIL = [-1]


def timed(sim, equil, steps, reps=3):
    sim.run(equil if EQUIL[0] < 0 else EQUIL[0])
    sim.sync()
    best = None
    for _ in range(reps):
        s0 = sim.state()["step"]
        sim.run(steps)
        sim.sync()
        ms = sim.last_run_ms()
        pop = sim.stats(s0, steps)["pop"].astype(np.float64).sum()
        rate = pop / (ms * 1e-3)
        if best is None or rate > best["walker_steps_per_s"]:
            best = {"ms_per_step": ms / steps, "walker_steps_per_s": rate, "mean_population": pop / steps}
    st = sim.stats(sim.state()["step"] - steps, steps)
    best["mean_vref_cm1"] = float(st["vref"].mean() / WN)
    return best


def collect(only=("c1", "c3", "c4", "c4a", "c5"), large_only=False, steps=100, equil=-1, scale=1.0):
    """Returns {config label: {ms_per_step, walker_steps_per_s, mean_population, mean_vref_cm1}}."""
    only = set(only)
    EQUIL[0] = equil
    sizes = (lambda small, large: (large,) if large_only else (small, large))
    mH, mO = Constants.mass("H"), Constants.mass("O")
    out = {}

    if "c1" in only:
        mu = Constants.reduced_mass("O-H")
        om = 3700 * WN
        for n in sizes(1000, int(1e6 * scale)):
            sim = K.DeviceSim(1, 1, [mu], n, 10.0, _capi.POT_HARMONIC, pot_params=[(0.5 * mu) * om ** 2], seed=1)
            sim.upload(np.zeros((n, 1, 1)))
            out[f"c1_ho_discrete_{n}"] = timed(sim, 200, steps * (10 if n < 10000 else 1))
            sim.close()

    if "c3" in only:
        for n in sizes(20000, int(1e6 * scale)):
            sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_H2O_PS, weighting="continuous", seed=2)
            sim.upload(np.broadcast_to(EQ * 1.01, (n, 3, 3)).copy())
            r = timed(sim, 300, steps)
            st = sim.stats(sim.state()["step"] - steps, steps)
            r["branched_per_step"] = float(st["births"].mean())
            out[f"c3_h2o_continuous_{n}"] = r
            sim.close()

    if "c4" in only:
        tab = water_table()
        for n in sizes(20000, int(1.25e6 * scale)):
            sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 1.0, _capi.POT_H2O_PS, trial=_capi.TRIAL_H2O_FD, seed=3)
            sim.set_trial_table(tab)
            sim.upload(np.broadcast_to(EQ * 1.01, (n, 3, 3)).copy())
            out[f"c4_h2o_impsamp_fd_{n}"] = timed(sim, 200, steps)
            sim.close()

    if "c4a" in only:
        spec = importlib.util.spec_from_file_location("call_trl_h2o_b200", os.path.join(SP, "FortPots", "Partridge_Schwenke_H2O", "call_trl_h2o.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for n in sizes(20000, int(1.25e6 * scale)):
            sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 1.0, _capi.POT_H2O_PS, trial=_capi.TRIAL_H2O_AN, seed=3)
            sim.set_trial_table(mod.packed_table_analytic())
            sim.upload(np.broadcast_to(EQ * 1.01, (n, 3, 3)).copy())
            out[f"c4a_h2o_impsamp_analytic_{n}"] = timed(sim, 200, steps)
            sim.close()

    if "c5" in only:
        w = np.load(os.path.join(SP, "TensorflowPots", "sample_h4o2_nn_packed.npy"))
        for n in sizes(int(1e6 * scale), int(1.25e7 * scale)):
            sim = K.DeviceSim(6, 3, [mO, mH, mH] * 2, n, 5.0, _capi.POT_NN_H4O2, seed=4)
            sim.set_nn_weights(w)
            sim.upload(np.broadcast_to(DIMER, (n, 6, 3)).copy())
            out[f"c5_dimer_nn_{n}"] = timed(sim, 50, max(steps // 4, 5), reps=2)
            sim.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="c1,c3,c4,c4a,c5")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--equil", type=int, default=-1, help="override the equilibration length (profiling runs)")
    ap.add_argument("--large-only", action="store_true")
    args = ap.parse_args()
    out = collect(args.only.split(","), args.large_only, args.steps, args.equil, 0.1 if args.quick else 1.0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
