#!/usr/bin/env python
"""BASELINE config 2 on the GPU: H2O / Partridge-Schwenke, discrete weighting, 20,000 walkers x 20,000 steps,
dt = 5 a.u., five seeds.  ZPE = mean(Vref[T//4:]) (the reference's 'Approximate ZPE', pyvibdmc.py:908) with a
blocked standard error, compared with the shipped tutorial runs (4634.1 +- 2.2 cm^-1, 8000 x 5000).
Also config 1 (1-D HO, 1000 x 5000, dt = 10; analytic 1850 cm^-1 + time-step bias).  Writes JSON to argv[1]."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from pyvibdmc_b200 import kernels as K, _capi  # noqa: E402

WN = 4.556335281212229e-6
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
AMU = 1.0 / 6.02213670000e23 / 9.10938970000e-28
M = np.array([1.00782503, 1.00782503, 15.99491462]) * AMU


def blocked_sem(x, nblocks=20):
    b = np.array_split(x, nblocks)
    means = np.array([c.mean() for c in b])
    return means.std(ddof=1) / np.sqrt(nblocks)


def run(kind, n0, T, dt, seed, rng=_capi.RNG_DEFAULT):
    if kind == "h2o":
        sim = K.DeviceSim(3, 3, M, n0, dt, _capi.POT_H2O_PS, seed=seed, rng_mode=rng, stats_ring=T + 8)
        sim.upload(np.repeat(EQ[None] * 1.01, n0, axis=0))
    else:
        mu = M[0] * M[2] / (M[0] + M[2])
        om = 3700.0 * WN
        sim = K.DeviceSim(1, 1, [mu], n0, dt, _capi.POT_HARMONIC, pot_params=[(0.5 * mu) * om ** 2], seed=seed, rng_mode=rng,
                          stats_ring=T + 8)
        sim.upload(np.zeros((n0, 1, 1)))
    t0 = time.time()
    sim.run(T)
    sim.sync()
    wall = time.time() - t0
    st = sim.stats(0, T)
    sim.close()
    v = st["vref"][T // 4:] / WN
    return {"seed": seed, "zpe": float(v.mean()), "sem": float(blocked_sem(v)), "wall_s": wall,
            "walker_steps_per_s": float(st["pop"].sum() / wall), "pop_mean": float(st["pop"].mean())}


def main():
    out = {"config2_h2o_20000x20000_dt5": [run("h2o", 20000, 20000, 5.0, s) for s in range(5)],
           "config2_boxmuller_fp64": [run("h2o", 20000, 20000, 5.0, s, _capi.RNG_FP64) for s in range(5)],
           "config2_fast_rng": [run("h2o", 20000, 20000, 5.0, s, _capi.RNG_FAST) for s in range(5)],
           "tutorial_h2o_8000x5000_dt5": [run("h2o", 8000, 5000, 5.0, 100 + s) for s in range(5)],
           "config1_ho_1000x5000_dt10": [run("ho", 1000, 5000, 10.0, s) for s in range(5)],
           "reference": {"shipped_tutorial_zpe_cm1": [4635.33, 4637.79, 4631.96, 4632.38, 4632.92], "mean": 4634.1, "std": 2.2,
                         "ho_analytic": 1850.0}}
    for k, v in out.items():
        if isinstance(v, list):
            z = np.array([r["zpe"] for r in v])
            print(k, "ZPE mean %.2f  std over seeds %.2f  mean blocked sem %.2f  wall %.2fs" %
                  (z.mean(), z.std(ddof=1), np.mean([r["sem"] for r in v]), np.mean([r["wall_s"] for r in v])))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
