#!/usr/bin/env python
"""The driver's bench shape: fresh ensemble, 5 warm-up steps, then ONE 20-step segment (repeated on fresh ensembles)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from pyvibdmc_b200 import _capi, kernels as K
from pyvibdmc_b200.simulation_utilities import Constants
eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
n = int(os.environ.get("AB_WALKERS", "1000000"))
mH, mO = Constants.mass("H"), Constants.mass("O")
start = np.broadcast_to(eq * 1.01, (n, 3, 3)).copy()
for rep in range(4):
    sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_H2O_PS, seed=7 + rep, capacity=int(1.5 * n) + 1024)
    if os.environ.get("AB_MODE"):
        sim.set_resident(int(os.environ["AB_MODE"]))
    sim.upload(start)
    sim.run(int(os.environ.get("AB_WARM", "5"))); sim.sync()
    sim.run(20); sim.sync()
    ms = sim.last_run_ms()
    st = sim.stats(0, 25 if int(os.environ.get("AB_WARM", "5")) == 5 else 20)
    print(json.dumps({"rep": rep, "ms20": ms, "us_per_step": 50 * ms, "pop_first_last": [float(st["pop"][0]), float(st["pop"][-1])], "pop_min": float(st["pop"].min())}))
    sim.close()
