#!/usr/bin/env python
"""Time pvd_sim_run segments of several lengths (H2O discrete, 1e6 walkers): per-step cost of short segments."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from pyvibdmc_b200 import _capi, kernels as K
from pyvibdmc_b200.simulation_utilities import Constants
eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
n = int(os.environ.get("AB_WALKERS", "1000000"))
mH, mO = Constants.mass("H"), Constants.mass("O")
sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_H2O_PS, seed=7)
if os.environ.get("AB_MODE"):
    sim.set_resident(int(os.environ["AB_MODE"]))
sim.upload(np.broadcast_to(eq * 1.01, (n, 3, 3)).copy())
sim.run(300); sim.sync()
for seg in (1, 2, 5, 20, 100, 400):
    best = 1e9
    for _ in range(5):
        sim.run(seg); sim.sync()
        best = min(best, sim.last_run_ms())
    print(json.dumps({"segment": seg, "ms": best, "us_per_step": 1e3 * best / seg}))
