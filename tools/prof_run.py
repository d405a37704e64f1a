#!/usr/bin/env python
"""Profile driver: H2O discrete, N walkers; one warm segment then one short resident launch (ncu: -k regex:k_run_discrete -s 1 -c 1)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from pyvibdmc_b200 import _capi, kernels as K
from pyvibdmc_b200.simulation_utilities import Constants

eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
n = int(os.environ.get("AB_WALKERS", "1000000"))
steps = int(os.environ.get("AB_STEPS", "20"))
mH, mO = Constants.mass("H"), Constants.mass("O")
sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_H2O_PS, seed=7)
if os.environ.get("AB_MODE"):
    sim.set_resident(int(os.environ["AB_MODE"]))
sim.upload(np.broadcast_to(eq * 1.01, (n, 3, 3)).copy())
sim.run(100)
sim.sync()
sim.run(steps)
sim.sync()
print(sim.last_run_ms() / steps, sim.state())
