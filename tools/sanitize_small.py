#!/usr/bin/env python
"""Small run of every kernel touched by the ziggurat / PDL work, for compute-sanitizer (memcheck, racecheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from pyvibdmc_b200 import _capi, kernels as K  # noqa: E402
from pyvibdmc_b200.simulation_utilities import Constants  # noqa: E402

eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
mH, mO = Constants.mass("H"), Constants.mass("O")
n = 5000
for weighting in ("discrete", "continuous"):
    for mode in (_capi.RNG_ZIGGURAT, _capi.RNG_FP64):
        sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_H2O_PS, seed=3, rng_mode=mode, weighting=weighting)
        sim.upload(np.broadcast_to(eq * 1.01, (n, 3, 3)).copy())
        sim.run(12)
        sim.sync()
        print(weighting, mode, sim.state()["n"], sim.state()["vref"])
        sim.close()
z = K.normals(20000, 18, seed=1, step=2, rng_mode=_capi.RNG_ZIGGURAT)
print("normals", float(z.std()))
b = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_EXTERNAL, seed=42)
start = np.broadcast_to(eq * 1.01, (n, 3, 3)).copy()
b.upload(start)
b.set_pots(K.pes_h2o(start))
for _ in range(4):
    cds = b.ext_move()
    b.ext_finish(K.pes_h2o(cds))
print("external", b.state()["n"])
b.close()
