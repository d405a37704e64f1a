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

# ---- round 2: importance sampling (finite-difference and analytic stencils), the tcgen05 NN surface, descendant weighting,
# DistIt descriptors.  Which discrete step kernel runs is chosen by the environment (tools/r02_sanitize.sh):
#   default: resident k_run_discrete; PVD_RUN_MAX_WALKERS=0: k_step_gather + k_gather_materialise;
#   PVD_NO_RESIDENT=1 PVD_NO_GATHER=1: k_step_discrete
import importlib.util  # noqa: E402
sp = os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials")
spec = importlib.util.spec_from_file_location("call_trl_h2o_b200", os.path.join(sp, "FortPots", "Partridge_Schwenke_H2O", "call_trl_h2o.py"))
trl = importlib.util.module_from_spec(spec)
spec.loader.exec_module(trl)
for trial, table in ((_capi.TRIAL_H2O_FD, trl.packed_table()), (_capi.TRIAL_H2O_AN, trl.packed_table_analytic())):
    s = K.DeviceSim(3, 3, [mH, mH, mO], 3000, 1.0, _capi.POT_H2O_PS, seed=5, trial=trial)
    s.set_trial_table(table)
    s.upload(np.broadcast_to(eq * 1.01, (3000, 3, 3)).copy())
    s.run(6)
    s.sync()
    print("imp samp", trial, s.state()["n"], s.state()["vref"])
    s.close()
dimer = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                  [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]]) / 0.529177
w = np.load(os.path.join(sp, "TensorflowPots", "sample_h4o2_nn_packed.npy"))
K.nn_h4o2_set_weights(w)
xs = dimer[None] + np.random.default_rng(0).normal(0, 0.05, size=(1000, 6, 3))
print("nn stand-alone", float(K.nn_h4o2(xs).mean()))
s = K.DeviceSim(6, 3, [mO, mH, mH] * 2, 1500, 5.0, _capi.POT_NN_H4O2, seed=6)
s.set_nn_weights(w)
s.upload(np.broadcast_to(dimer, (1500, 6, 3)).copy())
s.run(4)
s.sync()
print("nn sim", s.state()["n"], s.state()["vref"])
s.close()
s = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_H2O_PS, seed=8)
s.upload(np.broadcast_to(eq * 1.01, (n, 3, 3)).copy())
s.run(5)
n_parent = s.state()["n"]
s.dw_begin()
s.run(7)
dw = s.dw_end(n_parent)
print("descendant weights", int(dw.sum()), s.state()["n"])
s.close()
