#!/usr/bin/env python
"""Accuracy of the NN-PES kernels against a float64 evaluation of the same float32 weights (and the float32 oracle):
max and rms relative error over a cloud of water-dimer geometries.  usage: nn_accuracy.py [n]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from pyvibdmc_b200 import kernels as K  # noqa: E402

WN = 4.556335281212229e-6
dimer = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                  [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]]) / 0.529177
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
x = dimer[None] + np.random.default_rng(1).normal(0, 0.08, size=(n, 6, 3))
P = np.load(os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials", "TensorflowPots", "sample_h4o2_nn_packed.npy"))
K.nn_h4o2_set_weights(P)


def forward(desc, dtype):
    o = 0
    h = desc.astype(dtype)
    for (k, m) in ((15, 120), (120, 120), (120, 120)):
        W = P[o:o + k * m].reshape(k, m).astype(dtype); o += k * m
        b = P[o:o + m].astype(dtype); o += m
        z = h @ W + b
        h = z / (1 + np.exp(-z))
    W = P[o:o + 120].astype(dtype); o += 120
    return np.maximum(h @ W + P[o].astype(dtype), 0) * dtype(WN)


zs = np.array([8., 1, 1, 8, 1, 1])
iu = np.triu_indices(6, 1)
d = np.linalg.norm(x[:, iu[0]] - x[:, iu[1]], axis=2)
desc = (zs[iu[0]] * zs[iu[1]])[None] / d
ref64 = forward(desc.astype(np.float32), np.float64)
ref32 = forward(desc.astype(np.float32), np.float32).astype(np.float64)
out = {}
scale = np.abs(ref64).max()
for name, env in (("tc2_4terms", {"PVD_NN_TERMS": "4"}), ("tc2_3terms", {"PVD_NN_TERMS": "3"}), ("tc1", {"PVD_NN_TC1": "1"}), ("cuda_cores_fp32", {"PVD_NN_FP32": "1"})):
    for k in ("PVD_NN_TERMS", "PVD_NN_TC1", "PVD_NN_FP32"):
        os.environ.pop(k, None)
    os.environ.update(env)
    v = K.nn_h4o2(x)
    e = (v - ref64) / scale
    out[name] = {"max_err_over_max": float(np.abs(e).max()), "rms_err_over_max": float(np.sqrt((e ** 2).mean())), "kernel_ms": K.last_kernel_ms()}
e = (ref32 - ref64) / scale
out["numpy_float32"] = {"max_err_over_max": float(np.abs(e).max()), "rms_err_over_max": float(np.sqrt((e ** 2).mean()))}
out["max_energy_cm1"] = float(scale / WN)
print(json.dumps(out, indent=1))
