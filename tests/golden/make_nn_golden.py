#!/usr/bin/env python
"""Golden vectors that pin the NN-PES row independently of any float32 implementation: the shipped network
(sample_potentials/TensorflowPots/sample_h4o2_nn.h5, read with the in-tree HDF5 reader when /root/reference is present, else the
packed copy of the same weights) evaluated in FLOAT64 on 512 water-dimer geometries, next to a float32 NumPy evaluation (what
TensorFlow's float32 forward pass is: same precision, another summation order).  Run in the build container:
    python tests/golden/make_nn_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import dmc_oracle as O  # noqa: E402
from pyvibdmc_b200.simulation_utilities import h5lite  # noqa: E402

PACKED = os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials", "TensorflowPots", "sample_h4o2_nn_packed.npy")
H5 = "/root/reference/pyvibdmc/sample_potentials/TensorflowPots/sample_h4o2_nn.h5"


def unpack(p):
    o, out = 0, []
    for k, m in ((15, 120), (120, 120), (120, 120), (120, 1)):
        W = p[o:o + k * m].reshape(k, m); o += k * m
        b = p[o:o + m]; o += m
        out.append((W, b))
    return out


packed = np.load(PACKED)
source = "packed copy"
if os.path.exists(H5):
    w = h5lite.read_h5(H5)
    parts = []
    for layer in ("dense", "dense_1", "dense_2", "dense_3"):
        parts += [w[f"model_weights/{layer}/{layer}/kernel:0"].ravel(), w[f"model_weights/{layer}/{layer}/bias:0"].ravel()]
    from_h5 = np.concatenate(parts).astype(np.float32)
    assert np.array_equal(from_h5, packed), "the packed weights are not the reference's .h5"
    source = "reference .h5 (identical to the packed copy)"
dimer = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                  [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]]) / 0.529177
rng = np.random.default_rng(20261018)
coords = dimer[None] + rng.normal(0, 0.08, size=(512, 6, 3))
coords[0] = dimer
desc = O.coulomb_descriptor(coords, [8, 1, 1, 8, 1, 1])
e64 = O.nn_forward_f64(desc, unpack(packed))
e32 = O.nn_forward_f32(desc, unpack(packed))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nn_h4o2_f64_golden.npz"), coords=coords, e64=e64, e32=e32)
print("weights from:", source)
print("float32 NumPy vs float64: max |err| / max E = %.3g" % (np.abs(e32 - e64).max() / np.abs(e64).max()))
