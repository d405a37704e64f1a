#!/usr/bin/env python
"""Read the shipped Keras model (sample_potentials/TensorflowPots/sample_h4o2_nn.h5, saved by TF 2.4:
Sequential[Input(15) -> Dense(120,swish) x3 -> Dense(1,relu)], float32) with the in-tree HDF5
reader and pack the weights as one float32 vector
    [W0 (15x120) | b0 (120) | W1 (120x120) | b1 | W2 (120x120) | b2 | W3 (120) | b3 (1)]
(row-major (in,out) kernels, as Keras stores them) -> pyvibdmc_b200/sample_potentials/sample_h4o2_nn_packed.npy
Run in the build container only."""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from pyvibdmc_b200.simulation_utilities import h5lite  # noqa: E402

SRC = "/root/reference/pyvibdmc/sample_potentials/TensorflowPots/sample_h4o2_nn.h5"
w = h5lite.read_h5(SRC)
parts = []
for layer, shape in (("dense", (15, 120)), ("dense_1", (120, 120)), ("dense_2", (120, 120)), ("dense_3", (120, 1))):
    k = w[f"model_weights/{layer}/{layer}/kernel:0"]
    b = w[f"model_weights/{layer}/{layer}/bias:0"]
    assert k.shape == shape and b.shape == (shape[1],) and k.dtype == np.float32
    parts += [k.ravel(), b.ravel()]
packed = np.concatenate(parts).astype(np.float32)
assert packed.size == 31081
out = os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials", "sample_h4o2_nn_packed.npy")
np.save(out, packed)
print("wrote", out, packed.size, "floats")
