#!/usr/bin/env python
"""NN water-dimer potential: walkers/s of the tcgen05 kernel vs the float32 CUDA-core kernel (kernel time from CUDA
events inside pvd_nn_h4o2, host<->device copies excluded) and the tensor-core roofline fraction
(61 440 algorithmic flop / walker; 3 fp16 cross-term MMAs per product are executed for fp32-class accuracy)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from pyvibdmc_b200 import kernels as K  # noqa: E402

dimer = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                      [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]]) / 0.529177
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
x = dimer[None] + np.random.default_rng(0).normal(0, 0.1, size=(n, 6, 3))
w = np.load(os.path.join(os.path.dirname(K.__file__), "sample_potentials", "TensorflowPots", "sample_h4o2_nn_packed.npy"))
K.nn_h4o2_set_weights(w)
out = {}
for mode in ("tcgen05", "cuda_cores"):
    K.nn_config(path=mode)
    best = 1e9
    for _ in range(4):
        v = K.nn_h4o2(x)
        best = min(best, K.last_kernel_ms())
    out[mode] = {"ms": best, "walkers_per_s": n / (best * 1e-3), "algorithmic_tflops": n * 61440 / (best * 1e-3) / 1e12,
                 "mean_cm1": float(v.mean() / 4.556335281212229e-6)}
peaks = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["bf16_tflops"] if os.path.exists(peaks) else 1590.0
out["tensor_peak_tflops_bf16"] = peak
out["tcgen05"]["frac_of_bf16_peak_algorithmic"] = out["tcgen05"]["algorithmic_tflops"] / peak
# executed on the tensor cores: 3 cross terms of the two-piece split, N and K padded to 128 (layer 0: K padded to 16)
executed = 3 * 2.0 * 128 * (16 + 128 + 128)          # flop per walker
out["tcgen05"]["executed_tflops"] = executed * out["tcgen05"]["walkers_per_s"] / 1e12
out["tcgen05"]["frac_of_f16_peak_executed"] = out["tcgen05"]["executed_tflops"] / peak
print(json.dumps(out, indent=1))
