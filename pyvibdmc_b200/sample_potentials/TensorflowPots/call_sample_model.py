"""sample_h4o2_pot(cds, model, extra_args): Coulomb-matrix descriptor + the shipped 15-120-120-120-1
network on the GPU (replaces TensorflowPots/call_sample_model.py:4-9, which needs TensorFlow).
`model` may be the packed float32 weight vector (see load_packed_weights) or any object with
get_weights() returning the Keras list [W0,b0,W1,b1,W2,b2,W3,b3]."""
import os

import numpy as np

from pyvibdmc_b200 import kernels as _K, _capi

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_packed_weights(path=None):
    """float32 [W0(15x120) b0 W1(120x120) b1 W2(120x120) b2 W3(120) b3], extracted from sample_h4o2_nn.h5."""
    return np.load(path or os.path.join(_HERE, "sample_h4o2_nn_packed.npy"))


def _pack(model):
    if isinstance(model, np.ndarray):
        return np.ascontiguousarray(model, dtype=np.float32).ravel()
    return np.concatenate([np.asarray(w, dtype=np.float32).ravel() for w in model.get_weights()])


def sample_h4o2_pot(cds, model, extra_args=None):
    _K.nn_h4o2_set_weights(_pack(model))
    return _K.nn_h4o2(cds)


sample_h4o2_pot._pvd_builtin = lambda kw, model: {"potential": _capi.POT_NN_H4O2, "weights": _pack(model)}
