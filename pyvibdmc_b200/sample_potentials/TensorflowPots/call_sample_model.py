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


def load_keras_h5(path):
    """The Keras weight list [W0, b0, W1, b1, ...] of a Sequential model of Dense layers saved as .h5 (tf.keras model.save, the
    format of the reference's sample_h4o2_nn.h5), read with the in-tree HDF5 reader -- no TensorFlow, no h5py.  Layers are taken
    in Keras' naming order dense, dense_1, dense_2, ..."""
    from pyvibdmc_b200.simulation_utilities import h5lite
    w = h5lite.read_h5(path)
    names = sorted({k.split("/")[1] for k in w if k.startswith("model_weights/") and k.endswith("kernel:0")},
                   key=lambda n: int(n.split("_")[1]) if "_" in n else 0)
    out = []
    for n in names:
        out += [np.asarray(w[f"model_weights/{n}/{n}/kernel:0"]), np.asarray(w[f"model_weights/{n}/{n}/bias:0"])]
    return out


def _pack(model):
    """packed float32 vector from: the packed vector itself, a path to a Keras .h5 / packed .npy, a Keras weight list, or any
    object with get_weights()."""
    if isinstance(model, (str, os.PathLike)):
        model = np.load(model) if str(model).endswith(".npy") else load_keras_h5(model)
    if isinstance(model, np.ndarray):
        return np.ascontiguousarray(model, dtype=np.float32).ravel()
    weights = model if isinstance(model, (list, tuple)) else model.get_weights()
    shapes = [tuple(np.shape(w)) for w in weights]
    if shapes != [(15, 120), (120,), (120, 120), (120,), (120, 120), (120,), (120, 1), (1,)]:
        raise ValueError("the tensor-core NN path is built for the shipped 15-120-120-120-1 network (6 atoms, Coulomb descriptor); "
                         f"got layers {shapes}.  Any other model runs on the GPU through Potential_Direct(device=True) "
                         "(zero-copy CUDA tensors) with DistIt descriptors")
    return np.concatenate([np.asarray(w, dtype=np.float32).ravel() for w in weights])


def sample_h4o2_pot(cds, model, extra_args=None):
    _K.nn_h4o2_set_weights(_pack(model))
    return _K.nn_h4o2(cds)


sample_h4o2_pot._pvd_builtin = lambda kw, model: {"potential": _capi.POT_NN_H4O2, "weights": _pack(model)}
