"""Potential plug-in wrappers with the reference's constructors and getpot() contract
(reference: simulation_utilities/potential_manager.py:12-253).

Two kinds of user functions are handled:
  * the shipped sample potentials (pyvibdmc_b200/sample_potentials/...): their Python functions run
    the CUDA kernels and carry a `_pvd_builtin` descriptor, which lets DMC_Sim keep the whole
    propagation loop on the GPU (gpu_spec() below);
  * any other user function: called on the host exactly as the reference does (optionally through
    a multiprocessing pool); DMC_Sim then moves / weights / branches on the GPU and only hands the
    coordinates to the function and takes the energies back.
"""
import importlib
import multiprocessing as mp
import os
import sys
import time
from itertools import repeat

import numpy as np

__all__ = ['Potential', 'Potential_NoMP', 'NN_Potential', 'Potential_Direct']


def _import_function(directory, python_file, name):
    """chdir into `directory`, import `python_file`, return (function, previous cwd)."""
    here = os.getcwd()
    os.chdir(directory)
    try:
        sys.path.insert(0, os.getcwd())
        module = importlib.import_module(python_file.split(".")[0])
        return getattr(module, name)
    finally:
        os.chdir(here)


def _builtin_spec(fn, kwargs, model=None):
    """GPU descriptor of a shipped sample potential, or None for ordinary user functions."""
    spec = getattr(fn, "_pvd_builtin", None)
    if callable(spec):
        spec = spec(kwargs, model) if model is not None else spec(kwargs)
    return spec


_WORKER_FN = None


def _worker_init(directory, python_file, name):
    global _WORKER_FN
    os.chdir(directory)
    sys.path.insert(0, os.getcwd())
    _WORKER_FN = getattr(importlib.import_module(python_file.split(".")[0]), name)


def _worker_call(cds, kwargs=None):
    return _WORKER_FN(cds) if kwargs is None else _WORKER_FN(cds, kwargs)


class _TimestepMixin:
    def _init_timestep(self):
        if self.pass_timestep:
            self.ct = 0
            if self.pot_kwargs is None:
                self.pot_kwargs = {'timestep': 0}
            else:
                self.pot_kwargs['timestep'] = 0

    def _finish_call(self, v, start, timeit):
        if self.pass_timestep:
            self.pot_kwargs['timestep'] += 1
        return (v, time.time() - start) if timeit else v


class Potential(_TimestepMixin):
    """Potential(potential_function, potential_directory, python_file, num_cores=1, pass_timestep=False,
    pot_kwargs=None) -- the multiprocessing manager of the reference (potential_manager.py:12-104)."""

    def __init__(self, potential_function, potential_directory, python_file, num_cores=1, pass_timestep=False,
                 pot_kwargs=None):
        self.potential_function = potential_function
        self.python_file = python_file
        self.potential_directory = potential_directory
        self.num_cores = num_cores
        self.pass_timestep = pass_timestep
        self.pot_kwargs = pot_kwargs
        self._init_timestep()
        self._init_pool()

    def _init_pool(self):
        if self.num_cores <= 0:
            print('Weird number of cores specified. Defaulting to 1...')
            self.num_cores = 1
        self._pot = _import_function(self.potential_directory, self.python_file, self.potential_function)
        self._potPool = None
        if _builtin_spec(self._pot, self.pot_kwargs) is None:
            # host function: worker processes live in the potential directory, like the reference's
            ctx = mp.get_context("fork")
            self._potPool = ctx.Pool(self.num_cores, initializer=_worker_init,
                                     initargs=(self.potential_directory, self.python_file, self.potential_function))

    @property
    def pool(self):
        return self._potPool

    def gpu_spec(self):
        return _builtin_spec(self._pot, self.pot_kwargs)

    def getpot(self, cds, timeit=False):
        start = time.time()
        if self._potPool is not None:
            chunks = np.array_split(cds, self.num_cores)
            if self.pot_kwargs is not None:
                res = self._potPool.starmap(_worker_call, zip(chunks, repeat(self.pot_kwargs, len(chunks))))
            else:
                res = self._potPool.map(_worker_call, chunks)
            v = np.concatenate(res)
        else:                                          # shipped sample potential: one CUDA call
            v = self._pot(cds) if self.pot_kwargs is None else self._pot(cds, self.pot_kwargs)
        return self._finish_call(v, start, timeit)

    def mp_close(self):
        if self._potPool is not None:
            self._potPool.close()
            self._potPool.join()

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_potPool'] = None
        return state


class Potential_NoMP(_TimestepMixin):
    """Potential_NoMP(potential_function, potential_directory, python_file, pass_timestep=False, ch_dir=False,
    pot_kwargs=None) (potential_manager.py:107-174)."""

    def __init__(self, potential_function, potential_directory, python_file, pass_timestep=False, ch_dir=False,
                 pot_kwargs=None):
        self.potential_function = potential_function
        self.python_file = python_file
        self.pass_timestep = pass_timestep
        self.potential_directory = potential_directory
        self.pot_kwargs = pot_kwargs
        self._init_timestep()
        self.ch_dir = ch_dir
        self._curdir = os.getcwd()
        self._pot = _import_function(self.potential_directory, self.python_file, self.potential_function)

    def gpu_spec(self):
        return _builtin_spec(self._pot, self.pot_kwargs)

    def _call(self, cds):
        return self._pot(cds, self.pot_kwargs) if self.pot_kwargs is not None else self._pot(cds)

    def getpot(self, cds, timeit=False):
        start = time.time()
        if self.ch_dir:
            os.chdir(self.potential_directory)
            try:
                v = self._call(cds)
            finally:
                os.chdir(self._curdir)
        else:
            v = self._call(cds)
        return self._finish_call(v, start, timeit)


class NN_Potential(Potential_NoMP):
    """NN_Potential(potential_function, potential_directory, python_file, model, ch_dir=False, pot_kwargs=None,
    pass_timestep=False): user function signature f(cds, model[, kwargs]) (potential_manager.py:177-214)."""

    def __init__(self, potential_function, potential_directory, python_file, model, ch_dir=False, pot_kwargs=None,
                 pass_timestep=False):
        super().__init__(potential_function, potential_directory, python_file, pass_timestep, ch_dir, pot_kwargs)
        self.model = model
        self.ch_dir = ch_dir
        self.pot_kwargs = pot_kwargs

    def gpu_spec(self):
        return _builtin_spec(self._pot, self.pot_kwargs, self.model)

    def getpot(self, cds, timeit=False):
        start = time.time()
        v = self._pot(cds, self.model, self.pot_kwargs) if self.pot_kwargs is not None else self._pot(cds, self.model)
        return self._finish_call(v, start, timeit)


class Potential_Direct(_TimestepMixin):
    """Potential_Direct(potential_function: callable, pot_kwargs=None, pass_timestep=False) (potential_manager.py:217-253).

    device=True (B200 extension): the callable is a GPU model.  It receives the walkers as a CUDA tensor
    -- a zero-copy `torch.Tensor` view (n, atoms, 3) float64 of the simulation's buffer in HBM (`as_torch=False`: the raw
    `kernels.DeviceArray`, which speaks `__cuda_array_interface__`, for CuPy / numba) -- and returns the energies as a device
    array (torch / CuPy / anything with `__cuda_array_interface__` or `__dlpack__`), float64 Hartree, shape (n,).  Nothing
    crosses PCIe: the reference's NN_Potential exists precisely so that user models run on GPUs (:177-214)."""

    def __init__(self, potential_function, pot_kwargs=None, pass_timestep=False, device=False, as_torch=True):
        self.potential_function = potential_function
        self.pot_kwargs = pot_kwargs
        self.pass_timestep = pass_timestep
        self.device = bool(device)
        self.as_torch = bool(as_torch)
        self._init_timestep()

    def gpu_spec(self):
        return _builtin_spec(self.potential_function, self.pot_kwargs)

    def getpot(self, cds, timeit=False):
        start = time.time()
        if self.device and not hasattr(cds, "__cuda_array_interface__"):
            # host coordinates (e.g. the reference-style call on sim.walkers): through the GPU model and back
            import torch
            cds = torch.as_tensor(np.ascontiguousarray(cds, dtype=np.float64), device="cuda")
            v = self.potential_function(cds, self.pot_kwargs) if self.pot_kwargs is not None else self.potential_function(cds)
            return self._finish_call(v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v), start, timeit)
        v = self.potential_function(cds, self.pot_kwargs) if self.pot_kwargs is not None else self.potential_function(cds)
        return self._finish_call(v, start, timeit)

    def getpot_device(self, dev_array, timeit=False):
        """dev_array: kernels.DeviceArray of the walkers in HBM -> device array of energies (stays on the GPU)."""
        start = time.time()
        cds = dev_array.torch() if self.as_torch else dev_array
        v = self.potential_function(cds, self.pot_kwargs) if self.pot_kwargs is not None else self.potential_function(cds)
        return self._finish_call(v, start, timeit)
