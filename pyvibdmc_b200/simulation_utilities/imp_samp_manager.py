"""Importance-sampling plug-in wrappers with the reference's constructors and call_trial /
call_derivs contract (simulation_utilities/imp_samp_manager.py:11-224).

Shipped sample trial functions (pyvibdmc_b200/sample_potentials/.../harm_trial_wfn.py, call_trl_h2o.py)
carry a `_pvd_builtin_trial` descriptor: DMC_Sim then runs the complete importance-sampled step on
the GPU.  Other user functions are called on the host as the reference does."""
import importlib
import os
import sys

import numpy as np

from .imp_samp import ImpSamp
from .potential_manager import Potential, Potential_NoMP, NN_Potential

__all__ = ['ImpSampManager', 'ImpSampManager_NoMP']


def _load(directory, python_file, names):
    here = os.getcwd()
    os.chdir(directory)
    try:
        sys.path.insert(0, os.getcwd())
        module = importlib.import_module(python_file.split(".")[0])
        return [getattr(module, n) if n is not None else None for n in names]
    finally:
        os.chdir(here)


class _ImpBase:
    def _bind(self, directory, python_file, trial_name, deriv_name):
        self._trial_fn, self._deriv_fn = _load(directory, python_file, [trial_name, deriv_name])
        self.all_finite = deriv_name is None
        self.derivs = ImpSamp.finite_diff if self.all_finite else self._deriv_fn

    def gpu_spec(self):
        """Descriptor of a built-in trial wfn (None for user functions)."""
        # a built-in derivative function (e.g. the analytic water derivatives) decides the device trial variant
        spec = getattr(self._deriv_fn, "_pvd_builtin_trial", None) if not self.all_finite else None
        if spec is None:
            spec = getattr(self._trial_fn, "_pvd_builtin_trial", None)
        if spec is None:
            return None
        if callable(spec):
            spec = spec(self.trial_kwargs)
        spec = dict(spec)
        want_fd = self.all_finite
        if spec["fd"] != want_fd:
            return None          # e.g. analytic-only built-in asked for finite differences
        return spec

    def _call(self, fn, cds, kwargs):
        return fn(cds) if kwargs is None else fn(cds, kwargs)

    def _bump_timestep(self):
        if self.pass_timestep:
            self.ct += 1
            self.trial_kwargs['timestep'] = self.ct
            self.deriv_kwargs['timestep'] = self.ct


class ImpSampManager(_ImpBase):
    """ImpSampManager(trial_function, trial_directory, python_file, pot_manager, pass_timestep=False,
    new_pool_num_cores=None, deriv_function=None, trial_kwargs=None, deriv_kwargs=None)."""

    def __init__(self, trial_function, trial_directory, python_file, pot_manager, pass_timestep=False,
                 new_pool_num_cores=None, deriv_function=None, trial_kwargs=None, deriv_kwargs=None):
        self.trial_func = trial_function
        self.trial_dir = trial_directory
        self.python_file = python_file
        self.deriv_func = deriv_function
        self.trial_kwargs = trial_kwargs
        self.deriv_kwargs = deriv_kwargs
        self.pot_manager = pot_manager
        self.pass_timestep = pass_timestep
        self.nomp_pool_cores = new_pool_num_cores
        if self.pass_timestep:
            self.ct = 0
            self.trial_kwargs['timestep'] = 0
            self.deriv_kwargs['timestep'] = 0
        self.pool = getattr(pot_manager, 'pool', None) if isinstance(pot_manager, Potential) else None
        self.num_cores = getattr(pot_manager, 'num_cores', 1) if isinstance(pot_manager, (Potential, Potential_NoMP, NN_Potential)) else 1
        self._bind(trial_directory, python_file, trial_function, deriv_function)

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop('pool', None)
        state.pop('pot_manager', None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    def call_trial(self, cds):
        return self._call(self._trial_fn, cds, self.trial_kwargs)

    call_trial_no_mp = call_trial

    def call_derivs(self, cds):
        if self.all_finite:
            first, sec, psi = ImpSamp.finite_diff(cds, self.call_trial)
            derivz, sderivz = first / psi[:, np.newaxis, np.newaxis], sec / psi[:, np.newaxis, np.newaxis]
        else:
            derivz, sderivz = self._call(self._deriv_fn, cds, self.deriv_kwargs)
        self._bump_timestep()
        return derivz, sderivz


class ImpSampManager_NoMP(_ImpBase):
    """ImpSampManager_NoMP(trial_function, trial_directory, python_file, chdir=False, pass_timestep=False,
    deriv_function=None, trial_kwargs=None, deriv_kwargs=None)."""

    def __init__(self, trial_function, trial_directory, python_file, chdir=False, pass_timestep=False,
                 deriv_function=None, trial_kwargs=None, deriv_kwargs=None):
        self.trial_fuc = trial_function
        self.trial_dir = trial_directory
        self.python_file = python_file
        self.pass_timestep = pass_timestep
        self.deriv_func = deriv_function
        self.trial_kwargs = trial_kwargs
        self.deriv_kwargs = deriv_kwargs
        self.chdir = chdir
        if self.pass_timestep:
            self.ct = 0
            self.trial_kwargs['timestep'] = 0
            self.deriv_kwargs['timestep'] = 0
        self._curdir = os.getcwd()
        self._bind(trial_directory, python_file, trial_function, deriv_function)
        self.trial = self._trial_fn

    def call_imp_func(self, func, cds, func_kwargs=None):
        if self.chdir:
            os.chdir(self.trial_dir)
        try:
            return self._call(func, cds, func_kwargs)
        finally:
            if self.chdir:
                os.chdir(self._curdir)

    def call_trial(self, cds):
        return self.call_imp_func(self._trial_fn, cds, self.trial_kwargs)

    def call_derivs(self, cds):
        if self.all_finite:
            first, sec, psi = ImpSamp.finite_diff(cds, self.call_trial)
            derivz, sderivz = first / psi[:, np.newaxis, np.newaxis], sec / psi[:, np.newaxis, np.newaxis]
        else:
            derivz, sderivz = self.call_imp_func(self._deriv_fn, cds, self.deriv_kwargs)
        self._bump_timestep()
        return derivz, sderivz
