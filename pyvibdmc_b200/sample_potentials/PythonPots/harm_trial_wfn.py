"""GPU versions of the reference's 1-D Gaussian trial wave function and its analytic derivatives
(sample_potentials/PythonPots/harm_trial_wfn.py:6-40): same function names and return shapes."""
import numpy as np

from pyvibdmc_b200 import kernels as _K, _capi
from pyvibdmc_b200.simulation_utilities.Constants import Constants


def _alpha():
    return Constants.reduced_mass('O-H') * Constants.convert(3700, 'wavenumbers', to_AU=True)


def _drift(x):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 1, 1)
    return _K.trial_drift(_capi.TRIAL_HARM1D, x, np.array([_alpha()]))


def harmonic_oscillator(x, mass, omega):
    alpha = mass * omega
    return _K.trial_drift(_capi.TRIAL_HARM1D, np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 1, 1), np.array([alpha]))[1]


def trial_harm(x):
    """psi, shape (num_walkers,)"""
    return _drift(x)[1]


def derivative(x):
    """(psi'/psi, psi''/psi), each (num_walkers, 1, 1)"""
    d1, _, d2 = _drift(x)
    return d1, d2


def first_derivative(x):
    d1, psi, _ = _drift(x)
    return d1 * psi[:, np.newaxis, np.newaxis]


def second_derivative(x):
    _, psi, d2 = _drift(x)
    return d2 * psi[:, np.newaxis, np.newaxis]


trial_harm._pvd_builtin_trial = {"trial": _capi.TRIAL_HARM1D, "table": np.array([_alpha()]), "fd": False}
