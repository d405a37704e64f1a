"""Water product trial wave function on the GPU (replaces FortPots/Partridge_Schwenke_H2O/call_trl_h2o.py:7-78):
psi = interp(r_OH1) * interp(r_OH2) * Gaussian(theta), the O-H factors linearly interpolated on the shipped
5000-point grid exactly like np.interp (clamped outside [0.5, 4.0] bohr).  Derivatives come either from finite differences
(ImpSampManager(..., deriv_function=None): a 19-point stencil in registers) or analytically (deriv_function='dpsi_dx',
reference call_trl_h2o.py:101-149 with the ChainRuleHelper formulas of imp_samp_helper.py:10-209 written out on the device)."""
import os

import numpy as np

from pyvibdmc_b200 import kernels as _K, _capi
from pyvibdmc_b200.simulation_utilities.Constants import Constants

_HERE = os.path.dirname(os.path.abspath(__file__))
r1_eq = Constants.convert(0.95784, 'angstroms', to_AU=True)
r2_eq = Constants.convert(0.95783997, 'angstroms', to_AU=True)
theta_eq = np.deg2rad(104.5080029)
theta_freq = Constants.convert(1668.4590610594878, 'wavenumbers', to_AU=True)
inv_mh = 1 / Constants.mass('H')
inv_mo = 1 / Constants.mass('O')
_table = np.load(os.path.join(_HERE, "free_oh_wvfn_table.npy"))      # rows: grid (bohr), psi, psi', psi''


def gmat():
    return inv_mh / r1_eq ** 2 + inv_mh / r2_eq ** 2 + inv_mo * (
        1 / r1_eq ** 2 + 1 / r2_eq ** 2 - 2 * np.cos(theta_eq) / (r1_eq * r2_eq))


def packed_table():
    """[grid | psi | alpha_theta, theta_eq] as the C ABI expects (pvd_trial_drift / pvd_sim_set_trial_table)."""
    return np.concatenate([_table[0], _table[1], [theta_freq / gmat(), theta_eq]])


def packed_table_analytic():
    """[grid | psi | psi' | psi'' | alpha_theta, theta_eq] for PVD_TRIAL_H2O_AN."""
    return np.concatenate([_table[0], _table[1], _table[2], _table[3], [theta_freq / gmat(), theta_eq]])


def _check(ex_args):
    if ex_args is not None and (list(map(list, ex_args.get('dists', []))) != [[0, 2], [2, 1]] or
                                list(map(list, ex_args.get('angs', []))) != [[0, 2, 1]]):
        raise NotImplementedError("the built-in water trial wfn is compiled for dists=[[0,2],[2,1]], angs=[[0,2,1]] (atoms H,H,O)")


def trial_wavefunction(cds, ex_args=None, ret_pdt=True):
    _check(ex_args)
    if not ret_pdt:
        raise NotImplementedError("per-factor output is only needed by the analytic-derivative helper (out of scope)")
    cds = np.ascontiguousarray(cds, dtype=np.float64)
    return _K.trial_drift(_capi.TRIAL_H2O_FD, cds, packed_table(), ntab=_table.shape[1])[1]


def _spec(ex_args):
    _check(ex_args)
    return {"trial": _capi.TRIAL_H2O_FD, "table": packed_table(), "ntab": _table.shape[1], "fd": True}


trial_wavefunction._pvd_builtin_trial = _spec


def dpsi_dx(cds, ex_args=None):
    """(grad psi / psi, d2 psi / dx2 / psi), analytic (reference call_trl_h2o.py:101-149)."""
    _check(ex_args)
    cds = np.ascontiguousarray(cds, dtype=np.float64)
    d1, _, d2 = _K.trial_drift(_capi.TRIAL_H2O_AN, cds, packed_table_analytic(), ntab=_table.shape[1])
    return d1, d2


def _spec_analytic(ex_args):
    _check(ex_args)
    return {"trial": _capi.TRIAL_H2O_AN, "table": packed_table_analytic(), "ntab": _table.shape[1], "fd": False}


dpsi_dx._pvd_builtin_trial = _spec_analytic
