"""GPU versions of the reference's 1-D harmonic sample potentials
(sample_potentials/PythonPots/harmonicOscillator1D.py:5-49): same names, same call signatures,
same values (k = (0.5*m)*omega**2 is formed on the host in the reference's operation order; the
kernel evaluates k*(x*x) without FMA contraction, so results are bit-identical to NumPy)."""
import numpy as np

from pyvibdmc_b200 import kernels as _K, _capi
from pyvibdmc_b200.simulation_utilities.Constants import Constants


def _k(mass, omega):
    return 0.5 * mass * omega ** 2


def _harm(cds, k):
    cds = np.asarray(cds, dtype=np.float64)
    return np.squeeze(_K.pes_harmonic(cds.reshape(len(cds), -1), k))


def _spec(k):
    return {"potential": _capi.POT_HARMONIC, "k": float(k)}


def hydrogen_harm(cds):
    return _harm(cds, _k(Constants.mass('H', to_AU=True), Constants.convert(3600., 'wavenumbers', to_AU=True)))


def oh_stretch_harm(cds):
    """(N,1,1) -> (N,) Hartree; mu = reduced mass of O-H, omega = 3700 cm^-1."""
    return _harm(cds, _k(Constants.reduced_mass('O-H', to_AU=True), Constants.convert(3700., 'wavenumbers', to_AU=True)))


def oh_stretch_harm_with_arg(cds, extra_args):
    return _harm(cds, _k(extra_args['mass'], extra_args['freq']))


def oh_stretch_harm_loadtxt(cds):
    np.loadtxt('random.txt')          # exercises Potential_NoMP(ch_dir=True), like the reference's test helper
    return oh_stretch_harm(cds)


def n2_stretch_harm(cds):
    return _harm(cds, _k(Constants.reduced_mass('N-N', to_AU=True), Constants.convert(2750., 'wavenumbers', to_AU=True)))


def hcl_stretch_harm(cds):
    return _harm(cds, _k(Constants.reduced_mass('H-Cl', to_AU=True), Constants.convert(2850., 'wavenumbers', to_AU=True)))


def oh_stretch_harm_shifted(cds):
    cds = np.asarray(cds, dtype=np.float64) - Constants.convert(0.98, 'angstroms', to_AU=True)
    return oh_stretch_harm(cds)


hydrogen_harm._pvd_builtin = _spec(_k(Constants.mass('H', to_AU=True), Constants.convert(3600., 'wavenumbers', to_AU=True)))
oh_stretch_harm._pvd_builtin = _spec(_k(Constants.reduced_mass('O-H', to_AU=True), Constants.convert(3700., 'wavenumbers', to_AU=True)))
oh_stretch_harm_loadtxt._pvd_builtin = oh_stretch_harm._pvd_builtin
n2_stretch_harm._pvd_builtin = _spec(_k(Constants.reduced_mass('N-N', to_AU=True), Constants.convert(2750., 'wavenumbers', to_AU=True)))
hcl_stretch_harm._pvd_builtin = _spec(_k(Constants.reduced_mass('H-Cl', to_AU=True), Constants.convert(2850., 'wavenumbers', to_AU=True)))
oh_stretch_harm_with_arg._pvd_builtin = lambda kw: _spec(_k(kw['mass'], kw['freq']))
