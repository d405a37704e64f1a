"""GPU version of sample_potentials/PythonPots/morse_osc_1d.py:4-12."""
import numpy as np

from pyvibdmc_b200 import kernels as _K, _capi
from pyvibdmc_b200.simulation_utilities.Constants import Constants


def _params():
    mass = Constants.reduced_mass('O-H', to_AU=True)
    omega = Constants.convert(3704.5, 'wavenumbers', to_AU=True)
    omega_x = Constants.convert(75.3, 'wavenumbers', to_AU=True)
    de = (omega ** 2 / (4 * omega_x))
    alpha = np.sqrt(mass * (omega ** 2.) / 2. / de)
    return de, float(alpha)


def oh_stretch_morse(disp):
    de, alpha = _params()
    return _K.pes_morse1d(np.asarray(disp, dtype=np.float64), de, alpha).squeeze()


oh_stretch_morse._pvd_builtin = {"potential": _capi.POT_MORSE1D, "de": _params()[0], "alpha": _params()[1]}
