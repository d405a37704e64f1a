"""pyvibdmc_b200 -- B200-native implementation of PyVibDMC's walker-propagation path.

Drop-in surface (same names as the reference package `pyvibdmc`):
    DMC_Sim, dmc_restart, Potential, Potential_NoMP, NN_Potential, Potential_Direct,
    ImpSampManager, ImpSampManager_NoMP, Constants
The compute path is hand-written sm_100a CUDA behind a C ABI (include/pvd_b200.h); there is no
CPU fallback.
"""
from . import _capi, kernels  # noqa: F401
from .dmc_sim import DMC_Sim, dmc_restart  # noqa: F401
from .simulation_utilities import *  # noqa: F401,F403

__version__ = "0.1.0"
