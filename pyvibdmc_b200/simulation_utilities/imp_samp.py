"""ImpSamp: drift / Metropolis / local kinetic energy with the reference's call signatures
(simulation_utilities/imp_samp.py:5-76), evaluated by the CUDA kernels for any (walkers, atoms, dims) ensemble up to 16 atoms (compile-time
instances for 3 atoms x 3 dims and 1 x 1, a run-time-shaped kernel otherwise); DMC_Sim itself never calls these per step on the
built-in path -- the whole importance-sampled step runs in k_imp_move -- they are the plug-in level
entry points and what the parity tests exercise."""
import numpy as np

from .. import kernels as _K

__all__ = ['ImpSamp']


class ImpSamp:
    def __init__(self, imp_samp_manager):
        self.imp_manager = imp_samp_manager

    def trial(self, cds):
        return self.imp_manager.call_trial(cds)

    def drift(self, cds):
        """(grad psi / psi, psi, d2psi/dx2 / psi) -- imp_samp.py:21-26."""
        psi_t = self.trial(cds)
        deriv, sderiv = self.imp_manager.call_derivs(cds)
        return deriv, psi_t, sderiv

    @staticmethod
    def metropolis(sigma_trip, trial_x, trial_y, disp_x, disp_y, D_x, D_y, dt):
        """Acceptance ratios (imp_samp.py:29-47).  D_x = inv_mass * f_x as in the reference."""
        x = np.asarray(disp_x, dtype=np.float64)
        if x.ndim != 3:
            raise ValueError("metropolis expects (walkers, atoms, dims) arrays")
        natoms, ndim = x.shape[1:]
        sig = ImpSamp._per_atom(sigma_trip, natoms, ndim)
        # the kernel takes the drift f and inv_mass separately; pass D with inv_mass = 1
        return _K.metropolis(x, disp_y, D_x, D_y, np.asarray(trial_x).reshape(-1), np.asarray(trial_y).reshape(-1),
                             sig, np.ones(natoms), dt)

    @staticmethod
    def local_kin(inv_masses_trip, sec_deriv):
        """-1/2 sum (1/m) d2psi/psi (imp_samp.py:50-53)."""
        d2 = np.asarray(sec_deriv, dtype=np.float64)
        natoms, ndim = d2.shape[1:]
        return _K.local_kin(d2, ImpSamp._per_atom(inv_masses_trip, natoms, ndim))

    @staticmethod
    def _per_atom(trip, natoms, ndim):
        """The reference carries per-atom constants repeated over the dimensions ((atoms, dims) 'trip' arrays,
        pyvibdmc.py:236-240); the kernels take one value per atom."""
        t = np.asarray(trip, dtype=np.float64).reshape(-1)
        if t.size == natoms * ndim and ndim > 1:
            return np.ascontiguousarray(t.reshape(natoms, ndim)[:, 0])
        return np.ascontiguousarray(t[:natoms])

    @staticmethod
    def finite_diff(cds, trial_func):
        """Host finite differences for USER trial functions (imp_samp.py:56-76): dx = 1e-3 bohr, the
        coordinate array is walked in place and restored, exactly like the reference.  Built-in trial
        functions never come here (their stencil is evaluated in registers on the GPU)."""
        dx = 0.001
        first, sec = np.zeros(cds.shape), np.zeros(cds.shape)
        centre = trial_func(cds)
        for atom in range(cds.shape[1]):
            for xyz in range(cds.shape[-1]):
                cds[:, atom, xyz] -= dx
                minus = trial_func(cds)
                cds[:, atom, xyz] += 2. * dx
                plus = trial_func(cds)
                cds[:, atom, xyz] -= dx
                first[:, atom, xyz] = (plus - minus) / (2 * dx)
                sec[:, atom, xyz] = (minus - 2. * centre + plus) / dx ** 2
        return first, sec, centre
