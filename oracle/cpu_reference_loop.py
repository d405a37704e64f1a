"""ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's CPU path for the hot loop, restated so that it can be timed on the GPU box
(where /root/reference does not exist and the Fortran cannot be built):
  * loop body  = DMC_Sim.propagate, discrete weighting (pyvibdmc.py:701-876) -> oracle.dmc_oracle
  * potential  = Potential.getpot's multiprocessing path: np.array_split over num_cores,
                 Pool.map, np.concatenate (simulation_utilities/potential_manager.py:71-99)
                 around the C restatement of the Fortran PES.
Used only by bench.py (cpu_baseline leg and --impl reference).
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import dmc_oracle as O

EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])


def _pot_chunk(cds):
    return O.water_pot(cds)


class PoolPotential:
    """Potential(num_cores=C).getpot restated (potential_manager.py:59-64,83-88)."""

    def __init__(self, cores):
        O.build_c_oracle()
        self.cores = max(1, int(cores))
        self.pool = mp.get_context("fork").Pool(self.cores) if self.cores > 1 else None

    def __call__(self, cds):
        if self.pool is None:
            return O.water_pot(cds)
        return np.concatenate(self.pool.map(_pot_chunk, np.array_split(cds, self.cores)))

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def time_h2o_discrete(n_walkers, steps, warmup, cores=None, dt=5.0, seed=0):
    """Walker-steps/s of the reference-style CPU loop on H2O / PS, discrete weighting.
    Returns dict(value, seconds, walker_steps, cores, per_step_s)."""
    cores = cores or os.cpu_count() or 1
    pot = PoolPotential(cores)
    try:
        masses = np.array([O.mass('H'), O.mass('H'), O.mass('O')])
        sig = np.sqrt(dt / masses)
        rs = np.random.RandomState(seed)
        coords = np.repeat(EQ[None] * 1.01, n_walkers, axis=0)
        pots = pot(coords)
        vref = O.calc_vref(pots, n_walkers, 1 / (2 * dt))
        per_step, total = [], 0
        for t in range(warmup + steps):
            t0 = time.perf_counter()
            coords = coords + rs.normal(0.0, sig, size=(len(coords), 3, 3)).transpose(0, 2, 1)   # :540-547
            pots = pot(coords)                                                                   # :786-793
            u = rs.random_sample(len(coords))
            _, idx, _, _, _ = O.birth_or_death_discrete(pots, vref, dt, u, n_walkers)            # :391-431
            coords, pots = coords[idx], pots[idx]
            vref = O.calc_vref(pots, n_walkers, 1 / (2 * dt))                                    # :651-661
            if t >= warmup:
                per_step.append(time.perf_counter() - t0)
                total += len(coords)
        secs = float(np.sum(per_step))
        return dict(value=total / secs, seconds=secs, walker_steps=int(total), cores=cores, per_step_s=per_step)
    finally:
        pot.close()
