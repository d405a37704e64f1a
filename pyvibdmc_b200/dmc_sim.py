"""DMC_Sim / dmc_restart: the reference's user API (pyvibdmc/pyvibdmc.py:15-965) driving the B200 path.

Same constructor arguments, attribute names, output files (log text, pickle checkpoints, HDF5
wave-function dumps and sim_info) and exceptions as the reference.  What differs is where the
per-time-step work runs: the walker ensemble lives in HBM (kernels.DeviceSim) and every step --
displacement, potential, weighting / birth-death, Vref, descendant bookkeeping -- is CUDA.  The host
only intervenes at the reference's "special" steps (checkpoints, wave-function windows) and drains
the per-step statistics ring to write the log and the Vref / population histories.

Three stepping modes, chosen from the plug-ins handed in:
  * built-in potential (shipped samples: harmonic, Morse, Partridge-Schwenke water, NN water dimer)
    -> whole segments of time steps are enqueued at once, no host synchronisation per step;
  * any other potential callable -> move / weight / branch on the GPU, the callable is handed the
    coordinates once per step (the reference's plug-in contract: getpot(cds) -> V);
  * importance sampling with a built-in trial wave function -> fused GPU drift/Metropolis step.
"""
import copy
import os
import time

import numpy as np

from . import _capi, kernels
from .simulation_utilities.Constants import Constants, get_atomic_num
from .simulation_utilities.file_manager import FileManager
from .simulation_utilities.imp_samp import ImpSamp
from .simulation_utilities.sim_archive import SimArchivist
from .simulation_utilities.sim_logger import SimLogger

__all__ = ['DMC_Sim', 'dmc_restart']

MASSIVE = "Massive walker birth or death event!!!!!!! Dying..."
_NOT_PICKLED = ('potential', 'potential_info', 'impsamp_manager', 'impsamp', 'imp_info', 'adiabatic_dmc', 'ad_obs_func',
                '_world', '_rank',
                'fixed_node', 'fixed_node_func', 'g_mat', '_dev', '_potential_obj')


class DMC_Sim:
    """See the reference docstring (pyvibdmc.py:18-65) for the meaning of every argument; the
    keyword-only extras select the random stream and the device."""

    def __init__(self, sim_name="DMC_Sim", output_folder="exSimResults", weighting='discrete', num_walkers=10000,
                 num_timesteps=20000, equil_steps=2000, chkpt_every=1000, wfn_every=1000, desc_wt_steps=100, atoms=[],
                 delta_t=1, potential=None, masses=None, start_structures=None, start_cont_wts=None, branch_every=1,
                 log_every=1, cur_timestep=0, cont_wt_thresh=None, imp_samp=None, imp_samp_oned=False,
                 second_impsamp_displacement=False, excited_state_imp_samp=False, adiabatic_dmc=None, fixed_node=None,
                 DEBUG_alpha=None, DEBUG_save_desc_wt_tracker=None, DEBUG_save_training_every=None,
                 DEBUG_save_before_bod=False, DEBUG_mass_change=None, *, seed=None, rng='ziggurat', device=0, distributed=None):
        self.atoms = atoms
        self.sim_name = sim_name
        self.output_folder = output_folder
        self.num_walkers = num_walkers
        self.num_timesteps = num_timesteps
        self.potential_info = vars(potential)
        self.potential = potential.getpot
        self._potential_obj = potential
        self.weighting = weighting.lower()
        self.desc_wt_time_steps = desc_wt_steps
        self.branch_every = branch_every
        self.delta_t = delta_t
        self.start_structures = start_structures
        self.start_cont_wts = start_cont_wts
        self.masses = masses
        self.equil_steps = equil_steps
        self.chkpt_every = chkpt_every
        self.wfn_every = wfn_every
        self.log_every = log_every
        self.cur_timestep = cur_timestep
        self.cont_wt_thresh = cont_wt_thresh
        self.impsamp_manager = imp_samp
        self.imp1d = imp_samp_oned
        self.second_impsamp_displacement = second_impsamp_displacement
        self.adiabatic_dmc = adiabatic_dmc
        self.fixed_node = fixed_node
        self.excited_state_imp_samp = excited_state_imp_samp
        self._deb_training_every = DEBUG_save_training_every
        self._deb_save_before_bod = DEBUG_save_before_bod
        self._deb_desc_wt_tracker = DEBUG_save_desc_wt_tracker
        self._deb_alpha = DEBUG_alpha
        self._deb_mass_change = DEBUG_mass_change
        self._seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        if rng not in _capi.RNG_MODES:
            raise ValueError(f"rng must be one of {sorted(_capi.RNG_MODES)}")
        self._rng_mode = _capi.RNG_MODES[rng]
        self._device = int(device)
        self._dev = None
        # one process per GPU under torch.distributed (torchrun): every rank runs this same object, walkers are sharded,
        # rank 0 writes the reference's output files, the other ranks write theirs (identical) into a scratch folder
        self._distributed_arg = distributed
        self._detect_world(distributed)
        self._host_rng = np.random.default_rng(self._seed ^ 0x5DEECE66D)      # fixed-node recrossing draws (host side)
        if excited_state_imp_samp and (imp_samp_oned or second_impsamp_displacement):
            raise NotImplementedError("excited_state_imp_samp is implemented for 3-D atoms with the standard move only")
        # variants that call back into user Python once per step run in the per-step (hosted) mode:
        # the GPU moves, weights and branches; the host sees the coordinates like a user potential would
        self._hooked = bool(adiabatic_dmc is not None or fixed_node is not None
                            or (DEBUG_save_training_every is not None and DEBUG_save_before_bod))
        if self._hooked and imp_samp is not None:
            raise NotImplementedError("adiabatic_dmc / fixed_node / DEBUG_save_before_bod together with importance sampling "
                                      "are not implemented on the B200 path")
        self._initialize()

    # ------------------------------------------------------------------ set-up (pyvibdmc.py:132-297)
    def _schedule(self):
        T = self.num_timesteps
        self._prop_steps = np.arange(self.cur_timestep, T)
        self._branch_step = np.arange(0, T + self.branch_every, self.branch_every)
        self._chkpt_step = np.arange(self.chkpt_every, T + self.chkpt_every, self.chkpt_every)
        self._wfn_save_step = np.arange(self.equil_steps, T + self.wfn_every, self.wfn_every)
        self._desc_wt_save_step = self._wfn_save_step + self.desc_wt_time_steps
        if self._deb_training_every is not None:                                   # pyvibdmc.py:143-147
            self.deb_train_save_step = np.arange(0, T + self._deb_training_every, self._deb_training_every)
        else:
            self.deb_train_save_step = []
        self._log_steps = np.arange(0, T, self.log_every)

    def _initialize(self):
        self._schedule()
        self._prop_steps = np.arange(0, self.num_timesteps)
        self._who_from = None
        self._walker_pots = None
        self._vref_vs_tau = np.zeros(self.num_timesteps)
        self._pop_vs_tau = np.zeros(self.num_timesteps)

        if self.start_structures is None:
            raise Exception("Please supply a starting structure for your chemical system.")
        elif len(self.start_structures.shape) != 3:
            raise Exception("Start structure must have format (n,m,d), where n = 1 or num_walkers, m = num atoms, "
                            "d = dimensions (usually 3)")
        elif self.start_structures.shape[0] == 1:
            self._walker_coords = np.repeat(self.start_structures, self.num_walkers, axis=0)
        elif self.start_structures.shape[0] == self.num_walkers:
            self._walker_coords = self.start_structures
        else:
            print("WARNING: NUMBER OF STARTING GEOMETRIES DOES NOT EQUAL NUM_WALKERS VARIABLE. MAKE SURE THIS IS"
                  "INTENTIONAL.")
            self._walker_coords = self.start_structures

        if type(self.atoms) is not list:
            self.atoms = [self.atoms]
        if self.masses is None:
            if '-' in self.atoms[0]:                      # reduced mass: 1-D problem
                self.masses = np.array([Constants.reduced_mass(self.atoms[0])])
                self._atm_nums = get_atomic_num(self.atoms[0].split('-'))
            else:
                self.masses = np.array([Constants.mass(a) for a in self.atoms])
                self._atm_nums = get_atomic_num(self.atoms)
        elif isinstance(self.masses, (int, float)):
            self.masses = np.array([self.masses])
            self._atm_nums = [-1]
        elif isinstance(self.masses, (list, np.ndarray)):
            self.masses = np.array(self.masses)
            self._atm_nums = get_atomic_num(self.atoms)
        if len(self.masses) != len(self.atoms):
            raise Exception("Your number of atoms list does not match your number of masses you provided.")
        if self._walker_coords.shape[1] != len(self.atoms):
            raise Exception("Your number of atoms list does not match the shape of your walkers.")

        self._sigmas = np.sqrt(self.delta_t / self.masses)
        self._alpha = 1.0 / (2.0 * self.delta_t) if self._deb_alpha is None else self._deb_alpha

        FileManager.create_filesystem(self.output_folder)
        self._logger = SimLogger(f"{self.output_folder}/{self.sim_name}_log.txt", overwrite=(self.cur_timestep == 0))

        if self.weighting == 'continuous':
            self._thresh_upper = None
            self._cont_wts = self.start_cont_wts if self.start_cont_wts is not None else np.ones(self.num_walkers)
            if self.cont_wt_thresh is None:
                self._thresh_lower = 1 / self.num_walkers
            elif isinstance(self.cont_wt_thresh, (int, float)):
                self._thresh_lower = self.cont_wt_thresh
            elif isinstance(self.cont_wt_thresh, list):
                if len(self.cont_wt_thresh) == 2:
                    self._thresh_lower, self._thresh_upper = self.cont_wt_thresh
                elif len(self.cont_wt_thresh) == 1:
                    self._thresh_lower = self.cont_wt_thresh[0]
            else:
                raise ValueError("Invalid input for continuous weight threshold")
        else:
            self._cont_wts = None
        self._desc_wt = False
        # Mass change throughout simulation (pyvibdmc.py:236-247)
        if self._deb_mass_change is not None:
            change_every = self._deb_mass_change['change_every']
            self._mass_change_steps = np.arange(0, self.num_timesteps, change_every)[1:]
            self._factor_per_change = self._deb_mass_change['factor_per_change']
            if isinstance(self._factor_per_change, (int, float)):
                self._factor_per_change = np.repeat(self._factor_per_change, len(self._mass_change_steps))
            if len(self._factor_per_change) != len(self._mass_change_steps):
                raise ValueError("Number of mass change steps must be equal to mass changes you provide.")
            self._mass_counter = 0
        else:
            self._mass_change_steps = []
        if self.adiabatic_dmc is not None:                                         # pyvibdmc.py:276-291
            ad_lam = self.adiabatic_dmc['initial_lambda']
            ad_lam_dx = self.adiabatic_dmc['lambda_change']
            ad_eq_time = self.adiabatic_dmc['equil_time']
            self.ad_obs_func = self.adiabatic_dmc['observable_func']
            self.ad_lam_array = np.concatenate((np.zeros(ad_eq_time),
                                                np.arange(ad_lam, ad_lam + (ad_lam_dx * (self.num_timesteps - ad_eq_time)), ad_lam_dx)))
        if self.fixed_node is not None:                                            # pyvibdmc.py:292-294
            self.fixed_node_func = self.fixed_node['function']
            self.g_mat = self.fixed_node['g_matrix']
        self._pop_thresh = [self.num_walkers - self.num_walkers * 0.5, self.num_walkers + self.num_walkers * 0.5]

        if 'num_mpi' in self.potential_info.keys():
            self.potential(self._walker_coords)

        if self.impsamp_manager is not None:
            self._init_impsamp()

    def _init_impsamp(self):
        self.imp_info = vars(self.impsamp_manager)
        if self.delta_t != 1:
            print("WARNING! Using DT>1 for Imp Samp. Make sure this is not a mistake!!!!")
        self.f_x = self.psi_1 = self.psi_sec_der = None
        self.inv_masses_trip = (1 / np.repeat(self.masses, 3)).reshape(len(self.masses), 3)[np.newaxis, ...]
        self.sigma_trip = np.repeat(self._sigmas, 3).reshape(len(self.masses), 3)[np.newaxis, ...]
        self.impsamp = ImpSamp(self.impsamp_manager)
        self.eff_ts = np.zeros(self.num_timesteps)
        if self.imp1d:
            self.sigma_trip = self._sigmas
            self.inv_masses_trip = (1 / self.masses)[np.newaxis]

    def _init_restart(self, add_ts, impsamp):
        """Extend the histories and step markers for `add_ts` more steps (pyvibdmc.py:299-338)."""
        self.num_timesteps = self.num_timesteps + add_ts
        self._schedule()
        self._vref_vs_tau = np.concatenate((self._vref_vs_tau, np.zeros(add_ts)))
        self._pop_vs_tau = np.concatenate((self._pop_vs_tau, np.zeros(add_ts)))
        self._prop_steps = np.arange(self.cur_timestep, self.num_timesteps)
        if impsamp is not None:
            self.impsamp_manager = impsamp
            if self.delta_t != 1:
                raise ValueError("Delta tau cannot be anything but 1 for importance sampling DMC!!!!")
            old = getattr(self, 'eff_ts', np.zeros(0))
            self._init_impsamp()
            self.eff_ts[:min(len(old), len(self.eff_ts))] = old[:len(self.eff_ts)]
        else:
            self.impsamp_manager = None
        self.adiabatic_dmc = None

    def _detect_world(self, distributed=None):
        """One process per GPU under torch.distributed (torchrun): world size, rank, this rank's GPU and -- for ranks > 0 -- a
        scratch output folder (rank 0 writes the reference's files).  Called by the constructor AND by dmc_restart: the pickle
        of a checkpoint carries neither the world nor the device."""
        self._world, self._rank = 1, 0
        if distributed or distributed is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    self._world, self._rank = dist.get_world_size(), dist.get_rank()
            except ImportError:
                pass
            if distributed and self._world == 1:
                raise RuntimeError("distributed=True needs torch.distributed initialised with more than one rank")
        if self._world > 1 and "LOCAL_RANK" in os.environ:
            self._device = int(os.environ["LOCAL_RANK"])
        if self._rank > 0:
            import tempfile
            self.output_folder = tempfile.mkdtemp(prefix=f"pvd_rank{self._rank}_")
            os.makedirs(os.path.join(self.output_folder, "wfns"), exist_ok=True)
            os.makedirs(os.path.join(self.output_folder, "chkpts"), exist_ok=True)

    # ------------------------------------------------------------------ public properties
    @property
    def vref_vs_tau(self):
        vref_wvn = self._vref_vs_tau[:self.cur_timestep]
        return np.column_stack((np.arange(len(vref_wvn)), vref_wvn))

    @property
    def walkers(self):
        self._pull_walkers()
        if self.weighting == 'continuous':
            return self._walker_coords, self._cont_wts
        return self._walker_coords

    # ------------------------------------------------------------------ device plumbing
    def _specs(self):
        pot = getattr(self._potential_obj, 'gpu_spec', lambda: None)() if self._potential_obj is not None else None
        if getattr(self, '_hooked', False):
            pot = None                 # per-step mode: the plug-in's getpot is called like any user potential
        trial = None
        self._hosted_imp = False
        if self.impsamp_manager is not None:
            trial = getattr(self.impsamp_manager, 'gpu_spec', lambda: None)()
            if trial is None or pot is None:
                # ANY trial wave function / derivative function (the reference's plug-in contract, imp_samp_manager.py:92-139,
                # 197-224) or a user potential next to a shipped trial function: the host answers impsamp.drift once per
                # step, the GPU proposes, runs the Metropolis step, weights and branches (csrc/pvd_impext.cuh)
                if self.second_impsamp_displacement or self.excited_state_imp_samp:
                    raise NotImplementedError("second_impsamp_displacement / excited_state_imp_samp with a user trial wave function "
                                              "or a user potential: only the shipped sample functions run these variants on the B200 path")
                trial = {"trial": _capi.TRIAL_EXTERNAL, "table": None}
                self._hosted_imp = True
        return pot, trial

    def _ensure_device(self):
        if self._dev is not None:
            return self._dev
        pot, trial = self._specs()
        n_atoms, n_dim = self._walker_coords.shape[1], self._walker_coords.shape[2]
        pot_id, pot_params = _capi.POT_EXTERNAL, None
        if pot is not None:
            pot_id = pot["potential"]
            if pot_id == _capi.POT_HARMONIC:
                pot_params = np.full(n_atoms * n_dim, pot["k"])
            elif pot_id == _capi.POT_MORSE1D:
                pot_params = [pot["de"], pot["alpha"]]
        cap = int(1.5 * max(self.num_walkers, len(self._walker_coords))) + 1024
        if self._world > 1:
            if pot is None and (self.fixed_node is not None or self._deb_save_before_bod):
                raise NotImplementedError("a sharded DMC_Sim with a user potential has no fixed node / DEBUG_save_before_bod")
            from .distributed import ShardedDevice
            self._dev = ShardedDevice(n_atoms, n_dim, self.masses, self.num_walkers, self.delta_t, pot_id, weighting=self.weighting,
                                      alpha=self._alpha, seed=self._seed + 7919 * int(self.cur_timestep), rng_mode=self._rng_mode,
                                      trial=(trial["trial"] if trial else _capi.TRIAL_NONE), pot_params=pot_params,
                                      thresh_lower=getattr(self, '_thresh_lower', None), thresh_upper=getattr(self, '_thresh_upper', None),
                                      stats_ring=max(4096, min(1 << 20, int(self.num_timesteps) + 8)),
                                      imp_variant=(_capi.IMP_SECOND_DISPLACEMENT if self.second_impsamp_displacement else
                                                   _capi.IMP_EXCITED_STATE if self.excited_state_imp_samp else _capi.IMP_STANDARD),
                                      trial_table=(trial["table"] if trial else None))
            self._builtin = pot is not None and not self._hosted_imp
            self._pot_on_device = pot is not None
            if pot_id == _capi.POT_NN_H4O2:
                self._dev.set_nn_weights(pot["weights"])
            self._dev.upload(self._walker_coords, self._cont_wts)
            if self._hosted_imp:
                # user trial / derivative functions (imp_samp_manager.py:92-139): every rank's impsamp.drift sees its own shard;
                # first-step exception (pyvibdmc.py:760-769) as on one GPU
                s0, c0 = self._dev.local_slice(len(self._walker_coords))
                mine = self._walker_coords[s0:s0 + c0]
                f_x, psi_1, sec = self.impsamp.drift(mine)
                v0 = None if self._pot_on_device else np.asarray(self.potential(mine), dtype=np.float64)
                self._dev.imp_ext_init(f_x, psi_1, sec, v0)
            elif pot is None:
                # user potential (getpot plug-in, potential_manager.py:71-99): every rank evaluates its own shard
                s0, c0 = self._dev.local_slice(len(self._walker_coords))
                self._dev.set_pots(np.asarray(self.potential(self._walker_coords[s0:s0 + c0]), dtype=np.float64))
            self._dev_step0 = int(self.cur_timestep)
            self._host_stale = False
            return self._dev
        self._dev = kernels.DeviceSim(n_atoms, n_dim, self.masses, self.num_walkers, self.delta_t, pot_id,
                                      weighting=self.weighting, alpha=self._alpha, capacity=cap,
                                      seed=self._seed + 7919 * int(self.cur_timestep), rng_mode=self._rng_mode,
                                      trial=(trial["trial"] if trial else _capi.TRIAL_NONE), pot_params=pot_params,
                                      thresh_lower=getattr(self, '_thresh_lower', None),
                                      thresh_upper=getattr(self, '_thresh_upper', None), device=self._device,
                                      stats_ring=max(4096, min(1 << 20, int(self.num_timesteps) + 8)),
                                      imp_variant=(_capi.IMP_SECOND_DISPLACEMENT if self.second_impsamp_displacement else
                                                   _capi.IMP_EXCITED_STATE if self.excited_state_imp_samp else _capi.IMP_STANDARD))
        self._builtin = pot is not None and not self._hosted_imp      # whole segments of steps without the host
        self._pot_on_device = pot is not None
        if pot is not None and pot_id == _capi.POT_NN_H4O2:
            self._dev.set_nn_weights(pot["weights"])
        if trial and trial.get("table") is not None:
            self._dev.set_trial_table(trial["table"], trial.get("ntab"))
        self._dev.upload(self._walker_coords, self._cont_wts)
        if self._hosted_imp:
            # first-step exception (pyvibdmc.py:760-769): drift terms and V on the start ensemble, Vref from E_L
            f_x, psi_1, sec = self.impsamp.drift(self._walker_coords)
            v0 = None if self._pot_on_device else np.asarray(self.potential(self._walker_coords), dtype=np.float64)
            self._dev.imp_ext_init(f_x, psi_1, sec, v0)
        elif not self._builtin:
            if self._device_potential():
                self._dev.set_pots_device(self._potential_obj.getpot_device(self._dev.coords_device()))
            else:
                self._dev.set_pots(np.asarray(self.potential(self._walker_coords), dtype=np.float64))
        if getattr(self, '_desc_wt', False) and getattr(self, '_parent', None) is not None and len(self._who_from) == len(self._walker_coords):
            self._dev.dw_resume(self._who_from, self._parent, getattr(self, '_parent_wts', None))      # dmc_restart inside a window
            self._dw_n_parent = len(self._parent)
        self._dev_step0 = int(self.cur_timestep)      # propagation step that device step 0 corresponds to
        self._host_stale = False
        return self._dev

    def _device_potential(self):
        """A Potential_Direct(device=True) plug-in and no per-step host hook: coordinates and energies never leave HBM."""
        return (bool(getattr(self._potential_obj, 'device', False)) and hasattr(self._potential_obj, 'getpot_device')
                and not getattr(self, '_hooked', False) and self._world == 1)

    def _pull_walkers(self):
        """Refresh the host copies of the per-walker arrays from HBM (checkpoints, .walkers, end of run)."""
        if self._dev is None or not getattr(self, '_host_stale', False):
            return
        out = self._dev.download(who_from=self._desc_wt)
        self._walker_coords, self._walker_pots = out["coords"], out["pots"]
        if self.weighting == 'continuous':
            self._cont_wts = out["wts"]
        if self._desc_wt:
            self._who_from = out["who_from"]
        if self.impsamp_manager is not None:
            if getattr(self, '_hosted_imp', False):     # a pure function of the coordinates: ask the plug-in
                self.f_x, self.psi_1, self.psi_sec_der = self.impsamp.drift(self._walker_coords)
            else:
                self.f_x, self.psi_1, self.psi_sec_der = self._dev.download_imp()
        self._vref = self._dev.state(raise_on_error=False)["vref"]
        self._host_stale = False

    def _drain(self, first, count, events):
        """Copy `count` per-step records starting at propagation step `first` into the histories and the log."""
        st = self._dev.stats(first - self._dev_step0, count)
        state = self._dev.state(raise_on_error=False)
        done = state["step"] - (first - self._dev_step0)            # records that belong to completed steps
        done = max(0, min(count, done))
        per_step_s = events.get("seconds", 0.0) / max(count, 1)
        for k in range(done):
            step = first + k
            r = st[k]
            self._vref_vs_tau[step] = r["vref"]
            self._pop_vs_tau[step] = r["pop"]
            if self.impsamp_manager is not None:
                self.eff_ts[step] = (self.eff_ts[step - 1] if step > 0 else 0.0) + r["dt_eff"]
            log = step in self._log_set
            if log:
                self._logger.write_ts(step)
            if k == 0 and events.get("chkpt"):
                self._logger.write_chkpt(step)
            if k == 0 and events.get("wfn"):
                self._logger.write_wfn_save(step)
            if log:
                if self.impsamp_manager is not None:
                    prev = self._pop_vs_tau[step - 1] if step > 0 else 0
                    n_before = int(prev) if (self.weighting == 'discrete' and prev > 0) else len(self._walker_coords)
                    self._logger.write_rejections(int(r["rejected"]), n_before)
                    self._logger.write_imp_disp_time(per_step_s)
                self._logger.write_pot_time(step, events.get("pot_seconds", {}).get(step, per_step_s), r["v_max"], r["v_min"], r["v_avg"])
            if self.impsamp_manager is not None:
                self._logger.write_local(r["v_avg"])
            if log and (step % self.branch_every == 0):
                if self.weighting == 'discrete':
                    self._logger.write_branching(step, 'discrete', (int(r["births"]), int(r["deaths"]), int(r["pop"])))
                else:
                    self._logger.write_branching(step, 'continuous', (int(r["births"]), r["w_max"], r["w_min"]))
            if k == count - 1 and events.get("desc"):
                self._logger.write_desc_wt(step)
            if step % 10 == 0:
                self._logger.fl.flush()
        if done:
            self.cur_timestep = first + done - 1
        if state["err"]:
            self.cur_timestep = first + done
            raise ValueError(MASSIVE)

    # ------------------------------------------------------------------ the loop (pyvibdmc.py:701-876)
    def propagate(self):
        dev = self._ensure_device()
        T = int(self.num_timesteps)
        first = int(self._prop_steps[0]) if len(self._prop_steps) else T
        self._log_set = set(int(s) for s in self._log_steps)
        chk, wfn = set(int(s) for s in self._chkpt_step), set(int(s) for s in self._wfn_save_step)
        dw_end = set(int(s) for s in self._desc_wt_save_step)
        # host-side events: start-of-step {chkpt, wfn window opens}, end-of-step {window closes after step t: t+1 in dw_end}
        mass_steps = set(int(s) for s in self._mass_change_steps)
        train = set(int(s) for s in self.deb_train_save_step) if not self._deb_save_before_bod else set()
        starts = sorted(s for s in (chk | wfn | mass_steps) if first <= s < T)
        ends = sorted(set(s for s in dw_end if first < s <= T) | set(s + 1 for s in train if first <= s < T))
        self._desc_wt_history = []
        t = first
        self._logger.write_beginning(self.__dict__) if first < T else None
        while t < T:
            self.cur_timestep = t
            events = {}
            async_chk = False
            if t in chk:
                events["chkpt"] = True
                # resident single-GPU runs: the checkpoint travels to the host on a side stream while the next segment
                # of time steps is already running; it is pickled below, after that segment has been enqueued
                async_chk = (self._builtin and self._world == 1 and self.impsamp_manager is None
                             and hasattr(dev, "snapshot_begin"))
                if async_chk:
                    dev.snapshot_begin()
                    chk_desc = self._desc_wt                     # the pickle describes the run as it was when the snapshot was taken
                else:
                    self._pull_walkers()
                    self._write_chkpt(t)
            if t in wfn:
                events["wfn"] = True
                n_now = dev.state()["n"]
                self._desc_wts = np.zeros(n_now)
                dev.dw_begin()
                self._desc_wt = True
                self._dw_n_parent = n_now
            if t in mass_steps:                                                    # pyvibdmc.py:749-753
                self.masses = self.masses * self._factor_per_change[self._mass_counter]
                self._sigmas = np.sqrt(self.delta_t / self.masses)
                self._mass_counter += 1
                dev.set_masses(self.masses)
            nxt = min([s for s in starts if s > t] + [s for s in ends if s > t] + [T])
            if self._desc_wt and self._deb_desc_wt_tracker:
                nxt = t + 1                                                        # the tracker records the weights after every step
            tic = time.time()
            if self._builtin:
                dev.run(nxt - t, self.branch_every)
                self._finish_wfn_dump(dev)                       # the previous window's dump landed while these steps were enqueued
                if async_chk:
                    snap = dev.snapshot_wait(who_from=chk_desc)
                    self._walker_coords, self._walker_pots, self._vref = snap["coords"], snap["pots"], snap["vref"]
                    if self.weighting == 'continuous':
                        self._cont_wts = snap["wts"]
                    if chk_desc:
                        self._who_from = snap["who_from"]
                    now_desc, self._desc_wt = self._desc_wt, chk_desc     # (a window that opens at this very step opens after the checkpoint)
                    self._write_chkpt(t)
                    self._desc_wt = now_desc
                dev.sync()
            else:
                events["pot_seconds"] = self._run_external(dev, t, nxt)
            events["seconds"] = time.time() - tic
            self._host_stale = True
            if nxt in dw_end and self._desc_wt:
                events["desc"] = True
            self._drain(t, nxt - t, events)
            if (nxt - 1) in train:                                                 # training data after birth/death (pyvibdmc.py:840-845)
                out = dev.download()
                print(f'{out["coords"].shape} walkers collected')
                SimArchivist.save_h5(fname=f"{self.output_folder}/{self.sim_name}_training_{nxt - 1}ts.hdf5",
                                     keyz=['coords', 'pots'], valz=[out["coords"], out["pots"]])
            if self._desc_wt and self._deb_desc_wt_tracker:                        # pyvibdmc.py:849-852
                self._desc_wt_history.append(dev.dw_peek(self._dw_n_parent))
            if events.get("desc") and self._builtin and self._world == 1 and not self._deb_desc_wt_tracker and hasattr(dev, "dw_end_begin"):
                # asynchronous wave-function dump: descendant weights and parents travel to the host on a side stream while
                # the next segment of time steps runs; the file is written when they have landed (_finish_wfn_dump)
                self._desc_wt = False
                dev.dw_end_begin(self._dw_n_parent)
                self._pending_wfn = f"{self.output_folder}/wfns/{self.sim_name}_wfn_{nxt - self.desc_wt_time_steps}ts.hdf5"
            elif events.get("desc"):
                self._desc_wt = False
                self._desc_wts = dev.dw_end(self._dw_n_parent)
                self._parent, self._parent_wts = dev.dw_parent()
                fname = f"{self.output_folder}/wfns/{self.sim_name}_wfn_{nxt - self.desc_wt_time_steps}ts.hdf5"
                if self.weighting == 'continuous':
                    SimArchivist.save_h5(fname=fname, keyz=['coords', 'desc_wts', 'parent_wts'],
                                         valz=[self._parent, self._desc_wts, self._parent_wts])
                else:
                    SimArchivist.save_h5(fname=fname, keyz=['coords', 'desc_wts'], valz=[self._parent, self._desc_wts])
                if self._deb_desc_wt_tracker:                                      # pyvibdmc.py:868-870
                    np.save(f"{self.output_folder}/wfns/{self.sim_name}_desc_wt_tracker_{nxt - self.desc_wt_time_steps}ts.npy",
                            np.array(self._desc_wt_history))
                    self._desc_wt_history = []
            t = nxt
            self.cur_timestep = t - 1
        self._finish_wfn_dump(dev)
        self._pull_walkers()

    def _finish_wfn_dump(self, dev):
        """Write the wave-function file of a window whose arrays were sent to the host asynchronously (pyvibdmc.py:861-867)."""
        fname = getattr(self, '_pending_wfn', None)
        if fname is None:
            return
        self._pending_wfn = None
        self._desc_wts, self._parent, self._parent_wts = dev.dw_end_wait()
        if self.weighting == 'continuous':
            SimArchivist.save_h5(fname=fname, keyz=['coords', 'desc_wts', 'parent_wts'], valz=[self._parent, self._desc_wts, self._parent_wts])
        else:
            SimArchivist.save_h5(fname=fname, keyz=['coords', 'desc_wts'], valz=[self._parent, self._desc_wts])

    def _write_chkpt(self, t):
        if self._desc_wt and self._dev is not None and self._world == 1:
            self._parent, self._parent_wts = self._dev.dw_parent()         # an open window travels in the pickle (pyvibdmc.py:299-338)
        FileManager.delete_older_checkpoints(self.output_folder, self.sim_name, t)
        logger, self._logger = self._logger, None
        SimArchivist.chkpt(self, t)
        self._logger = logger

    def _run_external(self, dev, t0, t1):
        """User potential callable: the GPU moves / weights / branches, the callable sees the coordinates
        once per step (getpot contract, potential_manager.py:71-99)."""
        pot_seconds = {}
        train_before = set(int(s) for s in self.deb_train_save_step) if self._deb_save_before_bod else set()
        for step in range(t0, t1):
            if getattr(self, '_hosted_imp', False):
                # imp_move_randomly (pyvibdmc.py:549-612) around the plug-in call impsamp.drift(displaced_cds)
                y = dev.imp_ext_propose()
                f_y, psi_2, sec_y = self.impsamp.drift(y)
                dev.imp_ext_accept(f_y, psi_2, sec_y)
                do_branch = (step % self.branch_every) == 0
                if self._pot_on_device:
                    dev.imp_ext_finish(None, do_branch)
                else:
                    cds = (dev.download_local() if self._world > 1 else dev.download())["coords"]      # sharded: this rank's walkers
                    if step in self._log_set:
                        v, pot_seconds[step] = self.potential(cds, timeit=True)
                    else:
                        v = self.potential(cds)
                    dev.imp_ext_finish(np.array(v, dtype=np.float64), do_branch)
                if dev.state(raise_on_error=False)["err"]:
                    break
                continue
            if self._device_potential():
                # device-tensor plug-in: the callable gets a CUDA view of the moved walkers and returns a device array
                arr = dev.ext_move_device()
                if step in self._log_set:
                    v_dev, pot_seconds[step] = self._potential_obj.getpot_device(arr, timeit=True)
                else:
                    v_dev = self._potential_obj.getpot_device(arr)
                dev.ext_finish_device(v_dev, (step % self.branch_every) == 0)
                if dev.state(raise_on_error=False)["err"]:
                    break
                continue
            if self.fixed_node is not None:                                        # pyvibdmc.py:755-757
                q_beginning = self.fixed_node_func(dev.download()["coords"], step)
            cds = dev.ext_move()
            if step in self._log_set:
                v, pot_seconds[step] = self.potential(cds, timeit=True)
            else:
                v = self.potential(cds)
            v = np.array(v, dtype=np.float64)
            if self.fixed_node is not None:                                        # pyvibdmc.py:795-797
                self._recrossing(q_beginning, self.fixed_node_func(cds, step), cds, v)
            if step in train_before:                                               # pyvibdmc.py:799-804
                print(f'{cds.shape} walkers collected')
                SimArchivist.save_h5(fname=f"{self.output_folder}/{self.sim_name}_training_{step}ts.hdf5",
                                     keyz=['coords', 'pots'], valz=[cds, v])
            if self.adiabatic_dmc is not None:                                     # pyvibdmc.py:820-825
                v = v + self.ad_lam_array[step] * self.ad_obs_func(cds)
            do_branch = (step % self.branch_every) == 0
            dev.ext_finish(np.asarray(v, dtype=np.float64), do_branch)
            if dev.state(raise_on_error=False)["err"]:
                break
        return pot_seconds

    def _recrossing(self, q_1, q_2, cds, pots):
        """Fixed-node recrossing correction (pyvibdmc.py:684-699): walkers that probably crossed the node and came back
        get a prohibitive energy.  In the reference `cds` aliases the moved walkers (move_randomly works in place), so
        both G-matrix evaluations see the same coordinates; the same holds here."""
        try:
            m1 = 1 / self.g_mat(cds)
            m2 = 1 / self.g_mat(cds)
            summed = m1 + m2
            numerator = 2 * q_1 * q_2 * summed
            denominator = -2 * self.delta_t
        except Exception:
            numerator = -1 * 4 * q_1 * q_2
            sigma_q = np.sqrt(self.delta_t * self.g_mat)
            denominator = 2 * sigma_q ** 2
        p_recross = np.exp(numerator / denominator)
        randz = self._host_rng.random(size=len(cds))
        pots[randz < p_recross] = 10

    # ------------------------------------------------------------------ run / checkpoint (pyvibdmc.py:878-947)
    def run(self):
        try:
            print("Starting Simulation...")
        except OSError:
            pass
        dmc_time_start = time.time()
        throw_error = None
        try:
            self.propagate()
            FileManager.delete_older_checkpoints(self.output_folder, self.sim_name, self.cur_timestep)
        except Exception as e:
            import traceback
            print("ERROR! An error occurred while running the DMC simulation. Dumping a final checkpoint...")
            print("Ignore Approximate ZPE!!!")
            traceback.print_exc()
            throw_error = e
        finally:
            if self._logger is not None:
                self._logger.final_chkpt()
                self._logger.fl.close()
            self._logger = None
            try:
                self._pull_walkers()
                if self._desc_wt and self._dev is not None and self._world == 1:
                    self._parent, self._parent_wts = self._dev.dw_parent()     # an open window travels in the pickle
            except Exception:
                pass
            SimArchivist.chkpt(self, self.cur_timestep)
            _vref_wvn = Constants.convert(self._vref_vs_tau, "wavenumbers", to_AU=False)
            try:
                print("Simulation Complete")
                print('Approximate ZPE', np.average(_vref_wvn[len(_vref_wvn) // 4:]))
            except OSError:
                pass
            ts = self.eff_ts if self.impsamp_manager is not None else np.arange(len(_vref_wvn)) * self.delta_t
            SimArchivist.save_h5(fname=f"{self.output_folder}/{self.sim_name}_sim_info.hdf5",
                                 keyz=['vref_vs_tau', 'pop_vs_tau', 'atomic_nums', 'atomic_masses'],
                                 valz=[np.column_stack((ts, self._vref_vs_tau)), np.column_stack((ts, self._pop_vs_tau)),
                                       self._atm_nums, self.masses])
            if getattr(self, 'adiabatic_dmc', None) is not None:                   # pyvibdmc.py:924-925
                np.save(f"{self.output_folder}/{self.sim_name}_lambda.npy", self.ad_lam_array)
            finish = time.time() - dmc_time_start
        self._logger = SimLogger(f"{self.output_folder}/{self.sim_name}_log.txt")
        self._logger.finish_sim(finish)
        if throw_error is not None:
            raise throw_error

    def __deepcopy__(self, memodict={}):
        """Checkpoint copy: plug-ins and the device handle are not pickled (pyvibdmc.py:934-947)."""
        cls = self.__class__
        res = cls.__new__(cls)
        memodict[id(self)] = res
        for k, v in self.__dict__.items():
            if k == '_logger':
                res._logger = None            # open file handle; the reference nulls it around chkpt()
            elif k not in _NOT_PICKLED:
                setattr(res, k, copy.deepcopy(v, memodict))
        return res

    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if k not in _NOT_PICKLED}

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._dev = None
        self._potential_obj = None
        for k in ('adiabatic_dmc', 'fixed_node', 'impsamp_manager'):
            self.__dict__.setdefault(k, None)
        self.__dict__.setdefault('_world', 1)
        self.__dict__.setdefault('_rank', 0)


def dmc_restart(potential, chkpt_folder, sim_name, additional_timesteps=0, impsamp=None, imp_samp_oned=False,
                fixed_node=None):
    """Reload `{chkpt_folder}/chkpts/{sim_name}_*.pickle` and continue (pyvibdmc.py:949-965)."""
    dmc_sim = SimArchivist.reload_sim(chkpt_folder, sim_name)
    dmc_sim.imp1d = imp_samp_oned
    dmc_sim.fixed_node = fixed_node
    if fixed_node is not None:
        dmc_sim.fixed_node_func = fixed_node['function']
        dmc_sim.g_mat = fixed_node['g_matrix']
    dmc_sim._hooked = bool(fixed_node is not None or (dmc_sim._deb_training_every is not None and dmc_sim._deb_save_before_bod))
    dmc_sim._init_restart(additional_timesteps, impsamp)
    dmc_sim.potential = potential.getpot
    dmc_sim._potential_obj = potential
    dmc_sim.potential_info = vars(potential)
    dmc_sim._dev = None
    # under torchrun every rank reloads rank 0's pickle: world, rank, GPU and the scratch folder of ranks > 0 are this process's own
    dmc_sim._detect_world(getattr(dmc_sim, '_distributed_arg', None))
    if getattr(dmc_sim, '_desc_wt', False):
        # the checkpoint was written inside a descendant-weighting window: it is resumed (_who_from, _parent and _parent_wts are in
        # the pickle, pyvibdmc.py:299-338), unless the run is sharded (who_from holds global parent ids there)
        if dmc_sim._world > 1 or getattr(dmc_sim, '_parent', None) is None or getattr(dmc_sim, '_who_from', None) is None:
            print("WARNING: the checkpoint was written inside a descendant-weighting window; that window is dropped.")
            dmc_sim._desc_wt = False
    dmc_sim._logger = SimLogger(f"{dmc_sim.output_folder}/{dmc_sim.sim_name}_log.txt")
    FileManager.delete_future_checkpoints(chkpt_folder, sim_name, dmc_sim.cur_timestep)
    return dmc_sim
