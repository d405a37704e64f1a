"""Launched with torchrun by tests/test_gpu_multi.py: the drop-in DMC_Sim API on WORLD_SIZE GPUs (sharded walkers)."""
import glob
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import pyvibdmc_b200 as pv
    from pyvibdmc_b200.simulation_utilities import h5lite
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    out = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] == "ext":
        return ext_main(pv, h5lite, dist, rank, out)
    if len(sys.argv) > 2 and sys.argv[2] == "uimp":
        return uimp_main(pv, dist, rank, out)
    if len(sys.argv) > 2 and sys.argv[2] == "exc":
        return exc_main(pv, dist, rank, out)
    if len(sys.argv) > 2 and sys.argv[2] == "imp2":
        return imp2_main(pv, dist, rank, out)
    eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
    d = os.path.join(os.path.dirname(pv.__file__), "sample_potentials", "FortPots", "Partridge_Schwenke_H2O")
    pot = pv.Potential(potential_function='water_pot', python_file='h2o_potential.py', potential_directory=d, num_cores=1)
    sim = pv.DMC_Sim(sim_name="w2", output_folder=out, weighting='discrete', num_walkers=20000, num_timesteps=600, equil_steps=100,
                     chkpt_every=300, wfn_every=200, desc_wt_steps=50, atoms=['H', 'H', 'O'], delta_t=5, potential=pot,
                     start_structures=eq[None] * 1.01, log_every=100, seed=3)
    assert sim._world == dist.get_world_size() and (sim.output_folder == out) == (rank == 0)
    sim.run()
    walkers = sim.walkers
    if rank == 0:
        info = h5lite.read_h5(f"{out}/w2_sim_info.hdf5")
        wf = sorted(os.path.basename(p) for p in glob.glob(f"{out}/wfns/*.hdf5"))
        w = h5lite.read_h5(f"{out}/wfns/w2_wfn_300ts.hdf5")
        pop = info['pop_vs_tau'][:, 1]
        res = {"world": sim._world, "vref_shape": list(info['vref_vs_tau'].shape), "wfns": wf,
               "zpe": float(info['vref_vs_tau'][300:, 1].mean() / 4.556335281212229e-6),
               "desc_sum": float(w['desc_wts'].sum()), "pop_at_window_end": float(pop[349]), "n_parent": int(len(w['coords'])),
               "pop_at_window_start": float(pop[299]), "final_walkers": int(len(walkers)), "final_pop": float(pop[-1]),
               "chkpts": sorted(os.path.basename(p) for p in glob.glob(f"{out}/chkpts/*.pickle")),
               "log_has_steps": "Time step 500" in open(f"{out}/w2_log.txt").read()}
    # restart under torchrun: every rank reloads rank 0's checkpoint; world, rank, GPU and the scratch folder of
    # ranks > 0 are re-detected, the continued run is sharded again
    dist.barrier()
    sim2 = pv.dmc_restart(potential=pot, chkpt_folder=out, sim_name="w2", additional_timesteps=100)
    assert sim2._world == dist.get_world_size() and sim2._rank == rank and sim2._device == local
    assert (sim2.output_folder == out) == (rank == 0)
    sim2.run()
    n2 = len(sim2.walkers)
    if rank == 0:
        info2 = h5lite.read_h5(f"{out}/w2_sim_info.hdf5")
        res["restart_vref_shape"] = list(info2['vref_vs_tau'].shape)
        res["restart_final_walkers"] = int(n2)
        res["restart_final_pop"] = float(info2['pop_vs_tau'][-1, 1])
        res["restart_zpe"] = float(info2['vref_vs_tau'][300:, 1].mean() / 4.556335281212229e-6)
        print("\nRESULT " + json.dumps(res) + "\n", end="", flush=True)
    dist.destroy_process_group()


USER_TRIAL = """
import numpy as np

ALPHA = 1.4 * %r          # a Gaussian that is NOT the exact ground state: the estimator has a variance, walkers branch


def my_trial(cds):
    return np.exp(-0.5 * ALPHA * cds ** 2).squeeze()


def my_derivs(cds):
    x = cds
    return -ALPHA * x, (ALPHA ** 2 * x ** 2 - ALPHA)
"""


def uimp_main(pv, dist, rank, out):
    """A user trial wave function / derivative function (imp_samp_manager.py:92-139, 197-224) on a sharded run: every rank's
    impsamp.drift sees its own shard once per step, the acceptance fraction and Vref are global."""
    wn = 4.556335281212229e-6
    m, om = pv.Constants.reduced_mass('O-H'), 3700.0 * wn
    mine = os.path.join(out, f"trial_rank{rank}")
    os.makedirs(mine, exist_ok=True)
    with open(os.path.join(mine, "my_trial.py"), "w") as fh:
        fh.write(USER_TRIAL % (m * om))
    d = os.path.join(os.path.dirname(pv.__file__), "sample_potentials", "PythonPots")
    res = {"world": dist.get_world_size()}
    for tag, user_potential, derivs in (("builtin_pot", False, 'my_derivs'), ("user_pot", True, None)):
        imp = pv.ImpSampManager_NoMP(trial_function='my_trial', trial_directory=mine, python_file='my_trial.py', deriv_function=derivs)
        pot = pv.Potential_Direct(potential_function=lambda c: (0.5 * m * om ** 2 * c ** 2).reshape(len(c))) if user_potential else \
            pv.Potential(potential_function='oh_stretch_harm', python_file='harmonicOscillator1D.py', potential_directory=d, num_cores=1)
        sim = pv.DMC_Sim(sim_name=tag, output_folder=os.path.join(out, tag), num_walkers=6000, num_timesteps=700, equil_steps=100,
                         chkpt_every=350, wfn_every=300, desc_wt_steps=20, atoms=['O-H'], delta_t=5, potential=pot,
                         start_structures=np.zeros((1, 1, 1)), imp_samp=imp, imp_samp_oned=True, seed=3)
        assert sim._world == dist.get_world_size()
        sim.run()
        pop = sim._pop_vs_tau
        if rank == 0:
            from pyvibdmc_b200.simulation_utilities import h5lite
            tau = h5lite.read_h5(os.path.join(out, tag, f"{tag}_sim_info.hdf5"))['vref_vs_tau'][:, 0]      # effective time axis
        else:
            tau = np.arange(2.0)
        res[tag] = {"zpe": float(sim._vref_vs_tau[200:].mean() / wn), "pop_min": float(pop.min()), "pop_max": float(pop.max()),
                    "pop_std": float(pop.std()), "n": int(len(sim.walkers)), "final_pop": float(pop[-1]),
                    "tau_increasing": bool(np.all(np.diff(tau) > 0)), "tau_last": float(tau[-1]), "hosted": bool(sim._hosted_imp)}
        dist.barrier()
    if rank == 0:
        print("\nRESULT " + json.dumps(res) + "\n", end="", flush=True)
    dist.destroy_process_group()


def exc_main(pv, dist, rank, out):
    """excited_state_imp_samp (pyvibdmc.py:562-591, 608-611, 810-811) on a sharded run next to the same run on one GPU (rank 0)."""
    wn = 4.556335281212229e-6
    eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
    hd = os.path.join(os.path.dirname(pv.__file__), "sample_potentials", "FortPots", "Partridge_Schwenke_H2O")
    kw = {'dists': [[0, 2], [2, 1]], 'angs': [[0, 2, 1]]}
    res = {"world": dist.get_world_size()}

    def run(tag, distributed):
        pot = pv.Potential(potential_function='water_pot', python_file='h2o_potential.py', potential_directory=hd, num_cores=1)
        wimp = pv.ImpSampManager(trial_function='trial_wavefunction', trial_directory=hd, python_file='call_trl_h2o.py', pot_manager=pot,
                                 deriv_function='dpsi_dx', trial_kwargs=kw, deriv_kwargs=kw)
        sim = pv.DMC_Sim(sim_name=tag, output_folder=os.path.join(out, tag), num_walkers=8000, num_timesteps=700, equil_steps=100,
                         chkpt_every=400, wfn_every=300, desc_wt_steps=20, atoms=['H', 'H', 'O'], delta_t=1, potential=pot,
                         start_structures=eq[None] * 1.01, imp_samp=wimp, excited_state_imp_samp=True, seed=4, distributed=distributed)
        sim.run()
        pop = sim._pop_vs_tau
        return {"world": int(sim._world), "zpe": float(sim._vref_vs_tau[300:].mean() / wn), "pop_min": float(pop.min()),
                "pop_max": float(pop.max()), "n": int(len(sim.walkers)), "final_pop": float(pop[-1])}
    res["sharded"] = run("sh", None)
    dist.barrier()
    if rank == 0:
        res["single"] = run("one", False)
    dist.barrier()
    if rank == 0:
        print("\nRESULT " + json.dumps(res) + "\n", end="", flush=True)
    dist.destroy_process_group()


def imp2_main(pv, dist, rank, out):
    """second_impsamp_displacement (pyvibdmc.py:614-649) on a sharded run: with the exact trial function the local energy is the
    eigenvalue for every walker on every rank -> Vref is exact and nobody branches, for the standard and the second move type."""
    wn = 4.556335281212229e-6
    d = os.path.join(os.path.dirname(pv.__file__), "sample_potentials", "PythonPots")
    res = {"world": dist.get_world_size()}
    for tag, second in (("std", False), ("second", True)):
        imp = pv.ImpSampManager_NoMP(trial_function='trial_harm', trial_directory=d, python_file='harm_trial_wfn.py',
                                     deriv_function='derivative')
        pot = pv.Potential(potential_function='oh_stretch_harm', python_file='harmonicOscillator1D.py', potential_directory=d, num_cores=1)
        sim = pv.DMC_Sim(sim_name=tag, output_folder=out, num_walkers=4000, num_timesteps=300, equil_steps=100, chkpt_every=200,
                         wfn_every=100, desc_wt_steps=20, atoms=['O-H'], delta_t=5, potential=pot, start_structures=np.zeros((1, 1, 1)),
                         imp_samp=imp, imp_samp_oned=True, second_impsamp_displacement=second, seed=2)
        assert sim._world == dist.get_world_size()
        sim.run()
        walkers = sim.walkers
        res[tag] = {"vref_max_dev_cm1": float(np.abs(sim._vref_vs_tau / wn - 1850.0).max()), "pop_constant": bool((sim._pop_vs_tau == 4000).all()),
                    "walker_std": float(walkers.std()), "n": int(len(walkers))}
        dist.barrier()
    # a trial function that is NOT the eigenfunction (harmonic Gaussian on the Morse oscillator): walkers branch, with discrete and
    # with continuous weighting; E0 = omega/2 - omega_x/4 = 1833.4 cm-1
    for weighting in ("discrete", "continuous"):
        imp = pv.ImpSampManager_NoMP(trial_function='trial_harm', trial_directory=d, python_file='harm_trial_wfn.py',
                                     deriv_function='derivative')
        pot = pv.Potential(potential_function='oh_stretch_morse', python_file='morse_osc_1d.py', potential_directory=d, num_cores=1)
        sim = pv.DMC_Sim(sim_name="m" + weighting, output_folder=out, weighting=weighting, num_walkers=20000, num_timesteps=1200,
                         equil_steps=200, chkpt_every=600, wfn_every=400, desc_wt_steps=20, atoms=['O-H'], delta_t=5, potential=pot,
                         start_structures=np.zeros((1, 1, 1)), imp_samp=imp, imp_samp_oned=True, seed=8)
        sim.run()
        walkers = sim.walkers
        cds = walkers[0] if weighting == "continuous" else walkers
        pop = sim._pop_vs_tau
        res["morse_" + weighting] = {"zpe": float(sim._vref_vs_tau[300:].mean() / wn), "vref_std": float(sim._vref_vs_tau[300:].std() / wn),
                                     "pop_min": float(pop.min()), "pop_max": float(pop.max()), "n": int(len(cds)),
                                     "weight_sum": float(walkers[1].sum()) if weighting == "continuous" else float(len(cds)),
                                     "final_pop": float(pop[-1])}
        dist.barrier()
    if rank == 0:
        print("\nRESULT " + json.dumps(res) + "\n", end="", flush=True)
    dist.destroy_process_group()


def ext_main(pv, h5lite, dist, rank, out):
    """A user potential callable (the reference's getpot plug-in) on a sharded run: every rank evaluates its own shard."""
    wn = 4.556335281212229e-6
    mass = pv.Constants.reduced_mass('O-H')        # omega = 3700 cm-1: the sample harmonic oscillator (harmonicOscillator1D.py)
    calls = []

    def user_pot(cds):
        calls.append(len(cds))
        return (0.5 * mass * (3700.0 * wn) ** 2 * cds ** 2).reshape(len(cds))
    res = {"world": dist.get_world_size()}
    for weighting in ("discrete", "continuous"):
        del calls[:]
        sim = pv.DMC_Sim(sim_name=weighting, output_folder=out, weighting=weighting, num_walkers=20000, num_timesteps=1500,
                         equil_steps=300, chkpt_every=700, wfn_every=500, desc_wt_steps=50, atoms=['O-H'], delta_t=10,
                         potential=pv.Potential_Direct(potential_function=user_pot), start_structures=np.zeros((1, 1, 1)),
                         log_every=500, seed=21,
                         # the sharded run also takes the reference's debug switches: the per-step descendant-weight tracker
                         # (pyvibdmc.py:849-852, 868-870) and a non-default Vref feedback strength (DEBUG_alpha, pyvibdmc.py:651-661)
                         **({"DEBUG_save_desc_wt_tracker": True} if weighting == "discrete" else {"DEBUG_alpha": 0.03}))
        assert sim._world == dist.get_world_size()
        sim.run()
        walkers = sim.walkers                     # continuous weighting: (coords, weights), as in the reference
        n = len(walkers[0]) if weighting == "continuous" else len(walkers)
        wsum = float(walkers[1].sum()) if weighting == "continuous" else float(n)
        if rank == 0:
            info = h5lite.read_h5(f"{out}/{weighting}_sim_info.hdf5")
            w = h5lite.read_h5(f"{out}/wfns/{weighting}_wfn_800ts.hdf5")
            pop = info['pop_vs_tau'][:, 1]
            res[weighting] = {"zpe": float(info['vref_vs_tau'][400:, 1].mean() / wn), "final_walkers": int(n), "final_pop": float(pop[-1]),
                              "weight_sum": wsum, "pop_min": float(pop.min()), "pop_max": float(pop.max()), "calls": len(calls),
                              "shard_fraction": float(np.mean(calls)) / float(pop.mean()),
                              "desc_sum": float(w['desc_wts'].sum()), "pop_at_window_end": float(pop[849]),
                              "vref_shape": list(info['vref_vs_tau'].shape)}
            if weighting == "discrete":
                tr = np.load(f"{out}/wfns/discrete_desc_wt_tracker_800ts.npy")
                res[weighting]["tracker_ok"] = bool(tr.shape == (50, len(w['desc_wts'])) and np.array_equal(tr[-1], w['desc_wts'])
                                                    and np.array_equal(tr.sum(axis=1), pop[800:850]))
            else:
                res[weighting]["alpha"] = float(sim._alpha)
        dist.barrier()
    if rank == 0:
        print("\nRESULT " + json.dumps(res) + "\n", end="", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
