"""GPU tests (-m gpu) of the drop-in user API: DMC_Sim(...).run(), outputs on disk, restart.
They mirror the reference's own integration tests (pyvibdmc/tests/test_pyvibdmc.py, test_imp_samp.py)
but assert results instead of `assert True`."""
import glob
import os
import pickle

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
WN = 4.556335281212229e-6
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])


@pytest.fixture(scope="module")
def pv():
    import pyvibdmc_b200 as pv
    assert pv.kernels.device_count() > 0
    return pv


def sample_dir(pv, *parts):
    return os.path.join(os.path.dirname(pv.__file__), "sample_potentials", *parts)


def ho_potential(pv, cores=2):
    return pv.Potential(potential_function='oh_stretch_harm', python_file='harmonicOscillator1D.py',
                        potential_directory=sample_dir(pv, "PythonPots"), num_cores=cores)


def water_potential(pv):
    return pv.Potential(potential_function='water_pot', python_file='h2o_potential.py',
                        potential_directory=sample_dir(pv, "FortPots", "Partridge_Schwenke_H2O"), num_cores=2)


def read_h5(path):
    from pyvibdmc_b200.simulation_utilities import h5lite
    return h5lite.read_h5(path)


def test_config1_harmonic_full_run_outputs_and_zpe(pv, tmp_path):
    """BASELINE config 1: 1-D HO, discrete, 1000 walkers, 5000 steps, dt = 10 -> ZPE 1850 cm-1 (+ dt bias)."""
    out = str(tmp_path / "ho")
    sim = pv.DMC_Sim(sim_name="ho", output_folder=out, weighting='discrete', num_walkers=1000, num_timesteps=5000,
                     equil_steps=500, chkpt_every=1000, wfn_every=1000, desc_wt_steps=100, atoms=['O-H'], delta_t=10,
                     potential=ho_potential(pv), start_structures=np.zeros((1, 1, 1)), log_every=50, seed=11)
    sim.run()
    info = read_h5(f"{out}/ho_sim_info.hdf5")
    assert set(info) == {'vref_vs_tau', 'pop_vs_tau', 'atomic_nums', 'atomic_masses'}
    assert info['vref_vs_tau'].shape == (5000, 2) and info['pop_vs_tau'].shape == (5000, 2)
    assert np.array_equal(info['vref_vs_tau'][:, 0], np.arange(5000) * 10)
    zpe = info['vref_vs_tau'][1250:, 1].mean() / WN
    assert abs(zpe - 1852.5) < 20, zpe
    assert info['pop_vs_tau'][:, 1].min() > 500 and info['pop_vs_tau'][:, 1].max() < 1500
    wf = sorted(os.path.basename(p) for p in glob.glob(f"{out}/wfns/*.hdf5"))
    assert wf == sorted(f"ho_wfn_{t}ts.hdf5" for t in (500, 1500, 2500, 3500, 4500))
    w = read_h5(f"{out}/wfns/ho_wfn_1500ts.hdf5")
    assert set(w) == {'coords', 'desc_wts'} and w['coords'].shape[1:] == (1, 1) and len(w['coords']) == len(w['desc_wts'])
    assert w['desc_wts'].sum() == info['pop_vs_tau'][1599, 1]          # descendants after 100 steps == population then
    # only the final checkpoint remains, with the reference's attribute names
    ck = glob.glob(f"{out}/chkpts/*.pickle")
    assert [os.path.basename(c) for c in ck] == ["ho_4999.pickle"]
    obj = pickle.load(open(ck[0], "rb"))
    ref_attrs = set(golden("traj_ho_disc_golden.npz")["pickle_attrs"].tolist())
    assert ref_attrs - set(obj.__dict__) <= {'_mass_counter', '_factor_per_change'}, ref_attrs - set(obj.__dict__)
    assert obj._walker_coords.shape == (int(info['pop_vs_tau'][-1, 1]), 1, 1) and obj.cur_timestep == 4999
    k = 0.5 * pv.Constants.reduced_mass('O-H') * pv.Constants.convert(3700., 'wavenumbers') ** 2
    assert np.array_equal(obj._walker_pots, np.squeeze(k * obj._walker_coords ** 2))
    log = open(f"{out}/ho_log.txt").read()
    assert "Simulation ho starting at step 0" in log and "Birth/Death at time step 4950:" in log
    assert "Checkpointing, time step 1000" in log and "Simulation has finished." in log
    assert np.array_equal(sim.vref_vs_tau[:, 1], info['vref_vs_tau'][:4999, 1])

    # restart from the final checkpoint for 1000 more steps (reference tests/test_pyvibdmc.py:152-169)
    sim2 = pv.dmc_restart(potential=ho_potential(pv), chkpt_folder=out, sim_name="ho", additional_timesteps=1000)
    sim2.run()
    info2 = read_h5(f"{out}/ho_sim_info.hdf5")
    assert info2['vref_vs_tau'].shape == (6000, 2)
    assert np.array_equal(info2['vref_vs_tau'][:4999, 1], info['vref_vs_tau'][:4999, 1])
    assert abs(info2['vref_vs_tau'][5000:, 1].mean() / WN - 1852.5) < 40


def test_external_potential_equals_builtin(pv, tmp_path, oracle):
    """A user callable (host NumPy) and the built-in kernel give the same trajectory for the same seed."""
    res = []
    for tag, pot in (("builtin", ho_potential(pv)), ("external", pv.Potential_Direct(potential_function=oracle.oh_stretch_harm))):
        sim = pv.DMC_Sim(sim_name=tag, output_folder=str(tmp_path / tag), num_walkers=800, num_timesteps=120, equil_steps=20,
                         chkpt_every=50, wfn_every=40, desc_wt_steps=10, atoms=['O-H'], delta_t=10, potential=pot,
                         start_structures=np.zeros((1, 1, 1)), seed=5)
        sim.run()
        res.append((sim._vref_vs_tau.copy(), sim._pop_vs_tau.copy(), sim.walkers.copy(),
                    read_h5(str(tmp_path / tag / "wfns" / f"{tag}_wfn_60ts.hdf5"))))
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert np.allclose(res[0][0], res[1][0], rtol=1e-13, atol=0)
    assert np.array_equal(res[0][3]['desc_wts'], res[1][3]['desc_wts'])


def test_water_discrete_zpe(pv, tmp_path):
    out = str(tmp_path / "w")
    sim = pv.DMC_Sim(sim_name="water", output_folder=out, weighting='discrete', num_walkers=8000, num_timesteps=3000,
                     equil_steps=500, chkpt_every=500, wfn_every=1000, desc_wt_steps=300, atoms=['H', 'H', 'O'], delta_t=5,
                     potential=water_potential(pv), start_structures=EQ[None] * 1.01, log_every=100, seed=3)
    sim.run()
    info = read_h5(f"{out}/water_sim_info.hdf5")
    zpe = info['vref_vs_tau'][1000:, 1].mean() / WN
    assert abs(zpe - 4634.1) < 25, zpe                        # shipped tutorial runs: 4634.1 +- 2.2 (8000 x 5000)
    assert list(info['atomic_nums']) == [1, 1, 10]            # H, H, O in the reference's mass-table numbering
    w = read_h5(f"{out}/wfns/water_wfn_1500ts.hdf5")
    assert w['coords'].shape[1:] == (3, 3)


def test_continuous_weighting_with_wfn_dumps(pv, tmp_path):
    out = str(tmp_path / "c")
    sim = pv.DMC_Sim(sim_name="cont", output_folder=out, weighting='continuous', num_walkers=4000, num_timesteps=1200,
                     equil_steps=300, chkpt_every=400, wfn_every=300, desc_wt_steps=50, atoms=['H', 'H', 'O'], delta_t=5,
                     potential=water_potential(pv), start_structures=EQ[None] * 1.01, log_every=50, seed=9)
    sim.run()
    info = read_h5(f"{out}/cont_sim_info.hdf5")
    assert abs(info['pop_vs_tau'][-1, 1] - 4000) < 400
    w = read_h5(f"{out}/wfns/cont_wfn_600ts.hdf5")
    assert set(w) == {'coords', 'desc_wts', 'parent_wts'} and len(w['coords']) == 4000
    assert abs(w['desc_wts'].sum() - info['pop_vs_tau'][649, 1]) < 1e-6 * 4000
    coords, wts = sim.walkers
    assert coords.shape == (4000, 3, 3) and wts.shape == (4000,) and wts.min() > 0
    zpe = info['vref_vs_tau'][600:, 1].mean() / WN
    assert 4400 < zpe < 4850, zpe
    assert "Walkers Branched" in open(f"{out}/cont_log.txt").read()


def test_importance_sampling_runs(pv, tmp_path):
    d = sample_dir(pv, "PythonPots")
    imp = pv.ImpSampManager_NoMP(trial_function='trial_harm', trial_directory=d, python_file='harm_trial_wfn.py',
                                 deriv_function='derivative')
    sim = pv.DMC_Sim(sim_name="hoimp", output_folder=str(tmp_path / "i"), num_walkers=1000, num_timesteps=400, equil_steps=100,
                     chkpt_every=200, wfn_every=100, desc_wt_steps=20, atoms=['O-H'], delta_t=5, potential=ho_potential(pv),
                     start_structures=np.zeros((1, 1, 1)), imp_samp=imp, imp_samp_oned=True, seed=2)
    sim.run()
    # exact trial wfn => zero-variance estimator: Vref == hbar*omega/2 for every step, population never changes
    assert np.allclose(sim._vref_vs_tau / WN, 1850.0, atol=1e-6) and (sim._pop_vs_tau == 1000).all()
    hd = sample_dir(pv, "FortPots", "Partridge_Schwenke_H2O")
    wimp = pv.ImpSampManager(trial_function='trial_wavefunction', trial_directory=hd, python_file='call_trl_h2o.py',
                             pot_manager=water_potential(pv), trial_kwargs={'dists': [[0, 2], [2, 1]], 'angs': [[0, 2, 1]]},
                             deriv_kwargs={'dists': [[0, 2], [2, 1]], 'angs': [[0, 2, 1]]})
    sim = pv.DMC_Sim(sim_name="wimp", output_folder=str(tmp_path / "wi"), num_walkers=2000, num_timesteps=500, equil_steps=100,
                     chkpt_every=250, wfn_every=200, desc_wt_steps=20, atoms=['H', 'H', 'O'], delta_t=1,
                     potential=water_potential(pv), start_structures=EQ[None] * 1.01, imp_samp=wimp, seed=4)
    sim.run()
    info = read_h5(str(tmp_path / "wi" / "wimp_sim_info.hdf5"))
    assert np.all(np.diff(info['vref_vs_tau'][:, 0]) > 0) and info['vref_vs_tau'][-1, 0] <= 500.0   # effective time axis
    zpe = info['vref_vs_tau'][250:, 1].mean() / WN
    assert 4450 < zpe < 4850, zpe
    assert "Metropolis rejected" in open(str(tmp_path / "wi" / "wimp_log.txt")).read()


def test_massive_event_raises_and_still_writes_outputs(pv, tmp_path):
    out = str(tmp_path / "m")
    bad = np.repeat(EQ[None], 1000, axis=0)
    bad[:100] *= 3.0
    sim = pv.DMC_Sim(sim_name="boom", output_folder=out, num_walkers=1000, num_timesteps=50, equil_steps=10, chkpt_every=20,
                     wfn_every=20, desc_wt_steps=5, atoms=['H', 'H', 'O'], delta_t=200, potential=water_potential(pv),
                     start_structures=bad, seed=1)
    with pytest.raises(ValueError, match="Massive walker birth or death event!!!!!!! Dying..."):
        sim.run()
    assert os.path.exists(f"{out}/boom_sim_info.hdf5") and len(glob.glob(f"{out}/chkpts/boom_*.pickle")) == 1
    assert "Final checkpoint is written" in open(f"{out}/boom_log.txt").read()


def test_nn_water_dimer_dmc(pv, tmp_path):
    """BASELINE config 5 in miniature: (H2O)2 with the shipped NN surface (Coulomb descriptor + MLP on tcgen05)."""
    import importlib.util
    d = sample_dir(pv, "TensorflowPots")
    spec = importlib.util.spec_from_file_location("call_sample_model_b200", os.path.join(d, "call_sample_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = mod.load_packed_weights()
    pot = pv.NN_Potential(potential_function='sample_h4o2_pot', python_file='call_sample_model.py', potential_directory=d,
                          model=model, pot_kwargs={'descriptor': None, 'batch_size': 100})
    dimer = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                      [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]]) / 0.529177
    v0 = pot.getpot(dimer[None])
    assert 0.0 < v0[0] / WN < 50.0                               # ~7.6 cm-1 at the dimer minimum
    sim = pv.DMC_Sim(sim_name="dimer", output_folder=str(tmp_path / "nn"), num_walkers=5000, num_timesteps=300, equil_steps=100,
                     chkpt_every=150, wfn_every=100, desc_wt_steps=20, atoms=['O', 'H', 'H'] * 2, delta_t=5, potential=pot,
                     start_structures=dimer[None], seed=8, log_every=50)
    sim.run()
    assert np.isfinite(sim._vref_vs_tau).all() and 2500 < sim._pop_vs_tau[-1] < 7500
    zpe = sim._vref_vs_tau[150:].mean() / WN
    assert 5000 < zpe < 15000, zpe                               # two waters' worth of zero-point energy (~9800 cm-1)
    coords = sim.walkers
    assert coords.shape[1:] == (6, 3)
    assert np.allclose(sim._walker_pots, pot.getpot(coords), rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------ loop variants that call back into Python
def _ho_kw(pv, out, **extra):
    kw = dict(sim_name="v", output_folder=out, weighting='discrete', num_walkers=2000, num_timesteps=400, equil_steps=100,
              chkpt_every=1000, wfn_every=200, desc_wt_steps=20, atoms=['O-H'], delta_t=10, potential=ho_potential(pv, cores=1),
              start_structures=np.zeros((1, 1, 1)), log_every=100, seed=5)
    kw.update(extra)
    return kw


def test_adiabatic_dmc_shifts_the_reference_energy(pv, tmp_path):
    """pyvibdmc.py:820-825: V += lambda(t) * W(x).  With W = 1 the energy follows the lambda ramp."""
    lam0, dlam = 1000 * WN, 5 * WN
    out = str(tmp_path / "ad")
    sim = pv.DMC_Sim(**_ho_kw(pv, out, adiabatic_dmc={'initial_lambda': lam0, 'lambda_change': dlam, 'equil_time': 200,
                                                      'observable_func': lambda c: np.ones(len(c))}))
    assert np.array_equal(sim.ad_lam_array[:200], np.zeros(200)) and np.allclose(sim.ad_lam_array[200:], lam0 + dlam * np.arange(200))
    sim.run()
    v = sim.vref_vs_tau[:, 1] / WN
    assert abs(v[100:200].mean() - 1852) < 60 and abs(v[300:].mean() - (1852 + 1000 + 5 * 149.5)) < 80
    assert os.path.exists(f"{out}/v_lambda.npy") and np.array_equal(np.load(f"{out}/v_lambda.npy"), sim.ad_lam_array)


def test_fixed_node_gives_the_first_excited_state(pv, tmp_path):
    """pyvibdmc.py:684-699, 755-757, 795-797: node at x = 0 with the recrossing correction -> 3/2 hbar omega for the HO."""
    from pyvibdmc_b200.simulation_utilities import Constants
    mu = Constants.reduced_mass('O-H')

    def node(cds, step):
        return cds[:, 0, 0]

    def pot_with_node(cds):                       # walkers across the node die (infinite wall), like a user would set it up
        v = 0.5 * mu * (3700 * WN) ** 2 * cds[:, 0, 0] ** 2
        v[cds[:, 0, 0] < 0] = 10.0
        return v

    out = str(tmp_path / "fn")
    kw = _ho_kw(pv, out, num_walkers=4000, num_timesteps=600, delta_t=5, potential=pv.Potential_Direct(potential_function=pot_with_node),
                start_structures=np.full((1, 1, 1), 0.2), fixed_node={'function': node, 'g_matrix': 1.0 / mu})
    sim = pv.DMC_Sim(**kw)
    sim.run()
    zpe = sim.vref_vs_tau[200:, 1].mean() / WN
    assert abs(zpe - 1.5 * 3700) < 150, zpe
    assert (sim.walkers[:, 0, 0] > 0).mean() > 0.999


def test_debug_mass_change_training_dumps_and_tracker(pv, tmp_path):
    out = str(tmp_path / "dbg")
    sim = pv.DMC_Sim(**_ho_kw(pv, out, DEBUG_mass_change={'change_every': 200, 'factor_per_change': 4.0},
                              DEBUG_save_training_every=150, DEBUG_save_desc_wt_tracker=True))
    m0 = sim.masses.copy()
    sim.run()
    assert np.allclose(sim.masses, 4.0 * m0) and np.allclose(sim._sigmas, np.sqrt(10 / (4.0 * m0)))
    # heavier particle in the same potential: omega halves -> the zero-point energy halves
    v = sim.vref_vs_tau[:, 1] / WN
    assert abs(v[100:200].mean() - 1852) < 60 and abs(v[300:].mean() - 1852 / 2) < 60
    for t in (0, 150, 300):
        d = read_h5(f"{out}/v_training_{t}ts.hdf5")
        assert d['coords'].shape[1:] == (1, 1) and len(d['pots']) == len(d['coords']) and 1000 < len(d['coords']) < 3000
        k = 0.5 * m0[0] * (3700 * WN) ** 2
        assert np.allclose(d['pots'], k * d['coords'][:, 0, 0] ** 2, rtol=1e-12, atol=0)
    tr = np.load(f"{out}/wfns/v_desc_wt_tracker_100ts.npy")
    w = read_h5(f"{out}/wfns/v_wfn_100ts.hdf5")
    assert tr.shape == (20, len(w['desc_wts'])) and np.array_equal(tr[-1], w['desc_wts'])
    pop = sim._pop_vs_tau
    assert np.array_equal(tr.sum(axis=1), pop[100:120])       # every parent's descendants add up to the population, step by step


def test_training_dump_before_birth_death(pv, tmp_path):
    out = str(tmp_path / "bb")
    sim = pv.DMC_Sim(**_ho_kw(pv, out, num_timesteps=120, DEBUG_save_training_every=50, DEBUG_save_before_bod=True))
    sim.run()
    pop = sim._pop_vs_tau
    for t in (0, 50, 100):
        d = read_h5(f"{out}/v_training_{t}ts.hdf5")
        n_before = 2000 if t == 0 else int(pop[t - 1])
        assert len(d['coords']) == n_before                   # collected after the move, before birth/death of step t


def test_second_impsamp_displacement_through_dmc_sim(pv, tmp_path):
    """pyvibdmc.py:614-649 through the user API: exact trial wfn => zero-variance estimator also with the second move type."""
    d = sample_dir(pv, "PythonPots")
    imp = pv.ImpSampManager_NoMP(trial_function='trial_harm', trial_directory=d, python_file='harm_trial_wfn.py',
                                 deriv_function='derivative')
    sim = pv.DMC_Sim(sim_name="ho2", output_folder=str(tmp_path / "i2"), num_walkers=1000, num_timesteps=300, equil_steps=100,
                     chkpt_every=200, wfn_every=100, desc_wt_steps=20, atoms=['O-H'], delta_t=5, potential=ho_potential(pv),
                     start_structures=np.zeros((1, 1, 1)), imp_samp=imp, imp_samp_oned=True, second_impsamp_displacement=True, seed=2)
    sim.run()
    assert np.allclose(sim._vref_vs_tau / WN, 1850.0, atol=1e-6) and (sim._pop_vs_tau == 1000).all()
    assert sim.walkers.std() > 0.05           # the walkers do move (diffusion is never rejected)


def test_water_importance_sampling_with_analytic_derivatives(pv, tmp_path):
    """deriv_function='dpsi_dx' (reference call_trl_h2o.py:101-149): same physics as the finite-difference run, one psi evaluation."""
    hd = sample_dir(pv, "FortPots", "Partridge_Schwenke_H2O")
    kw = {'dists': [[0, 2], [2, 1]], 'angs': [[0, 2, 1]]}
    wimp = pv.ImpSampManager(trial_function='trial_wavefunction', trial_directory=hd, python_file='call_trl_h2o.py',
                             pot_manager=water_potential(pv), deriv_function='dpsi_dx', trial_kwargs=kw, deriv_kwargs=kw)
    assert wimp.gpu_spec()["trial"] == pv._capi.TRIAL_H2O_AN
    sim = pv.DMC_Sim(sim_name="wan", output_folder=str(tmp_path / "wa"), num_walkers=4000, num_timesteps=600, equil_steps=100,
                     chkpt_every=300, wfn_every=200, desc_wt_steps=20, atoms=['H', 'H', 'O'], delta_t=1,
                     potential=water_potential(pv), start_structures=EQ[None] * 1.01, imp_samp=wimp, seed=4)
    sim.run()
    zpe = sim.vref_vs_tau[300:, 1].mean() / WN
    assert 4500 < zpe < 4800, zpe


USER_TRIAL = '''
import numpy as np

ALPHA = 1.4 * %r          # a Gaussian that is NOT the exact ground state: the estimator has a variance, walkers branch


def my_trial(cds):
    return np.exp(-0.5 * ALPHA * cds ** 2).squeeze()


def my_derivs(cds):
    x = cds
    return -ALPHA * x, (ALPHA ** 2 * x ** 2 - ALPHA)
'''


@pytest.mark.parametrize("derivs", ["finite_difference", "user_function"])
@pytest.mark.parametrize("user_potential", [False, True])
def test_importance_sampling_with_user_trial_function(pv, tmp_path, derivs, user_potential, oracle):
    """any trial / derivative callable through ImpSampManager (imp_samp_manager.py:92-139, 197-224), with a shipped
    or a user potential: the reference's main importance-sampling use case runs on the GPU path (host answers drift once per step)."""
    m, om = oracle.reduced_mass('O-H'), 3700.0 * WN
    (tmp_path / "my_trial.py").write_text(USER_TRIAL % (m * om))
    imp = pv.ImpSampManager_NoMP(trial_function='my_trial', trial_directory=str(tmp_path), python_file='my_trial.py',
                                 deriv_function=None if derivs == "finite_difference" else 'my_derivs')
    pot = pv.Potential_Direct(potential_function=lambda c: (0.5 * m * om ** 2 * c ** 2).squeeze()) if user_potential else ho_potential(pv)
    name = f"uimp_{derivs[0]}{int(user_potential)}"
    sim = pv.DMC_Sim(sim_name=name, output_folder=str(tmp_path / name), num_walkers=3000, num_timesteps=700, equil_steps=100,
                     chkpt_every=350, wfn_every=300, desc_wt_steps=20, atoms=['O-H'], delta_t=5, potential=pot,
                     start_structures=np.zeros((1, 1, 1)), imp_samp=imp, imp_samp_oned=True, seed=3)
    sim.run()
    info = read_h5(str(tmp_path / name / f"{name}_sim_info.hdf5"))
    zpe = info['vref_vs_tau'][200:, 1].mean() / WN
    assert abs(zpe - 1850.0) < 25, zpe                               # exact 1850: importance sampling removes most of the noise
    assert np.all(np.diff(info['vref_vs_tau'][:, 0]) > 0) and info['vref_vs_tau'][-1, 0] <= 3500.0    # effective time axis
    pop = info['pop_vs_tau'][:, 1]
    assert pop.min() > 1500 and pop.max() < 4500 and pop.std() > 0    # not the zero-variance case: walkers do branch
    log = open(str(tmp_path / name / f"{name}_log.txt")).read()
    assert "Metropolis rejected" in log
    f_x, psi, sec = sim.f_x, sim.psi_1, sim.psi_sec_der
    assert f_x.shape == sim.walkers.shape and psi.shape == (len(sim.walkers),) and sec.shape == f_x.shape


def test_device_tensor_potential_plugin(pv, tmp_path, oracle):
    """Potential_Direct(device=True): the callable receives a zero-copy CUDA tensor view of the walkers and returns a CUDA
    tensor; the run is identical (same seed, same arithmetic in float64) to the same potential evaluated by the host path."""
    import torch
    m, om = oracle.reduced_mass('O-H'), 3700.0 * WN
    k = 0.5 * m * om ** 2
    seen = {}

    def gpu_model(cds):
        assert isinstance(cds, torch.Tensor) and cds.is_cuda and cds.dtype == torch.float64 and cds.shape[1:] == (1, 1)
        seen["ptr"] = cds.data_ptr()
        return (k * cds * cds).reshape(-1)

    runs = []
    for dev_flag in (True, False):
        pot = pv.Potential_Direct(potential_function=gpu_model, device=True) if dev_flag else \
            pv.Potential_Direct(potential_function=lambda c: (k * c * c).squeeze())
        name = f"devpot{int(dev_flag)}"
        sim = pv.DMC_Sim(sim_name=name, output_folder=str(tmp_path / name), num_walkers=2000, num_timesteps=300, equil_steps=50,
                         chkpt_every=150, wfn_every=100, desc_wt_steps=20, atoms=['O-H'], delta_t=10, potential=pot,
                         start_structures=np.zeros((1, 1, 1)), seed=6)
        sim.run()
        runs.append((sim._vref_vs_tau.copy(), sim._pop_vs_tau.copy(), sim.walkers.copy()))
    assert "ptr" in seen                                            # the model really ran on device memory
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    assert np.allclose(runs[0][0], runs[1][0], rtol=1e-13)
    zpe = runs[0][0][100:].mean() / WN
    assert abs(zpe - 1852) < 40, zpe


def test_device_array_interfaces(pv):
    """kernels.DeviceArray speaks __cuda_array_interface__; device_pointer accepts torch tensors, DLPack capsules' owners and DeviceArray."""
    import torch
    K, capi = pv.kernels, pv._capi
    sim = K.DeviceSim(3, 3, [1837.15, 1837.15, 29156.9], 500, 5.0, capi.POT_EXTERNAL, seed=1)
    start = EQ[None] * 1.01 + np.random.default_rng(0).normal(0, 0.01, (500, 3, 3))
    sim.upload(start)
    arr = sim.coords_device()
    t = arr.torch()
    assert t.shape == (500, 3, 3) and t.is_cuda and np.array_equal(t.cpu().numpy(), start)
    v = torch.from_numpy(K.pes_h2o(start)).cuda()
    sim.set_pots_device(v)
    moved = sim.ext_move_device().torch().clone()
    assert torch.all(torch.abs(moved - torch.from_numpy(start).cuda()) < 1.0)
    v2 = torch.from_numpy(K.pes_h2o(moved.cpu().numpy())).cuda()
    sim.ext_finish_device(v2.to(torch.float32).to(torch.float64))       # any float64 CUDA tensor; dtype / layout checked
    st = sim.state()
    assert st["step"] == 1 and 250 < st["n"] < 750
    with pytest.raises(ValueError):
        K.device_pointer(torch.zeros(4, dtype=torch.float64))          # a host tensor is refused
    sim.close()


@pytest.mark.parametrize("weighting", ["discrete", "continuous"])
def test_restart_inside_descendant_weighting_window(pv, tmp_path, weighting):
    """pyvibdmc.py:299-338: a checkpoint written while a descendant-weighting window is open carries _who_from / _parent(_wts);
    dmc_restart resumes the window and the wave function file of that window is written after the restart."""
    out = str(tmp_path / f"dwr_{weighting}")
    kw = dict(sim_name="dwr", output_folder=out, weighting=weighting, num_walkers=2000, num_timesteps=125, equil_steps=20,
              chkpt_every=1000, wfn_every=100, desc_wt_steps=50, atoms=['O-H'], delta_t=10, potential=ho_potential(pv),
              start_structures=np.zeros((1, 1, 1)), seed=13)
    sim = pv.DMC_Sim(**kw)
    sim.run()                                           # ends at step 124, inside the window opened at step 120 (equil + wfn_every)
    assert [os.path.basename(f) for f in glob.glob(f"{out}/wfns/*.hdf5")] == ["dwr_wfn_20ts.hdf5"]
    ck = pickle.load(open(glob.glob(f"{out}/chkpts/dwr_*.pickle")[0], "rb"))
    assert ck._desc_wt and len(ck._who_from) == len(ck._walker_coords) and ck._parent is not None
    n_parent = len(ck._parent)
    assert n_parent == (int(sim._pop_vs_tau[119]) if weighting == 'discrete' else 2000) and ck._who_from.max() < n_parent
    sim2 = pv.dmc_restart(potential=ho_potential(pv), chkpt_folder=out, sim_name="dwr", additional_timesteps=100)
    sim2.run()
    w = read_h5(f"{out}/wfns/dwr_wfn_120ts.hdf5")
    assert np.array_equal(w['coords'], ck._parent) and len(w['desc_wts']) == n_parent
    info = read_h5(f"{out}/dwr_sim_info.hdf5")
    if weighting == "discrete":
        assert w['desc_wts'].sum() == info['pop_vs_tau'][169, 1]         # descendants at the window's last step
        # walkers that had already died out before the checkpoint have no descendants afterwards either
        alive = np.zeros(n_parent, bool)
        alive[ck._who_from] = True
        assert np.all(w['desc_wts'][~alive] == 0) and w['desc_wts'][alive].sum() == w['desc_wts'].sum()
    else:
        assert abs(w['desc_wts'].sum() - info['pop_vs_tau'][169, 1]) < 1e-6 * 2000 and 'parent_wts' in w


def test_water_zpe_matches_shipped_reference_runs_statistically(pv):
    """north_star: "full runs must reproduce the reference's zero-point energy within combined statistical error".  Five seeds
    at the size of the reference's own shipped tutorial runs (8 000 walkers x 5 000 steps, dt = 5; their five
    `tutorial_water_*_sim_info.hdf5` give 4634.1 cm-1 with a run-to-run sigma of 2.2): the means must agree within
    three combined standard errors.  Deterministic: Philox streams and exact sums make every seed bit-reproducible."""
    K, capi = pv.kernels, pv._capi
    m = [pv.Constants.mass('H'), pv.Constants.mass('H'), pv.Constants.mass('O')]
    zpes = []
    for seed in range(5):
        sim = K.DeviceSim(3, 3, m, 8000, 5.0, capi.POT_H2O_PS, seed=1000 + seed)
        sim.upload(np.repeat(EQ[None] * 1.01, 8000, axis=0))
        sim.run(5000)
        zpes.append(sim.stats(0, 5000)["vref"][1000:].mean() / WN)
        sim.close()
    zpes = np.array(zpes)
    ours, sem_ours = zpes.mean(), zpes.std(ddof=1) / np.sqrt(len(zpes))
    ref, sem_ref = 4634.1, 2.2 / np.sqrt(5)
    sigma = np.sqrt(sem_ours ** 2 + sem_ref ** 2)
    assert abs(ours - ref) < 3.0 * sigma, (ours, sem_ours, zpes, sigma)
    assert zpes.std(ddof=1) < 6.0, zpes                       # single-run scatter of the same size as the reference's (2.2)
