#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, importable here through the test-only stubs in oracle/refstubs) and the
pinned C oracle.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every fixture stores the inputs next to the reference's outputs so that the oracle
(oracle/dmc_oracle.py) and the CUDA path can be replayed on identical inputs.
Reference entry points exercised (file:line under /root/reference/pyvibdmc):
  pyvibdmc.py:380-454 birth_or_death   :340-356 _branch   :651-661 calc_vref
  pyvibdmc.py:663-672 calc_desc_wts    :701-876 propagate :540-612 move/imp_move
  simulation_utilities/imp_samp.py:21-76 drift/metropolis/local_kin/finite_diff
  sample_potentials/PythonPots/harmonicOscillator1D.py:13-17, harm_trial_wfn.py:6-40
  sample_potentials/FortPots/Partridge_Schwenke_H2O/call_trl_h2o.py:63-78
  simulation_utilities/tensorflow_descriptors/distance_descriptors.py (coulomb; every DistIt variant: distit_golden.npz)
"""
import os, sys, ctypes, pickle, tempfile, shutil, warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "refstubs"), REF]
warnings.filterwarnings("ignore")
import pyvibdmc as pv                                             # noqa: E402  (the reference)
from pyvibdmc.simulation_utilities.imp_samp import ImpSamp        # noqa: E402

lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libpvd_oracle.so"))
lib.oracle_ps_h2o.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]


def water_pot(cds):
    cds = np.ascontiguousarray(cds, dtype=np.float64)
    v = np.empty(len(cds))
    lib.oracle_ps_h2o(cds.ctypes.data, len(cds), v.ctypes.data)
    return v


SAMPLE = f"{REF}/pyvibdmc/sample_sim_data"
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
WAT_ARGS = {'dists': [[0, 2], [2, 1]], 'angs': [[0, 2, 1]]}
TMP = tempfile.mkdtemp(prefix="pvd_golden_")


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print(f"{name}: " + ", ".join(f"{k}{np.shape(v)}" for k, v in kw.items()))


def raw_snapshot(s, t):
    raw = open(f"{SAMPLE}/tutorial_water_{s}_sim_info.hdf5", "rb").read()
    vref = np.frombuffer(raw, "<f8", 10000, 2048).reshape(5000, 2)
    pop = np.frombuffer(raw, "<f8", 10000, 82048).reshape(5000, 2)
    n = int(pop[t - 1, 1])
    w = open(f"{SAMPLE}/wfns/tutorial_water_{s}_wfn_{t}ts.hdf5", "rb").read()
    cds = np.frombuffer(w, "<f8", 9 * n, 2048).reshape(n, 3, 3).copy()
    return cds, vref[t - 1, 1], n


# ---------------------------------------------------------------- A. PES
def gen_pes():
    cds, vref, n = raw_snapshot(0, 500)
    save("ps_snapshot_s0_t500.npz", coords=cds, vref_expected=vref, num_walkers=8000, delta_t=5.0)
    # the 25-snapshot identity, condensed: per snapshot mean(V) from the pinned oracle + shipped Vref
    rows = []
    for s in range(5):
        for t in (500, 1500, 2500, 3500, 4500):
            c, vr, nn = raw_snapshot(s, t)
            rows.append((s, t, nn, vr, water_pot(c).mean()))
    rng = np.random.default_rng(1234)
    pick = cds[rng.choice(n, 1536, replace=False)]
    wild = EQ[None] + rng.normal(0, 0.25, size=(1536, 3, 3))          # far from equilibrium
    mild = EQ[None] * 1.01 + rng.normal(0, 0.03, size=(1024, 3, 3))
    geoms = np.concatenate([EQ[None], EQ[None] * 1.01, pick, wild, mild])
    save("ps_h2o_golden.npz", coords=geoms, v=water_pot(geoms), snapshot_table=np.array(rows))


# ---------------------------------------------------------------- B. HO
def gen_ho():
    with open(f"{SAMPLE}/chkpts/pytest_364.pickle", "rb") as fh:
        sim = pickle.load(fh)
    sys.path.insert(0, f"{REF}/pyvibdmc/sample_potentials/PythonPots")
    import harmonicOscillator1D as ho
    import harm_trial_wfn as htw
    import morse_osc_1d as mo
    cds = np.asarray(sim._walker_coords)
    assert np.array_equal(ho.oh_stretch_harm(cds), sim._walker_pots)
    x = np.random.default_rng(5).normal(0, 0.2, size=(2000, 1, 1))
    d1, d2 = htw.derivative(x)
    save("ho_golden.npz", pickle_coords=cds, pickle_pots=np.asarray(sim._walker_pots),
         x=x, v_oh=ho.oh_stretch_harm(x), v_morse=mo.oh_stretch_morse(x), psi=htw.trial_harm(x),
         dpsi=d1, d2psi=d2, mass=pv.Constants.reduced_mass('O-H'),
         omega=pv.Constants.convert(3700., 'wavenumbers', to_AU=True),
         pickle_attrs=np.array(sorted(sim.__dict__.keys())))


# ---------------------------------------------------------------- helpers: a reference sim object
def make_sim(weighting, n, atoms, start, dt=5.0, **kw):
    pot = pv.Potential_Direct(potential_function=water_pot if len(atoms) == 3 else None)
    return pv.DMC_Sim(sim_name="g", output_folder=os.path.join(TMP, f"o{np.random.randint(1 << 30)}"),
                      weighting=weighting, num_walkers=n, num_timesteps=10, equil_steps=2, chkpt_every=100,
                      wfn_every=5, desc_wt_steps=2, atoms=atoms, delta_t=dt, potential=pot,
                      start_structures=start, **kw)


# ---------------------------------------------------------------- C. discrete birth/death
def gen_branch_discrete():
    out = {}
    rng = np.random.default_rng(77)
    cases = [("water", 5000, 5.0, 0.004), ("bigdt", 3000, 60.0, 0.004), ("tiny", 1, 5.0, 0.0), ("odd", 777, 1.0, 0.02)]
    for name, n, dt, spread in cases:
        sim = make_sim('discrete', n, ["H", "H", "O"], np.tile(EQ, (n, 1, 1)), dt=dt)
        sim._walker_coords = np.arange(n * 9, dtype=float).reshape(n, 3, 3)
        v = 0.021 + spread * rng.standard_gamma(2.0, size=n) - spread * 2
        sim._walker_pots = v.copy()
        sim._vref = float(np.mean(v))
        sim._who_from = np.arange(n)
        sim._desc_wt = True
        seed = 1000 + n
        np.random.seed(seed); u = np.random.random(n)
        np.random.seed(seed)
        b, d, p = sim.birth_or_death()
        assert np.array_equal(sim._walker_coords[:, 0, 0], sim._who_from * 9.0)
        out.update({f"{name}_v": v, f"{name}_vref": sim._vref, f"{name}_dt": dt, f"{name}_u": u,
                    f"{name}_idx": sim._who_from.copy(), f"{name}_bdp": np.array([b, d, p]),
                    f"{name}_n0": n})
    # a massive event must raise ValueError (pyvibdmc.py:397-400 and :409-413)
    for tag, shift in (("massive_w", -40.0), ("massive_pop", -0.15)):
        n = 400
        sim = make_sim('discrete', n, ["H", "H", "O"], np.tile(EQ, (n, 1, 1)), dt=5.0)
        sim._walker_pots = np.full(n, 0.02 + shift); sim._vref = 0.02
        try:
            sim.birth_or_death(); raised = ""
        except ValueError as e:
            raised = str(e)
        out[f"{tag}_msg"] = np.array(raised); out[f"{tag}_shift"] = shift
    save("branch_discrete_golden.npz", **out)


# ---------------------------------------------------------------- D. continuous weighting
def gen_branch_continuous():
    out = {}
    rng = np.random.default_rng(99)
    cases = [("low", 4000, None), ("both", 4000, [0.05, 3.0]), ("ties", 512, [0.3])]
    for name, n, thresh in cases:
        sim = make_sim('continuous', n, ["H", "H", "O"], np.tile(EQ, (n, 1, 1)), dt=5.0, cont_wt_thresh=thresh)
        w0 = np.exp(rng.normal(0, 2.5 if name != "ties" else 1.0, size=n))
        if name == "low":
            w0[rng.choice(n, 40, replace=False)] = 1e-5          # below 1/N0
        if name == "ties":
            w0 = np.round(w0 * 4) / 4 + 0.25                     # many exactly equal weights
            w0[rng.choice(n, 60, replace=False)] = 0.0625
        v = 0.021 + 0.004 * rng.standard_normal(n)
        if name == "ties":
            v[:] = 0.021                                          # exp(0) = 1 keeps the ties exact
        sim._cont_wts = w0.copy(); sim._walker_pots = v.copy(); sim._vref = 0.021
        sim._walker_coords = np.arange(n * 9, dtype=float).reshape(n, 3, 3)
        sim._who_from = np.arange(n); sim._desc_wt = True
        nb, mx, mn = sim.birth_or_death()
        out.update({f"{name}_w0": w0, f"{name}_v": v, f"{name}_vref": 0.021, f"{name}_dt": 5.0,
                    f"{name}_lower": sim._thresh_lower,
                    f"{name}_upper": np.nan if sim._thresh_upper is None else sim._thresh_upper,
                    f"{name}_w": sim._cont_wts.copy(), f"{name}_src": sim._who_from.copy(),
                    f"{name}_vout": sim._walker_pots.copy(), f"{name}_stats": np.array([nb, mx, mn])})
        assert np.array_equal(sim._walker_coords[:, 0, 0], sim._who_from * 9.0)
        sim.calc_vref(); out[f"{name}_vref_after"] = sim._vref
        # descendant weights, continuous flavour (pyvibdmc.py:671-672)
        sim._desc_wts = np.zeros(n); sim.calc_desc_wts(); out[f"{name}_desc"] = sim._desc_wts.copy()
    save("branch_continuous_golden.npz", **out)


# ---------------------------------------------------------------- E. vref / desc wts (discrete)
def gen_vref_desc():
    rng = np.random.default_rng(3)
    n = 3000
    sim = make_sim('discrete', 2800, ["H", "H", "O"], np.tile(EQ, (2800, 1, 1)), dt=5.0)
    v = 0.021 + 0.004 * rng.standard_normal(n)
    sim._walker_pots = v; sim.calc_vref()
    who = np.sort(rng.integers(0, 2800, size=n))
    sim._who_from = who; sim._desc_wts = np.zeros(2800); sim.calc_desc_wts()
    save("vref_desc_golden.npz", v=v, n0=2800, alpha=sim._alpha, vref=sim._vref, who_from=who,
         desc_wts=sim._desc_wts.copy())


# ---------------------------------------------------------------- F. importance sampling pieces
def water_manager():
    d = f"{REF}/pyvibdmc/sample_potentials/FortPots/Partridge_Schwenke_H2O"
    return pv.ImpSampManager_NoMP(trial_function='trial_wavefunction', trial_directory=d,
                                  python_file='call_trl_h2o.py', chdir=True,
                                  trial_kwargs=WAT_ARGS, deriv_kwargs=WAT_ARGS)


def gen_impsamp():
    rng = np.random.default_rng(11)
    n = 1500
    cds = EQ[None] * 1.0 + rng.normal(0, 0.06, size=(n, 3, 3))
    cds[:20] = EQ[None] + rng.normal(0, 0.8, size=(20, 3, 3))       # some leave the [0.5,4] grid
    man = water_manager()
    imp = ImpSamp(man)
    f_x, psi, sec = imp.drift(cds.copy())
    masses = np.array([pv.Constants.mass(a) for a in ["H", "H", "O"]])
    dt = 1.0
    sig = np.sqrt(dt / masses)
    inv_m3 = (1 / np.repeat(masses, 3)).reshape(3, 3)[None]
    sig3 = np.repeat(sig, 3).reshape(3, 3)[None]
    disp = rng.normal(0, 1, size=(n, 3, 3)) * sig3
    d_x = inv_m3 * f_x
    y = cds + disp + d_x * dt
    f_y, psi_y, sec_y = imp.drift(y.copy())
    d_y = inv_m3 * f_y
    acc = ImpSamp.metropolis(sigma_trip=sig3, trial_x=psi, trial_y=psi_y, disp_x=cds, disp_y=y,
                             D_x=d_x, D_y=d_y, dt=dt)
    lk = ImpSamp.local_kin(inv_m3, sec)
    table = np.load(f"{REF}/pyvibdmc/sample_potentials/FortPots/Partridge_Schwenke_H2O/free_oh_wvfn_dense.npy")
    save("impsamp_water_golden.npz", coords=cds, psi=psi, f_x=f_x, sec=sec, disp=disp, y=y, psi_y=psi_y,
         f_y=f_y, sec_y=sec_y, acc=acc, local_kin=lk, masses=masses, dt=dt,
         grid_first=table[0, 0], grid_last=table[0, -1], grid_n=table.shape[1])
    # trial-wfn table itself is data the product needs (5000-pt grid, rows: r, psi, psi', psi'')
    np.save(os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials", "FortPots", "Partridge_Schwenke_H2O", "free_oh_wvfn_table.npy"), table)


def gen_impsamp_analytic():
    """Water trial wfn with the reference's ANALYTIC derivatives (call_trl_h2o.py:101-149 + ChainRuleHelper,
    imp_samp_helper.py:10-209) and a short trajectory that uses them."""
    rng = np.random.default_rng(12)
    n = 1500
    cds = EQ[None] * 1.0 + rng.normal(0, 0.06, size=(n, 3, 3))
    d = f"{REF}/pyvibdmc/sample_potentials/FortPots/Partridge_Schwenke_H2O"
    man = pv.ImpSampManager_NoMP(trial_function='trial_wavefunction', trial_directory=d, python_file='call_trl_h2o.py', chdir=True,
                                 deriv_function='dpsi_dx', trial_kwargs=WAT_ARGS, deriv_kwargs=WAT_ARGS)
    imp = ImpSamp(man)
    f_x, psi, sec = imp.drift(cds.copy())
    save("impsamp_water_analytic_golden.npz", coords=cds, psi=psi, f_x=f_x, sec=sec)
    table = np.load(f"{d}/free_oh_wvfn_dense.npy")
    # the analytic path needs the derivative rows of the shipped table too (rows: r, psi, psi', psi'')
    np.save(os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials", "FortPots", "Partridge_Schwenke_H2O", "free_oh_wvfn_table.npy"), table)
    wpot = pv.Potential_Direct(potential_function=water_pot)
    run_traj("h2o_imp_an", "discrete", 200, 12, ["H", "H", "O"], EQ[None] * 1.01, wpot, 1.0, 13, imp=man, equil=4, wfn=6, desc=3)


# ---------------------------------------------------------------- G. whole-loop trajectories with recorded RNG
class Recorder:
    """Records every np.random.normal / np.random.random draw the reference loop makes."""
    def __init__(self, seed):
        self.kinds, self.draws = [], []
        self._n, self._r = np.random.normal, np.random.random
        np.random.seed(seed)

    def __enter__(self):
        def normal(loc, scale, size=None):
            z = self._n(loc, scale, size=size); self.kinds.append("n"); self.draws.append(np.array(z)); return z

        def random(size=None):
            z = self._r(size); self.kinds.append("u"); self.draws.append(np.array(z)); return z
        np.random.normal, np.random.random = normal, random
        return self

    def __exit__(self, *a):
        np.random.normal, np.random.random = self._n, self._r


def run_traj(name, weighting, n0, T, atoms, start, pot, dt, seed, imp=None, **kw):
    folder = os.path.join(TMP, name)
    sim = pv.DMC_Sim(sim_name=name, output_folder=folder, weighting=weighting, num_walkers=n0, num_timesteps=T,
                     equil_steps=kw.pop("equil", 5), chkpt_every=kw.pop("chk", 1000), wfn_every=kw.pop("wfn", 10),
                     desc_wt_steps=kw.pop("desc", 4), atoms=atoms, delta_t=dt, potential=pot,
                     start_structures=start, imp_samp=imp, log_every=1, **kw)
    with Recorder(seed) as rec:
        sim.run()
    out = {"vref": sim._vref_vs_tau.copy(), "pop": sim._pop_vs_tau.copy(), "n0": n0, "T": T, "dt": dt,
           "final_coords": np.asarray(sim._walker_coords), "final_pots": np.asarray(sim._walker_pots),
           "kinds": np.array(rec.kinds), "masses": np.asarray(sim.masses)}
    # ragged draws -> flat + offsets
    flat = np.concatenate([np.ravel(d) for d in rec.draws])
    out["draw_flat"] = flat
    out["draw_sizes"] = np.array([d.size for d in rec.draws])
    if weighting == "continuous":
        out["final_wts"] = np.asarray(sim._cont_wts)
    if imp is not None:
        out["eff_ts"] = sim.eff_ts.copy()
    wf = sorted(f for f in os.listdir(folder + "/wfns") if f.endswith(".npz"))
    for f in wf:
        d = np.load(f"{folder}/wfns/{f}")
        tag = f.split("_wfn_")[1].split("ts")[0]
        for k in d.files:
            out[f"wfn{tag}_{k}"] = d[k]
    si = np.load(f"{folder}/{name}_sim_info.hdf5.npz")
    for k in si.files:
        out[f"siminfo_{k}"] = si[k]
    out["log_text"] = np.array(open(f"{folder}/{name}_log.txt").read())
    with open([f"{folder}/chkpts/{p}" for p in os.listdir(folder + "/chkpts")][0], "rb") as fh:
        ck = pickle.load(fh)
    out["pickle_attrs"] = np.array(sorted(ck.__dict__.keys()))
    save(f"traj_{name}_golden.npz", **out)


def gen_traj():
    sys.path.insert(0, f"{REF}/pyvibdmc/sample_potentials/PythonPots")
    import harmonicOscillator1D as ho
    hopot = pv.Potential_Direct(potential_function=ho.oh_stretch_harm)
    wpot = pv.Potential_Direct(potential_function=water_pot)
    run_traj("ho_disc", "discrete", 300, 60, ['O-H'], np.zeros((1, 1, 1)), hopot, 10.0, 1)
    run_traj("h2o_disc", "discrete", 256, 30, ["H", "H", "O"], EQ[None] * 1.01, wpot, 5.0, 2)
    run_traj("h2o_cont", "continuous", 256, 30, ["H", "H", "O"], EQ[None] * 1.01, wpot, 5.0, 3,
             cont_wt_thresh=[0.2, 4.0])
    run_traj("h2o_cont_low", "continuous", 256, 30, ["H", "H", "O"], EQ[None] * 1.01, wpot, 5.0, 6,
             cont_wt_thresh=0.3)
    run_traj("h2o_imp", "discrete", 200, 16, ["H", "H", "O"], EQ[None] * 1.01, wpot, 1.0, 4, imp=water_manager(),
             equil=4, wfn=6, desc=3)
    import harm_trial_wfn  # noqa: F401
    d = f"{REF}/pyvibdmc/sample_potentials/PythonPots"
    man = pv.ImpSampManager_NoMP(trial_function='trial_harm', trial_directory=d, python_file='harm_trial_wfn.py',
                                 deriv_function='derivative')
    run_traj("ho_imp", "discrete", 300, 40, ['O-H'], np.zeros((1, 1, 1)), hopot, 5.0, 5, imp=man,
             imp_samp_oned=True)


def gen_traj_variants():
    """Loop variants of SURVEY 8 f-3: second importance-sampling displacement (pyvibdmc.py:614-649)."""
    sys.path.insert(0, f"{REF}/pyvibdmc/sample_potentials/PythonPots")
    import harmonicOscillator1D as ho
    hopot = pv.Potential_Direct(potential_function=ho.oh_stretch_harm)
    wpot = pv.Potential_Direct(potential_function=water_pot)
    run_traj("h2o_imp2", "discrete", 200, 12, ["H", "H", "O"], EQ[None] * 1.01, wpot, 1.0, 8, imp=water_manager(),
             equil=4, wfn=6, desc=3, second_impsamp_displacement=True)
    d = f"{REF}/pyvibdmc/sample_potentials/PythonPots"
    man = pv.ImpSampManager_NoMP(trial_function='trial_harm', trial_directory=d, python_file='harm_trial_wfn.py',
                                 deriv_function='derivative')
    run_traj("ho_imp2", "discrete", 300, 40, ['O-H'], np.zeros((1, 1, 1)), hopot, 5.0, 9, imp=man,
             imp_samp_oned=True, second_impsamp_displacement=True)


def gen_traj_excited():
    """excited_state_imp_samp (pyvibdmc.py:562-591, 810-811): capped drift + vector score."""
    wpot = pv.Potential_Direct(potential_function=water_pot)
    run_traj("h2o_imp_exc", "discrete", 200, 12, ["H", "H", "O"], EQ[None] * 1.01, wpot, 1.0, 10, imp=water_manager(),
             equil=4, wfn=6, desc=3, excited_state_imp_samp=True)


# ---------------------------------------------------------------- H. NN descriptor
def gen_descriptor():
    from pyvibdmc.simulation_utilities.tensorflow_descriptors.distance_descriptors import DistIt
    rng = np.random.default_rng(21)
    dimer = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                      [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]])
    cds = (dimer[None] + rng.normal(0, 0.1, size=(512, 6, 3))) / 0.529177
    cds[0] = dimer / 0.529177                          # the reference's equilibrium dimer (tests/test_analysis.py:77-82)
    coul = DistIt([8, 1, 1] * 2, 'coulomb', force_numpy=True)
    save("descriptor_golden.npz", coords=cds, coulomb=np.asarray(coul.run(cds)), zs=np.array([8, 1, 1] * 2))
    # every DistIt variant (distance_descriptors.py:115-152,177-213): method x atom sorting x group sorting x full matrix
    eq = dimer / 0.529177
    cases, out = distit_cases(), {"coords": cds[:256], "eq_xyz": eq, "zs": np.array([8, 1, 1] * 2)}
    for name, kw in cases.items():
        d = DistIt([8, 1, 1] * 2, force_numpy=True, eq_xyz=eq if kw["method"] == "spf" else None, **kw)
        out[name] = np.asarray(d.run(cds[:256]))
    # larger molecules (column norms over 10 and 16 rows: NumPy still accumulates them sequentially)
    r2 = np.random.default_rng(5)
    eq10 = r2.normal(0, 2.0, size=(10, 3))
    out["coords10"], out["eq10"], out["zs10"] = eq10[None] + r2.normal(0, 0.3, size=(64, 10, 3)), eq10, np.array([8, 1, 1, 8, 1, 1, 8, 1, 1, 6])
    d = DistIt(out["zs10"], "spf", eq_xyz=eq10, sorted_atoms=[[0, 3, 6], [1, 2, 4, 5, 7, 8], [9]], sorted_groups=[[0, 1, 2], [3, 4, 5], [6, 7, 8]],
               full_mat=True, force_numpy=True)
    out["spf10_atoms_groups_full"] = np.asarray(d.run(out["coords10"]))
    out["coords16"] = r2.normal(0, 3.0, size=(16, 3))[None] + r2.normal(0, 0.3, size=(32, 16, 3))
    out["coulomb16_atoms"] = np.asarray(DistIt([1] * 16, "coulomb", sorted_atoms=[list(range(16))], force_numpy=True).run(out["coords16"]))
    # two groups of eight atoms: np.sum over a contiguous axis of length 8 switches to NumPy's unrolled pairwise order
    g8 = [list(range(8)), list(range(8, 16))]
    out["distance16_groups8"] = np.asarray(DistIt([1] * 16, "distance", sorted_groups=g8, force_numpy=True).run(out["coords16"]))
    out["coulomb16_groups8_full"] = np.asarray(DistIt([8, 1] * 8, "coulomb", sorted_groups=g8, full_mat=True, force_numpy=True).run(out["coords16"]))
    out["zs16"] = np.array([8, 1] * 8)
    # mirror-image groups: their totals agree to the last bits, so the swap depends on the ORDER in which NumPy adds the
    # norms (and the squares inside each norm) -- one third of these walkers change when that order is altered
    def mirrored(n, h, seed):
        r = np.random.default_rng(seed)
        half = r.normal(0, 2.0, size=(n, h, 3))
        half[:, :, 0] = np.abs(half[:, :, 0]) + 0.5
        other = half[:, r.permutation(h)].copy()
        other[:, :, 0] *= -1
        return np.concatenate([half, other], axis=1)
    for na, seed in ((16, 11), (10, 4), (6, 6)):
        c, g2 = mirrored(96, na // 2, seed), [list(range(na // 2)), list(range(na // 2, na))]
        out[f"mirror{na}"] = c
        out[f"mirror{na}_groups"] = np.asarray(DistIt([1] * na, "distance", sorted_groups=g2, force_numpy=True).run(c))
        out[f"mirror{na}_groups_one"] = np.asarray(DistIt([1] * na, "distance", sorted_groups=g2, force_numpy=True).run(c[:1]))
        out[f"mirror{na}_spf_full"] = np.asarray(DistIt([1] * na, "spf", eq_xyz=c[0], sorted_groups=g2, full_mat=True, force_numpy=True).run(c[1:]))
    # r_eq is made from TWO copies of eq_xyz (:75); this mirror-symmetric equilibrium structure sorts differently when only one
    # copy is pushed through sort_groups (NumPy adds a single walker's group totals in another order)
    c = mirrored(41, 8, 123)
    out["mirror16b"] = c
    out["mirror16b_spf_full"] = np.asarray(DistIt([1] * 16, "spf", eq_xyz=c[0], sorted_groups=[list(range(8)), list(range(8, 16))],
                                                  full_mat=True, force_numpy=True).run(c[1:]))
    save("distit_golden.npz", **out)


def distit_cases():
    """name -> DistIt keyword arguments; shared with tests/test_descriptors.py (kept literal there)."""
    sa1, sa2 = [[0], [1, 2], [3], [4, 5]], [[0, 3], [1, 2, 4, 5]]
    sg1, sg2 = [[0, 1, 2], [3, 4, 5]], [[1, 2], [4, 5]]
    return {"distance": dict(method="distance"), "spf": dict(method="spf"), "coulomb_full": dict(method="coulomb", full_mat=True),
            "distance_full": dict(method="distance", full_mat=True), "spf_full": dict(method="spf", full_mat=True),
            "coulomb_atoms": dict(method="coulomb", sorted_atoms=sa1), "coulomb_atoms2_full": dict(method="coulomb", sorted_atoms=sa2, full_mat=True),
            "distance_groups": dict(method="distance", sorted_groups=sg1), "coulomb_groups2_full": dict(method="coulomb", sorted_groups=sg2, full_mat=True),
            "spf_atoms_groups": dict(method="spf", sorted_atoms=sa1, sorted_groups=sg1),
            "spf_atoms2_full": dict(method="spf", sorted_atoms=sa2, full_mat=True),
            "coulomb_atoms_groups": dict(method="coulomb", sorted_atoms=sa1, sorted_groups=sg1)}


if __name__ == "__main__":
    try:
        gens = {"pes": gen_pes, "ho": gen_ho, "branch_discrete": gen_branch_discrete, "branch_continuous": gen_branch_continuous,
                "vref_desc": gen_vref_desc, "impsamp": gen_impsamp, "traj": gen_traj, "traj_variants": gen_traj_variants, "traj_excited": gen_traj_excited, "impsamp_analytic": gen_impsamp_analytic,
                "descriptor": gen_descriptor}
        for name in (sys.argv[1:] or list(gens)):          # default: everything; or name the generators to (re)run
            gens[name]()
    finally:
        shutil.rmtree(TMP, ignore_errors=True)
