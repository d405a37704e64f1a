"""CPU tests (-m "not gpu"): pin the oracle (oracle/) against the reference's shipped data
and against fixtures generated from the unmodified reference (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from conftest import golden

REF_DATA = "/root/reference/pyvibdmc/sample_sim_data"


def test_ps_snapshot_identity(oracle):
    """vref_vs_tau[t-1] == mean(V_PS(coords_t)) - alpha (N-N0)/N0 on a shipped snapshot."""
    g = golden("ps_snapshot_s0_t500.npz")
    cds = g["coords"]
    pred = oracle.water_pot(cds).mean() - (1 / (2 * float(g["delta_t"]))) * (len(cds) - int(g["num_walkers"])) / int(g["num_walkers"])
    assert abs(pred - float(g["vref_expected"])) <= 1e-14 * abs(float(g["vref_expected"]))


def test_ps_all_25_snapshot_means(oracle):
    """The condensed 25-snapshot table stores shipped Vref next to mean(V): identity to 1e-14 relative."""
    tab = golden("ps_h2o_golden.npz")["snapshot_table"]
    assert tab.shape == (25, 5)
    pred = tab[:, 4] - 0.1 * (tab[:, 2] - 8000) / 8000
    assert np.max(np.abs(pred - tab[:, 3]) / np.abs(tab[:, 3])) <= 1e-14


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference sample data only exist in the build container")
def test_ps_all_25_snapshots_from_reference_files(oracle):
    worst = 0.0
    for s in range(5):
        raw = open(f"{REF_DATA}/tutorial_water_{s}_sim_info.hdf5", "rb").read()
        vref = np.frombuffer(raw, "<f8", 10000, 2048).reshape(5000, 2)
        pop = np.frombuffer(raw, "<f8", 10000, 82048).reshape(5000, 2)
        for f in glob.glob(f"{REF_DATA}/wfns/tutorial_water_{s}_wfn_*ts.hdf5"):
            t = int(f.split("_wfn_")[1].split("ts")[0])
            n = int(pop[t - 1, 1])
            assert os.path.getsize(f) == 2048 + 80 * n
            cds = np.frombuffer(open(f, "rb").read(), "<f8", 9 * n, 2048).reshape(n, 3, 3)
            pred = oracle.water_pot(cds).mean() - 0.1 * (n - 8000) / 8000
            worst = max(worst, abs(pred - vref[t - 1, 1]) / abs(vref[t - 1, 1]))
    assert worst <= 1e-14


def test_ps_known_values(oracle):
    g = golden("ps_h2o_golden.npz")
    v = oracle.water_pot(g["coords"])
    assert np.array_equal(v, g["v"])
    # SURVEY 8c sanity values: V(eq) = -0.419 cm-1, V(1.01 eq) = 37.35 cm-1
    assert abs(v[0] / oracle.WAVENUMBERS + 0.4194) < 1e-3 and abs(v[1] / oracle.WAVENUMBERS - 37.3506) < 1e-3


def test_ho_pickle_and_trial(oracle):
    g = golden("ho_golden.npz")
    assert np.array_equal(oracle.oh_stretch_harm(g["pickle_coords"]), g["pickle_pots"])
    assert np.array_equal(oracle.oh_stretch_harm(g["x"]), g["v_oh"])
    assert np.array_equal(oracle.oh_stretch_morse(g["x"]), g["v_morse"])
    assert np.array_equal(oracle.trial_harm(g["x"]), g["psi"])
    d1, d2 = oracle.harm_derivs(g["x"])
    assert np.array_equal(d1, g["dpsi"]) and np.array_equal(d2, g["d2psi"])
    assert float(g["mass"]) == oracle.reduced_mass('O-H') == 1728.2567838775067


@pytest.mark.parametrize("case", ["water", "bigdt", "tiny", "odd"])
def test_branch_discrete_bit_exact(oracle, case):
    g = golden("branch_discrete_golden.npz")
    counts, idx, b, d, p = oracle.birth_or_death_discrete(g[f"{case}_v"], float(g[f"{case}_vref"]), float(g[f"{case}_dt"]),
                                                          g[f"{case}_u"], int(g[f"{case}_n0"]))
    assert np.array_equal(idx, g[f"{case}_idx"])
    assert [b, d, p] == list(g[f"{case}_bdp"])


@pytest.mark.parametrize("tag", ["massive_w", "massive_pop"])
def test_branch_discrete_guard(oracle, tag):
    g = golden("branch_discrete_golden.npz")
    assert str(g[f"{tag}_msg"]) == oracle.MASSIVE
    with pytest.raises(ValueError, match="Massive walker birth or death"):
        oracle.birth_or_death_discrete(np.full(400, 0.02 + float(g[f"{tag}_shift"])), 0.02, 5.0, np.full(400, 0.5), 400)


@pytest.mark.parametrize("case", ["low", "both", "ties"])
def test_branch_continuous(oracle, case):
    g = golden("branch_continuous_golden.npz")
    up = float(g[f"{case}_upper"])
    w, src, nb, mx, mn = oracle.branch_continuous(g[f"{case}_w0"], g[f"{case}_v"], float(g[f"{case}_vref"]),
                                                  float(g[f"{case}_dt"]), float(g[f"{case}_lower"]),
                                                  None if np.isnan(up) else up)
    assert np.array_equal(w, g[f"{case}_w"]) and np.array_equal(src, g[f"{case}_src"])
    assert [nb, mx, mn] == list(g[f"{case}_stats"])
    v = g[f"{case}_v"][src]
    assert np.array_equal(v, g[f"{case}_vout"])
    assert oracle.calc_vref(v, len(w), 0.1, w) == float(g[f"{case}_vref_after"])
    dw = oracle.desc_wts_continuous(w, src, len(w))
    assert np.allclose(dw, g[f"{case}_desc"], rtol=1e-13, atol=0)


def test_vref_and_desc_discrete(oracle):
    g = golden("vref_desc_golden.npz")
    assert oracle.calc_vref(g["v"], int(g["n0"]), float(g["alpha"])) == float(g["vref"])
    assert np.array_equal(oracle.desc_wts_discrete(g["who_from"], int(g["n0"])), g["desc_wts"])


def test_impsamp_water_pieces(oracle):
    g = golden("impsamp_water_golden.npz")
    table = np.load(os.path.join(os.path.dirname(__file__), "..", "pyvibdmc_b200", "sample_potentials",
                                 "FortPots", "Partridge_Schwenke_H2O", "free_oh_wvfn_table.npy"))
    trial = oracle.WaterTrial(table)
    assert np.allclose(trial(g["coords"]), g["psi"], rtol=1e-13, atol=1e-300)
    f, psi, sec = oracle.drift_fd(g["coords"], trial)
    # FD quotients amplify 1-ulp differences in psi by psi/dx^2 ~ 1e6: compare at 1e-8 relative to scale
    assert np.allclose(f, g["f_x"], rtol=1e-9, atol=1e-9 * np.abs(g["f_x"]).max())
    assert np.allclose(sec, g["sec"], rtol=1e-6, atol=1e-7 * np.nanmax(np.abs(g["sec"][20:])))
    m = g["masses"]
    inv_m3 = (1 / np.repeat(m, 3)).reshape(3, 3)[None]
    sig3 = np.repeat(np.sqrt(float(g["dt"]) / m), 3).reshape(3, 3)[None]
    acc = oracle.metropolis(sig3, g["psi"], g["psi_y"], g["coords"], g["y"], inv_m3 * g["f_x"], inv_m3 * g["f_y"],
                            float(g["dt"]))
    assert np.array_equal(acc, g["acc"])
    assert np.array_equal(oracle.local_kin(inv_m3, g["sec"]), g["local_kin"])


def test_descriptor(oracle):
    g = golden("descriptor_golden.npz")
    assert np.allclose(oracle.coulomb_descriptor(g["coords"], g["zs"]), g["coulomb"], rtol=1e-14)


def _replay(oracle, name, **kw):
    g = golden(f"traj_{name}_golden.npz")
    draws = oracle.Draws(g["draw_flat"], g["draw_sizes"])
    return g, draws


def test_traj_ho_discrete(oracle):
    g, draws = _replay(oracle, "ho_disc")
    out = oracle.dmc_loop(np.zeros((300, 1, 1)), g["masses"], float(g["dt"]), 300, 60, oracle.oh_stretch_harm, draws,
                          equil=5, wfn_every=10, desc_steps=4)
    assert np.array_equal(out["pop"], g["pop"]) and np.array_equal(out["vref"], g["vref"])
    assert np.array_equal(out["coords"], g["final_coords"])
    for t in (5, 15, 25, 35, 45, 55):
        assert np.array_equal(out["wfns"][t]["desc_wts"], g[f"wfn{t}_desc_wts"])
        assert np.array_equal(out["wfns"][t]["coords"], g[f"wfn{t}_coords"])


EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])


def test_traj_h2o_discrete(oracle):
    g, draws = _replay(oracle, "h2o_disc")
    out = oracle.dmc_loop(np.repeat(EQ[None] * 1.01, 256, 0), g["masses"], 5.0, 256, 30, oracle.water_pot, draws,
                          equil=5, wfn_every=10, desc_steps=4)
    assert np.array_equal(out["pop"], g["pop"]) and np.array_equal(out["vref"], g["vref"])
    assert np.array_equal(out["coords"], g["final_coords"])
    assert np.array_equal(out["wfns"][15]["desc_wts"], g["wfn15_desc_wts"])


@pytest.mark.parametrize("name,thresh", [("h2o_cont", (0.2, 4.0)), ("h2o_cont_low", (0.3, None))])
def test_traj_h2o_continuous(oracle, name, thresh):
    g, draws = _replay(oracle, name)
    out = oracle.dmc_loop(np.repeat(EQ[None] * 1.01, 256, 0), g["masses"], 5.0, 256, 30, oracle.water_pot, draws,
                          weighting='continuous', cont_thresh=thresh, equil=5, wfn_every=10, desc_steps=4)
    assert np.array_equal(out["vref"], g["vref"]) and np.array_equal(out["pop"], g["pop"])
    assert np.array_equal(out["wts"], g["final_wts"]) and np.array_equal(out["coords"], g["final_coords"])
    assert np.allclose(out["wfns"][15]["desc_wts"], g["wfn15_desc_wts"], rtol=1e-13)
    assert np.array_equal(out["wfns"][15]["parent_wts"], g["wfn15_parent_wts"])


def test_traj_h2o_impsamp(oracle):
    g, draws = _replay(oracle, "h2o_imp")
    table = np.load(os.path.join(os.path.dirname(__file__), "..", "pyvibdmc_b200", "sample_potentials",
                                 "FortPots", "Partridge_Schwenke_H2O", "free_oh_wvfn_table.npy"))
    out = oracle.dmc_loop(np.repeat(EQ[None] * 1.01, 200, 0), g["masses"], 1.0, 200, 16, oracle.water_pot, draws,
                          equil=4, wfn_every=6, desc_steps=3, trial=oracle.WaterTrial(table))
    assert np.array_equal(out["pop"], g["pop"])
    assert np.allclose(out["vref"], g["vref"], rtol=1e-9) and np.allclose(out["eff_ts"], g["eff_ts"], rtol=1e-14)


def test_traj_ho_impsamp_analytic(oracle):
    g, draws = _replay(oracle, "ho_imp")
    out = oracle.dmc_loop(np.zeros((300, 1, 1)), g["masses"], 5.0, 300, 40, oracle.oh_stretch_harm, draws,
                          equil=5, wfn_every=10, desc_steps=4, imp1d_derivs=(oracle.trial_harm, oracle.harm_derivs))
    assert np.array_equal(out["pop"], g["pop"]) and np.allclose(out["vref"], g["vref"], rtol=1e-12)


def test_traj_second_impsamp_displacement(oracle):
    """imp_move_randomly_second_type (pyvibdmc.py:614-649), replayed from the reference's recorded draws."""
    g, draws = _replay(oracle, "h2o_imp2")
    table = np.load(os.path.join(os.path.dirname(__file__), "..", "pyvibdmc_b200", "sample_potentials",
                                 "FortPots", "Partridge_Schwenke_H2O", "free_oh_wvfn_table.npy"))
    out = oracle.dmc_loop(np.repeat(EQ[None] * 1.01, 200, 0), g["masses"], 1.0, 200, 12, oracle.water_pot, draws,
                          equil=4, wfn_every=6, desc_steps=3, trial=oracle.WaterTrial(table), second_disp=True)
    assert np.array_equal(out["pop"], g["pop"])
    assert np.allclose(out["vref"], g["vref"], rtol=1e-9) and np.allclose(out["eff_ts"], g["eff_ts"], rtol=1e-14)
    g, draws = _replay(oracle, "ho_imp2")
    out = oracle.dmc_loop(np.zeros((300, 1, 1)), g["masses"], 5.0, 300, 40, oracle.oh_stretch_harm, draws,
                          equil=5, wfn_every=10, desc_steps=4, imp1d_derivs=(oracle.trial_harm, oracle.harm_derivs), second_disp=True)
    assert np.array_equal(out["pop"], g["pop"]) and np.allclose(out["vref"], g["vref"], rtol=1e-12)


def test_traj_excited_state_impsamp(oracle):
    """excited_state_imp_samp (pyvibdmc.py:562-591, 608-611, 810-811): capped drift and vector score."""
    g, draws = _replay(oracle, "h2o_imp_exc")
    table = np.load(os.path.join(os.path.dirname(__file__), "..", "pyvibdmc_b200", "sample_potentials",
                                 "FortPots", "Partridge_Schwenke_H2O", "free_oh_wvfn_table.npy"))
    out = oracle.dmc_loop(np.repeat(EQ[None] * 1.01, 200, 0), g["masses"], 1.0, 200, 12, oracle.water_pot, draws,
                          equil=4, wfn_every=6, desc_steps=3, trial=oracle.WaterTrial(table), excited=True)
    assert np.array_equal(out["pop"], g["pop"])
    assert np.allclose(out["vref"], g["vref"], rtol=1e-9) and np.allclose(out["eff_ts"], g["eff_ts"], rtol=1e-14)


def test_water_trial_analytic_derivatives(oracle):
    """Oracle restatement of dpsi_dx + ChainRuleHelper against the reference's own output, and the loop that uses it."""
    table = np.load(os.path.join(os.path.dirname(__file__), "..", "pyvibdmc_b200", "sample_potentials",
                                 "FortPots", "Partridge_Schwenke_H2O", "free_oh_wvfn_table.npy"))
    g = golden("impsamp_water_analytic_golden.npz")
    tr = oracle.WaterTrial(table)
    d1, psi, d2 = oracle.drift_analytic(g["coords"], tr)
    assert np.allclose(psi, g["psi"], rtol=1e-14) and np.allclose(d1, g["f_x"], rtol=1e-10, atol=1e-12)
    assert np.allclose(d2, g["sec"], rtol=1e-9, atol=1e-9)
    # analytic and finite-difference derivatives describe the same function
    f_fd, _, s_fd = oracle.drift_fd(g["coords"][:200], tr)
    # (to the accuracy of differencing a piecewise-linear table with dx = 1e-3; second differences of it are noise)
    assert np.allclose(d1[:200], f_fd, rtol=5e-3, atol=2e-3)
    gt, draws = _replay(oracle, "h2o_imp_an")
    out = oracle.dmc_loop(np.repeat(EQ[None] * 1.01, 200, 0), gt["masses"], 1.0, 200, 12, oracle.water_pot, draws,
                          equil=4, wfn_every=6, desc_steps=3, trial=tr, analytic=True)
    assert np.array_equal(out["pop"], gt["pop"])
    assert np.allclose(out["vref"], gt["vref"], rtol=1e-9) and np.allclose(out["eff_ts"], gt["eff_ts"], rtol=1e-14)


def test_nn_float64_golden_pins_the_float32_oracle(oracle):
    """a6: the float32 NumPy forward pass (the stand-in for TensorFlow's float32 one) against the float64 evaluation of the same
    weights; the weights themselves are the reference's sample_h4o2_nn.h5 bit for bit when it can be read here."""
    g = golden("nn_h4o2_f64_golden.npz")
    p = np.load(os.path.join(os.path.dirname(__file__), "..", "pyvibdmc_b200", "sample_potentials", "TensorflowPots", "sample_h4o2_nn_packed.npy"))
    o, w = 0, []
    for k, m in ((15, 120), (120, 120), (120, 120), (120, 1)):
        W = p[o:o + k * m].reshape(k, m); o += k * m
        b = p[o:o + m]; o += m
        w.append((W, b))
    desc = oracle.coulomb_descriptor(g["coords"], [8, 1, 1, 8, 1, 1])
    e64, e32 = oracle.nn_forward_f64(desc, w), oracle.nn_forward_f32(desc, w)
    assert np.allclose(e64, g["e64"], rtol=1e-13, atol=0)
    assert np.abs(e32 - g["e64"]).max() < 1.0e-6 * np.abs(g["e64"]).max()
    assert abs(e64[0] / 4.556335281212229e-6 - 7.62) < 0.05        # equilibrium water dimer (reference tests/test_analysis.py:77-82)
    h5 = "/root/reference/pyvibdmc/sample_potentials/TensorflowPots/sample_h4o2_nn.h5"
    if os.path.exists(h5):                                          # build container only: the packed copy IS the reference's model
        from pyvibdmc_b200.simulation_utilities import h5lite
        ww = h5lite.read_h5(h5)
        parts = []
        for layer in ("dense", "dense_1", "dense_2", "dense_3"):
            parts += [ww[f"model_weights/{layer}/{layer}/kernel:0"].ravel(), ww[f"model_weights/{layer}/{layer}/bias:0"].ravel()]
        assert np.array_equal(np.concatenate(parts).astype(np.float32), p)
