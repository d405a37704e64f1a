"""GPU parity tests (-m gpu): continuous weighting, importance sampling, NN potential."""
import os

import numpy as np
import pytest

from conftest import golden, ROOT

pytestmark = pytest.mark.gpu
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
WN = 4.556335281212229e-6
H2O_DIR = os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials", "FortPots", "Partridge_Schwenke_H2O")


@pytest.fixture(scope="module")
def K():
    from pyvibdmc_b200 import kernels
    assert kernels.device_count() > 0
    return kernels


def water_table():
    import importlib.util
    spec = importlib.util.spec_from_file_location("call_trl_h2o_b200", os.path.join(H2O_DIR, "call_trl_h2o.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.packed_table()


class Replay:
    def __init__(self, g):
        self.flat, self.sizes, self.k, self.off = g["draw_flat"], g["draw_sizes"], 0, 0

    def take(self, n):
        assert self.sizes[self.k] == n, (self.k, int(self.sizes[self.k]), n)
        out = self.flat[self.off:self.off + n]
        self.k += 1
        self.off += n
        return out

    def normal(self, n, a, d):
        return np.ascontiguousarray(self.take(n * a * d).reshape(n, d, a).transpose(0, 2, 1))


# ------------------------------------------------------------------ continuous weighting
@pytest.mark.parametrize("case", ["low", "ties"])
def test_branch_continuous_golden_exact(K, case):
    g = golden("branch_continuous_golden.npz")
    w, src, nb, mx, mn = K.branch_continuous(g[f"{case}_w0"], g[f"{case}_v"], float(g[f"{case}_vref"]), float(g[f"{case}_dt"]),
                                             float(g[f"{case}_lower"]), None)
    # CUDA's exp and glibc's exp differ in the last bit for some arguments, so updated weights agree to
    # 2 ulp rather than bitwise (the "ties" case has exp(0) = 1 and is bit-exact); who-copies-whom is exact.
    assert np.array_equal(src, g[f"{case}_src"])
    assert np.allclose(w, g[f"{case}_w"], rtol=4.5e-16, atol=0)
    if case == "ties":
        assert np.array_equal(w, g[f"{case}_w"])
    assert nb == int(g[f"{case}_stats"][0]) and np.allclose([mx, mn], g[f"{case}_stats"][1:], rtol=4.5e-16)


def test_branch_continuous_upper_threshold_multiset(K):
    """np.argpartition leaves the order of the killed set unspecified: parity as a multiset (SURVEY hard part 4)."""
    g = golden("branch_continuous_golden.npz")
    w, src, nb, mx, mn = K.branch_continuous(g["both_w0"], g["both_v"], float(g["both_vref"]), 5.0, float(g["both_lower"]),
                                             float(g["both_upper"]))
    assert np.allclose(np.sort(w), np.sort(g["both_w"]), rtol=4.5e-16, atol=0)
    assert np.array_equal(np.sort(src), np.sort(g["both_src"]))
    assert nb == int(g["both_stats"][0]) and np.allclose([mx, mn], g["both_stats"][1:], rtol=4.5e-16)
    assert abs(w.sum() - g["both_w"].sum()) < 1e-9


@pytest.mark.parametrize("n,spread", [(1, 0.1), (33, 0.5), (1000, 0.5), (4000, 2.5), (60000, 0.4), (300000, 0.6)])
def test_branch_continuous_vs_oracle(K, oracle, n, spread):
    rng = np.random.default_rng(n)
    w0 = np.exp(rng.normal(0, spread, size=n))
    if n > 1:
        w0[rng.choice(n, max(1, n // 80), replace=False)] = 1e-7
    v = 0.021 + 0.004 * rng.standard_normal(n)
    lower = 1.0 / n if n > 1 else 0.5
    wo, so, nbo, mxo, mno = oracle.branch_continuous(w0, v, 0.021, 5.0, lower, None)
    w, s, nb, mx, mn = K.branch_continuous(w0, v, 0.021, 5.0, lower, None)
    assert nb == nbo
    assert np.array_equal(s, so)                             # donors / copies exactly as the sequential reference loop
    assert np.allclose(w, wo, rtol=4.5e-16, atol=0)          # weights to 2 ulp (exp implementations differ in the last bit)
    assert np.allclose([mx, mn], [mxo, mno], rtol=4.5e-16)


@pytest.mark.parametrize("kind", ["one_bin_overfull", "narrow_band", "many_halvings"])
def test_branch_continuous_sort_paths(K, oracle, kind):
    """The three donor-ordering paths against the sequential reference loop: (a) more than 8192 candidates in one
    histogram bin (identical weights) -> single-CTA fallback sort; (b) thousands of candidates inside one 1/64-octave bin ->
    sub-bucket ranking; (c) K-th largest weight below half the largest -> halved pieces re-enter the top: exact replay."""
    rng = np.random.default_rng(7)
    n = 30000
    if kind == "one_bin_overfull":
        w0 = np.ones(n)
        v = np.full(n, 0.021)                                 # exp(0) = 1: weights stay identical -> ties broken by index
        kill = rng.choice(n, 700, replace=False)
    elif kind == "narrow_band":
        w0 = 1.0 + 0.004 * rng.random(n)                      # all inside one histogram bin (1/64 octave = 1.1 %)
        v = np.full(n, 0.021)
        kill = rng.choice(n, 5000, replace=False)
    else:
        w0 = np.exp(rng.normal(0, 0.2, size=n))
        w0[rng.choice(n, 5, replace=False)] = 64.0            # a few very heavy walkers must donate repeatedly
        v = np.full(n, 0.021)
        kill = rng.choice(n, 400, replace=False)
    w0[kill] = 1e-9
    wo, so, nbo, mxo, mno = oracle.branch_continuous(w0, v, 0.021, 5.0, 1.0 / n, None)
    w, s, nb, mx, mn = K.branch_continuous(w0, v, 0.021, 5.0, 1.0 / n, None)
    assert nb == nbo == len(kill)
    assert np.array_equal(s, so)
    assert np.array_equal(w, wo)                              # exp(0) is exact on both sides: bitwise
    assert (mx, mn) == (mxo, mno)


def test_replay_h2o_continuous_trajectory(K):
    from pyvibdmc_b200 import _capi
    g = golden("traj_h2o_cont_low_golden.npz")
    rp = Replay(g)
    sim = K.DeviceSim(3, 3, g["masses"], 256, 5.0, _capi.POT_H2O_PS, weighting="continuous", thresh_lower=0.3)
    sim.upload(np.repeat(EQ[None] * 1.01, 256, 0))
    T = 30
    vref, pop, wfns = np.zeros(T), np.zeros(T), {}
    for t in range(T):
        if t in (5, 15, 25):
            sim.dw_begin()
        sim.step_injected(rp.normal(256, 3, 3))
        st = sim.stats(t, 1)
        vref[t], pop[t] = st["vref"][0], st["pop"][0]
        if t + 1 in (9, 19, 29):
            parent, pw = sim.dw_parent()
            wfns[t + 1 - 4] = (parent, pw, sim.dw_end(256))
    assert np.allclose(vref, g["vref"], rtol=1e-10, atol=0) and np.allclose(pop, g["pop"], rtol=1e-12)
    out = sim.download()
    assert np.array_equal(out["coords"], g["final_coords"])
    assert np.allclose(out["wts"], g["final_wts"], rtol=1e-10)
    for t in (5, 15, 25):
        assert np.array_equal(wfns[t][0], g[f"wfn{t}_coords"])
        assert np.allclose(wfns[t][1], g[f"wfn{t}_parent_wts"], rtol=1e-10)
        assert np.allclose(wfns[t][2], g[f"wfn{t}_desc_wts"], rtol=1e-9, atol=1e-12)
    sim.close()


def test_continuous_free_running_properties(K, oracle):
    from pyvibdmc_b200 import _capi
    m = np.array([oracle.mass('H'), oracle.mass('H'), oracle.mass('O')])
    n0, T = 20000, 400
    sim = K.DeviceSim(3, 3, m, n0, 5.0, _capi.POT_H2O_PS, weighting="continuous", seed=11)
    sim.upload(EQ[None] * 1.01 + np.zeros((n0, 1, 1)))
    sim.run(T)
    st, out, stats = sim.state(), sim.download(), sim.stats(0, T)
    assert st["step"] == T and st["n"] == n0 and len(out["wts"]) == n0
    assert out["wts"].min() >= 1.0 / n0 * 0.5 and np.isfinite(out["wts"]).all()
    assert abs(stats["pop"][-1] - out["wts"].sum()) < 1e-8 * n0
    pred = np.average(out["pots"], weights=out["wts"]) - 0.1 * (out["wts"].sum() - n0) / n0
    assert abs(pred - st["vref"]) < 1e-11 * abs(pred)
    assert np.max(np.abs(out["pots"] - oracle.water_pot(out["coords"])) / np.maximum(np.abs(out["pots"]), WN)) < 1e-10
    zpe = stats["vref"][T // 2:].mean() / WN
    assert 4300 < zpe < 4900, zpe
    sim.close()


# ------------------------------------------------------------------ importance sampling
def test_water_trial_drift_vs_reference(K):
    from pyvibdmc_b200 import _capi
    g = golden("impsamp_water_golden.npz")
    f, psi, sec = K.trial_drift(_capi.TRIAL_H2O_FD, g["coords"], water_table())
    assert np.allclose(psi, g["psi"], rtol=1e-13, atol=1e-300)
    # finite differences amplify last-ulp differences of psi by psi/dx and psi/dx^2
    assert np.allclose(f, g["f_x"], rtol=1e-8, atol=1e-9 * np.abs(g["f_x"]).max())
    ok = np.isfinite(g["sec"])
    assert np.allclose(sec[ok], g["sec"][ok], rtol=1e-5, atol=1e-6 * np.nanmax(np.abs(g["sec"][20:])))


def test_metropolis_and_local_kin_vs_reference(K):
    g = golden("impsamp_water_golden.npz")
    m = g["masses"]
    sig = np.sqrt(float(g["dt"]) / m)
    acc = K.metropolis(g["coords"], g["y"], g["f_x"], g["f_y"], g["psi"], g["psi_y"], sig, 1 / m, float(g["dt"]))
    ok = np.isfinite(g["acc"])
    assert np.allclose(acc[ok], g["acc"][ok], rtol=1e-12, atol=1e-300)
    assert np.array_equal(K.local_kin(g["sec"][20:], 1 / m), g["local_kin"][20:])


def test_harm_trial_analytic_vs_reference(K):
    from pyvibdmc_b200 import _capi
    g = golden("ho_golden.npz")
    alpha = float(g["mass"]) * float(g["omega"])
    d1, psi, d2 = K.trial_drift(_capi.TRIAL_HARM1D, g["x"], np.array([alpha]))
    assert np.allclose(psi, g["psi"], rtol=1e-15) and np.allclose(d1, g["dpsi"], rtol=1e-14) and np.allclose(d2, g["d2psi"], rtol=1e-13)


def test_replay_h2o_impsamp_trajectory(K):
    from pyvibdmc_b200 import _capi
    g = golden("traj_h2o_imp_golden.npz")
    rp = Replay(g)
    sim = K.DeviceSim(3, 3, g["masses"], 200, 1.0, _capi.POT_H2O_PS, trial=_capi.TRIAL_H2O_FD)
    sim.set_trial_table(water_table())
    sim.upload(np.repeat(EQ[None] * 1.01, 200, 0))
    T, n = 16, 200
    vref, pop, dts = np.zeros(T), np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 3, 3)
        um = rp.take(n)
        ub = rp.take(n)
        sim.step_injected(disp, ub, um)
        st = sim.stats(t, 1)
        vref[t], pop[t], dts[t] = st["vref"][0], st["pop"][0], st["dt_eff"][0]
        n = int(pop[t])
    assert np.array_equal(pop, g["pop"])
    assert np.allclose(vref, g["vref"], rtol=1e-8)
    assert np.allclose(np.cumsum(dts), g["eff_ts"], rtol=1e-13)
    sim.close()


def test_replay_ho_impsamp_analytic_trajectory(K):
    from pyvibdmc_b200 import _capi
    g = golden("traj_ho_imp_golden.npz")
    rp = Replay(g)
    m, om = float(g["masses"][0]), 3700.0 * WN
    sim = K.DeviceSim(1, 1, g["masses"], 300, 5.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2], trial=_capi.TRIAL_HARM1D)
    sim.set_trial_table(np.array([m * om]))
    sim.upload(np.zeros((300, 1, 1)))
    T, n = 40, 300
    vref, pop = np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 1, 1)
        um, ub = rp.take(n), rp.take(n)
        sim.step_injected(disp, ub, um)
        st = sim.stats(t, 1)
        vref[t], pop[t] = st["vref"][0], st["pop"][0]
        n = int(pop[t])
    assert np.array_equal(pop, g["pop"]) and np.allclose(vref, g["vref"], rtol=1e-11)
    sim.close()


def test_replay_second_impsamp_displacement_trajectories(K):
    """imp_move_randomly_second_type (pyvibdmc.py:614-649) replayed with the reference's recorded draws."""
    from pyvibdmc_b200 import _capi
    g = golden("traj_h2o_imp2_golden.npz")
    rp = Replay(g)
    sim = K.DeviceSim(3, 3, g["masses"], 200, 1.0, _capi.POT_H2O_PS, trial=_capi.TRIAL_H2O_FD, imp_variant=_capi.IMP_SECOND_DISPLACEMENT)
    sim.set_trial_table(water_table())
    sim.upload(np.repeat(EQ[None] * 1.01, 200, 0))
    T, n = 12, 200
    vref, pop, dts = np.zeros(T), np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 3, 3)
        um, ub = rp.take(n), rp.take(n)
        sim.step_injected(disp, ub, um)
        st = sim.stats(t, 1)
        vref[t], pop[t], dts[t] = st["vref"][0], st["pop"][0], st["dt_eff"][0]
        n = int(pop[t])
    assert np.array_equal(pop, g["pop"])
    assert np.allclose(vref, g["vref"], rtol=1e-8) and np.allclose(np.cumsum(dts), g["eff_ts"], rtol=1e-13)
    sim.close()
    g = golden("traj_ho_imp2_golden.npz")
    rp = Replay(g)
    m, om = float(g["masses"][0]), 3700.0 * WN
    sim = K.DeviceSim(1, 1, g["masses"], 300, 5.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2], trial=_capi.TRIAL_HARM1D,
                      imp_variant=_capi.IMP_SECOND_DISPLACEMENT)
    sim.set_trial_table(np.array([m * om]))
    sim.upload(np.zeros((300, 1, 1)))
    T, n = 40, 300
    vref, pop = np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 1, 1)
        um, ub = rp.take(n), rp.take(n)
        sim.step_injected(disp, ub, um)
        st = sim.stats(t, 1)
        vref[t], pop[t] = st["vref"][0], st["pop"][0]
        n = int(pop[t])
    assert np.array_equal(pop, g["pop"]) and np.allclose(vref, g["vref"], rtol=1e-11)
    sim.close()


def test_replay_excited_state_impsamp_trajectory(K):
    """excited_state_imp_samp (pyvibdmc.py:562-591, 608-611, 810-811) replayed with the reference's draws."""
    from pyvibdmc_b200 import _capi
    g = golden("traj_h2o_imp_exc_golden.npz")
    rp = Replay(g)
    sim = K.DeviceSim(3, 3, g["masses"], 200, 1.0, _capi.POT_H2O_PS, trial=_capi.TRIAL_H2O_FD, imp_variant=_capi.IMP_EXCITED_STATE)
    sim.set_trial_table(water_table())
    sim.upload(np.repeat(EQ[None] * 1.01, 200, 0))
    T, n = 12, 200
    vref, pop, dts = np.zeros(T), np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 3, 3)
        um, ub = rp.take(n), rp.take(n)
        sim.step_injected(disp, ub, um)
        st = sim.stats(t, 1)
        vref[t], pop[t], dts[t] = st["vref"][0], st["pop"][0], st["dt_eff"][0]
        n = int(pop[t])
    assert np.array_equal(pop, g["pop"])
    assert np.allclose(vref, g["vref"], rtol=1e-8) and np.allclose(np.cumsum(dts), g["eff_ts"], rtol=1e-13)
    sim.close()


def water_table_analytic():
    import importlib.util
    spec = importlib.util.spec_from_file_location("call_trl_h2o_b200", os.path.join(H2O_DIR, "call_trl_h2o.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.packed_table_analytic()


def test_water_trial_analytic_derivatives_vs_reference(K):
    """a13: dpsi_dx (call_trl_h2o.py:101-149) with ChainRuleHelper (imp_samp_helper.py:10-209) on the device."""
    from pyvibdmc_b200 import _capi
    g = golden("impsamp_water_analytic_golden.npz")
    f, psi, sec = K.trial_drift(_capi.TRIAL_H2O_AN, g["coords"], water_table_analytic())
    assert np.allclose(psi, g["psi"], rtol=1e-13, atol=1e-300)
    assert np.allclose(f, g["f_x"], rtol=1e-9, atol=1e-11)
    assert np.allclose(sec, g["sec"], rtol=1e-8, atol=1e-8)
    # trajectory of the reference driven by these derivatives, replayed with its recorded draws
    gt = golden("traj_h2o_imp_an_golden.npz")
    rp = Replay(gt)
    sim = K.DeviceSim(3, 3, gt["masses"], 200, 1.0, _capi.POT_H2O_PS, trial=_capi.TRIAL_H2O_AN)
    sim.set_trial_table(water_table_analytic())
    sim.upload(np.repeat(EQ[None] * 1.01, 200, 0))
    T, n = 12, 200
    vref, pop, dts = np.zeros(T), np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 3, 3)
        um, ub = rp.take(n), rp.take(n)
        sim.step_injected(disp, ub, um)
        st = sim.stats(t, 1)
        vref[t], pop[t], dts[t] = st["vref"][0], st["pop"][0], st["dt_eff"][0]
        n = int(pop[t])
    assert np.array_equal(pop, gt["pop"])
    assert np.allclose(vref, gt["vref"], rtol=1e-9) and np.allclose(np.cumsum(dts), gt["eff_ts"], rtol=1e-13)
    sim.close()


def test_impsamp_free_running_zpe(K, oracle):
    """Importance-sampled water: local energy fluctuates far less than V; ZPE stays near 4634 cm-1."""
    from pyvibdmc_b200 import _capi
    m = np.array([oracle.mass('H'), oracle.mass('H'), oracle.mass('O')])
    sim = K.DeviceSim(3, 3, m, 4000, 1.0, _capi.POT_H2O_PS, trial=_capi.TRIAL_H2O_FD, seed=5)
    sim.set_trial_table(water_table())
    sim.upload(EQ[None] * 1.01 + np.zeros((4000, 1, 1)))
    sim.run(600)
    st = sim.stats(0, 600)
    assert sim.state()["step"] == 600
    assert (st["rejected"] >= 0).all() and st["rejected"].mean() < 0.2 * 4000
    zpe = st["vref"][300:].mean() / WN
    assert 4400 < zpe < 4900, zpe
    sim.close()


# ------------------------------------------------------------------ NN potential
def packed_nn():
    return np.load(os.path.join(ROOT, "pyvibdmc_b200", "sample_potentials", "TensorflowPots", "sample_h4o2_nn_packed.npy"))


def unpack_nn(p):
    out, off = [], 0
    for k, n in ((15, 120), (120, 120), (120, 120), (120, 1)):
        w = p[off:off + k * n].reshape(k, n); off += k * n
        b = p[off:off + n]; off += n
        out.append((w, b))
    return out


def test_coulomb_descriptor(K):
    g = golden("descriptor_golden.npz")
    assert np.allclose(K.coulomb_descriptor(g["coords"], g["zs"]), g["coulomb"], rtol=1e-14)


@pytest.mark.parametrize("path", ["tcgen05", "cuda_cores"])
def test_nn_h4o2_vs_float32_oracle(K, oracle, path, monkeypatch):
    """Both MLP kernels (tcgen05 with two-piece fp16 splitting; float32 FMA on CUDA cores) against the float32 oracle."""
    K.nn_config(path="cuda_cores" if path == "cuda_cores" else "tcgen05")
    g = golden("descriptor_golden.npz")
    p = packed_nn()
    K.nn_h4o2_set_weights(p)
    v = K.nn_h4o2(g["coords"])
    ref = oracle.nn_forward_f32(g["coulomb"], unpack_nn(p))
    # float32 network: summation order differs between GPU and NumPy; the fp16 two-piece split leaves ~2^-22 per product and
    # the fast swish (ex2.approx / rcp.approx) a few float32 ulp per activation: measured 2.5e-6 of the largest energy
    assert np.allclose(v, ref, rtol=1e-5, atol=5e-6 * np.abs(ref).max()), np.abs(v - ref).max() / np.abs(ref).max()
    assert (v >= 0).all() and v.dtype == np.float64
    assert abs(v[0] / WN - 7.62) < 0.05                         # equilibrium water dimer (reference tests/test_analysis.py:77-82)
    # ragged sizes and a physical sanity value: ~7.6 cm-1 at the water-dimer minimum
    dimer = np.array([[1.513632, -0.005249, -0.121857], [0.560102, 0.002812, 0.048059], [1.913196, 0.033035, 0.750687],
                      [-1.385643, 0.004325, 0.110302], [-1.750594, 0.746224, -0.382028], [-1.746613, -0.774680, -0.324277]])
    for n in (1, 63, 64, 65, 127, 128, 129, 1000):
        vv = K.nn_h4o2(g["coords"][:n] if n <= len(g["coords"]) else np.tile(g["coords"], (2, 1, 1))[:n])
        assert np.allclose(vv[:min(n, 512)], v[:min(n, 512)], rtol=1e-6, atol=1e-9)
    K.nn_config(path="tcgen05")


def test_nn_h4o2_pinned_against_float64_golden(K, oracle):
    """SURVEY 8 a6 pin: tests/golden/nn_h4o2_f64_golden.npz holds the shipped network (weights identical to the reference's
    sample_h4o2_nn.h5, checked by make_nn_golden.py) evaluated in float64 -- the value every float32 forward pass, TensorFlow's
    included, approximates -- and a float32 NumPy evaluation.  Error budgets relative to the largest energy of the set:
    float32 NumPy 6.9e-7 (recorded in the fixture); float32 FMA kernel: the same class; tcgen05 kernel (two-piece fp16 split,
    22 mantissa bits per factor): 2.5e-6 measured, asserted below 4e-6."""
    g = golden("nn_h4o2_f64_golden.npz")
    K.nn_h4o2_set_weights(packed_nn())
    scale = np.abs(g["e64"]).max()
    f32_err = np.abs(g["e32"] - g["e64"]).max() / scale
    assert f32_err < 1.0e-6
    errs = {}
    for path in ("tcgen05", "cuda_cores", "tcgen05_one_tile"):
        K.nn_config(path=path)
        v = K.nn_h4o2(g["coords"])
        errs[path] = np.abs(v - g["e64"]).max() / scale
    K.nn_config(path="tcgen05")
    assert errs["cuda_cores"] < 2.0e-6, errs                  # float32 arithmetic, float32-class error
    assert errs["tcgen05"] < 4.0e-6 and errs["tcgen05_one_tile"] < 4.0e-6, errs
    assert abs(K.nn_h4o2(g["coords"][:1])[0] / WN - 7.62) < 0.05
