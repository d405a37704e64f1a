"""GPU parity tests (-m gpu) of the device-resident loop (pvd_sim_*): reference trajectories are
replayed with the recorded random numbers injected, and free-running runs are checked through
size-independent properties."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
WN = 4.556335281212229e-6


@pytest.fixture(scope="module")
def K():
    from pyvibdmc_b200 import kernels
    assert kernels.device_count() > 0
    return kernels


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


def run_discrete_replay(K, g, sim, n0, T, a, d, equil, wfn_every, desc):
    rp = Replay(g)
    vref, pop, wfns = np.zeros(T), np.zeros(T), {}
    wfn_steps = set(range(equil, T + wfn_every, wfn_every))
    save_steps = {s + desc for s in wfn_steps}
    n = n0
    for t in range(T):
        if t in wfn_steps:
            sim.dw_begin()
            n_parent = n
        sim.step_injected(rp.normal(n, a, d), rp.take(n))
        st = sim.state()
        n = st["n"]
        vref[t], pop[t] = st["vref"], n
        if t + 1 in save_steps:
            wfns[t + 1 - desc] = (sim.dw_parent()[0], sim.dw_end(n_parent))
    return vref, pop, wfns


def test_replay_ho_discrete_trajectory(K):
    from pyvibdmc_b200 import _capi
    g = golden("traj_ho_disc_golden.npz")
    m, om = float(g["masses"][0]), 3700.0 * WN
    sim = K.DeviceSim(1, 1, g["masses"], 300, 10.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2])
    sim.upload(np.zeros((300, 1, 1)))
    vref, pop, wfns = run_discrete_replay(K, g, sim, 300, 60, 1, 1, 5, 10, 4)
    assert np.array_equal(pop, g["pop"])
    assert np.allclose(vref, g["vref"], rtol=1e-13, atol=0)
    out = sim.download()
    assert np.array_equal(out["coords"], g["final_coords"]) and np.array_equal(out["pots"], g["final_pots"])
    for t in (5, 15, 25, 35, 45, 55):
        assert np.array_equal(wfns[t][0], g[f"wfn{t}_coords"]) and np.array_equal(wfns[t][1], g[f"wfn{t}_desc_wts"])
    st = sim.stats(0, 60)
    assert np.array_equal(st["pop"], g["pop"]) and np.array_equal(st["step"], np.arange(60))
    sim.close()


def test_replay_h2o_discrete_trajectory(K):
    from pyvibdmc_b200 import _capi
    g = golden("traj_h2o_disc_golden.npz")
    sim = K.DeviceSim(3, 3, g["masses"], 256, 5.0, _capi.POT_H2O_PS)
    sim.upload(np.repeat(EQ[None] * 1.01, 256, 0))
    vref, pop, wfns = run_discrete_replay(K, g, sim, 256, 30, 3, 3, 5, 10, 4)
    assert np.array_equal(pop, g["pop"])
    assert np.allclose(vref, g["vref"], rtol=1e-11, atol=0)
    out = sim.download()
    assert np.array_equal(out["coords"], g["final_coords"])
    assert np.allclose(out["pots"], g["final_pots"], rtol=1e-10, atol=1e-16)
    for t in (5, 15, 25):
        assert np.array_equal(wfns[t][0], g[f"wfn{t}_coords"]) and np.array_equal(wfns[t][1], g[f"wfn{t}_desc_wts"])
    sim.close()


@pytest.mark.parametrize("rng_mode", [0, 2])
def test_external_potential_path_matches_fused(K, oracle, rng_mode):
    """Stepping with a host-side potential (plug-in path) == fused built-in path, same seed (Box-Muller and
    ziggurat: the fused kernel settles a walker's rare non-common-path normals after the common ones, the
    generic kernel on the spot -- same numbers)."""
    from pyvibdmc_b200 import _capi
    m = np.array([oracle.mass('H'), oracle.mass('H'), oracle.mass('O')])
    start = np.repeat(EQ[None] * 1.01, 2000, 0)
    a = K.DeviceSim(3, 3, m, 2000, 5.0, _capi.POT_H2O_PS, seed=42, rng_mode=rng_mode)
    a.upload(start)
    a.run(25)
    b = K.DeviceSim(3, 3, m, 2000, 5.0, _capi.POT_EXTERNAL, seed=42, rng_mode=rng_mode)
    b.upload(start)
    b.set_pots(K.pes_h2o(start))
    for _ in range(25):
        cds = b.ext_move()
        b.ext_finish(K.pes_h2o(cds))
    sa, sb = a.state(), b.state()
    assert sa["n"] == sb["n"] and sa["step"] == sb["step"] == 25
    assert abs(sa["vref"] - sb["vref"]) <= 1e-12 * abs(sa["vref"])
    assert np.array_equal(a.download()["coords"], b.download()["coords"])
    a.close(); b.close()


def test_free_running_h2o_properties(K, oracle):
    """20,000 walkers x 300 steps: population control, V consistency, Vref identity, determinism."""
    from pyvibdmc_b200 import _capi
    m = np.array([oracle.mass('H'), oracle.mass('H'), oracle.mass('O')])
    n0, T = 20000, 300
    runs = []
    for rep in range(2):
        sim = K.DeviceSim(3, 3, m, n0, 5.0, _capi.POT_H2O_PS, seed=7)
        sim.upload(EQ[None] * 1.01 + np.zeros((n0, 1, 1)))
        sim.run(T)
        st = sim.state()
        out = sim.download()
        stats = sim.stats(0, T)
        runs.append((st, out, stats))
        sim.close()
    st, out, stats = runs[0]
    assert st["step"] == T and 0.5 * n0 < st["n"] < 1.5 * n0 and len(out["coords"]) == st["n"]
    # carried energies are the PES of the carried coordinates
    assert np.max(np.abs(out["pots"] - oracle.water_pot(out["coords"])) / np.maximum(np.abs(out["pots"]), WN)) < 1e-10
    # Vref identity of the last step: mean(V) - alpha (N - N0)/N0
    pred = out["pots"].mean() - 0.1 * (st["n"] - n0) / n0
    assert abs(pred - st["vref"]) < 1e-12 * abs(pred)
    assert np.array_equal(stats["pop"][1:] - stats["pop"][:-1], (stats["births"] - stats["deaths"])[1:])
    assert stats["pop"][-1] == st["n"]
    zpe = stats["vref"][T // 2:].mean() / WN
    assert 4300 < zpe < 4900          # equilibrating towards 4634 cm-1
    # bit-reproducible for a fixed seed
    assert np.array_equal(runs[0][1]["coords"], runs[1][1]["coords"]) and runs[0][0] == runs[1][0]


def test_population_guard_raises_like_reference(K, oracle):
    from pyvibdmc_b200 import _capi
    m = np.array([oracle.mass('H'), oracle.mass('H'), oracle.mass('O')])
    # 10% absurdly stretched walkers drag Vref up; with dt=200 the others get weights e^(+8) > 1.5 N0 + 1
    sim = K.DeviceSim(3, 3, m, 1000, 200.0, _capi.POT_H2O_PS, seed=1)
    bad = EQ[None] + np.zeros((1000, 1, 1))
    bad[:100] *= 3.0
    sim.upload(bad)
    sim.run(3)
    with pytest.raises(_capi.MassiveEvent, match="Massive walker birth or death event!!!!!!! Dying..."):
        sim.state()
    assert sim.state(raise_on_error=False)["step"] == 0      # the failing step did not complete
    sim.close()
    # population collapse: every walker far above Vref's reach dies -> population guard
    sim = K.DeviceSim(3, 3, m, 1000, 5.0, _capi.POT_EXTERNAL, seed=1)
    sim.upload(EQ[None] + np.zeros((1000, 1, 1)))
    sim.set_pots(np.zeros(1000))
    sim.ext_move()
    sim.ext_finish(np.full(1000, 0.5))                       # w = exp(-2.5): ~92% die
    with pytest.raises(_capi.MassiveEvent):
        sim.state()
    sim.close()


def test_ho_zpe_free_running(K, oracle):
    """Config 1 (1-D HO, 1000 walkers, dt=10): ZPE -> 1850 cm-1 (+ small time-step bias)."""
    from pyvibdmc_b200 import _capi
    m, om = oracle.reduced_mass('O-H'), 3700.0 * WN
    sim = K.DeviceSim(1, 1, [m], 1000, 10.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2], seed=3)
    sim.upload(np.zeros((1000, 1, 1)))
    sim.run(5000)
    st = sim.stats(0, 5000)
    zpe = st["vref"][1250:].mean() / WN
    assert abs(zpe - 1852) < 25, zpe
    sim.close()


def test_async_snapshot_matches_download(K):
    """pvd_sim_snapshot_begin/_wait: the ensemble at the moment of _begin, copied on a side stream while later steps run."""
    from pyvibdmc_b200 import _capi
    eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
    amu = 1.0 / 6.02213670000e23 / 9.10938970000e-28
    m = np.array([1.00782503, 1.00782503, 15.99491462]) * amu
    for weighting in ("discrete", "continuous"):
        sim = K.DeviceSim(3, 3, m, 30000, 5.0, _capi.POT_H2O_PS, weighting=weighting, seed=3)
        sim.upload(np.repeat(eq[None] * 1.01, 30000, axis=0))
        sim.run(20)
        sim.dw_begin()
        sim.run(5)
        ref = sim.download(who_from=True)
        vref = sim.state()["vref"]
        sim.snapshot_begin()
        sim.run(40)                                  # the compute stream moves on while the copy is in flight
        snap = sim.snapshot_wait(who_from=True)
        assert sim.state()["step"] == 65
        for k in ("coords", "pots", "who_from") + (("wts",) if weighting == "continuous" else ()):
            assert np.array_equal(snap[k], ref[k]), k
        assert snap["vref"] == vref
        with pytest.raises(Exception):
            sim.snapshot_wait()                      # nothing in flight any more
        sim.close()
