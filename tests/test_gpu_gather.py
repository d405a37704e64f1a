"""GPU parity tests (-m gpu) of the deferred-compaction step (csrc/pvd_gather.cuh), the kernel pvd_sim_run uses for large
ensembles.  (1) Against the oracle at the size the benchmark runs at: injected displacements and uniforms, 1e6 walkers,
populations / np.repeat order bit-exact and Vref to 1e-12 (reference: DMC_Sim.birth_or_death pyvibdmc.py:380-431,
calc_vref :651-661).  (2) Against the self-compacting step kernel k_step_discrete (itself tied to the reference's
trajectories by test_gpu_sim.py): same Philox addressing, same arithmetic, same order -> identical bits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
WN = 4.556335281212229e-6
AMU = 1.0 / 6.02213670000e23 / 9.10938970000e-28
M_H2O = np.array([1.00782503, 1.00782503, 15.99491462]) * AMU
STAT_KEYS = ("vref", "pop", "v_avg", "v_max", "v_min", "births", "deaths", "step")
GATHER, PER_STEP = 3, 0           # pvd_sim_set_resident modes


@pytest.fixture(scope="module")
def K():
    from pyvibdmc_b200 import kernels
    assert kernels.device_count() > 0
    return kernels


def _h2o(K, n0, seed, mode, start=None, **kw):
    from pyvibdmc_b200 import _capi
    sim = K.DeviceSim(3, 3, M_H2O, n0, 5.0, _capi.POT_H2O_PS, seed=seed, **kw)
    sim.set_resident(mode)
    sim.upload(np.repeat(EQ[None] * 1.01, n0, axis=0) if start is None else start)
    return sim


def _same(a, b, T):
    sa, sb = a.state(), b.state()
    assert sa == sb, (sa, sb)
    ta, tb = a.stats(0, T), b.stats(0, T)
    for k in STAT_KEYS:
        assert np.array_equal(ta[k], tb[k]), k
    da, db = a.download(), b.download()
    assert np.array_equal(da["coords"], db["coords"]) and np.array_equal(da["pots"], db["pots"])


@pytest.mark.parametrize("n0,steps", [(1_000_000, 3), (33_000, 5), (1000, 6)])
def test_gather_step_vs_oracle_injected(K, oracle, n0, steps):
    """Injected draws at the benchmark's size: the device ensemble after every step is the oracle's, bit for bit."""
    rng = np.random.default_rng(2026)
    sig = np.sqrt(5.0 / M_H2O)
    start = EQ[None] * 1.01 + rng.normal(0, 0.03, size=(n0, 3, 3))
    sim = _h2o(K, n0, 1, GATHER, start=start)
    coords, pots = start.copy(), oracle.water_pot(start)
    vref = oracle.calc_vref(pots, n0, 0.1)
    assert abs(sim.state()["vref"] - vref) < 1e-12 * abs(vref)
    for t in range(steps):
        n = len(coords)
        disp = rng.normal(0, 1, size=(n, 3, 3)) * sig[None, :, None]
        u = rng.random(n)
        sim.step_injected(disp, u)
        coords = coords + disp
        pots = oracle.water_pot(coords)
        cnt, idx, _, _, _ = oracle.birth_or_death_discrete(pots, vref, 5.0, u, n0)
        coords, pots = coords[idx], pots[idx]
        vref = oracle.calc_vref(pots, n0, 0.1)
        st = sim.state()
        assert st["n"] == len(coords), (t, st["n"], len(coords))
        assert abs(st["vref"] - vref) < 1e-12 * abs(vref), (t, st["vref"], vref)
        out = sim.download()
        assert np.array_equal(out["coords"], coords), t               # np.repeat order, every copy in its place
        assert np.max(np.abs(out["pots"] - pots) / np.maximum(np.abs(pots), WN)) < 1e-10
    sim.close()


@pytest.mark.parametrize("n0,T", [(1000, 40), (20000, 60), (300000, 25), (1_000_000, 12)])
def test_gather_equals_per_step_h2o(K, n0, T):
    a, b = _h2o(K, n0, 11, GATHER), _h2o(K, n0, 11, PER_STEP)
    a.run(T)
    b.run(T)
    _same(a, b, T)
    a.close(); b.close()


def test_gather_segments_and_mixing(K):
    """Segments of any length (each ends with a materialisation; odd and even lengths end in different buffers), mixed
    with self-compacting steps and the resident kernel."""
    n0 = 50000
    a, b = _h2o(K, n0, 5, GATHER), _h2o(K, n0, 5, PER_STEP)
    for seg in (1, 7, 2, 1, 13):
        a.run(seg)
    a.set_resident(PER_STEP)
    a.run(3)
    a.set_resident(2)
    a.run(4)
    a.set_resident(GATHER)
    a.run(5)
    b.run(36)
    _same(a, b, 36)
    a.close(); b.close()


def test_gather_ragged_population(K):
    """Populations that are no multiple of the tile, far from N0 (many empty / double tiles: wide windows in the pull)."""
    rng = np.random.default_rng(3)
    n0 = 40001
    start = EQ[None] * 1.01 + rng.normal(0, 0.12, size=(n0, 3, 3))     # hot start: ~20 % of the walkers die in the first steps
    a, b = _h2o(K, n0, 8, GATHER, start=start), _h2o(K, n0, 8, PER_STEP, start=start)
    a.run(9)
    b.run(9)
    _same(a, b, 9)
    st = a.stats(0, 9)
    assert st["deaths"][:3].sum() > 0.05 * n0
    a.close(); b.close()


@pytest.mark.parametrize("rng", ["fp64", "fast"])
def test_gather_other_rng_modes(K, rng):
    from pyvibdmc_b200 import _capi
    a = _h2o(K, 5000, 3, GATHER, rng_mode=_capi.RNG_MODES[rng])
    b = _h2o(K, 5000, 3, PER_STEP, rng_mode=_capi.RNG_MODES[rng])
    a.run(30)
    b.run(30)
    _same(a, b, 30)
    a.close(); b.close()


def test_gather_harmonic_and_morse(K, oracle):
    from pyvibdmc_b200 import _capi
    m, om = oracle.reduced_mass('O-H'), 3700.0 * WN
    for pot, params, nc in ((_capi.POT_HARMONIC, [(0.5 * m) * om ** 2], 1), (_capi.POT_HARMONIC, [(0.5 * m) * om ** 2] * 3, 3),
                            (_capi.POT_MORSE1D, [0.18, 1.2], 1)):
        sims = []
        for mode in (GATHER, PER_STEP):
            s = K.DeviceSim(1, nc, [m], 3000, 10.0 if pot == _capi.POT_HARMONIC else 5.0, pot, pot_params=params, seed=9)
            s.set_resident(mode)
            s.upload(np.zeros((3000, 1, nc)) + (0.1 if pot == _capi.POT_MORSE1D else 0.0))
            s.run(200)
            sims.append(s)
        _same(sims[0], sims[1], 200)
        for s in sims:
            s.close()


def test_gather_descendant_weighting(K):
    """who_from is pulled with the walkers."""
    n0 = 20000
    outs = []
    for mode in (GATHER, PER_STEP):
        s = _h2o(K, n0, 21, mode)
        s.run(30)
        npar = s.state()["n"]
        s.dw_begin()
        s.run(25)
        outs.append((s.dw_end(npar), s.download(who_from=True)["who_from"], s.state()))
        s.run(10)
        outs[-1] += (s.state(),)
        s.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2] and outs[0][3] == outs[1][3]
    assert outs[0][0].sum() == outs[0][2]["n"]


def test_gather_branch_every(K):
    a, b = _h2o(K, 5000, 2, GATHER), _h2o(K, 5000, 2, PER_STEP)
    a.run(31, branch_every=3)
    b.run(31, branch_every=3)
    _same(a, b, 31)
    a.close(); b.close()


@pytest.mark.parametrize("good_steps", [0, 3])
def test_gather_population_guard(K, good_steps):
    """A step that fails inside a segment: the ensemble the host sees is the failing step's input (the reference raises
    before it touches the walker arrays, pyvibdmc.py:397-413), compacted, whichever buffer it was in."""
    from pyvibdmc_b200 import _capi
    res = []
    for mode in (GATHER, PER_STEP):
        sim = K.DeviceSim(3, 3, M_H2O, 1000, 5.0, _capi.POT_H2O_PS, seed=1)
        sim.set_resident(mode)
        sim.upload(np.repeat(EQ[None] * 1.01, 1000, axis=0))
        if good_steps:
            sim.run(good_steps)
        sim.set_masses(M_H2O * 1e-4)                # sigma x 100: the next move throws the walkers far up the wall
        sim.run(6)
        with pytest.raises(_capi.MassiveEvent, match="Massive walker birth or death event!!!!!!! Dying..."):
            sim.state()
        st = sim.state(raise_on_error=False)
        sim.run(3)                                  # a dead run stays dead, whatever is enqueued afterwards
        with pytest.raises(_capi.MassiveEvent):
            sim.state()
        res.append((st["step"], st["n"], sim.download()["coords"]))
        sim.close()
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2])


def test_gather_long_run_zpe(K):
    s = _h2o(K, 100000, 77, GATHER)
    s.run(1500)
    st = s.stats(0, 1500)
    assert s.state()["step"] == 1500
    assert np.array_equal(st["pop"][1:] - st["pop"][:-1], (st["births"] - st["deaths"])[1:])
    zpe = st["vref"][400:].mean() / WN
    assert abs(zpe - 4636) < 15, zpe
    s.close()
