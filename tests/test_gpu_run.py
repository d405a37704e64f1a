"""GPU parity tests (-m gpu) of the resident multi-step kernel (csrc/pvd_run.cuh) behind pvd_sim_run: one launch that
overlaps consecutive time steps must give, bit for bit, what one launch per time step gives (same Philox addressing,
same per-tile arithmetic, exact sums), and both are tied to the oracle by test_gpu_sim.py / test_gpu_bigparity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
WN = 4.556335281212229e-6
AMU = 1.0 / 6.02213670000e23 / 9.10938970000e-28
M_H2O = np.array([1.00782503, 1.00782503, 15.99491462]) * AMU
STAT_KEYS = ("vref", "pop", "v_avg", "v_max", "v_min", "births", "deaths", "step")


@pytest.fixture(scope="module")
def K():
    from pyvibdmc_b200 import kernels
    assert kernels.device_count() > 0
    return kernels


def _h2o(K, n0, seed, resident, **kw):
    from pyvibdmc_b200 import _capi
    sim = K.DeviceSim(3, 3, M_H2O, n0, 5.0, _capi.POT_H2O_PS, seed=seed, **kw)
    sim.set_resident(resident)
    sim.upload(np.repeat(EQ[None] * 1.01, n0, axis=0))
    return sim


def _same(a, b, T):
    sa, sb = a.state(), b.state()
    assert sa == sb, (sa, sb)
    ta, tb = a.stats(0, T), b.stats(0, T)
    for k in STAT_KEYS:
        assert np.array_equal(ta[k], tb[k]), k
    da, db = a.download(), b.download()
    assert np.array_equal(da["coords"], db["coords"]) and np.array_equal(da["pots"], db["pots"])


@pytest.mark.parametrize("n0,T", [(1000, 40), (20000, 60), (300000, 25)])
def test_resident_equals_per_step_h2o(K, n0, T):
    a, b = _h2o(K, n0, 11, True), _h2o(K, n0, 11, False)
    a.run(T)
    b.run(T)
    _same(a, b, T)
    a.close(); b.close()


def test_resident_segments_and_mixing(K):
    """Segments of any length (odd ones flip the ping-pong parity) and per-step launches in between."""
    n0 = 20000
    a, b = _h2o(K, n0, 5, True), _h2o(K, n0, 5, False)
    for seg in (1, 7, 2, 1, 13):
        a.run(seg)
    a.set_resident(False)
    a.run(3)
    a.set_resident(True)
    a.run(9)
    b.run(36)
    _same(a, b, 36)
    a.close(); b.close()


@pytest.mark.parametrize("rng", ["fp64", "fast"])
def test_resident_other_rng_modes(K, rng):
    from pyvibdmc_b200 import _capi
    a = _h2o(K, 5000, 3, True, rng_mode=_capi.RNG_MODES[rng])
    b = _h2o(K, 5000, 3, False, rng_mode=_capi.RNG_MODES[rng])
    a.run(30)
    b.run(30)
    _same(a, b, 30)
    a.close(); b.close()


def test_resident_harmonic_and_morse(K, oracle):
    from pyvibdmc_b200 import _capi
    m, om = oracle.reduced_mass('O-H'), 3700.0 * WN
    for pot, params, nc in ((_capi.POT_HARMONIC, [(0.5 * m) * om ** 2], 1), (_capi.POT_HARMONIC, [(0.5 * m) * om ** 2] * 3, 3),
                            (_capi.POT_MORSE1D, [0.18, 1.2], 1)):
        sims = []
        for resident in (True, False):
            s = K.DeviceSim(1, nc, [m], 1000, 10.0 if pot == _capi.POT_HARMONIC else 5.0, pot, pot_params=params, seed=9)
            s.set_resident(resident)
            s.upload(np.zeros((1000, 1, nc)) + (0.1 if pot == _capi.POT_MORSE1D else 0.0))
            s.run(300)
            sims.append(s)
        _same(sims[0], sims[1], 300)
        for s in sims:
            s.close()


def test_resident_descendant_weighting(K):
    """who_from travels through the resident kernel exactly as through the per-step kernel."""
    n0 = 20000
    outs = []
    for resident in (True, False):
        s = _h2o(K, n0, 21, resident)
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


def test_resident_branch_every(K):
    a, b = _h2o(K, 5000, 2, True), _h2o(K, 5000, 2, False)
    a.run(31, branch_every=3)
    b.run(31, branch_every=3)
    _same(a, b, 31)
    a.close(); b.close()


def test_resident_population_guard(K):
    """A step that fails inside a resident launch stops the launch; the state is the one the per-step path reports."""
    from pyvibdmc_b200 import _capi
    res = []
    for resident in (True, False):
        sim = K.DeviceSim(3, 3, M_H2O, 1000, 200.0, _capi.POT_H2O_PS, seed=1)
        sim.set_resident(resident)
        bad = EQ[None] + np.zeros((1000, 1, 1))
        bad[:100] *= 3.0
        sim.upload(bad)
        sim.run(6)
        with pytest.raises(_capi.MassiveEvent, match="Massive walker birth or death event!!!!!!! Dying..."):
            sim.state()
        st = sim.state(raise_on_error=False)
        sim.run(3)                                  # a dead run stays dead, whatever is enqueued afterwards
        with pytest.raises(_capi.MassiveEvent):
            sim.state()
        res.append((st["step"], st["n"], sim.download()["coords"]))
        sim.close()
    assert res[0][0] == res[1][0] == 0 and res[0][1] == res[1][1] == 1000
    assert np.array_equal(res[0][2], res[1][2])


def test_resident_long_run_zpe(K):
    """20 000 walkers x 3000 steps in one launch: population control holds and the energy equilibrates."""
    s = _h2o(K, 20000, 77, True)
    s.run(3000)
    st = s.stats(0, 3000)
    assert s.state()["step"] == 3000
    assert np.array_equal(st["pop"][1:] - st["pop"][:-1], (st["births"] - st["deaths"])[1:])
    zpe = st["vref"][750:].mean() / WN
    assert abs(zpe - 4636) < 25, zpe
    s.close()
