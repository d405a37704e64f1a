"""GPU parity tests (-m gpu): every stand-alone C-ABI entry point against the oracle and the
golden fixtures generated from the unmodified reference."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
PES_RTOL = 1e-10          # north_star: PES energies to 1e-10 relative


@pytest.fixture(scope="module")
def K():
    from pyvibdmc_b200 import kernels
    assert kernels.device_count() > 0, "GPU tests need a CUDA device"
    return kernels


def rel_err(a, b, floor):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


# ------------------------------------------------------------------ PES
def test_ps_h2o_golden(K):
    g = golden("ps_h2o_golden.npz")
    v = K.pes_h2o(g["coords"])
    # relative to max(|V|, 1 cm^-1): V crosses zero near the minimum (V(eq) = -0.42 cm^-1)
    assert rel_err(v, g["v"], 4.556e-6) <= PES_RTOL


def test_ps_h2o_folded_parameters_bit_exact(K, oracle):
    c, s = K.pes_h2o_params()
    co, so = oracle.ps_folded_params()
    assert np.array_equal(c, co) and np.array_equal(s, so)


def test_ps_h2o_snapshot_identity(K):
    """Shipped-data identity evaluated with the CUDA kernel."""
    g = golden("ps_snapshot_s0_t500.npz")
    cds = g["coords"]
    pred = K.pes_h2o(cds).mean() - 0.1 * (len(cds) - 8000) / 8000
    assert abs(pred - float(g["vref_expected"])) <= 1e-12 * abs(float(g["vref_expected"]))


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1000, 100003])
def test_ps_h2o_vs_oracle_ragged_sizes(K, oracle, n):
    rng = np.random.default_rng(n)
    cds = EQ[None] + rng.normal(0, 0.12, size=(n, 3, 3))
    assert rel_err(K.pes_h2o(cds), oracle.water_pot(cds), 4.556e-6) <= PES_RTOL


def test_ps_h2o_empty_and_invariances(K):
    assert K.pes_h2o(np.zeros((0, 3, 3))).shape == (0,)
    rng = np.random.default_rng(0)
    cds = EQ[None] + rng.normal(0, 0.1, size=(4096, 3, 3))
    v = K.pes_h2o(cds)
    # translation, H<->H exchange and rotation invariance (size-independent properties)
    assert np.allclose(K.pes_h2o(cds + np.array([3.0, -2.0, 0.5])), v, rtol=1e-9, atol=1e-13)
    assert np.allclose(K.pes_h2o(cds[:, [1, 0, 2]]), v, rtol=1e-11, atol=1e-15)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    assert np.allclose(K.pes_h2o(cds @ R.T), v, rtol=1e-9, atol=1e-13)


def test_ps_h2o_full_size_checksum(K, oracle):
    """BASELINE size (2e4 walkers and 1e6): sum over the batch against the oracle's sum."""
    rng = np.random.default_rng(7)
    cds = EQ[None] * 1.01 + rng.normal(0, 0.05, size=(1_000_000, 3, 3))
    v = K.pes_h2o(cds)
    vo = oracle.water_pot(cds[:200_000])
    assert rel_err(v[:200_000], vo, 4.556e-6) <= PES_RTOL
    assert np.isfinite(v).all()


def test_harmonic_bit_exact(K, oracle):
    g = golden("ho_golden.npz")
    k = (0.5 * float(g["mass"])) * float(g["omega"]) ** 2
    assert np.array_equal(K.pes_harmonic(g["pickle_coords"], k), g["pickle_pots"])
    assert np.array_equal(K.pes_harmonic(g["x"], k), g["v_oh"])
    x3 = np.random.default_rng(2).normal(0, 0.2, size=(1000, 1, 3))
    ref = k * x3[:, 0, 0] ** 2 + k * x3[:, 0, 1] ** 2 + k * x3[:, 0, 2] ** 2
    assert np.array_equal(K.pes_harmonic(x3, k), ref)


def test_morse(K):
    g = golden("ho_golden.npz")
    m, om, omx = float(g["mass"]), 3704.5 * 4.556335281212229e-6, 75.3 * 4.556335281212229e-6
    de = om ** 2 / (4 * omx)
    alpha = np.sqrt(m * (om ** 2.) / 2. / de)
    # 1 - exp(-a x) cancels near x = 0: one ulp of exp becomes ~1e-16/|a x| relative
    assert np.allclose(K.pes_morse1d(g["x"], de, alpha), g["v_morse"], rtol=1e-11, atol=1e-18)


# ------------------------------------------------------------------ RNG
def philox_py(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_philox_known_answers(K):
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(philox_py(ctr, key)) == out
        assert tuple(int(x) for x in K.philox(ctr, key)) == out


def test_fp64_normals_match_reference_arithmetic(K):
    """The hand-written -2 ln u / sqrt / sincos of the fp64 Box-Muller against mpmath-free NumPy long-double
    arithmetic on the same Philox bits (counter layout: slot, call<<24, step, purpose<<24)."""
    seed, step, n = 0x123456789ABCDEF, 77, 4096
    z = K.normals(n, 2, seed=seed, step=step, rng_mode=0)
    ld = np.longdouble
    worst = 0.0
    for i in range(0, n, 37):
        x, y, zz, w = philox_py((i & 0xFFFFFFFF, (i >> 32) | (0 << 24), step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFF),
                                (seed & 0xFFFFFFFF, seed >> 32))
        e = 32 if y == 0 else 32 - y.bit_length()
        f = ((x << 20) | (w & 0xFFFFF)) / ld(2 ** 52)
        u = (ld(1) + f) * ld(2.0) ** (-(e + 1))
        v = ld((zz << 12) | (w >> 20)) / ld(2 ** 44)
        rad = np.sqrt(ld(-2) * np.log(u))
        ref = np.array([rad * np.cos(2 * ld(np.pi) * v), rad * np.sin(2 * ld(np.pi) * v)], dtype=np.float64)
        worst = max(worst, np.max(np.abs(z[i] - ref)) / max(float(rad), 1.0))     # error relative to the radius
    assert worst < 4e-15, worst


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_normals_statistics(K, mode):
    from scipy import stats
    z = K.normals(2_000_000, 9, seed=1234, step=5, rng_mode=mode)
    assert z.shape == (2_000_000, 9) and np.isfinite(z).all()
    flat = z.ravel()
    n = flat.size
    assert abs(flat.mean()) < 5 / np.sqrt(n)
    assert abs(flat.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs(stats.kurtosis(flat)) < 5 * np.sqrt(24 / n)
    assert abs(stats.skew(flat)) < 5 * np.sqrt(6 / n)
    assert stats.kstest(flat[:400_000], "norm").pvalue > 1e-4
    # independence across components and walkers
    c = np.corrcoef(z[:200_000].T)
    assert np.abs(c - np.eye(9)).max() < 0.02
    assert abs(np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]) < 0.01
    # reproducible and stream-separated
    assert np.array_equal(z[:1000], K.normals(1000, 9, seed=1234, step=5, rng_mode=mode))
    assert not np.array_equal(z[:1000], K.normals(1000, 9, seed=1234, step=6, rng_mode=mode))
    assert not np.array_equal(z[:1000], K.normals(1000, 9, seed=1235, step=5, rng_mode=mode))


def test_displace_matches_sigma(K):
    m = np.array([1837.1526727546586, 1837.1526727546586, 29156.946166157002])
    sig = np.sqrt(5.0 / m)
    x0 = np.tile(EQ, (500_000, 1, 1))
    x1 = K.displace(x0, sig, seed=9, step=3)
    d = x1 - x0
    for a in range(3):
        assert abs(d[:, a].std() / sig[a] - 1) < 5e-3
        assert abs(d[:, a].mean()) < 5 * sig[a] / np.sqrt(1.5e6)
    z = K.normals(500_000, 9, seed=9, step=3).reshape(-1, 3, 3)
    assert np.array_equal(x1, x0 + sig[None, :, None] * z)      # same stream as pvd_normals, sigma per atom


# ------------------------------------------------------------------ discrete birth/death
@pytest.mark.parametrize("case", ["water", "bigdt", "tiny", "odd"])
def test_branch_discrete_golden_bit_exact(K, case):
    g = golden("branch_discrete_golden.npz")
    counts, idx, b, d, p = K.branch_discrete(g[f"{case}_v"], float(g[f"{case}_vref"]), float(g[f"{case}_dt"]),
                                             g[f"{case}_u"], int(g[f"{case}_n0"]))
    assert np.array_equal(idx, g[f"{case}_idx"])
    assert [b, d, p] == list(g[f"{case}_bdp"])
    assert np.array_equal(np.repeat(np.arange(len(counts)), counts), idx)


@pytest.mark.parametrize("n", [1, 31, 256, 257, 5000, 200_000, 1_000_000])
def test_branch_discrete_vs_oracle(K, oracle, n):
    rng = np.random.default_rng(n)
    v = 0.021 + 0.004 * rng.standard_normal(n)
    vref = float(v.mean())
    u = rng.random(n)
    co, io, bo, do_, po = oracle.birth_or_death_discrete(v, vref, 5.0, u, n)
    c, i, b, d, p = K.branch_discrete(v, vref, 5.0, u, n)
    assert np.array_equal(c, co) and np.array_equal(i, io) and (b, d, p) == (bo, do_, po)


@pytest.mark.parametrize("tag", ["massive_w", "massive_pop"])
def test_branch_discrete_guard(K, tag):
    from pyvibdmc_b200._capi import MassiveEvent
    g = golden("branch_discrete_golden.npz")
    with pytest.raises(MassiveEvent, match="Massive walker birth or death event!!!!!!! Dying..."):
        K.branch_discrete(np.full(400, 0.02 + float(g[f"{tag}_shift"])), 0.02, 5.0, np.full(400, 0.5), 400)
    with pytest.raises(ValueError):
        K.branch_discrete(np.array([np.nan, 0.02]), 0.02, 5.0, np.array([0.5, 0.5]), 2)


def test_vref_and_desc_wts(K, oracle):
    g = golden("vref_desc_golden.npz")
    vr = K.calc_vref(g["v"], int(g["n0"]), float(g["alpha"]))
    assert abs(vr - float(g["vref"])) <= 1e-14 * abs(float(g["vref"]))
    assert np.array_equal(K.desc_wts(g["who_from"], int(g["n0"])), g["desc_wts"])
    gc = golden("branch_continuous_golden.npz")
    dw = K.desc_wts(gc["low_src"], len(gc["low_w"]), gc["low_w"])
    assert np.allclose(dw, gc["low_desc"], rtol=1e-13)
    w = gc["low_w"]
    vr = K.calc_vref(gc["low_vout"], len(w), 0.1, w)
    assert abs(vr - float(gc["low_vref_after"])) <= 1e-13 * abs(float(gc["low_vref_after"]))
