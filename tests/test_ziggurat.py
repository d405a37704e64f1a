"""The ziggurat normal generator (PVD_RNG_ZIGGURAT) that replaces np.random.normal in move_randomly
(pyvibdmc.py:540-547).  The reference draws from NumPy's global MT19937 stream, which no counter-based
generator can reproduce, so the checks are: the table is a ziggurat (equal areas, closed), the device
generator equals a NumPy restatement of the algorithm on the same Philox bits, and the output is
standard normal (moments, fine-grained chi-square, tails beyond R)."""
import ctypes as C

import numpy as np
import pytest

N = 1024


def table():
    from pyvibdmc_b200 import _capi
    lib = _capi.load()
    x, f = np.zeros(N + 1), np.zeros(N + 1)
    assert lib.pvd_ziggurat_table(x.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p)) == 0
    return x, f


def test_table_is_a_ziggurat():
    """No GPU needed: equal-area layers that close at x = 0 and add up to the half-normal integral."""
    from scipy.special import erfc
    from pyvibdmc_b200 import _capi
    assert _capi.ZIGGURAT_LAYERS == N
    x, f = table()
    R = x[1]
    assert x[N] == 0.0 and f[N] == 1.0 and np.all(np.diff(x) < 0) and 3.9 < R < 4.2
    assert np.allclose(f, np.exp(-0.5 * x * x), rtol=1e-14, atol=0)      # f from the long-double x
    V = R * f[1] + np.sqrt(np.pi / 2) * erfc(R / np.sqrt(2))              # base strip: rectangle + tail
    assert abs(x[0] * f[1] - V) < 1e-15 * V                                  # x_0 = V / f(R)
    areas = x[1:N] * (f[2:] - f[1:N])                                        # layers 1 .. N-1
    assert np.max(np.abs(areas / V - 1)) < 2e-12                             # (f differences cancel: ~1e-13)
    # the rectangles overhang the curve by the rejected parts of the wedges only: 0.2 % of the total area
    assert 1.0 < N * V / np.sqrt(np.pi / 2) < 1.003


# ------------------------------------------------------------------ NumPy restatement on the same Philox bits
def philox_np(c0, c1, c2, c3, k0, k1):
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    m32 = np.uint64(0xFFFFFFFF)
    c = [np.asarray(v, dtype=np.uint64) for v in np.broadcast_arrays(c0, c1, c2, c3)]
    for r in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        ka, kb = np.uint64((k0 + r * W0) & 0xFFFFFFFF), np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ ka, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ kb, p0 & m32]
    return c


def draw(seed, slot, step, purpose, call):
    slot = np.asarray(slot, dtype=np.uint64)
    return philox_np(slot & np.uint64(0xFFFFFFFF), (slot >> np.uint64(32)) | np.uint64(call << 24), np.uint64(step & 0xFFFFFFFF),
                     np.uint64(((step >> 32) & 0xFFFFFF) | (purpose << 24)), seed & 0xFFFFFFFF, seed >> 32)


def decode(lo, hi):
    layer = (lo & np.uint64(N - 1)).astype(np.int64)
    sign = np.where((lo >> np.uint64(10)) & np.uint64(1), -1.0, 1.0)
    mant = (hi << np.uint64(20)) | (lo >> np.uint64(12))
    return layer, mant.astype(np.float64) * 2.0 ** -52, sign


def u53(lo, hi, open_=False):
    r = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return (r.astype(np.float64) + 1.0) * 2.0 ** -53 if open_ else r.astype(np.float64) * 2.0 ** -53


def ziggurat_numpy(seed, step, n, nc, x, f):
    """z[i, c] as the device defines it; also returns the mask of normals that left the common path."""
    z = np.zeros((n, nc))
    slow = np.zeros((n, nc), dtype=bool)
    slots = np.arange(n)
    ratio = x[1:] / x[:-1]
    R = x[1]
    for c in range(nc):
        r = draw(seed, slots, step, 0, c >> 1)
        lo, hi = (r[2], r[3]) if c & 1 else (r[0], r[1])
        layer, mag, sign = decode(lo, hi)
        z[:, c] = sign * (mag * x[layer])
        slow[:, c] = ~(mag < ratio[layer])
        for i in np.nonzero(slow[:, c])[0]:
            l, m, s = int(layer[i]), float(mag[i]), float(sign[i])
            attempt = 0
            while True:
                if m < ratio[l]:
                    z[i, c] = s * (m * x[l])
                    break
                q = [int(v) for v in draw(seed, i, step, 16 + c, attempt)]
                q = [np.uint64(v) for v in q]
                if l == 0:
                    more = attempt + 1
                    while True:
                        xx = -np.log(u53(q[0], q[1], True)) / R
                        yy = -np.log(u53(q[2], q[3], True))
                        if yy + yy > xx * xx:
                            break
                        q = [np.uint64(int(v)) for v in draw(seed, i, step, 16 + c, more)]
                        more += 1
                    z[i, c] = s * (R + xx)
                    break
                xv = m * x[l]
                if f[l] + u53(q[0], q[1]) * (f[l + 1] - f[l]) < np.exp(-0.5 * xv * xv):
                    z[i, c] = s * xv
                    break
                lr, mr, sr = decode(np.uint64(q[2]), np.uint64(q[3]))
                l, m, s = int(lr), float(mr), float(sr)
                attempt += 1
    return z, slow


def test_numpy_restatement_is_standard_normal():
    """No GPU needed: the algorithm the device implements (restated above on the same Philox bits and the same
    table) produces standard normals -- chi-square over 512 equiprobable bins, tails, moments at 1.8e6 draws."""
    from scipy import stats
    x, f = table()
    z, slow = ziggurat_numpy(0xC0FFEE, 11, 200_000, 9, x, f)
    z = z.ravel()
    n, nb = z.size, 512
    counts = np.bincount(np.searchsorted(stats.norm.ppf(np.arange(1, nb) / nb), z), minlength=nb)
    chi2 = ((counts - n / nb) ** 2 / (n / nb)).sum()
    assert stats.chi2.sf(chi2, nb - 1) > 1e-4, chi2
    for t in (2.0, 3.0, 4.0):
        p = 2 * stats.norm.sf(t)
        assert abs((np.abs(z) > t).sum() - n * p) < 5 * np.sqrt(n * p) + 1
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert 0.003 < slow.mean() < 0.0056


@pytest.fixture(scope="module")
def K():
    from pyvibdmc_b200 import kernels
    assert kernels.device_count() > 0, "GPU tests need a CUDA device"
    return kernels


@pytest.mark.gpu
def test_device_generator_equals_numpy_restatement(K):
    seed, step, n, nc = 0x0123456789ABCDEF, (5 << 32) | 77, 120_000, 9
    x, f = table()
    z = K.normals(n, nc, seed=seed, step=step, rng_mode=2)
    ref, slow = ziggurat_numpy(seed, step, n, nc, x, f)
    assert np.array_equal(z[~slow], ref[~slow])                       # common path: bit-exact
    assert 0.003 < slow.mean() < 0.0056                               # 1 - 0.9957 of the normals
    # wedges / tail: same accept-reject decisions; the tail's two logs may differ in the last bits
    assert np.allclose(z[slow], ref[slow], rtol=1e-13, atol=0)
    assert np.array_equal(z[slow & (np.abs(ref) < x[1])], ref[slow & (np.abs(ref) < x[1])])
    assert (np.abs(ref) > x[1]).sum() > 20                            # the tail branch was exercised
    # odd component counts and the fused kernels' deferred handling give the same numbers
    assert np.array_equal(K.normals(5000, 3, seed=seed, step=step, rng_mode=2), z[:5000, :3])
    assert np.array_equal(K.normals(5000, 18, seed=seed, step=step, rng_mode=2)[:, :9], z[:5000])


@pytest.mark.gpu
def test_ziggurat_distribution_fine_grained(K):
    """1.8e7 normals: equiprobable-bin chi-square (sensitive to a wrong wedge or layer edge) and tail counts."""
    from scipy import stats
    z = K.normals(2_000_000, 9, seed=99, step=3, rng_mode=2).ravel()
    n = z.size
    nb = 4096
    edges = stats.norm.ppf(np.arange(1, nb) / nb)
    counts = np.bincount(np.searchsorted(edges, z), minlength=nb)
    chi2 = ((counts - n / nb) ** 2 / (n / nb)).sum()
    assert stats.chi2.sf(chi2, nb - 1) > 1e-4, chi2
    for t in (3.0, 4.0, table()[0][1], 4.5, 5.0):
        p = 2 * stats.norm.sf(t)
        got = (np.abs(z) > t).sum()
        assert abs(got - n * p) < 5 * np.sqrt(n * p) + 1, (t, got, n * p)
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs((z ** 4).mean() - 3) < 5 * np.sqrt(96 / n)
