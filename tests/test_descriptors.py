"""DistIt descriptors (tensorflow_descriptors/distance_descriptors.py): distance / Coulomb / SPF, atoms sorted within
sub-lists by column norm, groups swapped by summed column norms, upper triangle or full matrix.  The golden fixture
holds the unmodified reference's output (tests/golden/make_golden.py: gen_descriptor) for every variant below; the
oracle restatement and the CUDA kernel behind the drop-in DistIt class must both reproduce it bit for bit."""
import numpy as np
import pytest

from conftest import golden

SA1, SA2 = [[0], [1, 2], [3], [4, 5]], [[0, 3], [1, 2, 4, 5]]
SG1, SG2 = [[0, 1, 2], [3, 4, 5]], [[1, 2], [4, 5]]
CASES = {"distance": dict(method="distance"), "spf": dict(method="spf"), "coulomb_full": dict(method="coulomb", full_mat=True),
         "distance_full": dict(method="distance", full_mat=True), "spf_full": dict(method="spf", full_mat=True),
         "coulomb_atoms": dict(method="coulomb", sorted_atoms=SA1), "coulomb_atoms2_full": dict(method="coulomb", sorted_atoms=SA2, full_mat=True),
         "distance_groups": dict(method="distance", sorted_groups=SG1), "coulomb_groups2_full": dict(method="coulomb", sorted_groups=SG2, full_mat=True),
         "spf_atoms_groups": dict(method="spf", sorted_atoms=SA1, sorted_groups=SG1),
         "spf_atoms2_full": dict(method="spf", sorted_atoms=SA2, full_mat=True),
         "coulomb_atoms_groups": dict(method="coulomb", sorted_atoms=SA1, sorted_groups=SG1)}


def same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_distit_matches_reference(oracle, name):
    g = golden("distit_golden.npz")
    kw = CASES[name]
    out = oracle.distit(g["coords"][:64], g["zs"], eq_xyz=g["eq_xyz"] if kw["method"] == "spf" else None, **kw)
    assert same(out, g[name][:64])


LARGE = {"spf10_atoms_groups_full": ("coords10", "zs10", dict(method="spf", sorted_atoms=[[0, 3, 6], [1, 2, 4, 5, 7, 8], [9]],
                                                                sorted_groups=[[0, 1, 2], [3, 4, 5], [6, 7, 8]], full_mat=True)),
         "coulomb16_atoms": ("coords16", None, dict(method="coulomb", sorted_atoms=[list(range(16))])),
         # groups of eight: np.sum switches to its unrolled pairwise order at eight summands
         "distance16_groups8": ("coords16", None, dict(method="distance", sorted_groups=[list(range(8)), list(range(8, 16))])),
         "coulomb16_groups8_full": ("coords16", "zs16", dict(method="coulomb", sorted_groups=[list(range(8)), list(range(8, 16))],
                                                             full_mat=True))}


@pytest.mark.parametrize("name", sorted(LARGE))
def test_oracle_distit_larger_molecules(oracle, name):
    g = golden("distit_golden.npz")
    ck, zk, kw = LARGE[name]
    zs = g[zk] if zk else [1] * 16
    assert same(oracle.distit(g[ck], zs, eq_xyz=g["eq10"] if kw["method"] == "spf" else None, **kw), g[name])


@pytest.mark.parametrize("gs", [2, 3, 7, 8, 9, 15, 16])
def test_oracle_group_sum_order_is_numpys(oracle, gs):
    """The group totals decide the swap, so their last bit matters: np.sum over a contiguous axis adds left to right below
    eight summands and in its unrolled pairwise order from eight on (distance_descriptors.py:146 sums such an axis)."""
    rng = np.random.default_rng(gs)
    a = rng.random((20000, gs)) * rng.choice([1e-3, 1.0, 1e3], size=(20000, gs))
    assert np.array_equal(np.sum(a, axis=1), np.array([oracle._numpy_sum_order(list(row)) for row in a]))


def _mirror_variants(g, na):
    """(keyword arguments, coordinates, reference output) of the mirror-image cases in the golden file."""
    c, g2 = g[f"mirror{na}"], [list(range(na // 2)), list(range(na // 2, na))]
    extra = []
    if na == 16:       # an equilibrium structure whose own sort differs between one and two copies (the reference uses two, :75)
        extra = [(dict(method="spf", eq_xyz=g["mirror16b"][0], sorted_groups=g2, full_mat=True), g["mirror16b"][1:], g["mirror16b_spf_full"])]
    return extra + [(dict(method="distance", sorted_groups=g2), c, g[f"mirror{na}_groups"]),
            (dict(method="distance", sorted_groups=g2), c[:1], g[f"mirror{na}_groups_one"]),      # a single walker sums differently
            (dict(method="spf", eq_xyz=c[0], sorted_groups=g2, full_mat=True), c[1:], g[f"mirror{na}_spf_full"])]


@pytest.mark.parametrize("na", [6, 10, 16])
def test_oracle_distit_mirror_image_groups(oracle, na):
    """Groups that are mirror images of each other: which one comes first hangs on the last bit of the totals, i.e. on
    NumPy's summation order inside distance_descriptors.py:146 (contiguous axis for the norms, strided for the totals)."""
    for kw, c, ref in _mirror_variants(golden("distit_golden.npz"), na):
        assert same(oracle.distit(c, [1] * na, **kw), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("na", [6, 10, 16])
def test_device_distit_mirror_image_groups(na):
    from pyvibdmc_b200.simulation_utilities.tensorflow_descriptors import DistIt
    for kw, c, ref in _mirror_variants(golden("distit_golden.npz"), na):
        assert same(np.asarray(DistIt([1] * na, **kw).run(c)), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LARGE))
def test_device_distit_larger_molecules(name):
    from pyvibdmc_b200.simulation_utilities.tensorflow_descriptors import DistIt
    g = golden("distit_golden.npz")
    ck, zk, kw = LARGE[name]
    zs = g[zk] if zk else [1] * 16
    assert same(np.asarray(DistIt(zs, eq_xyz=g["eq10"] if kw["method"] == "spf" else None, **kw).run(g[ck])), g[name])


def test_distit_argument_errors():
    from pyvibdmc_b200.simulation_utilities.tensorflow_descriptors import DistIt
    with pytest.raises(ValueError, match="eq_xyz is not set but using spf"):
        DistIt([8, 1, 1], "spf")
    with pytest.raises(ValueError, match="Please put all atoms in sorted_atoms list"):
        DistIt([8, 1, 1], "distance", sorted_atoms=[[0], [1]])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_distit_matches_reference(name):
    from pyvibdmc_b200.simulation_utilities.tensorflow_descriptors import DistIt
    g = golden("distit_golden.npz")
    kw = CASES[name]
    d = DistIt(g["zs"], eq_xyz=g["eq_xyz"] if kw["method"] == "spf" else None, **kw)
    assert same(np.asarray(d.run(g["coords"])), g[name])
    assert np.asarray(d.run(g["coords"][:0])).shape == (0,) + g[name].shape[1:]        # empty batch
    assert same(np.asarray(d.run(g["coords"][:37])), g[name][:37])                     # ragged size


@pytest.mark.gpu
def test_device_distit_larger_molecule_vs_oracle(oracle):
    """10 atoms, three sorted sub-lists and three groups of three, against the oracle restatement."""
    from pyvibdmc_b200.simulation_utilities.tensorflow_descriptors import DistIt
    rng = np.random.default_rng(5)
    eq = rng.normal(0, 2.0, size=(10, 3))
    cds = eq[None] + rng.normal(0, 0.3, size=(300, 10, 3))
    zs = [8, 1, 1, 8, 1, 1, 8, 1, 1, 6]
    sa, sg = [[0, 3, 6], [1, 2, 4, 5, 7, 8], [9]], [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
    for method in ("distance", "coulomb", "spf"):
        for full in (False, True):
            kw = dict(method=method, sorted_atoms=sa, sorted_groups=sg, full_mat=full)
            ref = oracle.distit(cds, zs, eq_xyz=eq if method == "spf" else None, **kw)
            out = np.asarray(DistIt(zs, eq_xyz=eq if method == "spf" else None, **kw).run(cds))
            assert same(out, ref), (method, full)


def _random_distit_case(rng):
    """A random molecule of 3-16 atoms with random sorted sub-lists / groups / method / matrix form; structures are often
    mirror-symmetric with noise 0, 1e-12, 1e-6 or 0.1, so that the sorts are decided by exact or last-bit ties."""
    na = int(rng.integers(3, 17))
    sg = sa = None
    if rng.random() < 0.7:
        ng = int(rng.integers(2, min(5, na)))
        gs = int(rng.integers(1, max(na // ng, 1) + 1))
        perm = rng.permutation(na)[:ng * gs]
        sg = [perm[i * gs:(i + 1) * gs].tolist() for i in range(ng)]
    if rng.random() < 0.6:
        perm, sa, i = rng.permutation(na).tolist(), [], 0
        while i < na:
            ln = int(rng.integers(1, 5))
            sa.append(sorted(perm[i:i + ln]))
            i += ln
    method = ['distance', 'coulomb', 'spf'][int(rng.integers(3))]
    n = int(rng.choice([1, 2, 5, 40]))
    base = rng.normal(0, 2.0, size=(na, 3))
    if rng.random() < 0.6:
        h = na // 2
        base[h:2 * h] = base[:h][rng.permutation(h)] * np.array([-1, 1, 1])
        base[:h, 0] += 0.7
        base[h:2 * h, 0] -= 0.7
    cds = base[None] + rng.normal(0, 1, size=(n, na, 3)) * rng.choice([0, 1e-12, 1e-6, 0.1])
    eq = base + rng.normal(0, 1e-3, size=base.shape) if method == 'spf' else None
    return rng.choice([1, 6, 8], size=na).tolist(), cds, eq, dict(method=method, sorted_atoms=sa, sorted_groups=sg,
                                                                  full_mat=bool(rng.random() < 0.5))


@pytest.mark.gpu
def test_device_distit_random_tie_prone_cases_vs_oracle(oracle):
    """The kernel against the oracle on the inputs of tools/fuzz_oracle_vs_reference.py (which holds the oracle to the
    reference): same bits, including the single-walker summation order and the lower-index rule on exact ties."""
    from pyvibdmc_b200.simulation_utilities.tensorflow_descriptors import DistIt
    rng = np.random.default_rng(2024)
    for it in range(80):
        zs, cds, eq, kw = _random_distit_case(rng)
        ref = oracle.distit(cds, zs, eq_xyz=eq, **kw)
        out = np.asarray(DistIt(zs, eq_xyz=eq, **kw).run(cds))
        assert same(out, ref), (it, len(zs), len(cds), kw)
