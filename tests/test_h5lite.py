"""CPU tests of the HDF5 writer/reader used for sim_info / wave-function dumps (SURVEY f-1).
The writer must produce exactly what h5py produced for the reference's shipped tutorial files."""
import glob
import os
import struct

import numpy as np
import pytest

from conftest import golden

REF = "/root/reference/pyvibdmc/sample_sim_data"


def test_siminfo_layout_matches_h5py_bytes(tmp_path):
    from pyvibdmc_b200.simulation_utilities import h5lite
    g = golden("h5_siminfo_golden.npz")
    vref, pop = np.zeros((5000, 2)), np.zeros((5000, 2))
    f = tmp_path / "s.hdf5"
    h5lite.write_h5(str(f), ['vref_vs_tau', 'pop_vs_tau', 'atomic_nums', 'atomic_masses'],
                    [vref, pop, g["atomic_nums"], g["atomic_masses"]], mtime=int(g["mtime"]))
    raw = f.read_bytes()
    assert len(raw) == int(g["size"])
    assert raw[:2048] == g["head"].tobytes()                 # superblock, B-tree, heap, SNOD, object headers
    assert raw[162048:] == g["tail"].tobytes()               # small-data block + last object header
    back = h5lite.read_h5(str(f))
    assert back["vref_vs_tau"].shape == (5000, 2) and np.array_equal(back["atomic_nums"], g["atomic_nums"])


def test_wfn_layout_matches_h5py_bytes(tmp_path):
    from pyvibdmc_b200.simulation_utilities import h5lite
    g = golden("h5_wfn_golden.npz")
    n = int(g["n"])
    f = tmp_path / "w.hdf5"
    h5lite.write_h5(str(f), ['coords', 'desc_wts'], [np.zeros((n, 3, 3)), np.zeros(n)], mtime=int(g["mtime"]))
    raw = f.read_bytes()
    assert len(raw) == int(g["size"]) == 2048 + 80 * n
    assert raw[:2048] == g["head"].tobytes()


@pytest.mark.parametrize("shapes", [[(30, 2), (30, 2), (3,), (3,)], [(7, 1, 1), (7,), (7,)], [(1,)], [(300, 6, 3), (300,)]])
def test_round_trip_small_and_ragged(tmp_path, shapes):
    from pyvibdmc_b200.simulation_utilities import h5lite
    rng = np.random.default_rng(0)
    keys = [f"k{i}" for i in range(len(shapes))]
    vals = [rng.normal(size=s) for s in shapes]
    vals[-1] = (vals[-1] * 10).astype(np.int64)
    f = tmp_path / "r.hdf5"
    with h5lite.File(str(f), "w") as hf:
        for k, v in zip(keys, vals):
            hf.create_dataset(k, data=v)
    with h5lite.File(str(f), "r") as hf:
        for k, v in zip(keys, vals):
            assert np.array_equal(hf[k], v) and hf[k].dtype == v.dtype
    raw = f.read_bytes()
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)       # end-of-file address in the superblock


def test_python_lists_become_int64_like_h5py(tmp_path):
    from pyvibdmc_b200.simulation_utilities import h5lite
    f = tmp_path / "l.hdf5"
    h5lite.write_h5(str(f), ["atomic_nums"], [[1, 1, 8]])
    assert h5lite.read_h5(str(f))["atomic_nums"].dtype == np.int64


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sample data only exist in the build container")
def test_every_shipped_file_is_reproduced_byte_for_byte(tmp_path):
    from pyvibdmc_b200.simulation_utilities import h5lite
    files = glob.glob(f"{REF}/*.hdf5") + glob.glob(f"{REF}/wfns/*.hdf5")
    assert len(files) == 30
    for p in files:
        ref = open(p, "rb").read()
        d = h5lite.read_h5(p)
        keys = ['vref_vs_tau', 'pop_vs_tau', 'atomic_nums', 'atomic_masses'] if 'vref_vs_tau' in d else ['coords', 'desc_wts']
        first = d[keys[0]]
        mt = struct.unpack_from("<I", ref, 800 + 16 + 8 + (8 + 16 * first.ndim) + 32 + 16 + 32 + 8 + 4)[0]
        out = tmp_path / "x.hdf5"
        h5lite.write_h5(str(out), keys, [d[k] for k in keys], mtime=mt)
        assert out.read_bytes() == ref, p
