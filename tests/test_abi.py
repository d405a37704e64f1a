"""CPU tests: the C-ABI library builds, loads and exports every symbol declared in include/pvd_b200.h."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "pvd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pvd_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    from pyvibdmc_b200 import build, _capi
    path = build.build_library()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"declared in pvd_b200.h but not exported: {missing}"
    assert sorted(_capi.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert handle.pvd_abi_version() == 1


def test_struct_layouts_match_header():
    from pyvibdmc_b200 import _capi
    # pvd_config: 9 int32 (+pad) + 2 int64 + 4 double + uint64 + 16 + 48 doubles + int64 + 2 int32
    assert ctypes.sizeof(_capi.PvdConfig) == 40 + 16 + 32 + 8 + 8 * 16 + 8 * 48 + 8 + 8
    assert ctypes.sizeof(_capi.StepStats) == 96


def test_enumerations_match_header():
    from pyvibdmc_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "pvd_b200.h")).read()
    rng = dict((k, int(v)) for k, v in re.findall(r"(PVD_RNG_[A-Z0-9]+)\s*=\s*(\d+)", hdr))
    assert rng == {"PVD_RNG_FP64": _capi.RNG_FP64, "PVD_RNG_FAST": _capi.RNG_FAST, "PVD_RNG_ZIGGURAT": _capi.RNG_ZIGGURAT}
    assert int(re.search(r"#define PVD_ZIGGURAT_LAYERS (\d+)", hdr).group(1)) == _capi.ZIGGURAT_LAYERS
    assert _capi.RNG_MODES["ziggurat"] == _capi.RNG_DEFAULT


def test_no_device_fails_loudly():
    """Without a GPU every compute entry point must raise (no silent CPU fallback)."""
    import numpy as np
    from pyvibdmc_b200 import kernels, _capi
    if kernels.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_capi.PvdError):
        kernels.pes_h2o(np.zeros((4, 3, 3)) + np.eye(3))
    with pytest.raises(_capi.PvdError):
        kernels.normals(8, 9, seed=1, rng_mode=_capi.RNG_ZIGGURAT)
    from pyvibdmc_b200.simulation_utilities.tensorflow_descriptors import DistIt
    with pytest.raises(_capi.PvdError):
        DistIt([8, 1, 1], "distance").run(np.zeros((4, 3, 3)) + np.eye(3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyvibdmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src, f"{f} reads the reference at run time"
