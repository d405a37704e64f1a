"""ctypes binding of libpvd_b200.so (C ABI declared in include/pvd_b200.h).

There is no CPU fallback: if the CUDA library cannot be loaded, or no device is present when a
compute entry point is called, a PvdError is raised.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

__all__ = ["lib", "PvdError", "MassiveEvent", "check", "PvdConfig", "StepStats", "load"]

MASSIVE_MSG = "Massive walker birth or death event!!!!!!! Dying..."
PVD_OK, PVD_E_CUDA, PVD_E_ARG, PVD_E_MASSIVE, PVD_E_STATE, PVD_E_NODEVICE = range(6)
POT_EXTERNAL, POT_HARMONIC, POT_H2O_PS, POT_MORSE1D, POT_NN_H4O2 = range(5)
TRIAL_NONE, TRIAL_HARM1D, TRIAL_H2O_FD, TRIAL_H2O_AN, TRIAL_EXTERNAL = range(5)
IMP_STANDARD, IMP_SECOND_DISPLACEMENT, IMP_EXCITED_STATE = range(3)
WEIGHT_DISCRETE, WEIGHT_CONTINUOUS = 0, 1
RNG_FP64, RNG_FAST, RNG_ZIGGURAT = 0, 1, 2
RNG_DEFAULT = RNG_ZIGGURAT          # exact normals at the lowest cost (include/pvd_b200.h)
RNG_MODES = {'ziggurat': RNG_ZIGGURAT, 'fp64': RNG_FP64, 'fast': RNG_FAST}
ZIGGURAT_LAYERS = 1024
MAX_ATOMS, MAX_COMP, MAX_WORLD = 16, 48, 8
NSUMS = 16 + 4 * MAX_WORLD
# layout of the per-step reduction vector (csrc/pvd_step.cuh): the floating sums travel as three 44-bit chunks of their exact
# 128-bit fixed-point value (quantum 2^-80) so that a SUM over ranks is exact in any order
SUM_CV, SUM_C, SUM_V, SUM_BIRTHS, SUM_DEATHS, SUM_NIN, SUM_ERR, SUM_NACC, SUM_EXT = 0, 3, 6, 9, 10, 11, 12, 13, 16


def sums_encode(x):
    """Host restatement of sum_put_double (pvd_step.cuh): a double -> three exactly summable doubles."""
    from fractions import Fraction
    i = int(Fraction(float(x)) * (1 << 80))            # truncation towards -inf below 2^-80 (exact for |x| >= 2^-27)
    i &= (1 << 128) - 1
    c0, c1, c2 = i & ((1 << 44) - 1), (i >> 44) & ((1 << 44) - 1), i >> 88
    if c2 >= 1 << 39:
        c2 -= 1 << 40
    return [float(c0) * 2.0 ** -80, float(c1) * 2.0 ** -36, float(c2) * 2.0 ** 8]


def sums_decode(c):
    """Host restatement of sum_get: the (reduced) chunks -> the double nearest to the exact total."""
    from fractions import Fraction
    tot = Fraction(c[0]) + Fraction(c[1]) + Fraction(c[2])
    return float(tot)


class PvdError(RuntimeError):
    pass


class MassiveEvent(ValueError):
    """Same exception type and text as the reference's population guard (pyvibdmc.py:400,413,717)."""


class PvdConfig(C.Structure):
    _fields_ = [("natoms", C.c_int32), ("ndim", C.c_int32), ("weighting", C.c_int32), ("potential", C.c_int32),
                ("trial", C.c_int32), ("rng_mode", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32),
                ("world_size", C.c_int32), ("num_walkers", C.c_int64), ("capacity", C.c_int64),
                ("delta_t", C.c_double), ("alpha", C.c_double), ("thresh_lower", C.c_double),
                ("thresh_upper", C.c_double), ("seed", C.c_uint64), ("masses", C.c_double * MAX_ATOMS),
                ("pot_params", C.c_double * MAX_COMP), ("stats_ring", C.c_int64), ("imp_variant", C.c_int32),
                ("reserved_", C.c_int32)]


class StepStats(C.Structure):
    _fields_ = [("vref", C.c_double), ("pop", C.c_double), ("v_avg", C.c_double), ("v_max", C.c_double),
                ("v_min", C.c_double), ("w_max", C.c_double), ("w_min", C.c_double), ("dt_eff", C.c_double),
                ("births", C.c_int64), ("deaths", C.c_int64), ("rejected", C.c_int64), ("step", C.c_int64)]


STATS_DTYPE = np.dtype([("vref", "f8"), ("pop", "f8"), ("v_avg", "f8"), ("v_max", "f8"), ("v_min", "f8"),
                        ("w_max", "f8"), ("w_min", "f8"), ("dt_eff", "f8"), ("births", "i8"), ("deaths", "i8"),
                        ("rejected", "i8"), ("step", "i8")])
assert STATS_DTYPE.itemsize == C.sizeof(StepStats)

_P = C.c_void_p
_I64, _I32, _F64, _U64 = C.c_int64, C.c_int32, C.c_double, C.c_uint64

# name -> (restype, argtypes); must list every function declared in include/pvd_b200.h
SIGNATURES = {
    "pvd_abi_version": (C.c_int, []),
    "pvd_sizeof_config": (C.c_int, []),
    "pvd_sizeof_step_stats": (C.c_int, []),
    "pvd_last_error": (C.c_char_p, []),
    "pvd_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pvd_set_device": (C.c_int, [C.c_int]),
    "pvd_measure_fp64_peak": (C.c_int, [C.POINTER(_F64)]),
    "pvd_last_kernel_ms": (C.c_int, [C.POINTER(_F64)]),
    "pvd_launch_count": (_I64, []),
    "pvd_pes_h2o": (C.c_int, [_P, _I64, _P]),
    "pvd_pes_h2o_params": (C.c_int, [_P, _P]),
    "pvd_pes_harmonic": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "pvd_pes_morse1d": (C.c_int, [_P, _I64, _F64, _F64, _P]),
    "pvd_displace": (C.c_int, [_P, _I64, _I32, _I32, _P, _U64, _U64, _I32]),
    "pvd_normals": (C.c_int, [_P, _I64, _I32, _U64, _U64, _I32]),
    "pvd_philox_kat": (C.c_int, [_P, _P, _P]),
    "pvd_ziggurat_table": (C.c_int, [_P, _P]),
    "pvd_distit": (C.c_int, [_P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _I32, _P, _I32, _I32, _I32, _P]),
    "pvd_branch_discrete": (C.c_int, [_P, _I64, _F64, _F64, _P, _I64, _P, _P, _I64, _P]),
    "pvd_branch_continuous": (C.c_int, [_P, _P, _I64, _F64, _F64, _F64, _F64, _P, _P]),
    "pvd_calc_vref": (C.c_int, [_P, _P, _I64, _I64, _F64, C.POINTER(_F64)]),
    "pvd_desc_wts": (C.c_int, [_P, _P, _I64, _I64, _P]),
    "pvd_trial_drift": (C.c_int, [_I32, _P, _I64, _I32, _I32, _P, _I64, _P, _P, _P]),
    "pvd_metropolis": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _F64, _P]),
    "pvd_local_kin": (C.c_int, [_P, _I64, _I32, _I32, _P, _P]),
    "pvd_nn_h4o2_set_weights": (C.c_int, [_P, _I64]),
    "pvd_nn_config": (C.c_int, [_I32, _I32, _I32]),
    "pvd_nn_h4o2": (C.c_int, [_P, _I64, _P]),
    "pvd_coulomb_descriptor": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "pvd_sim_create": (C.c_int, [C.POINTER(PvdConfig), C.POINTER(_P)]),
    "pvd_sim_destroy": (C.c_int, [_P]),
    "pvd_sim_set_stream": (C.c_int, [_P, _P]),
    "pvd_sim_upload": (C.c_int, [_P, _P, _I64, _P]),
    "pvd_sim_set_pots": (C.c_int, [_P, _P, _I64]),
    "pvd_sim_init_finalize": (C.c_int, [_P]),
    "pvd_sim_set_trial_table": (C.c_int, [_P, _P, _I64]),
    "pvd_sim_set_nn_weights": (C.c_int, [_P, _P, _I64]),
    "pvd_sim_run": (C.c_int, [_P, _I64, _I32]),
    "pvd_sim_set_resident": (C.c_int, [_P, _I32]),
    "pvd_sim_step_injected": (C.c_int, [_P, _P, _P, _P]),
    "pvd_sim_ext_move": (C.c_int, [_P, _P, C.POINTER(_I64)]),
    "pvd_sim_ext_finish": (C.c_int, [_P, _P, _I64, _I32]),
    "pvd_sim_sums_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "pvd_sim_set_sums_ptr": (C.c_int, [_P, _P]),
    "pvd_sim_step_local": (C.c_int, [_P, _I32]),
    "pvd_sim_step_finalize": (C.c_int, [_P]),
    "pvd_sim_snapshot_begin": (C.c_int, [_P]),
    "pvd_sim_snapshot_wait": (C.c_int, [_P, _P, _P, _P, _P, _I64, C.POINTER(_I64), C.POINTER(_F64)]),
    "pvd_sim_mailbox_handle": (C.c_int, [_P, _P]),
    "pvd_sim_mailbox_connect": (C.c_int, [_P, _P, _I32]),
    "pvd_sim_run_mailbox": (C.c_int, [_P, _I64, _I32]),
    "pvd_sim_imp_move_local": (C.c_int, [_P]),
    "pvd_sim_imp_branch_local": (C.c_int, [_P, _I32]),
    "pvd_sim_dw_begin": (C.c_int, [_P, _I64]),
    "pvd_sim_dw_end": (C.c_int, [_P, _P, _I64]),
    "pvd_sim_dw_peek": (C.c_int, [_P, _P, _I64]),
    "pvd_sim_set_masses": (C.c_int, [_P, _P, _I32]),
    "pvd_sim_dw_parent": (C.c_int, [_P, _P, _P, C.POINTER(_I64)]),
    "pvd_sim_sync": (C.c_int, [_P]),
    "pvd_sim_state": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_F64), C.POINTER(_I64), C.POINTER(_I32)]),
    "pvd_sim_download": (C.c_int, [_P, _P, _P, _P, _P, _I64, C.POINTER(_I64)]),
    "pvd_sim_stats": (C.c_int, [_P, _I64, _I64, _P]),
    "pvd_sim_last_run_ms": (C.c_int, [_P, C.POINTER(_F64)]),
    "pvd_sim_download_imp": (C.c_int, [_P, _P, _P, _P, _I64]),
    "pvd_sim_dw_resume": (C.c_int, [_P, _P, _I64, _P, _P, _I64]),
    "pvd_sim_dw_end_begin": (C.c_int, [_P, _I64]),
    "pvd_sim_dw_end_wait": (C.c_int, [_P, _P, _P, _P, _I64]),
    "pvd_sim_ext_move_device": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(_I64)]),
    "pvd_sim_ext_finish_device": (C.c_int, [_P, _P, _I64, _I32]),
    "pvd_sim_coords_device": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(_I64)]),
    "pvd_sim_set_pots_device": (C.c_int, [_P, _P, _I64]),
    "pvd_sim_imp_ext_init": (C.c_int, [_P, _P, _P, _P, _P]),
    "pvd_sim_imp_ext_propose": (C.c_int, [_P, _P, _P, C.POINTER(_I64)]),
    "pvd_sim_imp_ext_accept": (C.c_int, [_P, _P, _P, _P, _I64, _P]),
    "pvd_sim_imp_ext_finish": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "pvd_sim_export_tail": (C.c_int, [_P, _I64, _P, _P, _P, _P]),
    "pvd_sim_import": (C.c_int, [_P, _I64, _P, _P, _P, _P]),
    "pvd_sim_export_tail_device": (C.c_int, [_P, _I64, C.POINTER(C.c_void_p), C.POINTER(_I32)]),
    "pvd_sim_import_device": (C.c_int, [_P, _I64, _P, _I32]),
}

_lib = None


def load(rebuild_if_stale=True):
    """Load (building first if the sources are newer) the in-tree CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PVD_B200_LIB") or _build.LIB      # developer knob: A/B a differently built library
    if path == _build.LIB and rebuild_if_stale and _build.is_stale():
        try:
            _build.build_library()
        except Exception as e:  # no nvcc on the box and no prebuilt library
            if not os.path.exists(path):
                raise PvdError(f"libpvd_b200.so is missing and could not be built: {e}") from e
            import warnings
            warnings.warn(f"libpvd_b200.so is older than its sources and could not be rebuilt ({e}): using the stale library", RuntimeWarning)
    if not os.path.exists(path):
        raise PvdError(f"CUDA extension not found at {path}; run `python -m pyvibdmc_b200.build`")
    try:
        handle = C.CDLL(path)
    except OSError as e:
        raise PvdError(f"cannot load {path}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError here == ABI drift: fail loudly
        fn.restype, fn.argtypes = res, args
    if handle.pvd_abi_version() != 1:
        raise PvdError("libpvd_b200.so ABI version mismatch")
    if handle.pvd_sizeof_config() != C.sizeof(PvdConfig) or handle.pvd_sizeof_step_stats() != C.sizeof(StepStats):
        raise PvdError("libpvd_b200.so struct layout differs from the ctypes binding (stale build?)")
    _lib = handle
    return _lib


class _Lazy:
    def __getattr__(self, name):
        return getattr(load(), name)


lib = _Lazy()


def last_error():
    return load().pvd_last_error().decode("utf-8", "replace")


def check(rc):
    """Translate a C-ABI status into the reference's exception behaviour."""
    if rc == PVD_OK:
        return
    msg = last_error()
    if rc == PVD_E_MASSIVE:
        raise MassiveEvent(msg)
    if rc == PVD_E_ARG:
        raise ValueError(msg)
    raise PvdError(f"[pvd error {rc}] {msg}")


def ptr(a):
    """Raw pointer of a C-contiguous NumPy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
