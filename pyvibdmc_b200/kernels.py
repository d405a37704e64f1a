"""NumPy-facing wrappers of the C ABI: the plug-in level calls (host arrays in, host arrays out)
and DeviceSim, the HBM-resident walker ensemble that DMC_Sim drives.

Every function here runs on the GPU through libpvd_b200.so; nothing falls back to the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, f64, lib, ptr

__all__ = ["pes_h2o", "pes_h2o_params", "pes_harmonic", "pes_morse1d", "displace", "normals", "philox",
           "branch_discrete", "branch_continuous", "calc_vref", "desc_wts", "trial_drift", "metropolis",
           "local_kin", "nn_h4o2", "nn_h4o2_set_weights", "coulomb_descriptor", "fp64_peak", "last_kernel_ms",
           "launch_count", "device_count", "DeviceSim"]


def device_count():
    n = C.c_int(0)
    lib.pvd_device_count(C.byref(n))
    return n.value


def fp64_peak():
    v = C.c_double(0)
    check(lib.pvd_measure_fp64_peak(C.byref(v)))
    return v.value


def last_kernel_ms():
    v = C.c_double(0)
    check(lib.pvd_last_kernel_ms(C.byref(v)))
    return v.value


def launch_count():
    return int(lib.pvd_launch_count())


# ----------------------------------------------------------------------------- potentials
def pes_h2o(cds):
    """GPU water_pot(cds): (N,3,3) bohr, atoms H,H,O -> (N,) Hartree
    (reference: FortPots/Partridge_Schwenke_H2O/h2o_potential.py:6-7)."""
    cds = f64(cds)
    if cds.ndim != 3 or cds.shape[1:] != (3, 3):
        raise ValueError("water_pot expects an (N,3,3) array ordered H,H,O")
    v = np.empty(len(cds))
    check(lib.pvd_pes_h2o(ptr(cds), len(cds), ptr(v)))
    return v


def pes_h2o_params():
    c, s = np.empty(245), np.empty(8)
    check(lib.pvd_pes_h2o_params(ptr(c), ptr(s)))
    return c, s


def pes_harmonic(cds, k):
    """v = sum_c k[c] x[c]^2 (harmonicOscillator1D.py:13-17 with k = (0.5*m)*omega**2)."""
    cds = f64(cds)
    n = len(cds)
    x = cds.reshape(n, -1)
    k = f64(np.broadcast_to(np.asarray(k, dtype=np.float64), (x.shape[1],)))
    v = np.empty(n)
    check(lib.pvd_pes_harmonic(ptr(x), n, x.shape[1], ptr(k), ptr(v)))
    return v


def pes_morse1d(x, de, alpha):
    x = f64(x).reshape(-1)
    v = np.empty(len(x))
    check(lib.pvd_pes_morse1d(ptr(x), len(x), float(de), float(alpha), ptr(v)))
    return v


# ----------------------------------------------------------------------------- random displacement
def displace(cds, sigmas, seed, step=0, rng_mode=_capi.RNG_DEFAULT):
    """move_randomly (pyvibdmc.py:540-547) on the GPU; returns the displaced copy."""
    out = f64(cds).copy()
    n, a, d = out.shape
    sig = f64(np.asarray(sigmas).reshape(-1))
    check(lib.pvd_displace(ptr(out), n, a, d, ptr(sig), int(seed), int(step), int(rng_mode)))
    return out


def normals(n, ncomp, seed, step=0, rng_mode=_capi.RNG_DEFAULT):
    z = np.empty((n, ncomp))
    check(lib.pvd_normals(ptr(z), n, ncomp, int(seed), int(step), int(rng_mode)))
    return z


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    o = np.zeros(4, dtype=np.uint32)
    check(lib.pvd_philox_kat(ptr(c), ptr(k), ptr(o)))
    return o


# ----------------------------------------------------------------------------- weighting
def branch_discrete(v, vref, dt, u, n0):
    """Discrete birth_or_death (pyvibdmc.py:391-431) with injected uniforms.
    Returns counts, walker_idx (== np.repeat(arange(n), counts)), births, deaths, final_pop."""
    v, u = f64(v), f64(u)
    n = len(v)
    counts = np.empty(n, dtype=np.int32)
    cap = int(1.5 * max(n, n0)) + 64
    idx = np.empty(cap, dtype=np.int64)
    st = np.zeros(3, dtype=np.int64)
    check(lib.pvd_branch_discrete(ptr(v), n, float(vref), float(dt), ptr(u), int(n0), ptr(counts), ptr(idx), cap, ptr(st)))
    return counts.astype(np.int64), idx[:st[2]].copy(), int(st[0]), int(st[1]), int(st[2])


def branch_continuous(w, v, vref, dt, lower, upper=None):
    """Continuous birth_or_death + _branch (pyvibdmc.py:432-454, 340-356).
    Returns w_out, src, n_branched, max_w, min_w."""
    w = f64(w).copy()
    v = f64(v)
    src = np.empty(len(w), dtype=np.int64)
    st = np.zeros(3)
    check(lib.pvd_branch_continuous(ptr(w), ptr(v), len(w), float(vref), float(dt), float(lower),
                                    float("nan") if upper is None else float(upper), ptr(src), ptr(st)))
    return w, src, int(st[0]), float(st[1]), float(st[2])


def calc_vref(v, n0, alpha, wts=None):
    v = f64(v)
    w = None if wts is None else f64(wts)
    out = C.c_double(0)
    check(lib.pvd_calc_vref(ptr(v), ptr(w), len(v), int(n0), float(alpha), C.byref(out)))
    return out.value


def desc_wts(who_from, n_parent, wts=None):
    who = np.ascontiguousarray(who_from, dtype=np.int64)
    w = None if wts is None else f64(wts)
    out = np.zeros(int(n_parent))
    check(lib.pvd_desc_wts(ptr(who), ptr(w), len(who), int(n_parent), ptr(out)))
    return out


# ----------------------------------------------------------------------------- device arrays for plug-ins
class DeviceArray:
    """A float64 C-ordered array in HBM owned by a simulation handle, handed to device-tensor plug-ins.  It speaks
    `__cuda_array_interface__` (version 3), so `torch.as_tensor(a, device='cuda')`, `cupy.asarray(a)` and numba see the
    memory without a copy; `.torch()` does the former."""

    def __init__(self, ptr_, shape, device=0):
        self.ptr, self.shape, self.device = int(ptr_ or 0), tuple(int(x) for x in shape), int(device)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f8", "data": (self.ptr, False), "version": 3, "strides": None, "stream": None}

    def torch(self):
        import torch
        return torch.as_tensor(self, device=torch.device("cuda", self.device))

    def __len__(self):
        return self.shape[0]


def device_pointer(arr):
    """(pointer, number of float64 elements) of a device array: anything with `__cuda_array_interface__` (torch, CuPy,
    numba, DeviceArray) or `__dlpack__`; float64, contiguous."""
    if not hasattr(arr, "__cuda_array_interface__") and hasattr(arr, "__dlpack__"):
        import torch
        arr = torch.utils.dlpack.from_dlpack(arr)
    if hasattr(arr, "is_cuda"):                  # a torch tensor: make the dtype / layout right without leaving the device
        import torch
        if not arr.is_cuda:
            raise ValueError("a device-tensor potential must return its energies on the GPU")
        arr = arr.detach().to(torch.float64).contiguous().reshape(-1)
        device_pointer._keep = arr               # keep the converted tensor alive until the C call has copied it
    cai = arr.__cuda_array_interface__
    if cai["typestr"] != "<f8":
        raise ValueError("device energies must be float64")
    if cai.get("strides") is not None:
        item, expect = 8, []
        for d in reversed(cai["shape"]):
            expect.append(item)
            item *= d
        if tuple(cai["strides"]) != tuple(reversed(expect)):
            raise ValueError("device energies must be contiguous")
    n = 1
    for d in cai["shape"]:
        n *= int(d)
    return int(cai["data"][0]), n


# ----------------------------------------------------------------------------- importance sampling
def trial_drift(trial, cds, table, ntab=None):
    """ImpSamp.drift for a built-in trial wfn: returns (grad psi/psi, psi, d2psi/psi).
    table: HARM1D -> [alpha]; H2O_FD -> [grid(ntab) | psi(ntab) | alpha_theta, theta_eq];
    H2O_AN -> [grid | psi | psi' | psi'' | alpha_theta, theta_eq]."""
    cds = f64(cds)
    n, a, d = cds.shape
    table = f64(table).reshape(-1)
    psi, dlog, d2 = np.empty(n), np.empty((n, a, d)), np.empty((n, a, d))
    if ntab is None:
        ntab = ((table.size - 2) // 2 if trial == _capi.TRIAL_H2O_FD else (table.size - 2) // 4 if trial == _capi.TRIAL_H2O_AN
                else table.size)
    check(lib.pvd_trial_drift(int(trial), ptr(cds), n, a, d, ptr(table), int(ntab), ptr(psi), ptr(dlog), ptr(d2)))
    return dlog, psi, d2


def metropolis(x, y, fx, fy, psi_x, psi_y, sigma, inv_mass, dt):
    x, y, fx, fy, psi_x, psi_y = map(f64, (x, y, fx, fy, psi_x, psi_y))
    n, a, d = x.shape
    sigma, inv_mass = f64(sigma), f64(inv_mass)
    acc = np.empty(n)
    check(lib.pvd_metropolis(ptr(x), ptr(y), ptr(fx), ptr(fy), ptr(psi_x), ptr(psi_y), n, a, d, ptr(sigma),
                             ptr(inv_mass), float(dt), ptr(acc)))
    return acc


def local_kin(d2, inv_mass):
    d2 = f64(d2)
    n, a, d = d2.shape
    inv_mass = f64(inv_mass)
    ke = np.empty(n)
    check(lib.pvd_local_kin(ptr(d2), n, a, d, ptr(inv_mass), ptr(ke)))
    return ke


# ----------------------------------------------------------------------------- NN potential
NN_PATHS = {"tcgen05": 0, "tcgen05_one_tile": 1, "cuda_cores": 2}


def nn_config(path=None, terms=None, threads=None):
    """Which kernel evaluates the NN PES (process-wide): path 'tcgen05' (default) | 'tcgen05_one_tile' | 'cuda_cores' (float32 FMA,
    float32-accurate cross-check); terms 3 | 4 cross terms of the fp16 split; threads 512 | 1024."""
    check(lib.pvd_nn_config(-1 if path is None else NN_PATHS.get(path, path), -1 if terms is None else int(terms),
                            -1 if threads is None else int(threads)))


def nn_h4o2_set_weights(packed):
    p = np.ascontiguousarray(packed, dtype=np.float32)
    check(lib.pvd_nn_h4o2_set_weights(ptr(p), p.size))


def nn_h4o2(cds):
    cds = f64(cds)
    v = np.empty(len(cds))
    check(lib.pvd_nn_h4o2(ptr(cds), len(cds), ptr(v)))
    return v


def coulomb_descriptor(cds, zs):
    cds = f64(cds)
    n, a, _ = cds.shape
    zs = f64(zs)
    out = np.empty((n, a * (a - 1) // 2))
    check(lib.pvd_coulomb_descriptor(ptr(cds), n, a, ptr(zs), ptr(out)))
    return out


# ----------------------------------------------------------------------------- resident simulation
class DeviceSim:
    """HBM-resident walker ensemble + per-step kernels (C handle pvd_sim)."""

    def __init__(self, natoms, ndim, masses, num_walkers, delta_t, potential, weighting="discrete", alpha=None,
                 capacity=None, seed=0, rng_mode=_capi.RNG_DEFAULT, trial=_capi.TRIAL_NONE, pot_params=None,
                 thresh_lower=None, thresh_upper=None, device=0, rank=0, world_size=1, stats_ring=1 << 16,
                 imp_variant=_capi.IMP_STANDARD):
        cfg = _capi.PvdConfig()
        if not (1 <= int(natoms) <= _capi.MAX_ATOMS) or not (1 <= int(ndim) <= 3) or int(natoms) * int(ndim) > _capi.MAX_COMP:
            raise ValueError(f"DeviceSim: {natoms} atoms x {ndim} dimensions is outside what the device path holds "
                             f"(1..{_capi.MAX_ATOMS} atoms, 1..3 dimensions, at most {_capi.MAX_COMP} components)")
        if len(np.asarray(masses).reshape(-1)) < int(natoms):
            raise ValueError(f"DeviceSim: {natoms} atoms need {natoms} masses")
        cfg.natoms, cfg.ndim = int(natoms), int(ndim)
        cfg.weighting = _capi.WEIGHT_CONTINUOUS if weighting == "continuous" else _capi.WEIGHT_DISCRETE
        cfg.potential, cfg.trial, cfg.rng_mode = int(potential), int(trial), int(rng_mode)
        cfg.device, cfg.rank, cfg.world_size = int(device), int(rank), int(world_size)
        cfg.num_walkers = int(num_walkers)
        local = -(-int(num_walkers) // int(world_size))
        cfg.capacity = int(capacity) if capacity else int(1.5 * local) + 1024
        cfg.delta_t = float(delta_t)
        cfg.alpha = 1.0 / (2.0 * float(delta_t)) if alpha is None else float(alpha)
        cfg.thresh_lower = 1.0 / num_walkers if thresh_lower is None else float(thresh_lower)
        cfg.thresh_upper = float("nan") if thresh_upper is None else float(thresh_upper)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        m = np.asarray(masses, dtype=np.float64).reshape(-1)
        for i in range(natoms):
            cfg.masses[i] = m[i]
        if pot_params is not None:
            for i, p in enumerate(np.asarray(pot_params, dtype=np.float64).reshape(-1)):
                cfg.pot_params[i] = p
        cfg.stats_ring = int(stats_ring)
        cfg.imp_variant = int(imp_variant)
        self.cfg = cfg
        self.natoms, self.ndim = int(natoms), int(ndim)
        self.capacity = cfg.capacity
        self._h = C.c_void_p(None)
        check(lib.pvd_sim_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.pvd_sim_destroy(self._h)
            self._h = C.c_void_p(None)

    __del__ = close

    # -- set-up
    def set_stream(self, cuda_stream):
        check(lib.pvd_sim_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def upload(self, coords, wts=None):
        coords = f64(coords)
        w = None if wts is None else f64(wts)
        check(lib.pvd_sim_upload(self._h, ptr(coords), len(coords), ptr(w)))

    def set_pots(self, v):
        v = f64(v)
        check(lib.pvd_sim_set_pots(self._h, ptr(v), len(v)))

    def init_finalize(self):
        check(lib.pvd_sim_init_finalize(self._h))

    def set_trial_table(self, table, ntab=None):
        t = f64(table).reshape(-1)
        if ntab is None:
            ntab = ((t.size - 2) // 2 if self.cfg.trial == _capi.TRIAL_H2O_FD else
                    (t.size - 2) // 4 if self.cfg.trial == _capi.TRIAL_H2O_AN else t.size)
        check(lib.pvd_sim_set_trial_table(self._h, ptr(t), int(ntab)))

    def set_nn_weights(self, packed):
        p = np.ascontiguousarray(packed, dtype=np.float32)
        check(lib.pvd_sim_set_nn_weights(self._h, ptr(p), p.size))

    # -- stepping
    def run(self, nsteps, branch_every=1):
        check(lib.pvd_sim_run(self._h, int(nsteps), int(branch_every)))

    def set_resident(self, enable=True):
        """How run() segments execute (same numbers in every mode): True/1 chosen by ensemble size (default), False/0 one
        self-compacting launch per time step, 2 the resident multi-step kernel always, 3 deferred-compaction steps always."""
        check(lib.pvd_sim_set_resident(self._h, int(enable)))

    def step_injected(self, disp, u_branch=None, u_metro=None):
        disp = f64(disp)
        ub = None if u_branch is None else f64(u_branch)
        um = None if u_metro is None else f64(u_metro)
        check(lib.pvd_sim_step_injected(self._h, ptr(disp), ptr(ub), ptr(um)))

    def ext_move(self):
        out = np.empty((self.capacity, self.natoms, self.ndim))
        n = C.c_int64(0)
        check(lib.pvd_sim_ext_move(self._h, ptr(out), C.byref(n)))
        return out[:n.value]

    def ext_finish(self, v, do_branch=True):
        v = f64(v)
        check(lib.pvd_sim_ext_finish(self._h, ptr(v), len(v), 1 if do_branch else 0))

    # ---- walkers to / from another shard without the host
    def export_tail_device(self, count):
        """Pack the last `count` walkers (coordinates, V, weight, who_from, importance-sampling companions) into one device
        buffer and drop them from this shard; returns a DeviceArray (count, ncols)."""
        p, nc = C.c_void_p(None), C.c_int32(0)
        check(lib.pvd_sim_export_tail_device(self._h, int(count), C.byref(p), C.byref(nc)))
        return DeviceArray(p.value, (int(count), nc.value), self.cfg.device)

    def import_device(self, payload):
        """Append walkers packed by export_tail_device on another shard (any device array (count, ncols))."""
        ptr_, n = device_pointer(payload)
        cols = int(payload.shape[1])
        check(lib.pvd_sim_import_device(self._h, n // cols, C.c_void_p(ptr_), cols))

    # ---- device-tensor potential plug-in: coordinates and energies stay in HBM
    def ext_move_device(self):
        """Move the walkers; returns a DeviceArray (n, atoms, dims) float64 view of them in HBM (valid until the next call)."""
        p, n = C.c_void_p(None), C.c_int64(0)
        check(lib.pvd_sim_ext_move_device(self._h, C.byref(p), C.byref(n)))
        return DeviceArray(p.value, (n.value, self.natoms, self.ndim), self.cfg.device)

    def coords_device(self):
        p, n = C.c_void_p(None), C.c_int64(0)
        check(lib.pvd_sim_coords_device(self._h, C.byref(p), C.byref(n)))
        return DeviceArray(p.value, (n.value, self.natoms, self.ndim), self.cfg.device)

    def ext_finish_device(self, v_dev, do_branch=True):
        ptr_, n = device_pointer(v_dev)
        check(lib.pvd_sim_ext_finish_device(self._h, C.c_void_p(ptr_), n, 1 if do_branch else 0))

    def set_pots_device(self, v_dev):
        ptr_, n = device_pointer(v_dev)
        check(lib.pvd_sim_set_pots_device(self._h, C.c_void_p(ptr_), n))

    # ---- importance sampling with a user trial wave function (trial=_capi.TRIAL_EXTERNAL): the host supplies psi and derivatives
    def imp_ext_init(self, fx, psi, sec, v=None):
        """Drift terms (and, for an external potential, V) of the start ensemble: first-step exception (pyvibdmc.py:760-769)."""
        fx, psi, sec = f64(fx), f64(psi), f64(sec)
        vv = None if v is None else f64(v)
        check(lib.pvd_sim_imp_ext_init(self._h, ptr(fx), ptr(psi), ptr(sec), ptr(vv)))

    def imp_ext_propose(self, disp=None):
        """coords + disps + (1/m) f_x dt for every walker (pyvibdmc.py:593), on the host, for impsamp.drift.
        disp: injected displacements (parity replays) instead of the Philox normals."""
        out = np.empty((self.capacity, self.natoms, self.ndim))
        n = C.c_int64(0)
        dd = None if disp is None else f64(disp)
        check(lib.pvd_sim_imp_ext_propose(self._h, ptr(dd), ptr(out), C.byref(n)))
        return out[:n.value]

    def imp_ext_accept(self, fy, psi_y, sec_y, u_metro=None):
        """Metropolis step with the drift terms of the displaced walkers (pyvibdmc.py:597-612)."""
        fy, psi_y, sec_y = f64(fy), f64(psi_y), f64(sec_y)
        um = None if u_metro is None else f64(u_metro)
        check(lib.pvd_sim_imp_ext_accept(self._h, ptr(fy), ptr(psi_y), ptr(sec_y), len(psi_y), ptr(um)))

    def imp_ext_finish(self, v=None, do_branch=True, u_branch=None):
        """E_L = V + T_L, weighting / branching, Vref.  v=None: the configured built-in potential, on the GPU."""
        vv = None if v is None else f64(v)
        ub = None if u_branch is None else f64(u_branch)
        check(lib.pvd_sim_imp_ext_finish(self._h, ptr(vv), 0 if vv is None else len(vv), 1 if do_branch else 0, ptr(ub)))

    def sums_ptr(self):
        p = C.c_void_p(None)
        check(lib.pvd_sim_sums_ptr(self._h, C.byref(p)))
        return p.value

    def set_sums_ptr(self, device_ptr):
        check(lib.pvd_sim_set_sums_ptr(self._h, C.c_void_p(int(device_ptr) if device_ptr else None)))

    def step_local(self, do_branch=1):
        check(lib.pvd_sim_step_local(self._h, int(do_branch)))

    def step_finalize(self):
        check(lib.pvd_sim_step_finalize(self._h))

    def mailbox_handle(self):
        buf = C.create_string_buffer(64)
        check(lib.pvd_sim_mailbox_handle(self._h, buf))
        return buf.raw

    def mailbox_connect(self, handles):
        blob = b"".join(handles)
        check(lib.pvd_sim_mailbox_connect(self._h, C.c_char_p(blob), len(handles)))

    def run_mailbox(self, nsteps, branch_every=1):
        check(lib.pvd_sim_run_mailbox(self._h, int(nsteps), int(branch_every)))

    def imp_move_local(self):
        check(lib.pvd_sim_imp_move_local(self._h))

    def imp_branch_local(self, do_branch=1):
        check(lib.pvd_sim_imp_branch_local(self._h, int(do_branch)))

    # -- descendant weighting
    def dw_begin(self, global_offset=0):
        check(lib.pvd_sim_dw_begin(self._h, int(global_offset)))

    def dw_resume(self, who_from, parent, parent_wts=None):
        """Re-open a descendant-weighting window from checkpointed who_from / parent arrays (dmc_restart inside a window)."""
        who = np.ascontiguousarray(who_from, dtype=np.int64)
        par = f64(parent)
        pw = None if parent_wts is None else f64(parent_wts)
        check(lib.pvd_sim_dw_resume(self._h, ptr(who), len(who), ptr(par), ptr(pw), len(par)))

    def dw_end_begin(self, n_parent):
        """Close the window and start the asynchronous transfer of (descendant weights, parent ensemble) to the host."""
        check(lib.pvd_sim_dw_end_begin(self._h, int(n_parent)))
        self._dump_n = int(n_parent)

    def dw_end_wait(self):
        """-> (desc_wts, parent coords, parent weights or None) of the dump started by dw_end_begin."""
        n = self._dump_n
        desc, par = np.empty(n), np.empty((n, self.natoms, self.ndim))
        pw = np.empty(n) if self.cfg.weighting == _capi.WEIGHT_CONTINUOUS else None
        check(lib.pvd_sim_dw_end_wait(self._h, ptr(desc), ptr(par), ptr(pw), n))
        return desc, par, pw

    def dw_end(self, n_parent):
        out = np.zeros(int(n_parent))
        check(lib.pvd_sim_dw_end(self._h, ptr(out), int(n_parent)))
        return out

    def dw_peek(self, n_parent):
        out = np.zeros(int(n_parent))
        check(lib.pvd_sim_dw_peek(self._h, ptr(out), int(n_parent)))
        return out

    def set_masses(self, masses):
        m = f64(masses).reshape(-1)
        check(lib.pvd_sim_set_masses(self._h, ptr(m), len(m)))

    def dw_parent(self):
        n = C.c_int64(0)
        check(lib.pvd_sim_dw_parent(self._h, None, None, C.byref(n)))
        xyz = np.empty((n.value, self.natoms, self.ndim))
        w = np.empty(n.value) if self.cfg.weighting == _capi.WEIGHT_CONTINUOUS else None
        check(lib.pvd_sim_dw_parent(self._h, ptr(xyz), ptr(w), C.byref(n)))
        return xyz, w

    # -- queries
    def sync(self):
        check(lib.pvd_sim_sync(self._h))

    def state(self, raise_on_error=True):
        n, vref, step, err = C.c_int64(0), C.c_double(0), C.c_int64(0), C.c_int32(0)
        rc = lib.pvd_sim_state(self._h, C.byref(n), C.byref(vref), C.byref(step), C.byref(err))
        if raise_on_error:
            check(rc)
        return dict(n=n.value, vref=vref.value, step=step.value, err=err.value)

    def download(self, who_from=False, out=None):
        """Walkers, energies (weights, who_from) -> host.  `out`: optional {'coords': (cap,A,D) f64, 'pots': (cap,) f64} host
        buffers to fill (e.g. views of pinned memory) instead of allocating fresh arrays."""
        n = C.c_int64(0)
        check(lib.pvd_sim_download(self._h, None, None, None, None, 0, C.byref(n)))
        nn = n.value
        if out is not None:
            xyz, pots = out["coords"][:nn], out["pots"][:nn]
            assert xyz.flags.c_contiguous and pots.flags.c_contiguous and xyz.dtype == np.float64 and pots.dtype == np.float64
        else:
            xyz, pots = np.empty((nn, self.natoms, self.ndim)), np.empty(nn)
        w = np.empty(nn) if self.cfg.weighting == _capi.WEIGHT_CONTINUOUS else None
        who = np.empty(nn, dtype=np.int64) if who_from else None
        check(lib.pvd_sim_download(self._h, ptr(xyz), ptr(pots), ptr(w), ptr(who), nn, C.byref(n)))
        return dict(coords=xyz, pots=pots, wts=w, who_from=who)

    def snapshot_begin(self):
        """Start an asynchronous copy of the ensemble as it is now (stream order) to pinned host memory on a side stream."""
        check(lib.pvd_sim_snapshot_begin(self._h))

    def snapshot_wait(self, who_from=False):
        """Collect the snapshot started by snapshot_begin (does not synchronise the compute stream)."""
        n, vref = C.c_int64(0), C.c_double(0)
        check(lib.pvd_sim_snapshot_wait(self._h, None, None, None, None, 0, C.byref(n), C.byref(vref)))
        nn = n.value
        xyz, pots = np.empty((nn, self.natoms, self.ndim)), np.empty(nn)
        w = np.empty(nn) if self.cfg.weighting == _capi.WEIGHT_CONTINUOUS else None
        who = np.empty(nn, dtype=np.int64) if who_from else None
        check(lib.pvd_sim_snapshot_wait(self._h, ptr(xyz), ptr(pots), ptr(w), ptr(who), nn, C.byref(n), C.byref(vref)))
        return dict(coords=xyz, pots=pots, wts=w, who_from=who, vref=vref.value)

    def stats(self, first_step, count):
        out = np.zeros(int(count), dtype=_capi.STATS_DTYPE)
        if count:
            check(lib.pvd_sim_stats(self._h, int(first_step), int(count), ptr(out)))
        return out

    def last_run_ms(self):
        v = C.c_double(0)
        check(lib.pvd_sim_last_run_ms(self._h, C.byref(v)))
        return v.value

    def download_imp(self):
        st = self.state()
        n = st["n"]
        fx = np.empty((n, self.natoms, self.ndim))
        psi = np.empty(n)
        sec = np.empty((n, self.natoms, self.ndim))
        check(lib.pvd_sim_download_imp(self._h, ptr(fx), ptr(psi), ptr(sec), n))
        return fx, psi, sec
