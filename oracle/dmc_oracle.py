"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by pyvibdmc_b200).

CPU restatement (NumPy) of the reference's per-timestep walker-propagation path.  Each
function cites the reference lines it follows (paths relative to /root/reference/pyvibdmc).
Pinned by tests/test_oracle_*.py against fixtures generated from the unmodified reference
(tests/golden/make_golden.py) and against the shipped tutorial data.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
WAVENUMBERS = 4.556335281212229e-6          # simulation_utilities/Constants.py:66
ANGSTROMS = 1 / 0.529177                    # Constants.py:67
AMU = 1.0 / 6.02213670000e23 / 9.10938970000e-28   # Constants.py:68
MASS_AMU = {'H': 1.00782503, 'D': 2.01410178, 'O': 15.99491462}     # Constants.py:3-4 (subset)


def mass(atom):                              # Constants.py:86-99
    return MASS_AMU[atom] * AMU


def reduced_mass(pair):                      # Constants.py:101-119
    a, b = pair.split('-')
    m1, m2 = mass(a), mass(b)
    return m1 * m2 / (m1 + m2)


# --------------------------------------------------------------------------- C PES oracle
_lib = None


def build_c_oracle():
    """Compile oracle/ps_h2o_oracle.c (gcc) if the shared object is missing/out of date."""
    so = os.path.join(_HERE, "_build", "libpvd_oracle.so")
    src = os.path.join(_HERE, "ps_h2o_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _clib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c_oracle())
        _lib.oracle_ps_h2o.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        _lib.oracle_ps_folded.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def water_pot(cds):
    """sample_potentials/FortPots/Partridge_Schwenke_H2O/h2o_potential.py:6-7 -> calc_h2o_pot.f + h2opes_v2.f
    (C restatement in ps_h2o_oracle.c).  cds: (N,3,3) bohr, atoms [H,H,O] -> (N,) Hartree."""
    cds = np.ascontiguousarray(cds, dtype=np.float64)
    v = np.empty(len(cds))
    _clib().oracle_ps_h2o(cds.ctypes.data, len(cds), v.ctypes.data)
    return v


def ps_folded_params():
    c = np.empty(245)
    s = np.empty(8)
    _clib().oracle_ps_folded(c.ctypes.data, s.ctypes.data)
    return c, s


# --------------------------------------------------------------------------- 1-D sample potentials / trial wfns
def oh_stretch_harm(cds):                    # PythonPots/harmonicOscillator1D.py:13-17
    m = reduced_mass('O-H')
    omega = 3700. * WAVENUMBERS
    return np.squeeze(0.5 * m * omega ** 2 * cds ** 2)


def oh_stretch_morse(disp):                  # PythonPots/morse_osc_1d.py:4-12
    m = reduced_mass('O-H')
    omega = 3704.5 * WAVENUMBERS
    omega_x = 75.3 * WAVENUMBERS
    de = omega ** 2 / (4 * omega_x)
    alpha = np.sqrt(m * (omega ** 2.) / 2. / de)
    return (de * np.square(1 - np.exp(-alpha * disp))).squeeze()


def trial_harm(x):                           # PythonPots/harm_trial_wfn.py:6-16
    alpha = reduced_mass('O-H') * 3700 * WAVENUMBERS
    return ((alpha / np.pi) ** 0.25 * np.exp(-alpha * x ** 2 / 2)).squeeze()


def harm_derivs(x):                          # harm_trial_wfn.py:19-40 -> (psi'/psi, psi''/psi)
    alpha = reduced_mass('O-H') * 3700 * WAVENUMBERS
    pref = (alpha / np.pi) ** 0.25
    e = np.exp(-alpha * x ** 2 / 2)
    trl = trial_harm(x)[:, None, None]
    return pref * (-alpha * x) * e / trl, pref * (alpha ** 2 * x ** 2 - alpha) * e / trl


# --------------------------------------------------------------------------- water product trial wfn
class WaterTrial:
    """FortPots/Partridge_Schwenke_H2O/call_trl_h2o.py:7-78 with kwargs
    {'dists':[[0,2],[2,1]],'angs':[[0,2,1]]} (tests/test_imp_samp.py:143-144)."""
    r1_eq = 0.95784 * ANGSTROMS
    r2_eq = 0.95783997 * ANGSTROMS
    theta_eq = np.deg2rad(104.5080029)
    theta_freq = 1668.4590610594878 * WAVENUMBERS

    def __init__(self, table):               # table rows: grid, psi  (free_oh_wvfn_dense.npy[0:2])
        self.grid, self.wfn = np.asarray(table[0]), np.asarray(table[1])
        if len(table) >= 4:                  # rows psi', psi'' (analytic derivatives, call_trl_h2o.py:17-18)
            self.dwfn, self.d2wfn = np.asarray(table[2]), np.asarray(table[3])
        inv_mh, inv_mo = 1 / mass('H'), 1 / mass('O')
        g = inv_mh / self.r1_eq ** 2 + inv_mh / self.r2_eq ** 2 + inv_mo * (
            1 / self.r1_eq ** 2 + 1 / self.r2_eq ** 2 - 2 * np.cos(self.theta_eq) / (self.r1_eq * self.r2_eq))
        self.alpha = self.theta_freq / g     # :38-46

    def __call__(self, cds):
        cds = np.asarray(cds)
        r1 = np.linalg.norm(cds[:, 0] - cds[:, 2], axis=1)        # dists [0,2]
        r2 = np.linalg.norm(cds[:, 2] - cds[:, 1], axis=1)        # dists [2,1]
        v1, v2 = cds[:, 0] - cds[:, 2], cds[:, 1] - cds[:, 2]     # angs [0,2,1]
        th = np.arccos(np.einsum('ij,ij->i', v1, v2) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1)))
        ang = (self.alpha / np.pi) ** 0.25 * np.exp(-self.alpha * (th - self.theta_eq) ** 2 / 2)
        return np.interp(r1, self.grid, self.wfn) * np.interp(r2, self.grid, self.wfn) * ang


    def derivs_analytic(self, cds):
        """(grad psi / psi, d2 psi / dx2 / psi): call_trl_h2o.py:101-149 (dpsi_dx) with the chain-rule formulas of
        imp_samp_helper.py:10-209 written out for dists [[0,2],[2,1]], angs [[0,2,1]].  Needs the 4-row table."""
        cds = np.asarray(cds, dtype=np.float64)
        H1, H2, O = cds[:, 0], cds[:, 1], cds[:, 2]
        r1 = np.linalg.norm(H1 - O, axis=1)
        r2 = np.linalg.norm(O - H2, axis=1)
        v1, v2 = H1 - O, H2 - O
        th = np.arccos(np.einsum('ij,ij->i', v1, v2) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1)))
        cth = np.cos(th)
        pref, al, x = (self.alpha / np.pi) ** 0.25, self.alpha, th - self.theta_eq
        e = np.exp(-al * x ** 2 / 2)
        trl = np.array([np.interp(r1, self.grid, self.wfn), np.interp(r2, self.grid, self.wfn), pref * e])
        dpsi = np.array([np.interp(r1, self.grid, self.dwfn), np.interp(r2, self.grid, self.dwfn), pref * (-al * x) * e]) / trl
        d2psi = np.array([np.interp(r1, self.grid, self.d2wfn), np.interp(r2, self.grid, self.d2wfn),
                          pref * (al ** 2 * x ** 2 - al) * e]) / trl
        n = len(cds)
        dint, d2int = np.zeros((3, n, 3, 3)), np.zeros((3, n, 3, 3))
        # bond lengths: dr/dx = +-(xa - xb)/r ; d2r/dx2 = 1/r - (1/r) (dr/dx)^2        (:40-66)
        for q, (a, b, r) in enumerate(((0, 2, r1), (2, 1, r2))):
            diff = cds[:, a] - cds[:, b]
            dint[q][:, a] = diff / r[:, None]
            dint[q][:, b] = -1 * diff / r[:, None]
            for atom in (a, b):
                d2int[q][:, atom] = (1 / r)[:, None] - (1 / r)[:, None] * dint[q][:, atom] ** 2
        dra, drc, d2ra, d2rc = dint[0], dint[1], d2int[0], d2int[1]
        rab, rcb = np.linalg.norm(O - H1, axis=1), np.linalg.norm(O - H2, axis=1)
        # d cos(theta) / dx   (:68-107), vertex = O
        dc = np.zeros((n, 3, 3))
        a1, a2, a3 = H2 - O, H1 - O, 2 * O - H1 - H2
        dc[:, 0] = a1 / (rab * rcb)[:, None] - (cth / rab)[:, None] * dra[:, 0]
        dc[:, 1] = a2 / (rab * rcb)[:, None] - (cth / rcb)[:, None] * drc[:, 1]
        dc[:, 2] = a3 / (rab * rcb)[:, None] - (cth / rab)[:, None] * dra[:, 2] - (cth / rcb)[:, None] * drc[:, 2]
        # d2 cos(theta) / dx2 (:109-160)
        d2c = np.zeros((n, 3, 3))
        d2c[:, 0] = ((-2 * a1) / (rab ** 2 * rcb)[:, None] * dra[:, 0] + (2 * cth / rab ** 2)[:, None] * dra[:, 0] ** 2
                     + (-1 * cth / rab)[:, None] * d2ra[:, 0])
        d2c[:, 1] = ((-2 * a2) / (rab * rcb ** 2)[:, None] * drc[:, 1] + (2 * cth / rcb ** 2)[:, None] * drc[:, 1] ** 2
                     + (-1 * cth / rcb)[:, None] * d2rc[:, 1])
        d2c[:, 2] = (2 / (rab * rcb)[:, None] + (-1 * cth / rab)[:, None] * d2ra[:, 2] + (-1 * cth / rcb)[:, None] * d2rc[:, 2]
                     + (2 * cth / rab ** 2)[:, None] * dra[:, 2] ** 2 + (2 * cth / rcb ** 2)[:, None] * drc[:, 2] ** 2
                     + (-2 * a3 / (rab ** 2 * rcb)[:, None]) * dra[:, 2] + (-2 * a3 / (rab * rcb ** 2)[:, None]) * drc[:, 2]
                     + (2 * cth / (rab * rcb))[:, None] * dra[:, 2] * drc[:, 2])
        # theta from cos(theta) (:162-209)
        dth_dc = -1 / np.sqrt(1 - cth ** 2)
        d2th_dc2 = -1 * cth / ((1 - cth ** 2) ** 1.5)
        dint[2] = dth_dc[:, None, None] * dc
        d2int[2] = dc ** 2 * d2th_dc2[:, None, None] + d2c * dth_dc[:, None, None]
        # chain rule for a direct-product wave function (:10-38)
        w1, w2 = dpsi[:, :, None, None], d2psi[:, :, None, None]
        dp = (dint * w1).sum(axis=0)
        t3 = 2 * sum(dint[q] * dint[(q + 1) % 3] * (w1[q] * w1[(q + 1) % 3]) for q in range(3))
        d2p = (dint ** 2 * w2).sum(axis=0) + (d2int * w1).sum(axis=0) + t3
        return dp, d2p


def drift_analytic(cds, trial):
    """ImpSamp.drift with a derivative function (imp_samp_manager.py:206-214): the plug-in already returns ratios."""
    d1, d2 = trial.derivs_analytic(cds)
    return d1, trial(np.asarray(cds)), d2


# --------------------------------------------------------------------------- importance-sampling math
def finite_diff(cds, trial):                 # simulation_utilities/imp_samp.py:56-76
    dx = 0.001
    cds = np.array(cds, dtype=np.float64)
    first, sec = np.zeros(cds.shape), np.zeros(cds.shape)
    psi0 = trial(cds)
    for a in range(cds.shape[1]):
        for d in range(cds.shape[2]):
            cds[:, a, d] -= dx
            pm = trial(cds)
            cds[:, a, d] += 2. * dx
            pp = trial(cds)
            cds[:, a, d] -= dx
            first[:, a, d] = (pp - pm) / (2 * dx)
            sec[:, a, d] = (pm - 2. * psi0 + pp) / dx ** 2
    return first, sec, psi0


def drift_fd(cds, trial):
    """ImpSamp.drift (imp_samp.py:21-26) through ImpSampManager_NoMP.call_derivs with
    deriv_function=None (imp_samp_manager.py:197-205): FD derivatives divided by psi."""
    psi = trial(np.asarray(cds))
    f, s, p0 = finite_diff(cds, trial)
    return f / p0[:, None, None], psi, s / p0[:, None, None]


def metropolis(sigma_trip, trial_x, trial_y, disp_x, disp_y, d_x, d_y, dt):   # imp_samp.py:29-47
    ratio = (trial_y / trial_x) ** 2
    t1 = np.exp(-1 * (disp_x - disp_y - d_y * dt) ** 2 / (2 * sigma_trip ** 2))
    t2 = np.exp(-1 * (disp_y - disp_x - d_x * dt) ** 2 / (2 * sigma_trip ** 2))
    acc = t1 / t2
    if acc.shape[-1] == 1:
        acc = acc.squeeze() * ratio.squeeze()
    else:
        acc = np.prod(np.prod(acc, axis=1), axis=1) * ratio
    acc = np.array(acc)
    acc[np.where(trial_x * trial_y <= 0)[0]] = 0.0
    return acc


def local_kin(inv_masses_trip, sec):         # imp_samp.py:50-53
    return -0.5 * np.sum(np.sum(inv_masses_trip * sec, axis=1), axis=1)


# --------------------------------------------------------------------------- weighting
MASSIVE = "Massive walker birth or death event!!!!!!! Dying..."


def birth_or_death_discrete(v, vref, dt, u, n0):
    """pyvibdmc.py:391-431.  Returns (counts, walker_idx, births, deaths, final_pop)."""
    lo, hi = n0 - n0 * 0.5, n0 + n0 * 0.5                         # :250-251
    w = np.exp(-1. * (v - vref) * dt)                             # :393
    if not np.all(np.isfinite(w)) or np.any(w > hi + 1):          # :397-400
        raise ValueError(MASSIVE)
    counts = np.floor(w).astype(np.int64)                         # :402
    counts += u < (w - counts)                                    # :403
    deaths = int(np.count_nonzero(counts == 0))
    births = int(np.sum(np.maximum(counts - 1, 0)))
    pop = int(np.sum(counts))
    if pop < lo or pop > hi:                                      # :409-413
        raise ValueError(MASSIVE)
    return counts, np.repeat(np.arange(len(counts)), counts), births, deaths, pop


def branch_continuous(w, v, vref, dt, lower, upper=None):
    """pyvibdmc.py:433-454 + _branch :340-356.  Returns (w_out, src, n_branched, max_w, min_w);
    src[i] = index of the walker whose coords/V/who_from walker i holds afterwards."""
    w = np.array(w, dtype=np.float64) * np.exp(-1.0 * (v - vref) * dt)
    src = np.arange(len(w))

    def branch(kill):
        for k in kill:
            donor = int(np.argmax(w))
            src[k] = src[donor]
            w[donor] /= 2.0
            w[k] = w[donor]
    kill = np.where(w < lower)[0]
    branch(kill)
    nk = len(kill)
    if upper is not None:
        n_above = int(np.sum(w > upper))
        kill_up = np.argpartition(w, n_above)[:n_above]
        branch(kill_up)
        nk += len(kill_up)
    return w, src, nk, float(np.amax(w)), float(np.amin(w))


def calc_vref(v, n0, alpha, wts=None):       # pyvibdmc.py:651-661
    if wts is None:
        v_bar, correction = np.average(v), (len(v) - n0) / n0
    else:
        v_bar, correction = np.average(v, weights=wts), (np.sum(wts) - n0) / n0
    return v_bar - (alpha * correction)


def desc_wts_discrete(who_from, n_parent):   # pyvibdmc.py:667-669
    out = np.zeros(n_parent)
    uq, ct = np.unique(who_from, return_counts=True)
    out[uq] = ct
    return out


def desc_wts_continuous(wts, who_from, n_parent):   # pyvibdmc.py:670-672 (O(N) restatement of the O(N^2) loop)
    return np.bincount(who_from, weights=wts, minlength=n_parent).astype(np.float64)


# --------------------------------------------------------------------------- NN PES (descriptor + MLP)
def coulomb_descriptor(cds, zs):
    """tensorflow_descriptors/distance_descriptors.py:102-113,154-168 (unsorted, upper triangle,
    itertools.combinations order): Z_i Z_j / r_ij."""
    cds = np.asarray(cds)
    zs = np.asarray(zs, dtype=np.float64)
    pairs = list(itertools.combinations(range(cds.shape[1]), 2))
    i0 = [p[0] for p in pairs]
    i1 = [p[1] for p in pairs]
    r = np.linalg.norm(cds[:, i0] - cds[:, i1], axis=2)
    return zs[i0] * zs[i1] / r


def _left_to_right(v):
    acc = 0.0
    for x in v:
        acc = acc + x
    return acc


def _numpy_sum_order(v):
    """np.sum over a contiguous axis (numpy/core/src/umath/loops_utils.h.src, pairwise_sum): fewer than 8 summands are
    added left to right; 8 to 128 go into eight interleaved partial sums combined as ((0+1)+(2+3))+((4+5)+(6+7)), the
    n % 8 leftovers added afterwards.  Checked against np.sum on 2e5 random rows for every length up to 16."""
    n = len(v)
    if n < 8:
        return _left_to_right(v)
    r = list(v[:8])
    i = 8
    while i < n - n % 8:
        for k in range(8):
            r[k] = r[k] + v[i + k]
        i += 8
    acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for x in v[i:]:
        acc = acc + x
    return acc


def distit(cds, zs, method, eq_xyz=None, sorted_atoms=None, sorted_groups=None, full_mat=False, _r_eq=None):
    """DistIt.run (tensorflow_descriptors/distance_descriptors.py:177-213) restated walker by walker with explicit loops:
    pair distances in itertools.combinations order (:102-113), Coulomb dressing Z_i Z_j / r and diagonal 0.5 Z^2.4
    (:154-168), full matrix (:88-100), atoms of each `sorted_atoms` sub-list reordered by descending column norm
    (:115-132), whole `sorted_groups` swapped by descending sum of their column norms (:134-152), SPF 1 - r_eq / r with
    r_eq taken from the equilibrium structure sorted the same way (:72-86, 191-213).  Column norms for
    the atom sort accumulate left to right, as NumPy does for a reduction over a non-last axis; the group sort's norms and
    totals follow the memory layout of the reference's fancy-indexed temporaries (see the comment there)."""
    cds = np.asarray(cds, dtype=np.float64)
    zs = np.asarray(zs)
    method = method.lower()
    n, na = cds.shape[0], cds.shape[1]
    pairs = list(itertools.combinations(range(na), 2))
    sort = sorted_atoms is not None or sorted_groups is not None
    zf = zs.astype(np.float64) if method == 'coulomb' else None
    if method == 'coulomb':
        rest = np.ones((na, na))
        np.fill_diagonal(rest, 0.5 * zs ** 0.4)
        skel = np.outer(zs, zs) * rest
        diag = np.diag(skel)
    if method == 'spf' and _r_eq is None:
        if eq_xyz is None:
            raise ValueError("eq_xyz is not set but using spf. Fix!")
        eq = np.repeat(np.asarray(eq_xyz, dtype=np.float64)[None], 2, axis=0)     # :75: two copies (matters for :146's order)
        if sort:
            _r_eq = distit(eq, zs, 'distance', None, sorted_atoms, sorted_groups, True)[0]
        else:
            _r_eq = distit(eq, zs, 'distance')[0]
    out = []
    for w in range(n):
        x = cds[w]
        vec = np.empty(len(pairs))
        for p, (i, j) in enumerate(pairs):
            d = x[i] - x[j]
            r = np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
            vec[p] = skel[i, j] / r if method == 'coulomb' else r
        if not sort and not full_mat:
            out.append(1 - _r_eq / vec if method == 'spf' else vec)
            continue
        m = np.zeros((na, na))
        for p, (i, j) in enumerate(pairs):
            m[i, j] = m[j, i] = vec[p]
        if method == 'coulomb':
            m[np.arange(na), np.arange(na)] = diag

        def colnorm(mat):
            nrm = np.zeros(na)
            for j in range(na):
                acc = 0.0
                for i in range(na):
                    acc = acc + mat[i, j] * mat[i, j]
                nrm[j] = np.sqrt(acc)
            return nrm
        if not sort:
            out.append(m)           # :196-197: an unsorted full matrix is returned as is (for 'spf' too: plain distances)
            continue
        if sorted_atoms is not None:
            nrm = colnorm(m)
            order = sorted(range(na), key=lambda a: (-nrm[a], a))
            inds = np.zeros(na, dtype=int)
            for lst in sorted_atoms:
                members = [a for a in order if a in lst]
                for pos, a in zip(lst, members):
                    inds[pos] = a
            m = m[np.ix_(inds, inds)]
        if sorted_groups is not None:
            # :146 `sum(norm(d_mat[:, :, g], axis=1), axis=1)`: the fancy-indexed array is laid out (member, walker, row) in
            # memory, so the norm reduces its CONTIGUOUS axis (np.sum's unrolled order from 8 atoms on, unlike :125) and
            # the group total then reduces a strided one: left to right, except for a single walker, where the walker
            # axis collapses and the members are contiguous again
            nrm = [np.sqrt(_numpy_sum_order([m[i, a] * m[i, a] for i in range(na)])) for a in range(na)]
            tot = []
            for g in sorted_groups:
                tot.append(_numpy_sum_order([nrm[a] for a in g]) if n == 1 else _left_to_right([nrm[a] for a in g]))
            rank = sorted(range(len(sorted_groups)), key=lambda g: (-tot[g], g))
            inds = np.arange(na)
            flat = [a for g in sorted_groups for a in g]
            newflat = [a for g in rank for a in sorted_groups[g]]
            for pos, a in zip(flat, newflat):
                inds[pos] = a
            m = m[np.ix_(inds, inds)]
        if full_mat:
            with np.errstate(invalid='ignore', divide='ignore'):
                out.append(1 - _r_eq / m if method == 'spf' else m)
        else:
            v = np.array([m[i, j] for (i, j) in pairs])
            out.append(1 - np.array([_r_eq[i, j] for (i, j) in pairs]) / v if method == 'spf' else v)
    return np.array(out)


def nn_forward_f32(desc, weights):
    """Keras Sequential[Dense(120,swish)x3, Dense(1,relu)] in float32
    (sample_potentials/TensorflowPots/sample_h4o2_nn.h5; call_sample_model.py:4-9).
    PARITY UNPINNED at the TensorFlow boundary: TF is absent from this image and the
    reference has no test asserting predicted values (tests/test_pot.py:66-93)."""
    x = np.asarray(desc, dtype=np.float32)
    for li, (k, b) in enumerate(weights):
        z = x @ k.astype(np.float32) + b.astype(np.float32)
        if li < len(weights) - 1:
            sig = np.where(z >= 0, 1 / (1 + np.exp(-np.abs(z))), np.exp(-np.abs(z)) / (1 + np.exp(-np.abs(z))))
            x = (z * sig).astype(np.float32)
        else:
            x = np.maximum(z, 0).astype(np.float32)
    return x.reshape(-1).astype(np.float64) * WAVENUMBERS


def nn_forward_f64(desc, weights):
    """The same network evaluated in float64 on the float32 weights and the float32-rounded descriptor: the value every
    float32 implementation (TensorFlow's, the NumPy one above, the tcgen05 kernel) approximates; the pin of tests/golden/
    nn_h4o2_f64_golden.npz."""
    x = np.asarray(desc, dtype=np.float32).astype(np.float64)
    for li, (k, b) in enumerate(weights):
        z = x @ k.astype(np.float64) + b.astype(np.float64)
        x = z / (1.0 + np.exp(-z)) if li < len(weights) - 1 else np.maximum(z, 0.0)
    return x.reshape(-1) * WAVENUMBERS


# --------------------------------------------------------------------------- the loop itself
class Draws:
    """Random-number source for dmc_loop: either NumPy's legacy global stream (like the
    reference) or a replay of recorded draws (tests/golden traj_* fixtures)."""

    def __init__(self, flat=None, sizes=None, seed=None):
        self.flat, self.sizes, self.k, self.off = flat, sizes, 0, 0
        self.rs = np.random.RandomState(seed) if flat is None else None

    def _take(self, n):
        assert self.sizes[self.k] == n, (self.k, self.sizes[self.k], n)
        out = self.flat[self.off:self.off + n]
        self.k += 1
        self.off += n
        return out

    def normal(self, sigmas, shape_nad):     # pyvibdmc.py:544-546 : drawn as (N,D,A) then transposed
        n, a, d = shape_nad
        if self.rs is not None:
            return self.rs.normal(0.0, sigmas, size=(n, d, a)).transpose(0, 2, 1)
        return self._take(n * a * d).reshape(n, d, a).transpose(0, 2, 1)   # recorded in draw order (N,D,A)

    def uniform(self, n):
        if self.rs is not None:
            return self.rs.random_sample(n)
        return self._take(n)


def dmc_loop(coords, masses, dt, n0, nsteps, potential, draws, weighting='discrete', cont_thresh=(None, None),
             wts=None, equil=None, wfn_every=None, desc_steps=None, trial=None, imp1d_derivs=None, second_disp=False, excited=False, analytic=False):
    """Restatement of DMC_Sim.propagate (pyvibdmc.py:701-876) for the BASELINE configs:
    no checkpoints/logging, branch_every=1.  Returns dict(vref, pop, coords, pots, wts, wfns, eff_ts)."""
    coords = np.array(coords, dtype=np.float64)
    masses = np.asarray(masses, dtype=np.float64)
    sig = np.sqrt(dt / masses)                                    # :199
    alpha = 1.0 / (2.0 * dt)                                      # :201
    T = nsteps
    vref_t, pop_t = np.zeros(T), np.zeros(T)
    cont = weighting == 'continuous'
    if cont:
        wts = np.ones(n0) if wts is None else np.array(wts, dtype=np.float64)
        lower = 1 / n0 if cont_thresh[0] is None else cont_thresh[0]
        upper = cont_thresh[1]
    wfn_steps = set() if equil is None else set(range(equil, T + wfn_every, wfn_every))
    save_steps = {s + desc_steps for s in wfn_steps} if wfn_steps else set()
    desc_on, who, parent, parent_w, wfns = False, None, None, None, {}
    imp = trial is not None or imp1d_derivs is not None
    if imp:
        oned = imp1d_derivs is not None
        inv_m3 = (1 / masses)[None] if oned else (1 / np.repeat(masses, 3)).reshape(len(masses), 3)[None]
        sig3 = sig if oned else np.repeat(sig, 3).reshape(len(masses), 3)[None]
        eff = np.zeros(T)

        def drift(c):
            if oned:
                psi_fn, der_fn = imp1d_derivs
                d1, d2 = der_fn(c)
                return d1, psi_fn(c), d2
            return drift_analytic(c, trial) if analytic else drift_fd(c, trial)
        f_x = psi1 = sec = None
    vscore = None
    for t in range(T):
        if t in wfn_steps:                                        # :739-745
            parent = coords.copy()
            parent_w = None if not cont else wts.copy()
            who = np.arange(len(coords))
            desc_on = True
        if t == 0:                                                # :760-769
            pots = potential(coords)
            if imp:
                _, _, sec0 = drift(coords)
                sec = sec0
                pots = pots + local_kin(inv_m3, sec)
            vref = calc_vref(pots, n0, alpha, wts if cont else None)
        if not imp:                                               # :540-547
            coords = coords + draws.normal(sig, coords.shape)
            dt_eff = dt
        else:                                                     # :549-612 ; second_disp: :614-649
            disps = draws.normal(sig, coords.shape) if second_disp else None
            if second_disp:
                coords = coords + disps                           # the diffusion part is never rejected
                f_x, psi1, sec = drift(coords)
            elif f_x is None:
                f_x, psi1, sec = drift(coords)
            if not second_disp:
                disps = draws.normal(sig, coords.shape)
            d_x = inv_m3 * f_x
            if excited:                                           # :562-574 capped drift, first vector score
                d_x2 = d_x + 1e-50
                ms = 1 / inv_m3[:, :, 0]
                v2 = np.linalg.norm(d_x2, axis=2) ** 2
                factor = np.divide(-1 + np.sqrt(1 + 2 * ms * v2), ms * v2)
                d_x2 = factor[:, :, None] * d_x2
                if vscore is None:
                    numer = np.sum(np.linalg.norm(d_x2, axis=2) ** 2 * masses[None, :], axis=1)
                    denom = np.sum(np.linalg.norm(d_x, axis=2) ** 2 * masses[None, :], axis=1)
                    vscore = np.sqrt(numer / denom)
                d_x = d_x2
            y = coords + d_x * dt if second_disp else coords + disps + d_x * dt
            f_y, psi2, sec_y = drift(y)
            d_y = inv_m3 * f_y
            if excited:                                           # :581-591 (note: the factor scales d_y, not the shifted copy)
                d_y2 = d_y + 1e-50
                ms = 1 / inv_m3[:, :, 0]
                v2 = np.linalg.norm(d_y2, axis=2) ** 2
                factor = np.divide(-1 + np.sqrt(1 + 2 * ms * v2), ms * v2)
                d_y2 = factor[:, :, None] * d_y
                numer = np.sum(np.linalg.norm(d_y2, axis=2) ** 2 * masses[None, :], axis=1)
                denom = np.sum(np.linalg.norm(d_y, axis=2) ** 2 * masses[None, :], axis=1)
                vscore_new = np.sqrt(numer / denom)
                d_y = d_y2
            acc = metropolis(sig3, psi1, psi2, coords, y, d_x, d_y, dt)
            u = draws.uniform(len(coords))
            ok = acc > u
            coords = np.where(ok[:, None, None], y, coords)
            f_x = np.where(ok[:, None, None], f_y, f_x)
            psi1 = np.where(ok, psi2, psi1)
            sec = np.where(ok[:, None, None], sec_y, sec)
            if excited:
                vscore = np.where(ok, vscore_new, vscore)         # :608-609
            dt_eff = dt * (np.count_nonzero(ok) / len(coords))    # :603, :372-378
            eff[t] = dt_eff if t == 0 else eff[t - 1] + dt_eff
        pots = potential(coords)                                  # :786-793
        if imp:
            pots = pots + local_kin(inv_m3, sec)                  # :807-809
            if excited:
                pots = vref - (vref - pots) * vscore              # :810-811
        if not cont:                                              # :391-431
            u = draws.uniform(len(coords))
            _, idx, _, _, _ = birth_or_death_discrete(pots, vref, dt_eff, u, n0)
            coords, pots = coords[idx], pots[idx]
            if imp:
                f_x, psi1, sec = f_x[idx], psi1[idx], sec[idx]
                if excited:
                    vscore = vscore[idx]                          # :427-428
            if desc_on:
                who = who[idx]
            vref = calc_vref(pots, n0, alpha)
            pop_t[t] = len(coords)
        else:                                                     # :432-454
            wts, src, _, _, _ = branch_continuous(wts, pots, vref, dt_eff, lower, upper)
            coords, pots = coords[src], pots[src]
            if imp:
                f_x, psi1, sec = f_x[src], psi1[src], sec[src]
            if desc_on:
                who = who[src]
            vref = calc_vref(pots, n0, alpha, wts)
            pop_t[t] = np.sum(wts)
        vref_t[t] = vref
        if t + 1 in save_steps:                                   # :856-869
            desc_on = False
            dw = desc_wts_discrete(who, len(parent)) if not cont else desc_wts_continuous(wts, who, len(parent))
            wfns[t + 1 - desc_steps] = dict(coords=parent, desc_wts=dw, parent_wts=parent_w)
    return dict(vref=vref_t, pop=pop_t, coords=coords, pots=pots, wts=wts if cont else None, wfns=wfns,
                eff_ts=eff if imp else None)
