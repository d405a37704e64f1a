"""GPU parity tests (-m gpu) of importance sampling with a USER trial wave function (csrc/pvd_impext.cuh, C-ABI
pvd_sim_imp_ext_*): the host answers ImpSamp.drift like the reference's plug-in (imp_samp_manager.py:92-139, 197-224),
the GPU runs the rest of imp_move_randomly (pyvibdmc.py:549-612).  Reference trajectories recorded with every random
draw (tests/golden/make_golden.py) are replayed through this path: populations bit-exact."""
import os

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
WN = 4.556335281212229e-6


@pytest.fixture(scope="module")
def K():
    from pyvibdmc_b200 import kernels
    assert kernels.device_count() > 0
    return kernels


class Replay:
    def __init__(self, g):
        self.flat, self.sizes, self.k, self.off = g["draw_flat"], g["draw_sizes"], 0, 0

    def take(self, n):
        assert self.sizes[self.k] == n, (self.k, int(self.sizes[self.k]), n)
        out = self.flat[self.off:self.off + n]
        self.k += 1
        self.off += n
        return out

    def normal(self, n, a, d):
        return np.ascontiguousarray(self.take(n * a * d).reshape(n, d, a).transpose(0, 2, 1))


def test_replay_ho_impsamp_hosted_trial(K, oracle):
    """traj_ho_imp (reference run with harm_trial_wfn.trial_harm / derivatives): the trial function is evaluated on the
    host (oracle restatement of the plug-in), built-in harmonic potential on the GPU."""
    from pyvibdmc_b200 import _capi
    g = golden("traj_ho_imp_golden.npz")
    rp = Replay(g)
    m, om = float(g["masses"][0]), 3700.0 * WN
    drift = lambda c: (oracle.harm_derivs(c)[0], oracle.trial_harm(c), oracle.harm_derivs(c)[1])
    sim = K.DeviceSim(1, 1, g["masses"], 300, 5.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2], trial=_capi.TRIAL_EXTERNAL)
    start = np.zeros((300, 1, 1))
    sim.upload(start)
    sim.imp_ext_init(*drift(start))
    T, n = 40, 300
    vref, pop = np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 1, 1)
        um, ub = rp.take(n), rp.take(n)
        y = sim.imp_ext_propose(disp)
        sim.imp_ext_accept(*drift(y), u_metro=um)
        sim.imp_ext_finish(None, True, u_branch=ub)
        st = sim.stats(t, 1)
        vref[t], pop[t] = st["vref"][0], st["pop"][0]
        n = int(pop[t])
    assert np.array_equal(pop, g["pop"]) and np.allclose(vref, g["vref"], rtol=1e-11)
    sim.close()


def test_replay_h2o_impsamp_hosted_trial_and_potential(K, oracle):
    """traj_h2o_imp (water product trial wfn, finite-difference derivatives): trial function AND potential on the host."""
    from pyvibdmc_b200 import _capi
    g = golden("traj_h2o_imp_golden.npz")
    rp = Replay(g)
    import pyvibdmc_b200
    table = np.load(os.path.join(os.path.dirname(pyvibdmc_b200.__file__), "sample_potentials", "FortPots", "Partridge_Schwenke_H2O",
                                 "free_oh_wvfn_table.npy"))
    trial = oracle.WaterTrial(table)
    drift = lambda c: oracle.drift_fd(c, trial)
    sim = K.DeviceSim(3, 3, g["masses"], 200, 1.0, _capi.POT_EXTERNAL, trial=_capi.TRIAL_EXTERNAL)
    start = np.repeat(EQ[None] * 1.01, 200, 0)
    sim.upload(start)
    sim.imp_ext_init(*drift(start), v=oracle.water_pot(start))
    T, n = 16, 200
    vref, pop, dts = np.zeros(T), np.zeros(T), np.zeros(T)
    for t in range(T):
        disp = rp.normal(n, 3, 3)
        um, ub = rp.take(n), rp.take(n)
        y = sim.imp_ext_propose(disp)
        sim.imp_ext_accept(*drift(y), u_metro=um)
        cds = sim.download()["coords"]
        sim.imp_ext_finish(oracle.water_pot(cds), True, u_branch=ub)
        st = sim.stats(t, 1)
        vref[t], pop[t], dts[t] = st["vref"][0], st["pop"][0], st["dt_eff"][0]
        n = int(pop[t])
    assert np.array_equal(pop, g["pop"])
    assert np.allclose(vref, g["vref"], rtol=1e-9)
    assert np.allclose(np.cumsum(dts), g["eff_ts"], rtol=1e-13)
    sim.close()


def test_hosted_equals_builtin_free_running(K, oracle):
    """Free-running (Philox) steps: the hosted path with the shipped harmonic trial function on the host draws the same numbers
    as the fused kernel with the same function on the device."""
    from pyvibdmc_b200 import _capi
    m, om = oracle.reduced_mass('O-H'), 3700.0 * WN
    drift = lambda c: (oracle.harm_derivs(c)[0], oracle.trial_harm(c), oracle.harm_derivs(c)[1])
    a = K.DeviceSim(1, 1, [m], 2000, 5.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2], trial=_capi.TRIAL_HARM1D, seed=5)
    a.set_trial_table(np.array([m * om]))
    b = K.DeviceSim(1, 1, [m], 2000, 5.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2], trial=_capi.TRIAL_EXTERNAL, seed=5)
    start = np.zeros((2000, 1, 1))
    a.upload(start)
    b.upload(start)
    b.imp_ext_init(*drift(start))
    a.run(30)
    for _ in range(30):
        y = b.imp_ext_propose()
        b.imp_ext_accept(*drift(y))
        b.imp_ext_finish(None, True)
    sa, sb = a.stats(0, 30), b.stats(0, 30)
    assert np.array_equal(sa["pop"], sb["pop"]) and np.allclose(sa["vref"], sb["vref"], rtol=1e-11)
    assert np.allclose(sa["dt_eff"], sb["dt_eff"], rtol=1e-14)
    a.close(); b.close()


def test_metropolis_and_local_kin_any_shape(K, oracle):
    """ImpSamp.metropolis / local_kin are shape generic in the reference (imp_samp.py:29-53; a last axis of length 1 means the
    one-dimensional problem there): 5 atoms x 3, 4 x 2, 12 x 3 next to the unrolled 3 x 3 and 1 x 1."""
    rng = np.random.default_rng(4)
    for na, nd in ((5, 3), (4, 2), (12, 3), (3, 3), (1, 1)):
        n = 777
        masses = rng.uniform(1000, 30000, na)
        sig = np.sqrt(1.0 / masses)
        x, y = rng.normal(0, 1, (n, na, nd)), rng.normal(0, 1, (n, na, nd))
        y = x + 0.05 * y
        fx, fy = rng.normal(0, 3, (n, na, nd)), rng.normal(0, 3, (n, na, nd))
        psx, psy = rng.normal(0, 1, n), rng.normal(0, 1, n)
        sec = rng.normal(0, 5, (n, na, nd))
        inv = (1 / masses)
        shape = (1, na, nd)
        inv_trip = np.broadcast_to(inv[None, :, None], shape)
        sig_trip = np.broadcast_to(sig[None, :, None], shape)
        ref = oracle.metropolis(sig_trip, psx, psy, x, y, inv_trip * fx, inv_trip * fy, 1.0)
        got = K.metropolis(x, y, fx, fy, psx, psy, sig, inv, 1.0)
        assert np.allclose(got, ref, rtol=1e-11, atol=0), (na, nd)
        assert np.array_equal(got == 0.0, ref == 0.0)
        assert np.allclose(K.local_kin(sec, inv), oracle.local_kin(inv_trip, sec), rtol=1e-14, atol=0), (na, nd)
        # the reference-facing class (same signature as imp_samp.py:29-53: 'trip' constants, D = inv_mass * f)
        from pyvibdmc_b200.simulation_utilities.imp_samp import ImpSamp
        cls = ImpSamp.metropolis(np.array(sig_trip[0]), psx, psy, x, y, inv_trip * fx, inv_trip * fy, 1.0)
        assert np.allclose(cls, ref, rtol=1e-11, atol=0), (na, nd)
        assert np.allclose(ImpSamp.local_kin(np.array(inv_trip[0]), sec), oracle.local_kin(inv_trip, sec), rtol=1e-14, atol=0), (na, nd)


def test_device_to_device_walker_transfer(K, oracle):
    """Rebalancing payloads: export_tail_device / import_device move whole walkers -- coordinates, V, who_from and the
    importance-sampling companions -- between two simulation handles without touching the host; same result as the host path."""
    from pyvibdmc_b200 import _capi
    m, om = oracle.reduced_mass('O-H'), 3700.0 * WN
    mk = lambda seed: K.DeviceSim(1, 1, [m], 4000, 5.0, _capi.POT_HARMONIC, pot_params=[(0.5 * m) * om ** 2], trial=_capi.TRIAL_HARM1D,
                                  seed=seed, capacity=9000)
    a, b = mk(1), mk(2)
    for s in (a, b):
        s.set_trial_table(np.array([1.3 * m * om]))       # not the exact ground state: the walkers differ
        s.upload(np.random.default_rng(7).normal(0, 0.05, (3000, 1, 1)))
        s.run(5)
    da, db = a.download(), b.download()
    fa, pa, _ = a.download_imp()
    fb, pb, _ = b.download_imp()
    na, nb, count = len(da["coords"]), len(db["coords"]), 700
    payload = a.export_tail_device(count)
    assert payload.shape == (count, 1 + 3 + 1 + 2)
    b.import_device(payload.torch())
    assert a.state()["n"] == na - count and b.state()["n"] == nb + count
    da2, db2 = a.download(), b.download()
    fb2, pb2, _ = b.download_imp()
    assert np.array_equal(da2["coords"], da["coords"][:na - count])
    assert np.array_equal(db2["coords"], np.concatenate([db["coords"], da["coords"][na - count:]]))
    assert np.array_equal(db2["pots"], np.concatenate([db["pots"], da["pots"][na - count:]]))
    assert np.array_equal(fb2, np.concatenate([fb, fa[na - count:]])) and np.array_equal(pb2, np.concatenate([pb, pa[na - count:]]))
    b.run(3)                                              # the enlarged shard keeps running
    assert b.state()["step"] == 8
    a.close(); b.close()


def test_chain_rule_helper_on_cuda_tensors():
    """ChainRuleHelper(coords, torch) with CUDA tensors: every method runs on the device and reproduces the
    unmodified reference class (golden vectors from imp_samp_helper.py:10-209)."""
    import torch
    from pyvibdmc_b200.simulation_utilities.imp_samp_helper import ChainRuleHelper
    g = golden("chain_rule_golden.npz")
    dev = torch.device("cuda", 0)
    for tag in ("w", "p"):
        cds = torch.from_numpy(g[f"{tag}_cds"]).to(dev)
        h = ChainRuleHelper(cds, torch)
        pairs, ang = [list(p) for p in g[f"{tag}_pairs"]], list(g[f"{tag}_ang"])
        dr = [h.dr_dx(p) for p in pairs]
        d2r = [h.d2r_dx2(p) for p in pairs]
        dth, d2th = h.dth_dx(ang), h.d2th_dx2(ang)
        dpsi, d2psi = torch.from_numpy(g[f"{tag}_dpsi"]).to(dev), torch.from_numpy(g[f"{tag}_d2psi"]).to(dev)
        got = {"dr0": dr[0], "d2r1": d2r[1], "dc": h.dcth_dx(ang), "d2c": h.d2cth_dx2(ang), "dth": dth, "d2th": d2th,
               "jac": h.dpsidx(dpsi, [dr[0], dr[1], dth]), "lap": h.d2psidx2(d2psi, [d2r[0], d2r[1], d2th], dpsi, [dr[0], dr[1], dth])}
        for k, v in got.items():
            assert v.is_cuda
            ref = g[f"{tag}_{k}"]
            assert np.allclose(v.cpu().numpy(), ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max()), (tag, k)
