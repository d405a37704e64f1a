"""GPU test (-m gpu, needs >= 2 devices): walker sharding with the per-step NCCL all-reduce and rebalancing."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _result(stdout):
    """The worker's RESULT line; the ranks share one pipe, so another rank's output may follow on the same line."""
    line = [l for l in stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.JSONDecoder().raw_decode(line[7:])[0]


def test_two_gpu_sharded_run():
    from pyvibdmc_b200 import kernels
    if kernels.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(here, "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    out = _result(res.stdout)
    assert out["world"] == 2 and out["step"] == 600 and out["same_on_all_ranks"] and out["births_minus_deaths_ok"]
    assert sum(out["pops"]) == int(out["global_pop_last"]) and 20000 < sum(out["pops"]) < 60000
    assert abs(out["pops"][0] - out["pops"][1]) < 0.2 * sum(out["pops"])          # rebalancing keeps shards level
    assert 4350 < out["zpe"] < 4900
    assert out["mailbox_equals_nccl"]                                            # NVLink mailbox exchange == NCCL all-reduce, bit for bit
    assert out["dw_ok"]                                                          # descendant weights of all parents, all-reduced
    imp = out["imp"]                                                             # importance sampling with the global acceptance fraction
    assert imp["same"] and 4350 < imp["zpe"] < 4900 and 0.9 < imp["dt_eff_mean"] <= 1.0 and 4000 < imp["pop_last"] < 12000
    assert 0 < imp["rejected_mean"] < 0.1 * 8000
    cont = out["cont"]                                                           # continuous weighting, sharded: both exchanges, same bits
    assert cont["same"] and cont["mailbox_equals_nccl"] and 4350 < cont["zpe"] < 4900
    assert 0.9 * 16000 < cont["pop_mean"] < 1.1 * 16000 and cont["branched_mean"] > 0


def test_two_gpu_dmc_sim_drop_in(tmp_path):
    """The user API itself under torchrun: walkers sharded over 2 GPUs, rank 0 writes the reference's files."""
    from pyvibdmc_b200 import kernels
    if kernels.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path / "w2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29618", os.path.join(here, "multi_gpu_dmcsim_worker.py"), out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    r = _result(res.stdout)
    assert r["world"] == 2 and r["vref_shape"] == [600, 2] and r["log_has_steps"]
    assert r["wfns"] == ["w2_wfn_100ts.hdf5", "w2_wfn_300ts.hdf5", "w2_wfn_500ts.hdf5"]
    assert r["n_parent"] == int(r["pop_at_window_start"]) and r["desc_sum"] == r["pop_at_window_end"]
    assert r["final_walkers"] == int(r["final_pop"]) and 4400 < r["zpe"] < 4850
    assert len(r["chkpts"]) >= 1
    # dmc_restart under torchrun stays sharded (world / rank / GPU re-detected) and extends the histories
    assert r["restart_vref_shape"] == [700, 2] and r["restart_final_walkers"] == int(r["restart_final_pop"]) and 4400 < r["restart_zpe"] < 4850


def test_two_gpu_dmc_sim_user_potential(tmp_path):
    """A user potential callable (getpot plug-in, potential_manager.py:71-99) on a sharded run: each rank's callable sees its own
    shard of the moved walkers once per step, Vref and the population are global."""
    from pyvibdmc_b200 import kernels
    if kernels.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path / "e2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29619", os.path.join(here, "multi_gpu_dmcsim_worker.py"), out, "ext"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    r = _result(res.stdout)
    assert r["world"] == 2
    for weighting in ("discrete", "continuous"):
        w = r[weighting]
        assert w["vref_shape"] == [1500, 2] and abs(w["zpe"] - 1852.5) < 15, w       # HO 1850 cm-1 + time-step bias (config 1)
        assert w["calls"] == 1501 and 0.4 < w["shard_fraction"] < 0.6, w             # start ensemble + one call per step, half the walkers
        assert 15000 < w["pop_min"] and w["pop_max"] < 25000, w
        assert abs(w["desc_sum"] - w["pop_at_window_end"]) < 1e-6 * 20000, w
    assert r["discrete"]["final_walkers"] == int(r["discrete"]["final_pop"])
    assert r["discrete"]["tracker_ok"] and r["continuous"]["alpha"] == 0.03        # DEBUG_save_desc_wt_tracker / DEBUG_alpha, sharded
    assert r["continuous"]["final_walkers"] == 20000 and abs(r["continuous"]["weight_sum"] - r["continuous"]["final_pop"]) < 1e-6 * 20000


def test_two_gpu_second_impsamp_displacement(tmp_path):
    """pyvibdmc.py:614-649 sharded: exact trial function => zero-variance estimator on both ranks, standard and second move type."""
    from pyvibdmc_b200 import kernels
    if kernels.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29620", os.path.join(here, "multi_gpu_dmcsim_worker.py"), str(tmp_path / "i2"), "imp2"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    r = _result(res.stdout)
    assert r["world"] == 2
    for tag in ("std", "second"):
        assert r[tag]["vref_max_dev_cm1"] < 1e-6 and r[tag]["pop_constant"] and r[tag]["n"] == 4000 and r[tag]["walker_std"] > 0.05, r
    # harmonic trial function on the Morse oscillator: branching under importance sampling, discrete and continuous, sharded
    for weighting in ("discrete", "continuous"):
        m = r["morse_" + weighting]
        assert abs(m["zpe"] - 1833.4) < 6 and m["vref_std"] > 0, m
        assert 17000 < m["pop_min"] and m["pop_max"] < 23000 and abs(m["weight_sum"] - m["final_pop"]) < 1e-6 * 20000, m
    assert r["morse_continuous"]["n"] == 20000 and r["morse_discrete"]["n"] == int(r["morse_discrete"]["final_pop"])


def test_two_gpu_excited_state_imp_samp(tmp_path):
    """excited_state_imp_samp (pyvibdmc.py:562-591, 608-611, 810-811) sharded over 2 GPUs agrees with the same run on one GPU."""
    from pyvibdmc_b200 import kernels
    if kernels.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join(here, "multi_gpu_dmcsim_worker.py"), str(tmp_path / "ex"), "exc"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    r = _result(res.stdout)
    sh, one = r["sharded"], r["single"]
    assert sh["world"] == 2 and one["world"] == 1, r
    assert abs(sh["zpe"] - one["zpe"]) < 25 and 4500 < sh["zpe"] < 4800, r
    assert sh["n"] == int(sh["final_pop"]) and 6000 < sh["pop_min"] and sh["pop_max"] < 10000, r


def test_two_gpu_user_trial_function(tmp_path):
    """Any trial / derivative callable through ImpSampManager (imp_samp_manager.py:92-139, 197-224) on a sharded run, with a shipped
    and with a user potential: drift terms per shard from the host, global acceptance fraction, global Vref."""
    from pyvibdmc_b200 import kernels
    if kernels.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29622", os.path.join(here, "multi_gpu_dmcsim_worker.py"), str(tmp_path / "ui"), "uimp"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    r = _result(res.stdout)
    assert r["world"] == 2
    for tag in ("builtin_pot", "user_pot"):
        w = r[tag]
        assert w["hosted"] and abs(w["zpe"] - 1850.0) < 25, r
        assert 4000 < w["pop_min"] and w["pop_max"] < 8000 and w["pop_std"] > 0 and w["n"] == int(w["final_pop"]), r
        assert w["tau_increasing"] and w["tau_last"] <= 3500.0, r                   # effective time axis (rejections shorten the step)
