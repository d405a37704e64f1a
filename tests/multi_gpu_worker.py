"""Launched with torchrun by tests/test_gpu_multi.py: sharded H2O discrete DMC over WORLD_SIZE GPUs."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from pyvibdmc_b200 import _capi
    from pyvibdmc_b200.distributed import ShardedSim, shard_bounds
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
    amu = 1.0 / 6.02213670000e23 / 9.10938970000e-28
    masses = np.array([1.00782503, 1.00782503, 15.99491462]) * amu
    n0, T = 40000, 600
    sim = ShardedSim(3, 3, masses, n0, 5.0, _capi.POT_H2O_PS, seed=17, rebalance_every=100)
    start, count = shard_bounds(n0, world)[rank]
    sim.upload(np.repeat(eq[None] * 1.01, count, axis=0))
    sim.run(T)
    torch.cuda.synchronize()
    st = sim.state()
    stats = sim.stats(0, T)
    pops = sim.populations()
    # every rank must hold identical global histories
    h = torch.tensor(np.concatenate([stats["vref"], stats["pop"]]), device=sim.device)
    hmax, hmin = h.clone(), h.clone()
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
    same = bool(torch.equal(hmax, hmin))
    # ---- the NVLink mailbox exchange and the NCCL all-reduce must give bit-identical histories
    ref = ShardedSim(3, 3, masses, n0, 5.0, _capi.POT_H2O_PS, seed=17, rebalance_every=100, collective="nccl")
    ref.upload(np.repeat(eq[None] * 1.01, count, axis=0))
    ref.run(T)
    torch.cuda.synchronize()
    rstats = ref.stats(0, T)
    same_as_nccl = bool(sim.collective == "mailbox" and ref.collective == "nccl" and np.array_equal(rstats["vref"], stats["vref"])
                        and np.array_equal(rstats["pop"], stats["pop"]))
    ref.close()

    # ---- descendant weighting across shards: weights of all parents sum to the global population at window end
    off, n_par = sim.dw_begin()
    sim.run(60)
    dw = sim.dw_end()
    pops_dw = sim.populations()
    par_xyz, _ = sim.dw_parent()
    dw_ok = bool(len(dw) == n_par and abs(dw.sum() - sum(pops_dw)) < 0.5 and (dw >= 0).all() and (dw == np.round(dw)).all())
    npar_local = torch.tensor([len(par_xyz)], device=sim.device)
    dist.all_reduce(npar_local)
    dw_ok = dw_ok and int(npar_local.item()) == n_par
    sim.close()

    # ---- importance sampling across shards (global acceptance fraction -> dt_eff)
    import importlib.util
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyvibdmc_b200", "sample_potentials", "FortPots", "Partridge_Schwenke_H2O")
    spec = importlib.util.spec_from_file_location("call_trl_h2o_b200", os.path.join(d, "call_trl_h2o.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n1, T1 = 8000, 300
    imp = ShardedSim(3, 3, masses, n1, 1.0, _capi.POT_H2O_PS, seed=23, rebalance_every=100, trial=_capi.TRIAL_H2O_FD, trial_table=mod.packed_table())
    s1, c1 = shard_bounds(n1, world)[rank]
    imp.upload(np.repeat(eq[None] * 1.01, c1, axis=0))
    imp.run(T1)
    torch.cuda.synchronize()
    ist = imp.stats(0, T1)
    h2 = torch.tensor(np.concatenate([ist["vref"], ist["pop"], ist["dt_eff"]]), device=imp.device)
    h2max, h2min = h2.clone(), h2.clone()
    dist.all_reduce(h2max, op=dist.ReduceOp.MAX)
    dist.all_reduce(h2min, op=dist.ReduceOp.MIN)
    imp_same = bool(torch.equal(h2max, h2min))
    imp_out = {"same": imp_same, "zpe": float(ist["vref"][T1 // 2:].mean() / 4.556335281212229e-6), "dt_eff_mean": float(ist["dt_eff"][5:].mean()),
               "pop_last": float(ist["pop"][-1]), "rejected_mean": float(ist["rejected"][5:].mean())}
    imp.close()
    # ---- continuous weighting across shards (per-shard branching; the six kernels of the branching tail and both exchanges)
    n2, T2 = 16000, 400
    cont = {}
    for coll in ("mailbox", "nccl"):
        cs = ShardedSim(3, 3, masses, n2, 5.0, _capi.POT_H2O_PS, weighting="continuous", seed=31, rebalance_every=0, collective=coll)
        s2, c2 = shard_bounds(n2, world)[rank]
        cs.upload(np.repeat(eq[None] * 1.01, c2, axis=0))
        cs.run(T2)
        torch.cuda.synchronize()
        cst = cs.stats(0, T2)
        cont[coll] = (cst["vref"].copy(), cst["pop"].copy(), cst["births"].copy(), cs.collective)
        cs.close()
    h3 = torch.tensor(np.concatenate([cont["mailbox"][0], cont["mailbox"][1]]), device=torch.device("cuda", torch.cuda.current_device()))
    h3max, h3min = h3.clone(), h3.clone()
    dist.all_reduce(h3max, op=dist.ReduceOp.MAX)
    dist.all_reduce(h3min, op=dist.ReduceOp.MIN)
    cont_out = {"same": bool(torch.equal(h3max, h3min)),
                "mailbox_equals_nccl": bool(cont["mailbox"][3] == "mailbox" and cont["nccl"][3] == "nccl"
                                            and all(np.array_equal(cont["mailbox"][k], cont["nccl"][k]) for k in range(3))),
                "zpe": float(cont["mailbox"][0][T2 // 2:].mean() / 4.556335281212229e-6),
                "pop_mean": float(cont["mailbox"][1][T2 // 2:].mean()), "branched_mean": float(cont["mailbox"][2][T2 // 2:].mean())}
    sim = None
    if rank == 0:
        out = {"world": world, "cont": cont_out, "dw_ok": dw_ok, "imp": imp_out, "mailbox_equals_nccl": same_as_nccl, "step": st["step"], "pops": pops, "global_pop_last": float(stats["pop"][-1]), "same_on_all_ranks": same,
               "zpe": float(stats["vref"][T // 2:].mean() / 4.556335281212229e-6), "births_minus_deaths_ok":
               bool(np.array_equal(np.diff(stats["pop"]), (stats["births"] - stats["deaths"])[1:]))}
        print("\nRESULT " + json.dumps(out) + "\n", end="", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
