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
    if rank == 0:
        out = {"world": world, "step": st["step"], "pops": pops, "global_pop_last": float(stats["pop"][-1]), "same_on_all_ranks": same,
               "zpe": float(stats["vref"][T // 2:].mean() / 4.556335281212229e-6), "births_minus_deaths_ok":
               bool(np.array_equal(np.diff(stats["pop"]), (stats["births"] - stats["deaths"])[1:]))}
        print("RESULT " + json.dumps(out))
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
