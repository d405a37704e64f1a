"""Walker sharding across GPUs (one process per GPU, torch.distributed / NCCL as plumbing).

The reference has no multi-device path for the loop itself (only Pool.map / MPI futures around the
potential call: simulation_utilities/potential_manager.py:59-99, mpi_potential_manager.py:69-74).
Here every rank owns a contiguous shard of the ensemble in its own HBM and runs the same step
kernels; the only per-step exchange is one all-reduce of PVD_NSUMS doubles
[sum c*V, sum c, births, deaths, sum V, n_in, err, n_acc, (vmin,vmax,wmin,wmax) x ranks]
from which every rank derives the same Vref / population.  Shard populations drift
apart under discrete weighting, so every `rebalance_every` steps walkers are moved from the
fullest to the emptiest shards (plan_rebalance: pure host logic, tested on CPU with gloo).
"""
import os

import numpy as np

from . import _capi, kernels

__all__ = ["plan_rebalance", "shard_bounds", "ShardedSim", "ShardedDevice"]


def shard_bounds(n_total, world):
    """Contiguous split of n_total walkers over `world` ranks (first ranks take the remainder)."""
    base, rem = divmod(int(n_total), int(world))
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int)
    return [(int(s), int(c)) for s, c in zip(starts, counts)]


def plan_rebalance(pops, tolerance=0.05):
    """Transfers [(src_rank, dst_rank, count), ...] that bring every shard within `tolerance` of the mean.

    Deterministic and identical on every rank (it only depends on the all-gathered populations):
    repeatedly move walkers from the fullest shard to the emptiest one."""
    pops = [int(p) for p in pops]
    world = len(pops)
    total = sum(pops)
    target = [total // world + (1 if r < total % world else 0) for r in range(world)]
    if total == 0 or max(abs(p - t) for p, t in zip(pops, target)) <= tolerance * max(total / world, 1.0):
        return []
    surplus = [p - t for p, t in zip(pops, target)]
    moves = []
    while True:
        src = max(range(world), key=lambda r: (surplus[r], -r))
        dst = min(range(world), key=lambda r: (surplus[r], r))
        count = min(surplus[src], -surplus[dst])
        if count <= 0:
            break
        moves.append((src, dst, count))
        surplus[src] -= count
        surplus[dst] += count
    return moves


class ShardedSim:
    """DeviceSim on every rank + the per-step all-reduce.  Requires torch.distributed to be initialised
    (backend nccl) and one CUDA device per rank."""

    def __init__(self, natoms, ndim, masses, num_walkers, delta_t, potential, weighting="discrete", seed=0,
                 rng_mode=_capi.RNG_DEFAULT, pot_params=None, thresh_lower=None, thresh_upper=None, rebalance_every=250,
                 capacity=None, stats_ring=1 << 14, trial=_capi.TRIAL_NONE, trial_table=None, collective="mailbox", alpha=None,
                 imp_variant=_capi.IMP_STANDARD):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.local_rank = int(os.environ.get("LOCAL_RANK", self.rank % max(torch.cuda.device_count(), 1)))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.stream = torch.cuda.Stream(device=self.device)
        self.num_walkers = int(num_walkers)
        local = -(-self.num_walkers // self.world)
        self.sim = kernels.DeviceSim(natoms, ndim, masses, num_walkers, delta_t, potential, weighting=weighting, alpha=alpha,
                                     seed=int(seed) + 1000003 * self.rank, rng_mode=rng_mode, pot_params=pot_params,
                                     thresh_lower=thresh_lower, thresh_upper=thresh_upper, device=self.local_rank,
                                     rank=self.rank, world_size=self.world, trial=trial, imp_variant=imp_variant,
                                     capacity=capacity or int(1.6 * local) + 4096, stats_ring=stats_ring)
        self.sim.set_stream(self.stream.cuda_stream)
        self.imp = trial != _capi.TRIAL_NONE
        # a user potential (the reference's getpot plug-in): every rank evaluates its own shard on the host between
        # ext_move() and ext_finish(); the per-step exchange is the NCCL all-reduce (the host is in the loop anyway)
        self.external = int(potential) == _capi.POT_EXTERNAL
        self.hosted = trial == _capi.TRIAL_EXTERNAL
        if self.imp and not self.hosted:
            self.sim.set_trial_table(trial_table)
        self.sums = torch.zeros(_capi.NSUMS, dtype=torch.float64, device=self.device)
        self._gate = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.sim.set_sums_ptr(self.sums.data_ptr())
        # per-step exchange: "mailbox" = peer stores over NVLink fused into the step kernel's last CTA (no collective kernel),
        # "nccl" = all-reduce between the step kernel and the finalisation.  Importance sampling needs two exchanges -> nccl.
        self.collective = "nccl" if (self.imp or self.external or self.world == 1) else collective
        if self.collective == "mailbox":
            handles = [None] * self.world
            dist.all_gather_object(handles, self.sim.mailbox_handle())
            self.sim.mailbox_connect(handles)
        self.rebalance_every = int(rebalance_every)
        self.steps_done = 0
        self.natoms, self.ndim = natoms, ndim

    def upload(self, coords_local, wts_local=None):
        with self.torch.cuda.stream(self.stream):
            self.sim.upload(coords_local, wts_local)
            if self.external or self.hosted:
                return                                 # the energies / drift terms of the start ensemble follow: set_pots() / imp_ext_init()
            self.dist.all_reduce(self.sums)
            self.sim.init_finalize()

    def set_pots(self, v_local):
        """Energies of this rank's shard of the start ensemble (user potential) -> global Vref."""
        with self.torch.cuda.stream(self.stream):
            self.sim.set_pots(v_local)
            self.dist.all_reduce(self.sums)
            self.sim.init_finalize()

    def ext_move(self):
        """Displace this rank's shard; the moved coordinates (n_local, atoms, dims) for the user potential."""
        return self.sim.ext_move()

    def ext_finish(self, v_local, do_branch=True):
        """Energies of the moved shard -> weights / branching per shard, global Vref and population."""
        with self.torch.cuda.stream(self.stream):
            self.sim.ext_finish(v_local, do_branch)
            self.dist.all_reduce(self.sums)
            self.sim.step_finalize()
            self.steps_done += 1
            if self.rebalance_every and self.steps_done % self.rebalance_every == 0:
                self.rebalance()

    # -- importance sampling with a user trial wave function (PVD_TRIAL_EXTERNAL): every rank's impsamp.drift sees its own shard
    def imp_ext_init(self, fx, psi, sec, v=None):
        with self.torch.cuda.stream(self.stream):
            self.sim.imp_ext_init(fx, psi, sec, v)
            self.dist.all_reduce(self.sums)
            self.sim.init_finalize()

    def imp_ext_propose(self):
        return self.sim.imp_ext_propose()

    def imp_ext_accept(self, fy, psi_y, sec_y):
        """Metropolis step per shard; the acceptance fraction that scales the time step is global (pyvibdmc.py:372-378, 603)."""
        with self.torch.cuda.stream(self.stream):
            self.sim.imp_ext_accept(fy, psi_y, sec_y)
            self.dist.all_reduce(self.sums)

    def imp_ext_finish(self, v=None, do_branch=True):
        with self.torch.cuda.stream(self.stream):
            self.sim.imp_ext_finish(v, do_branch)
            self.dist.all_reduce(self.sums)
            self.sim.step_finalize()
            self.steps_done += 1
            if self.rebalance_every and self.steps_done % self.rebalance_every == 0:
                self.rebalance()

    def run(self, nsteps, branch_every=1):
        if self.collective == "mailbox":
            done = 0
            while done < int(nsteps):                  # whole segments are enqueued by one call; stop at rebalancing points
                k = int(nsteps) - done
                if self.rebalance_every:
                    k = min(k, self.rebalance_every - self.steps_done % self.rebalance_every)
                # the mailbox exchange spins on the device for the peers' messages: line the ranks up first (a stream-ordered
                # all-reduce: the host may have spent seconds on a checkpoint since the last segment)
                with self.torch.cuda.stream(self.stream):
                    self.dist.all_reduce(self._gate)
                self.sim.run_mailbox(k, branch_every)
                done += k
                self.steps_done += k
                if self.rebalance_every and self.steps_done % self.rebalance_every == 0:
                    with self.torch.cuda.stream(self.stream):
                        self.rebalance()
            return
        with self.torch.cuda.stream(self.stream):
            for _ in range(int(nsteps)):
                do_branch = 1 if branch_every == 1 else -int(branch_every)
                if self.imp:
                    # the acceptance fraction that scales dt is global (pyvibdmc.py:372-378): one more tiny all-reduce
                    self.sim.imp_move_local()
                    self.dist.all_reduce(self.sums)
                    self.sim.imp_branch_local(do_branch)
                else:
                    self.sim.step_local(do_branch)
                self.dist.all_reduce(self.sums)
                self.sim.step_finalize()
                self.steps_done += 1
                if self.rebalance_every and self.steps_done % self.rebalance_every == 0:
                    self.rebalance()

    def populations(self):
        torch, dist = self.torch, self.dist
        n = torch.tensor([self.sim.state(raise_on_error=False)["n"]], dtype=torch.int64, device=self.device)
        out = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, n)
        return [int(v) for v in out.tolist()]                   # one device-to-host read for the whole world

    def rebalance(self, tolerance=0.05):
        """Move whole walkers (coords, V, weight, who_from and the importance-sampling companions) from over- to under-populated
        shards, GPU to GPU: packed on the device, NCCL send / recv over NVLink on device pointers, unpacked on the device."""
        torch, dist = self.torch, self.dist
        moves = plan_rebalance(self.populations(), tolerance)
        for src, dst, count in moves:
            if self.rank == src:
                payload = self.sim.export_tail_device(count).torch()
                dist.send(payload, dst)
                torch.cuda.synchronize(self.device)             # the buffer belongs to the simulation handle: sent before it is re-used
            elif self.rank == dst:
                ncols = self.natoms * self.ndim + 3
                if self.sim.cfg.trial != _capi.TRIAL_NONE:
                    ncols += self.natoms * self.ndim + 2 + (1 if self.sim.cfg.imp_variant == _capi.IMP_EXCITED_STATE else 0)
                buf = torch.empty((count, ncols), dtype=torch.float64, device=self.device)
                dist.recv(buf, src)
                self.sim.import_device(buf)
        return moves

    # -- descendant weighting across shards: who_from holds GLOBAL parent ids (rank-major order)
    def dw_begin(self):
        """Open a descendant-weighting window (pyvibdmc.py:739-747); returns (global offset of this shard's parents, N_parent)."""
        pops = self.populations()
        offset = sum(pops[:self.rank])
        with self.torch.cuda.stream(self.stream):
            self.sim.dw_begin(offset)
        self._dw_total = sum(pops)
        return offset, self._dw_total

    def dw_end(self, close_window=True):
        """Close the window: descendant weights of ALL parents, identical on every rank (pyvibdmc.py:663-672, 856-869).
        close_window=False: the same sums with the window left open (DEBUG_save_desc_wt_tracker, pyvibdmc.py:849-852)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            local = self.sim.dw_end(self._dw_total) if close_window else self.sim.dw_peek(self._dw_total)
            t = torch.from_numpy(local).to(self.device)
            self.dist.all_reduce(t)
        self.stream.synchronize()
        return t.cpu().numpy()

    def dw_parent(self):
        """This shard's slice of the parent ensemble (coords, weights or None); concatenate in rank order for the global one."""
        return self.sim.dw_parent()

    def state(self):
        return self.sim.state()

    def stats(self, first, count):
        return self.sim.stats(first, count)

    def close(self):
        self.sim.close()


class ShardedDevice:
    """`kernels.DeviceSim`-shaped facade over ShardedSim, used by DMC_Sim when it runs under torch.distributed:
    every rank executes the same DMC_Sim code; counts, histories and downloaded arrays are GLOBAL (rank-major order),
    so each rank could write the reference's output files (DMC_Sim lets rank 0 write into the real folder)."""

    def __init__(self, natoms, ndim, masses, num_walkers, delta_t, potential, weighting="discrete", alpha=None, capacity=None,
                 seed=0, rng_mode=_capi.RNG_DEFAULT, trial=_capi.TRIAL_NONE, pot_params=None, thresh_lower=None, thresh_upper=None,
                 device=0, stats_ring=1 << 16, imp_variant=_capi.IMP_STANDARD, trial_table=None, rebalance_every=250):
        self.ss = ShardedSim(natoms, ndim, masses, num_walkers, delta_t, potential, weighting=weighting, seed=seed, rng_mode=rng_mode,
                             pot_params=pot_params, thresh_lower=thresh_lower, thresh_upper=thresh_upper, stats_ring=stats_ring,
                             trial=trial, trial_table=trial_table, rebalance_every=rebalance_every, imp_variant=imp_variant, alpha=alpha)     # alpha: DEBUG_alpha (Vref feedback)
        self.natoms, self.ndim, self.cfg = natoms, ndim, self.ss.sim.cfg
        self.rank, self.world = self.ss.rank, self.ss.world

    # -- helpers
    def _gather_rows(self, a):
        """Concatenate per-rank arrays with the same trailing shape along axis 0, rank-major, on every rank."""
        torch, dist = self.ss.torch, self.ss.dist
        a = np.ascontiguousarray(a)
        n = torch.tensor([a.shape[0]], dtype=torch.int64, device=self.ss.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        m = max(sizes + [1])
        pad = torch.zeros((m,) + a.shape[1:], dtype=torch.from_numpy(a[:0]).dtype, device=self.ss.device)
        if a.shape[0]:
            pad[:a.shape[0]] = torch.from_numpy(a).to(self.ss.device)
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad)
        return np.concatenate([b[:k].cpu().numpy() for b, k in zip(bufs, sizes)], axis=0)

    # -- set-up
    def set_nn_weights(self, packed):
        self.ss.sim.set_nn_weights(packed)

    def set_trial_table(self, table, ntab=None):
        pass                                   # handed to ShardedSim at construction

    def upload(self, coords, wts=None):
        start, count = shard_bounds(len(coords), self.world)[self.rank]
        self.ss.upload(np.ascontiguousarray(coords[start:start + count]), None if wts is None else np.ascontiguousarray(wts[start:start + count]))

    def local_slice(self, n_global):
        """(start, count) of this rank's shard of a global start ensemble (what upload() keeps)."""
        return shard_bounds(n_global, self.world)[self.rank]

    def set_pots(self, v_local):
        self.ss.set_pots(np.ascontiguousarray(v_local, dtype=np.float64))

    def ext_move(self):
        return self.ss.ext_move()

    def imp_ext_init(self, fx, psi, sec, v=None):
        self.ss.imp_ext_init(fx, psi, sec, v)

    def imp_ext_propose(self):
        return self.ss.imp_ext_propose()

    def imp_ext_accept(self, fy, psi_y, sec_y):
        self.ss.imp_ext_accept(fy, psi_y, sec_y)

    def imp_ext_finish(self, v=None, do_branch=True):
        self.ss.imp_ext_finish(v, do_branch)

    def download_local(self):
        """This rank's shard (what a user potential / trial function is evaluated on)."""
        return self.ss.sim.download()

    def ext_finish(self, v_local, do_branch=True):
        self.ss.ext_finish(np.ascontiguousarray(v_local, dtype=np.float64), do_branch)

    def set_masses(self, masses):
        self.ss.sim.set_masses(masses)

    # -- stepping
    def run(self, nsteps, branch_every=1):
        self.ss.run(nsteps, branch_every)

    def sync(self):
        self.ss.stream.synchronize()

    # -- descendant weighting
    def dw_begin(self):
        self.ss.dw_begin()

    def dw_end(self, n_parent):
        return self.ss.dw_end()

    def dw_peek(self, n_parent):
        return self.ss.dw_end(close_window=False)

    def dw_parent(self):
        xyz, w = self.ss.dw_parent()
        return self._gather_rows(xyz), (None if w is None else self._gather_rows(w))

    # -- queries (global views)
    def state(self, raise_on_error=True):
        st = self.ss.sim.state(raise_on_error=raise_on_error)
        st["n"] = sum(self.ss.populations())
        return st

    def stats(self, first_step, count):
        return self.ss.stats(first_step, count)

    def download(self, who_from=False, out=None):
        d = self.ss.sim.download(who_from=who_from)
        return {k: (None if v is None else self._gather_rows(v)) for k, v in d.items()}

    def download_imp(self):
        return tuple(self._gather_rows(a) for a in self.ss.sim.download_imp())

    def close(self):
        self.ss.close()
