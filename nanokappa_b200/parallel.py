"""Particle shards across the GPUs of one box.

Within a timestep every particle is independent given the previous subvolume temperatures; the only
coupling is the per-subvolume / per-reservoir sums.  So:

* particles are split by index block across ranks (geometry, mode tables and LUTs are replicated);
* every reservoir's (Q*J) emission table is split by mode range (``mode_range``) so that each rank
  injects a deterministic share -- ids and random draws are keyed by (step, reservoir, mode, copy),
  hence the union over ranks is exactly the single-GPU emission;
* per step ONE all-reduce(sum, f64) of the accumulator vector ``[sum e (S), count (S), sum v e (3S),
  per reservoir N_leaving / E_bal / flux, emitted, absorbed]`` (<= a few KB) between the two halves of
  the step: ``step_local`` (stream + emit + boundary kernels) and ``step_finalize`` (T_sv, results).

``ShardedEngine`` is the torch.distributed plumbing around ``Engine``; the collective is NCCL over
NVLink on GPUs (``backend='nccl'``).  ``reduce_fn`` can be replaced (tests use gloo on host copies).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check


def mode_range(rank, world, n_modes):
    """Contiguous share [lo, hi) of the flat mode index owned by `rank` (same split as nk_set_rank)."""
    return (n_modes * rank) // world, (n_modes * (rank + 1)) // world


def shard_bounds(rank, world, n_particles):
    """Index block [lo, hi) of the initial particles owned by `rank`."""
    return (n_particles * rank) // world, (n_particles * (rank + 1)) // world


class ShardedEngine:
    def __init__(self, engine, rank=None, world=None, group=None):
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        check(engine.ctx, engine.L.nk_set_rank(engine.ctx, self.rank, self.world), "nk_set_rank")
        ptr, ln = C.c_void_p(), C.c_int64()
        check(engine.ctx, engine.L.nk_acc_buffer(engine.ctx, C.byref(ptr), C.byref(ln)), "nk_acc_buffer")

        class _Acc:
            __cuda_array_interface__ = {"shape": (ln.value,), "typestr": "<f8", "data": (ptr.value, False), "version": 3}
        self.acc = torch.as_tensor(_Acc(), device=engine.device)

        self.fused = False
        self._synced = False

    def enable_fused_exchange(self):
        """Replace the per-step NCCL all-reduce by the in-kernel exchange: every rank exports its mailbox as a
        CUDA IPC handle, imports the peers' (NVLink peer mapping) and from then on the block that closes a step
        stores the rank's sums into all mailboxes, waits for the peers and adds them in rank order
        (csrc: nk_exchange_sums).  Returns False (and keeps NCCL) if peer mapping is not available."""
        if self.world == 1:
            return False
        eng = self.engine
        mine = (C.c_ubyte * 64)()
        try:
            check(eng.ctx, eng.L.nk_comm_export(eng.ctx, mine), "nk_comm_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine), group=self.group)
            for r, h in enumerate(handles):
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                check(eng.ctx, eng.L.nk_comm_import(eng.ctx, r, buf), "nk_comm_import")
            check(eng.ctx, eng.L.nk_comm_enable(eng.ctx, 1), "nk_comm_enable")
            ok = 1
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device=eng.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)      # all ranks or none
        self.fused = bool(flag.item())
        if not self.fused:
            eng.L.nk_comm_enable(eng.ctx, 0)
        return self.fused

    def step(self, n=1):
        if self.fused and not self._synced:
            torch.cuda.synchronize(self.engine.device)
            dist.barrier(group=self.group)      # the in-kernel wait is bounded: start the first step together
            self._synced = True
        if self.fused or self.world == 1:
            self.engine.step(n)
            return
        for _ in range(int(n)):
            self.engine.step_local()
            dist.all_reduce(self.acc, group=self.group)
            self.engine.step_finalize()
