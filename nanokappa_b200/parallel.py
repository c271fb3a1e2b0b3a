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

    def step(self, n=1):
        for _ in range(int(n)):
            self.engine.step_local()
            if self.world > 1:
                dist.all_reduce(self.acc, group=self.group)
            self.engine.step_finalize()
