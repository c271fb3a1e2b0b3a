"""Particle shards across the GPUs of one box.

Within a timestep every particle is independent given the previous subvolume temperatures; the only
coupling is the per-subvolume / per-reservoir sums.  So:

* particles are split by index block across ranks (geometry, mode tables and LUTs are replicated);
* every rank advances the whole (R, Q*J) emission table; the copies an entry emits are dealt round-robin over the ranks
  (``emission_owner``: copy k of an entry that has emitted ``fire`` particles before belongs to rank
  ``(fire + k + mode) % world``), so each rank injects 1/world of EVERY mode -- its per-mode particle numbers stay in
  balance with what it absorbs, which keeps the per-mode slot pools (``Engine.sort_by_mode``) and the live counts level.
  Ids and random draws are keyed by (step, reservoir, mode, copy), hence the union over ranks is exactly the single-GPU
  emission;
* per step ONE all-reduce(sum, f64) of the accumulator vector ``[sum e (S), count (S), sum v e (3S),
  per reservoir N_leaving / E_bal / flux, emitted, absorbed]`` (<= a few KB) between the two halves of
  the step: ``step_local`` (stream + emit + boundary kernels) and ``step_finalize`` (T_sv, results).

* live counts drift apart because absorption depends on position: ``ShardedEngine.rebalance`` (every ~100 steps, outside
  a timestep) all-gathers the live counts and migrates whole particle rows from the fullest to the emptiest ranks with
  point-to-point NCCL sends over NVLink.  Which rank owns a particle has no effect on the physics (ids and draws are
  keyed by the particle, sums are global), so results are unchanged.

``ShardedEngine`` is the torch.distributed plumbing around ``Engine``; the collective is NCCL over
NVLink on GPUs (``backend='nccl'``).  ``reduce_fn`` can be replaced (tests use gloo on host copies).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check


def emission_owner(fire, copy, mode, world):
    """Rank that injects copy `copy` (0-based) of a reservoir-table entry of flat mode `mode` that has emitted `fire`
    particles before this step (csrc: nk_emit_owner; the device keeps `fire` modulo 256)."""
    return (fire % 256 + copy + mode) % world


F64_FIELDS = ("px", "py", "pz", "tc", "occ", "cx", "cy", "cz", "pid")      # pid travels as its bit pattern
I32_FIELDS = ("mode", "omode", "cfacet")


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the host staging buffers it
    allocates afterwards (first touch) and the threads that scan / patch them sit next to the GPU's PCIe root port.
    With one rank per GPU and host-resident particle arrays, cross-socket DMA is otherwise the bottleneck.  Returns the
    node, or None when the topology is not exposed (containers, single-node hosts): then nothing is changed."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "{:04x}:{:02x}:{:02x}.0".format(getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def rebalance_plan(counts):
    """Deterministic transfer list [(src, dst, n)] that brings every rank to total // W (+1 for the first total % W ranks).
    Every rank computes the same plan from the all-gathered live counts."""
    counts = [int(c) for c in counts]
    W, total = len(counts), sum(counts)
    target = [total // W + (1 if r < total % W else 0) for r in range(W)]
    surplus = [[r, counts[r] - target[r]] for r in range(W) if counts[r] > target[r]]
    deficit = [[r, target[r] - counts[r]] for r in range(W) if counts[r] < target[r]]
    plan, i, j = [], 0, 0
    while i < len(surplus) and j < len(deficit):
        n = min(surplus[i][1], deficit[j][1])
        plan.append((surplus[i][0], deficit[j][0], n))
        surplus[i][1] -= n
        deficit[j][1] -= n
        if surplus[i][1] == 0:
            i += 1
        if deficit[j][1] == 0:
            j += 1
    return plan


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
        self.fused_error = None
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
            self.fused_error = None
        except Exception as e:              # e.g. no peer access between the GPUs: keep the NCCL all-reduce
            ok = 0
            self.fused_error = str(e)
        flag = torch.tensor([ok], device=eng.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)      # all ranks or none
        self.fused = bool(flag.item())
        if not self.fused:
            eng.L.nk_comm_enable(eng.ctx, 0)
        return self.fused

    def close(self):
        """Collective: unmap the peers' mailboxes on every rank, then destroy the context (a rank must not free its
        mailbox while a peer still has it mapped)."""
        if self.fused:
            torch.cuda.synchronize(self.engine.device)
            self.engine.L.nk_comm_enable(self.engine.ctx, 0)
            self.fused = False
        if self.world > 1:
            dist.barrier(group=self.group)
        self.engine.close()

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

    # ---- periodic rebalance of the live particles ------------------------------------------------------
    def live_counts(self):
        _, alive = self.engine.slot_count()
        if self.world == 1:
            return [alive]
        t = torch.zeros(self.world, dtype=torch.int64, device=self.engine.device)
        t[self.rank] = alive
        dist.all_reduce(t, group=self.group)
        return [int(v) for v in t.tolist()]

    def extract_for(self, plan):
        """Remove the rows this rank sends under `plan` -> {dst: (f64 block (9, n), i32 block (3, n))}.  The rows are
        a strided sample of the (mode-ordered, compacted) shard, so the mode mix of both sides stays representative."""
        eng = self.engine
        mine = [(dst, n) for src, dst, n in plan if src == self.rank and n > 0]
        if not mine:
            return {}
        eng.flush_relaxation()
        eng.sort_by_mode(pools=False)           # live particles in [0, n_live), free list dropped
        n_live, _ = eng.slot_count()
        n_out = sum(n for _, n in mine)
        if n_out > n_live:
            raise ValueError("rebalance plan asks for more particles than this rank holds")
        t = eng.t
        pick = torch.div(torch.arange(n_out, device=eng.device, dtype=torch.int64) * n_live, n_out, rounding_mode="floor")
        out, lo = {}, 0
        for dst, n in mine:
            rows = pick[lo:lo + n]
            f = torch.stack([t[k][rows] if k != "pid" else t[k][rows].view(torch.float64) for k in F64_FIELDS])
            i = torch.stack([t[k][rows] for k in I32_FIELDS])
            out[dst] = (f.contiguous(), i.contiguous())
            lo += n
        t["mode"][pick] = -1
        eng.sort_by_mode()                      # compacts the holes away
        return out

    def insert_from(self, blocks):
        """Append the rows received from other ranks ([(f64 block, i32 block)]) and restore the mode order."""
        eng = self.engine
        blocks = [b for b in blocks if b[0].shape[1] > 0]
        if not blocks:
            return 0
        eng.flush_relaxation()
        eng.sort_by_mode(pools=False)
        n_live, _ = eng.slot_count()
        n_in = sum(b[0].shape[1] for b in blocks)
        if n_live + n_in > eng.cap:
            raise ValueError(f"rebalance: {n_live} + {n_in} particles exceed the capacity {eng.cap} of rank {self.rank}")
        t, lo = eng.t, n_live
        for f, i in blocks:
            n = f.shape[1]
            for a, k in enumerate(F64_FIELDS):
                t[k][lo:lo + n] = f[a] if k != "pid" else f[a].view(torch.int64)
            for a, k in enumerate(I32_FIELDS):
                t[k][lo:lo + n] = i[a]
            lo += n
        torch.cuda.synchronize(eng.device)
        check(eng.ctx, eng.L.nk_set_slot_count(eng.ctx, lo), "nk_set_slot_count")
        eng.sort_by_mode()
        return n_in

    def rebalance(self, tolerance=0.02):
        """All-gather the live counts; if the spread exceeds `tolerance` x mean, migrate rows (NCCL send / recv) so that
        every rank holds total / world particles.  Call between timesteps.  Returns the number of rows this rank moved."""
        if self.world == 1:
            return 0
        counts = self.live_counts()
        mean = sum(counts) / self.world
        if mean == 0 or (max(counts) - min(counts)) <= tolerance * mean:
            return 0
        plan = rebalance_plan(counts)
        dev = self.engine.device
        out = self.extract_for(plan)
        ops, recv = [], []
        for src, dst, n in plan:
            if n == 0:
                continue
            if src == self.rank:
                f, i = out[dst]
                ops += [dist.P2POp(dist.isend, f, dst, group=self.group), dist.P2POp(dist.isend, i, dst, group=self.group)]
            elif dst == self.rank:
                f = torch.empty((len(F64_FIELDS), n), dtype=torch.float64, device=dev)
                i = torch.empty((len(I32_FIELDS), n), dtype=torch.int32, device=dev)
                recv.append((f, i))
                ops += [dist.P2POp(dist.irecv, f, src, group=self.group), dist.P2POp(dist.irecv, i, src, group=self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        moved = sum(b[0].shape[1] for b in out.values()) + self.insert_from(recv)
        if self.fused:
            self._synced = False                # ranks leave the maintenance pass at different times
        return moved
