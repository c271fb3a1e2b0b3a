"""Launched under torchrun on >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_multi_gpu_check.py

Every rank owns a block of the fixture's particles and every world-th copy that a reservoir-table entry emits.  The same shards are
stepped twice -- per-step NCCL all-reduce between nk_step_local / nk_step_finalize, and the fused in-kernel
exchange over NVLink peer memory -- and both must agree with each other (integers identical, temperatures to
1e-13) and, gathered on rank 0, with the single-context run of the whole population; the fused run, whose exchange adds
the ranks' fixed-point sums exactly, must reproduce the single-context temperatures and energies BIT FOR BIT."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanokappa_b200.engine import Engine            # noqa: E402
from nanokappa_b200.parallel import ShardedEngine, shard_bounds   # noqa: E402
from oracle import gen_golden                        # noqa: E402

STEPS, SEED = 40, 9


def make(tb, st, rows, dev):
    J = tb["omega"].shape[1]
    eng = Engine(dev, seed=SEED)
    eng.set_tables(tb, res_counter=st.res_counter)
    eng.allocate(2 * st.positions.shape[0] + 64)
    eng.load_particles(st.positions[rows], (st.modes[:, 0] * J + st.modes[:, 1])[rows], st.occupation[rows], ids=st.ids[rows],
                       omodes=st.omega_modes[rows], n_timesteps=st.n_timesteps[rows], collision_facets=st.collision_facets[rows],
                       collision_positions=st.collision_positions[rows])
    eng.set_sv_temperature(st.subvol_temperature)
    eng.set_timestep(0)
    return eng


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name in ("c2_crossplane", "c1_mixed"):
        tb, st, _ = gen_golden.load_fixture(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        lo, hi = shard_bounds(rank, world, st.positions.shape[0])
        runs = {}
        for mode in ("nccl", "fused"):
            sh = ShardedEngine(make(tb, st, slice(lo, hi), local), rank, world)
            if mode == "fused":
                assert sh.enable_fused_exchange(), "peer mailboxes could not be mapped"
            sh.step(STEPS)
            runs[mode] = (sh.engine.particles(), sh.engine.results())
            sh.close()
        (pn, rn), (pf, rf) = runs["nccl"], runs["fused"]
        for k in ("ids", "modes", "collision_facets", "positions", "n_timesteps"):
            ok &= bool(np.array_equal(pn[k], pf[k], equal_nan=True))
        ok &= bool(np.array_equal(rn["subvol_N_p"], rf["subvol_N_p"]))
        ok &= bool(np.allclose(rn["subvol_temperature"], rf["subvol_temperature"], rtol=1e-13, atol=0))
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: pf[k] for k in ("ids", "modes", "collision_facets")})
        if rank == 0:
            single = make(tb, st, slice(0, st.positions.shape[0]), local)
            single.step(STEPS)
            ps, rs = single.particles(), single.results()
            ids = np.concatenate([g["ids"] for g in gathered]); order = np.argsort(ids)
            ok &= bool(np.array_equal(ids[order], ps["ids"]))
            ok &= bool(np.array_equal(np.concatenate([g["modes"] for g in gathered])[order], ps["modes"]))
            ok &= bool(np.array_equal(np.concatenate([g["collision_facets"] for g in gathered])[order], ps["collision_facets"]))
            ok &= bool(np.array_equal(rf["subvol_N_p"], rs["subvol_N_p"]))
            ok &= bool(np.allclose(rf["subvol_temperature"], rs["subvol_temperature"], rtol=1e-12, atol=0))
            # the fused exchange adds the ranks' fixed-point sums exactly: the sharded run IS the single-GPU run
            exact = bool(np.array_equal(rf["subvol_temperature"], rs["subvol_temperature"]) and np.array_equal(rf["subvol_energy"], rs["subvol_energy"]))
            ok &= exact
            print(f"[{name}] world={world} fused==nccl==single: {ok}  fused bit-identical to single: {exact}  N_p={rs['N_p']}", flush=True)
        # unbalanced shards (rank 0 owns 70 %), rebalanced by NCCL point-to-point migration half way: same union
        n_all = st.positions.shape[0]
        cut = [0] + [int(n_all * (0.7 + 0.3 * r / (world - 1))) for r in range(world)]
        cut[-1] = n_all
        sh = ShardedEngine(make(tb, st, slice(cut[rank], cut[rank + 1]), local), rank, world)
        assert sh.enable_fused_exchange(), sh.fused_error
        sh.step(STEPS // 2)
        before = sh.live_counts()
        moved = sh.rebalance(tolerance=0.01)
        after = sh.live_counts()
        sh.step(STEPS - STEPS // 2)
        pr, rr = sh.engine.particles(), sh.engine.results()
        ok &= bool(max(after) - min(after) <= 1 and sum(after) == sum(before) and (moved > 0 or world == 1))
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: pr[k] for k in ("ids", "modes", "collision_facets")})
        if rank == 0:
            ids = np.concatenate([g["ids"] for g in gathered]); order = np.argsort(ids)
            ok &= bool(np.array_equal(ids[order], ps["ids"]))
            ok &= bool(np.array_equal(np.concatenate([g["modes"] for g in gathered])[order], ps["modes"]))
            ok &= bool(np.array_equal(np.concatenate([g["collision_facets"] for g in gathered])[order], ps["collision_facets"]))
            ok &= bool(np.array_equal(rr["subvol_N_p"], rs["subvol_N_p"]))
            ok &= bool(np.allclose(rr["subvol_temperature"], rs["subvol_temperature"], rtol=1e-12, atol=0))
            print(f"[{name}] rebalance {before} -> {after}: {ok}", flush=True)
        sh.close()
    flag = torch.tensor([int(ok)], device=torch.device("cuda", local))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
