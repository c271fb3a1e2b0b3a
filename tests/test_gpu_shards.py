"""Particle shards on the real kernels: two contexts (ranks 0 and 1 of 2, emulated on one GPU) own half of
the particles and half of every reservoir's mode table; their accumulator vectors are summed between
nk_step_local and nk_step_finalize exactly as the NCCL all-reduce does in nanokappa_b200.parallel.  The union of
the shards must equal the single-context run: identical census / modes / facets, temperatures to 1e-12."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import gen_golden

pytestmark = pytest.mark.gpu
SEED = 77


def _engine(tb, st, rows, rank=None, world=None):
    from nanokappa_b200.engine import Engine
    from nanokappa_b200._lib import check
    J = tb["omega"].shape[1]
    eng = Engine(0, seed=SEED)
    eng.set_tables(tb, res_counter=st.res_counter)
    eng.allocate(2 * st.positions.shape[0] + 64)
    eng.load_particles(st.positions[rows], (st.modes[:, 0] * J + st.modes[:, 1])[rows], st.occupation[rows], ids=st.ids[rows],
                       omodes=st.omega_modes[rows], n_timesteps=st.n_timesteps[rows], collision_facets=st.collision_facets[rows],
                       collision_positions=st.collision_positions[rows])
    eng.set_sv_temperature(st.subvol_temperature)
    eng.set_timestep(0)
    if world is not None:
        check(eng.ctx, eng.L.nk_set_rank(eng.ctx, rank, world), "nk_set_rank")
    return eng


def _acc(eng):
    from nanokappa_b200._lib import check
    ptr, ln = C.c_void_p(), C.c_int64()
    check(eng.ctx, eng.L.nk_acc_buffer(eng.ctx, C.byref(ptr), C.byref(ln)), "nk_acc_buffer")

    class _A:
        __cuda_array_interface__ = {"shape": (ln.value,), "typestr": "<f8", "data": (ptr.value, False), "version": 3}
    return torch.as_tensor(_A(), device=eng.device)


@pytest.mark.parametrize("name", ["c2_crossplane", "c1_mixed", "c5_box_grid_radial", "c7_fixed_rate", "c8_one_to_one"])
def test_two_shards_equal_one_context(name, golden_dir):
    tb, st, _ = gen_golden.load_fixture(os.path.join(golden_dir, name + ".npz"))
    n = st.positions.shape[0]
    single = _engine(tb, st, slice(0, n))
    shards = [_engine(tb, st, slice(0, n // 2), 0, 2), _engine(tb, st, slice(n // 2, n), 1, 2)]
    accs = [_acc(e) for e in shards]
    steps = 25
    single.step(steps)
    for _ in range(steps):
        for e in shards:
            e.step_local()
        torch.cuda.synchronize()
        total = accs[0] + accs[1]
        for a in accs:
            a.copy_(total)
        for e in shards:
            e.step_finalize()
    ps = single.particles()
    parts = [e.particles() for e in shards]
    ids = np.concatenate([p["ids"] for p in parts])
    assert np.unique(ids).shape[0] == ids.shape[0]
    order = np.argsort(ids)
    assert np.array_equal(ids[order], ps["ids"])
    cat = lambda k: np.concatenate([p[k] for p in parts])[order]
    assert np.array_equal(cat("modes"), ps["modes"])
    assert np.array_equal(cat("collision_facets"), ps["collision_facets"])
    assert np.allclose(cat("positions"), ps["positions"], rtol=1e-13, atol=1e-9, equal_nan=True)
    assert np.allclose(cat("occupation"), ps["occupation"], rtol=1e-9, atol=0)
    rs = single.results()
    for e in shards:
        r = e.results()
        assert np.array_equal(r["subvol_N_p"], rs["subvol_N_p"]) and np.array_equal(r["N_leaving"], rs["N_leaving"])
        assert np.allclose(r["subvol_temperature"], rs["subvol_temperature"], rtol=1e-12, atol=0)
    assert sum(e.slot_count()[1] for e in shards) == single.slot_count()[1]


def test_rebalance_between_emulated_ranks_keeps_the_union(golden_dir):
    """SURVEY 8e periodic rebalance: an 80/20 split of the particles, migrated to 50/50 half way through the run by
    ShardedEngine.extract_for / insert_from (the NCCL send/recv in between is replaced by handing the blocks over),
    must still give the single-context census, modes, facets and temperatures."""
    from nanokappa_b200.parallel import ShardedEngine, rebalance_plan
    tb, st, _ = gen_golden.load_fixture(os.path.join(golden_dir, "c1_mixed.npz"))
    n = st.positions.shape[0]
    cut = int(0.8 * n)
    single = _engine(tb, st, slice(0, n))
    shards = [ShardedEngine(_engine(tb, st, slice(0, cut)), 0, 2), ShardedEngine(_engine(tb, st, slice(cut, n)), 1, 2)]
    accs = [_acc(s.engine) for s in shards]

    def run(k):
        for _ in range(k):
            for s in shards:
                s.engine.step_local()
            torch.cuda.synchronize()
            total = accs[0] + accs[1]
            for a in accs:
                a.copy_(total)
            for s in shards:
                s.engine.step_finalize()

    single.step(24)
    run(12)
    counts = [s.engine.slot_count()[1] for s in shards]
    plan = rebalance_plan(counts)
    assert plan and plan[0][0] == 0 and plan[0][1] == 1
    out = shards[0].extract_for(plan)
    assert shards[1].extract_for(plan) == {}
    moved = shards[1].insert_from([out[1]])
    after = [s.engine.slot_count()[1] for s in shards]
    assert moved == plan[0][2] and sum(after) == sum(counts) and abs(after[0] - after[1]) <= 1
    run(12)
    ps = single.particles()
    parts = [s.engine.particles() for s in shards]
    ids = np.concatenate([p["ids"] for p in parts])
    order = np.argsort(ids)
    assert np.array_equal(ids[order], ps["ids"])
    cat = lambda k: np.concatenate([p[k] for p in parts])[order]
    assert np.array_equal(cat("modes"), ps["modes"]) and np.array_equal(cat("omega_modes"), ps["omega_modes"])
    assert np.array_equal(cat("collision_facets"), ps["collision_facets"])
    assert np.allclose(cat("positions"), ps["positions"], rtol=1e-13, atol=1e-9, equal_nan=True)
    assert np.allclose(cat("occupation"), ps["occupation"], rtol=1e-9, atol=0)
    rs = single.results()
    for s in shards:
        r = s.engine.results()
        assert np.array_equal(r["subvol_N_p"], rs["subvol_N_p"])
        assert np.allclose(r["subvol_temperature"], rs["subvol_temperature"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", ["c2_crossplane", "c8_one_to_one"])
def test_checkpoint_of_two_shards_continues_exactly(name, golden_dir):
    """a sharded run that is checkpointed in the middle of a convergence window (step 7 of 10) and restored into
    fresh contexts WITHOUT rebuilding the tables (Engine.checkpoint / Engine.restore: particles, reservoir counters and
    deal counters, window accumulators, N_leaving, results block) must continue exactly like the uninterrupted run --
    including the reservoir balances of the convergence row at step 10 and, for one_to_one, the re-emission that depends on
    the previous step's absorbed counts."""
    tb, st, _ = gen_golden.load_fixture(os.path.join(golden_dir, name + ".npz"))
    n = st.positions.shape[0]

    def make():
        sh = [_engine(tb, st, slice(0, n // 2), 0, 2), _engine(tb, st, slice(n // 2, n), 1, 2)]
        return sh, [_acc(e) for e in sh]

    def run(shards, accs, k):
        for _ in range(k):
            for e in shards:
                e.step_local()
            torch.cuda.synchronize()
            total = accs[0] + accs[1]
            for a in accs:
                a.copy_(total)
            for e in shards:
                e.step_finalize()

    ref, ref_acc = make()
    run(ref, ref_acc, 15)
    a, a_acc = make()
    run(a, a_acc, 7)
    saved = [e.checkpoint() for e in a]
    assert all(int(z["current_timestep"]) == 7 for z in saved)
    b, b_acc = make()                       # fresh contexts: tables set, rank set, nothing else
    for e, z in zip(b, saved):
        e.restore(z)
    run(b, b_acc, 8)
    for e_ref, e in zip(ref, b):
        p0, p1 = e_ref.particles(), e.particles()
        for f in p0:
            assert np.array_equal(p0[f], p1[f], equal_nan=True), f"{f} differs after the restart"
        r0, r1 = e_ref.results(), e.results()
        for f in ("subvol_temperature", "subvol_energy", "subvol_N_p", "subvol_heat_flux", "res_energy_balance", "res_heat_flux", "N_leaving"):
            assert np.array_equal(r0[f], r1[f]), f"{f} differs after the restart"
        s0, s1 = e_ref.run_state(), e.run_state()
        for f in s0:
            assert np.array_equal(s0[f], s1[f], equal_nan=True), f"run state {f} differs after the restart"      # kappa_sv is 0/0 in flat slices


def test_host_buffer_calls_recycle_holes_at_two_percent_headroom(golden_dir):
    """step-by-step integration through nk_advance_host with only 2 % spare capacity.  Absorbed particles leave
    holes (mode = -1) in the caller's arrays; the census of every call puts them on the free-slot ring, so emitted particles
    reuse them and the slot range stays bounded over 200 calls (it used to grow by every emission until NK_ERR_CAPACITY)."""
    from nanokappa_b200.engine import Engine
    from nanokappa_b200._lib import check
    tb, st, _ = gen_golden.load_fixture(os.path.join(golden_dir, "c2_crossplane.npz"))
    n0 = st.positions.shape[0]
    J = tb["omega"].shape[1]
    eng = Engine(0, seed=SEED)
    eng.set_tables(tb, res_counter=st.res_counter)
    eng.allocate(int(n0 * 1.02))
    eng.load_particles(st.positions, st.modes[:, 0] * J + st.modes[:, 1], st.occupation, ids=st.ids, omodes=st.omega_modes,
                       n_timesteps=st.n_timesteps, collision_facets=st.collision_facets, collision_positions=st.collision_positions)
    eng.set_sv_temperature(st.subvol_temperature)
    eng.set_timestep(0)
    names = ("px", "py", "pz", "tc", "occ", "mode", "omode", "cfacet", "cx", "cy", "cz", "pid")
    host = {k: eng.t[k].cpu().clone().pin_memory() for k in names}
    S = eng.S
    Tsv = np.zeros(S); Esv = np.zeros(S); Nsv = np.zeros(S, dtype=np.int64)
    hp = lambda k: C.c_void_p(host[k].data_ptr())
    n, peak = n0, n0
    for call in range(200):
        n_out = C.c_int64()
        check(eng.ctx, eng.L.nk_advance_host(eng.ctx, n, 1, *[hp(k) for k in names], C.byref(n_out), Tsv.ctypes.data_as(C.c_void_p),
                                             Esv.ctypes.data_as(C.c_void_p), Nsv.ctypes.data_as(C.c_void_p)), "nk_advance_host")
        n = n_out.value
        peak = max(peak, n)
        live = int((host["mode"][:n] >= 0).sum())
        assert live == int(Nsv.sum())
    ids = host["pid"][:n][host["mode"][:n] >= 0].numpy()
    emitted_alive = int((ids >= 2 ** 62).sum())                 # reservoir particles carry ids above 2^62
    assert emitted_alive > 0.05 * n0, "the run did not emit enough particles to exercise the recycling"
    assert peak <= eng.cap and peak <= int(n0 * 1.02) + 512, f"slot range grew to {peak} of capacity {eng.cap}"
    assert np.unique(ids).shape[0] == ids.shape[0]
