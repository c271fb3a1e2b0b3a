"""GPU parity: the CUDA path through the C ABI against the oracle restatement on identical inputs
and identical keyed random draws.  Integer results (facet, subvolume, mode, particle census) must be
bit-exact; floating-point per-particle state within 1e-9 and per-SV T / energy / flux / kappa within
1e-6 relative (north_star tolerance; observed ~1e-13)."""
import os

import numpy as np
import pytest

from oracle import gen_golden, nk_oracle as nko

pytestmark = pytest.mark.gpu

FIXTURES = sorted(gen_golden.CONFIGS)
SEED = 2024
RTOL_PARTICLE = 1e-9
RTOL_SV = 1e-6


def _engine(tb, st, cap_factor=2.0, seed=SEED):
    from nanokappa_b200.engine import Engine
    J = tb["omega"].shape[1]
    eng = Engine(0, seed=seed)
    eng.set_tables(tb, res_counter=st.res_counter)
    eng.allocate(int(st.positions.shape[0] * cap_factor) + 64)
    eng.load_particles(st.positions, st.modes[:, 0] * J + st.modes[:, 1], st.occupation, ids=st.ids,
                       omodes=st.omega_modes, n_timesteps=st.n_timesteps, collision_facets=st.collision_facets,
                       collision_positions=st.collision_positions)
    eng.set_sv_temperature(st.subvol_temperature)
    eng.set_timestep(0)
    return eng


def _load(name, golden_dir):
    return gen_golden.load_fixture(os.path.join(golden_dir, name + ".npz"))


def _close(name, got, want, rtol, atol=0.0):
    got = np.asarray(got, dtype=float); want = np.asarray(want, dtype=float)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    both_nan = np.isnan(got) & np.isnan(want)
    same_inf = np.isinf(want) & (got == want)
    ok = both_nan | same_inf | (np.abs(got - want) <= atol + rtol * np.abs(want))
    assert ok.all(), f"{name}: {np.count_nonzero(~ok)} of {ok.size} outside rtol {rtol}; worst {np.nanmax(np.abs(got - want)[~ok])}"


@pytest.mark.parametrize("name", ["c1_mixed", "c2_crossplane"])
def test_find_boundary_operator(name, golden_dir):
    tb, st, _ = _load(name, golden_dir)
    eng = _engine(tb, st)
    r = np.random.default_rng(1)
    lo, hi = tb["bounds"]
    x = lo + r.random((20000, 3)) * (hi - lo)
    v = r.standard_normal((20000, 3)) * 50
    v[:200, 1] = 0.0; v[200:400, 2] = 0.0; v[400:500, 0] = 0.0          # axis-parallel rays
    x = np.vstack([x, st.positions, st.collision_positions[np.isfinite(st.collision_positions).all(axis=1)][:1000]])
    v = np.vstack([v, st.group_vel, st.group_vel[np.isfinite(st.collision_positions).all(axis=1)][:1000]])
    with np.errstate(all="ignore"):
        xc0, tc0, fc0 = nko.find_boundary(tb, x, v)
    xc, tc, fc = eng.find_boundary(x, v)
    assert np.array_equal(fc, fc0), f"{np.count_nonzero(fc != fc0)} facet mismatches"
    _close("tc", tc, tc0, 1e-12)
    _close("xc", xc, xc0, 1e-12, atol=1e-9)
    assert (fc0 == -1).sum() < x.shape[0]


def test_find_boundary_empty(golden_dir):
    tb, st, _ = _load("c1_mixed", golden_dir)
    eng = _engine(tb, st)
    xc, tc, fc = eng.find_boundary(np.zeros((0, 3)), np.zeros((0, 3)))
    assert xc.shape == (0, 3) and tc.shape == (0,) and fc.shape == (0,)


@pytest.mark.parametrize("name", FIXTURES)
def test_init_collisions_matches_reference_init(name, golden_dir):
    """nk_init_collisions against the arrays the reference produced in Population.__init__."""
    tb, st, _ = _load(name, golden_dir)
    eng = _engine(tb, st)
    eng.init_collisions()
    p = eng.particles(flush=False)
    assert np.array_equal(p["collision_facets"], st.collision_facets)
    _close("n_timesteps", p["n_timesteps"], st.n_timesteps, 1e-12)
    _close("collision_positions", p["collision_positions"], st.collision_positions, 1e-12, atol=1e-9)


@pytest.mark.parametrize("name", ["c1_mixed", "c2_crossplane"])
def test_table_operators(name, golden_dir):
    tb, st, _ = _load(name, golden_dir)
    eng = _engine(tb, st)
    r = np.random.default_rng(2)
    lo, hi = tb["bounds"]
    x = lo - 0.05 * (hi - lo) + r.random((50000, 3)) * (hi - lo) * 1.1
    sv, counts = eng.classify(x, counts=True)
    sv0 = nko.classify(tb, x)
    assert np.array_equal(sv, sv0)
    assert np.array_equal(counts, np.bincount(sv0, minlength=tb["sv_centres"].shape[0]))
    Q, J = tb["omega"].shape
    T = 200 + 200 * r.random(20000)
    q = r.integers(0, Q, 20000); j = r.integers(0, J, 20000)
    om = tb["omega"][q, j]
    _close("occupation", eng.calculate_occupation(T, om), nko.calculate_occupation(tb, T, om), 1e-13)
    assert (eng.calculate_occupation(np.array([0.0, -1.0, 300.0]), np.array([1.0, 1.0, 0.0])) == 0).all()
    _close("lifetime", eng.lifetime_function(np.stack([T, q, j], axis=1)),
           nko.lifetime_function(tb, T, np.stack([q, j], axis=1)), 1e-14)
    Tg = tb["T_grid"]
    Tn = np.concatenate([Tg[:-1], Tg[1:]])            # exactly on the grid nodes
    qn = r.integers(0, Q, Tn.shape[0]); jn = r.integers(0, J, Tn.shape[0])
    _close("lifetime@nodes", eng.lifetime_function(np.stack([Tn, qn, jn], axis=1)),
           nko.lifetime_function(tb, Tn, np.stack([qn, jn], axis=1)), 1e-14)
    E = np.concatenate([tb["energy_array"][[0, 5, -1]], tb["energy_array"][0] + r.random(5000) * np.ptp(tb["energy_array"]),
                        [tb["energy_array"][0] - 1.0, tb["energy_array"][-1] + 1.0]])
    _close("T(E)", eng.temperature_function(E), nko.temperature_function(tb, E), 1e-14)
    Tq = np.concatenate([[-3.0, 0.0, 1000.0, 1500.0], 1000 * r.random(5000)])
    _close("E(T)", eng.crystal_energy_function(Tq), nko.crystal_energy_function(tb, Tq), 1e-14)
    T_sv = 298 + 4 * r.random(tb["sv_centres"].shape[0])
    eng.set_sv_temperature(T_sv)
    _close("particle T", eng.particle_temperature(x), nko.particle_temperature(tb, T_sv, x), 1e-14)


@pytest.mark.parametrize("name", ["c5_box_grid_radial", "c6_cylinder_voronoi_radial"])
def test_particle_temperature_radial(name, golden_dir):
    """--temp_interp radial: the device's cubic RBF field (weights factorised at set-up, coefficients refreshed with
    T_sv) against scipy's RBFInterpolator, inside and outside the mesh, after set_sv_temperature and after steps."""
    tb, st, _ = _load(name, golden_dir)
    eng = _engine(tb, st)
    r = np.random.default_rng(5)
    lo, hi = tb["bounds"]
    x = lo - 0.05 * (hi - lo) + r.random((20000, 3)) * (hi - lo) * 1.1
    T_sv = 298 + 4 * r.random(tb["sv_centres"].shape[0])
    eng.set_sv_temperature(T_sv)
    _close("particle T (radial)", eng.particle_temperature(x), nko.particle_temperature(tb, T_sv, x), 1e-9)
    eng.set_sv_temperature(st.subvol_temperature)
    eng.step(3)
    T_now = eng.results()["subvol_temperature"]
    _close("particle T after steps", eng.particle_temperature(x), nko.particle_temperature(tb, T_now, x), 1e-9)
    sv = eng.classify(x)
    assert np.array_equal(sv, nko.classify(tb, x))


def _compare_step(k, eng, st, tb):
    p = eng.particles()
    order = np.argsort(st.ids)
    assert np.array_equal(p["ids"], st.ids[order]), f"step {k}: particle census differs (gpu {p['ids'].shape[0]}, oracle {st.ids.shape[0]})"
    if not np.array_equal(p["modes"], st.modes[order]):
        bad = np.nonzero((p["modes"] != st.modes[order]).any(axis=1))[0]
        J = tb["omega"].shape[1]
        raise AssertionError(f"step {k}: {bad.shape[0]} mode indices differ; ids {p['ids'][bad][:5]} gpu {p['modes'][bad][:5].tolist()} "
                             f"oracle {st.modes[order][bad][:5].tolist()} gpu omode {p['omega_modes'][bad][:5]} oracle omode {st.omega_modes[order][bad][:5]} "
                             f"facets gpu {p['collision_facets'][bad][:5]} oracle {st.collision_facets[order][bad][:5]}")
    assert np.array_equal(p["omega_modes"], st.omega_modes[order]), f"step {k}: omega-carrying modes differ"
    assert np.array_equal(p["collision_facets"], st.collision_facets[order]), f"step {k}: collision facets differ"
    _close(f"step {k} positions", p["positions"], st.positions[order], RTOL_PARTICLE, atol=1e-9)
    _close(f"step {k} occupation", p["occupation"], st.occupation[order], RTOL_PARTICLE)
    _close(f"step {k} n_timesteps", p["n_timesteps"], st.n_timesteps[order], RTOL_PARTICLE, atol=1e-9)
    _close(f"step {k} collision_positions", p["collision_positions"], st.collision_positions[order], RTOL_PARTICLE, atol=1e-9)
    res = eng.results()
    assert np.array_equal(res["subvol_N_p"], st.subvol_N_p), f"step {k}: per-SV particle counts differ"
    assert np.array_equal(res["N_leaving"], st.N_leaving), f"step {k}: absorbed counts differ"
    _close(f"step {k} T_sv", res["subvol_temperature"], st.subvol_temperature, RTOL_SV)
    _close(f"step {k} E_sv", res["subvol_energy"], st.subvol_energy, RTOL_SV)
    _close(f"step {k} res_counter", eng.res_counter().reshape(st.res_counter.shape), st.res_counter, 1e-12, atol=1e-12)
    return res


@pytest.mark.parametrize("name", FIXTURES)
def test_step_parity_fixed_draws(name, golden_dir):
    tb, st, _ = _load(name, golden_dir)
    eng = _engine(tb, st)
    rng = nko.KeyedRNG(SEED)
    conv = {}
    n_steps = 30
    with np.errstate(all="ignore"):
        for k in range(1, n_steps + 1):
            nko.run_timestep(tb, st, rng, on_convergence=lambda s: conv.update(
                subvol_heat_flux=s.subvol_heat_flux.copy(), res_heat_flux=s.res_heat_flux.copy(),
                res_energy_balance=s.res_energy_balance.copy(),
                subvol_kappa=None if s.subvol_kappa is None else s.subvol_kappa.copy(), kappa=s.kappa))
            eng.step(1)
            if k in (1, 2, 5, 10, 20, 30):
                res = _compare_step(k, eng, st, tb)
                if k % tb["n_dt_to_conv"] == 0:
                    scale = np.abs(conv["subvol_heat_flux"]).max()
                    _close(f"step {k} heat flux", res["subvol_heat_flux"], conv["subvol_heat_flux"], RTOL_SV, atol=RTOL_SV * scale)
                    _close(f"step {k} res flux", res["res_heat_flux"], conv["res_heat_flux"], RTOL_SV, atol=RTOL_SV * np.abs(conv["res_heat_flux"]).max())
                    _close(f"step {k} res balance", res["res_energy_balance"], conv["res_energy_balance"], RTOL_SV, atol=RTOL_SV * np.abs(conv["res_energy_balance"]).max())
                    if not tb["sv_slice"]:
                        continue            # per-connection kappa of grid / voronoi subvolumes is computed on the host
                    _close(f"step {k} heat flux", res["subvol_heat_flux"], conv["subvol_heat_flux"], RTOL_SV, atol=RTOL_SV * scale)
                    # kappa_s = -phi_s dx / (T[s+1] - T[s-1]) is ill-conditioned where the profile is still flat
                    # (dT ~ 1e-10 K in the cold slices): compare kappa_s * dT (= -phi_s dx, well conditioned) and
                    # allow what a 2e-9 K disagreement in T (7e-12 relative) does to the quotient.
                    Tpad = np.concatenate(([tb["res_T"][0]], st.subvol_temperature, [tb["res_T"][-1]]))
                    dT = Tpad[2:] - Tpad[:-2]
                    kref, kgpu = conv["subvol_kappa"], res["subvol_kappa"]
                    num_ref, num_gpu = kref * dT, kgpu * dT
                    tol = RTOL_SV * np.abs(num_ref).max() + 2e-9 * np.abs(kref) + 1e-300
                    bad = np.abs(num_gpu - num_ref) > tol
                    assert not bad.any(), f"step {k} kappa_sv: {kgpu[bad]} vs {kref[bad]} (dT {dT[bad]})"
                    _close(f"step {k} kappa", res["kappa"], conv["kappa"], RTOL_SV)
                    _close(f"step {k} res flux", res["res_heat_flux"], conv["res_heat_flux"], RTOL_SV, atol=RTOL_SV * np.abs(conv["res_heat_flux"]).max())
                    _close(f"step {k} res balance", res["res_energy_balance"], conv["res_energy_balance"], RTOL_SV, atol=RTOL_SV * np.abs(conv["res_energy_balance"]).max())
    assert eng.timestep() == n_steps


def test_step_parity_from_an_empty_population(golden_dir):
    """Edge case: no particle at all at step 0 (and therefore empty subvolumes for many steps).  The reservoirs fill the film
    from both ends; the reference's per-SV normalisation divides by N_s = 0 in the empty slices, so their temperatures are
    NaN (Population.py:716-728) and stay NaN once the first particles arrive -- the CUDA path must reproduce the census, the
    per-particle state and the NaN pattern, not crash on the empty hit / emission lists of step 1."""
    tb, st, _ = _load("c2_crossplane", golden_dir)
    for f in ("positions", "modes", "omega", "group_vel", "occupation", "n_timesteps", "collision_facets", "collision_positions",
              "collision_cond", "temperatures", "ids", "omega_modes", "subvol_id", "energies"):
        v = getattr(st, f)
        if v is not None:
            setattr(st, f, v[:0].copy())
    from nanokappa_b200.engine import Engine
    eng = Engine(0, seed=SEED)
    eng.set_tables(tb, res_counter=st.res_counter)
    eng.allocate(4096)
    J = tb["omega"].shape[1]
    eng.load_particles(st.positions, st.modes[:, 0] * J + st.modes[:, 1], st.occupation, ids=st.ids, omodes=st.omega_modes,
                       n_timesteps=st.n_timesteps, collision_facets=st.collision_facets, collision_positions=st.collision_positions)
    eng.set_sv_temperature(st.subvol_temperature)
    eng.set_timestep(0)
    assert eng.slot_count() == (0, 0)
    rng = nko.KeyedRNG(SEED)
    with np.errstate(all="ignore"):
        for k in range(1, 16):
            nko.run_timestep(tb, st, rng)
            eng.step(1)
            if k in (1, 2, 5, 10, 15):
                res = _compare_step(k, eng, st, tb)
                assert np.array_equal(np.isnan(res["subvol_temperature"]), np.isnan(st.subvol_temperature))
    assert st.ids.shape[0] > 0 and np.isnan(st.subvol_temperature).any()


def test_multi_step_call_equals_single_steps(golden_dir):
    """nk_step(n) in one call == n calls, and flushing the deferred relaxation in between is neutral."""
    tb, st, _ = _load("c1_mixed", golden_dir)
    a = _engine(tb, st.copy()); b = _engine(tb, st.copy())
    a.step(12)
    for _ in range(12):
        b.step(1); b.flush_relaxation()
    pa, pb = a.particles(), b.particles()
    for f in pa:
        assert np.array_equal(pa[f], pb[f], equal_nan=True), f
    ra, rb = a.results(), b.results()
    assert np.array_equal(ra["subvol_N_p"], rb["subvol_N_p"])
    _close("T", ra["subvol_temperature"], rb["subvol_temperature"], 1e-13)


def test_slot_recycling_and_census(golden_dir):
    """Absorbed particles free their slot, emitted ones reuse it: slots stay bounded, live count = N_p."""
    tb, st, _ = _load("c2_crossplane", golden_dir)
    eng = _engine(tb, st, cap_factor=1.2)
    n0 = st.positions.shape[0]
    eng.step(60)
    slots, alive = eng.slot_count()
    res = eng.results()
    assert alive == res["N_p"]
    assert slots <= n0 * 1.1 + 64
    p = eng.particles()
    assert p["ids"].shape[0] == alive and np.unique(p["ids"]).shape[0] == alive


def test_capacity_overflow_is_reported(golden_dir):
    tb, st, _ = _load("c2_crossplane", golden_dir)
    from nanokappa_b200._lib import NkError
    eng = _engine(tb, st, cap_factor=1.0)
    eng.allocate(st.positions.shape[0])
    J = tb["omega"].shape[1]
    eng.load_particles(st.positions, st.modes[:, 0] * J + st.modes[:, 1], st.occupation, ids=st.ids,
                       n_timesteps=np.full(st.positions.shape[0], 1e9), collision_facets=st.collision_facets,
                       collision_positions=st.collision_positions)        # nobody ever leaves, emission must overflow
    eng.set_sv_temperature(st.subvol_temperature)
    eng.step(400)          # capacity is rounded up to whole 512-slot tiles: needs > 144 net emissions
    with pytest.raises(NkError):
        eng.slot_count()


@pytest.mark.parametrize("name", ["c2_crossplane", "c4_cylinder_voronoi", "c1_mixed", "c8_one_to_one"])
def test_step_kernel_variants_agree(name, golden_dir, monkeypatch):
    """Every implementation of a step -- streaming kernel with direct arithmetic or per-(mode, subvolume) tables, rare path
    with one thread per item or with block-cooperative triangle tiles, particles in creation order or ordered by mode with
    per-mode slot pools -- must give the same particles: integers identical, occupations identical to the last bit (they
    share one arithmetic), T_sv to 1e-13."""
    tb, st, _ = _load(name, golden_dir)
    results = {}
    keys = ("NK_STEP_TAB", "NK_RARE_TILED", "NK_POOL_MIN")
    for label, env in (("direct", {"NK_STEP_TAB": "0"}), ("tables", {"NK_STEP_TAB": "force"}),
                       ("tiled", {"NK_STEP_TAB": "0", "NK_RARE_TILED": "1"}),
                       ("pools", {"NK_STEP_TAB": "0", "NK_POOL_MIN": "0"}), ("pools_tiled_tables", {"NK_STEP_TAB": "force", "NK_RARE_TILED": "1", "NK_POOL_MIN": "0"})):
        for k in keys:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = _engine(tb, st.copy())
        if label.startswith("pools"):
            eng.sort_by_mode()
        eng.step(11)
        if label.startswith("pools"):
            eng.sort_by_mode()              # a maintenance pass in the middle of the run is neutral, too
        eng.step(12)
        results[label] = (eng.particles(), eng.results())
        eng.close()
    p0, r0 = results["direct"]
    for label, (p, r) in results.items():
        for f in ("ids", "modes", "omega_modes", "collision_facets", "positions", "n_timesteps", "occupation"):
            assert np.array_equal(p[f], p0[f], equal_nan=True), f"{label}: {f} differs from the direct kernel"
        assert np.array_equal(r["subvol_N_p"], r0["subvol_N_p"])
        _close(f"{label} T_sv", r["subvol_temperature"], r0["subvol_temperature"], 1e-13)


def test_sort_by_mode_orders_compacts_and_pools(golden_dir, monkeypatch):
    """nk_sort_by_mode: live particles ordered by mode, every mode region followed by its spare slots, nothing lost; with
    pools the slots freed by absorption are reused by emitted particles of the same mode, so the order stays exact."""
    import torch
    monkeypatch.setenv("NK_POOL_MIN", "0")
    tb, st, _ = _load("c2_crossplane", golden_dir)
    eng = _engine(tb, st.copy(), cap_factor=3.0)
    before = eng.particles(flush=False)
    eng.sort_by_mode()
    n, alive = eng.slot_count()
    assert alive == before["ids"].shape[0] and n > alive                 # spare slots were added
    after = eng.particles(flush=False)
    for f in before:
        assert np.array_equal(before[f], after[f], equal_nan=True), f
    md = eng.t["mode"][:n].cpu().numpy()
    live = md[md >= 0]
    assert (np.diff(live) >= 0).all(), "live particles are not ordered by mode"
    # every run of one mode is followed by free slots (its pool) before the next mode starts
    starts = np.nonzero(np.diff(np.concatenate(([-2], md))) != 0)[0]
    assert (md[starts[1::2]] == -1).all() if md[starts[0]] >= 0 else True
    eng.step(40)
    n2, alive2 = eng.slot_count()
    md2 = eng.t["mode"][:n2].cpu().numpy()
    live2 = md2[md2 >= 0]
    foreign = np.count_nonzero(np.diff(live2) < 0)
    assert n2 == n, "slot range grew although every mode had spare slots"
    assert foreign <= 0.02 * alive2, f"{foreign} of {alive2} particles sit outside the region of their mode after 40 steps"
    eng.sort_by_mode(pools=False)
    n3, alive3 = eng.slot_count()
    assert n3 == alive3 == alive2
    assert bool((eng.t["mode"][:n3] >= 0).all()) and bool((torch.diff(eng.t["mode"][:n3]) >= 0).all())


def test_energy_table_on_device_matches_numpy(golden_dir):
    """Set-up helper nk_energy_table against Phonon.calculate_crystal_energy (NumPy) and the reference's table."""
    import ctypes as C
    from nanokappa_b200 import _lib
    tb, st, _ = _load("c2_crossplane", golden_dir)
    L = _lib.lib()
    om = np.ascontiguousarray(tb["omega"].reshape(-1)); act = np.ascontiguousarray((np.abs(tb["group_vel"]).sum(axis=2) != 0).reshape(-1), dtype=np.uint8)
    T = np.ascontiguousarray(tb["T_array"]); out = np.zeros_like(T)
    Q = tb["omega"].shape[0]
    zero = tb["hbar"] * tb["omega"].sum() / 2 / (Q * tb["volume_unitcell"])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.nk_energy_table(0, om.shape[0], p(om), p(act), T.shape[0], p(T), tb["hbar"], tb["kb"], Q * tb["volume_unitcell"], zero, p(out))
    assert rc == 0, L.nk_last_error(None)
    _close("E(T) table", out, tb["energy_array"], 1e-13)


def test_error_conventions_of_the_c_abi(golden_dir):
    """Every entry point returns 0 / <0 and leaves the message in nk_last_error (the Python layer raises NkError with
    it, as upstream raises bare Exception('...')): calls out of order, bad arguments, capacity violations."""
    import ctypes as C
    from nanokappa_b200._lib import lib, NkError
    from nanokappa_b200.engine import Engine
    L = lib()
    msg = lambda ctx: L.nk_last_error(ctx).decode()
    ctx = C.c_void_p()
    assert L.nk_create(10 ** 6, C.byref(ctx)) != 0 and "device" in L.nk_last_error(None).decode()
    assert L.nk_create(0, C.byref(ctx)) == 0
    assert L.nk_step(ctx, 1) != 0 and "bound" in msg(ctx)                      # nothing set up yet
    c3 = (C.c_double * 9)(*range(9)); v3 = (C.c_double * 3)(1, 1, 1)
    assert L.nk_set_subvols(ctx, 0, c3, v3, 0, 0, 0) != 0 and "n_subvols" in msg(ctx)
    assert L.nk_set_subvols(ctx, 3, c3, v3, 0, 0, 1) != 0 and "slice" in msg(ctx)       # linear needs slices
    assert L.nk_set_subvols(ctx, 3, c3, v3, 1, 0, 2) != 0 and "radial" in msg(ctx)      # radial on slices is singular upstream
    assert L.nk_set_subvols(ctx, 3, c3, v3, 0, 0, 7) != 0
    assert L.nk_set_subvols(ctx, 3, c3, v3, 0, 0, 0) == 0
    assert L.nk_set_rbf(ctx, 3, (C.c_int32 * 3)(0, 1, 2), v3, v3, c3) != 0 and "RADIAL" in msg(ctx)
    assert L.nk_set_reservoir_mode(ctx, 1, None) != 0 and "nk_set_reservoirs" in msg(ctx)
    assert L.nk_set_rank(ctx, 2, 2) != 0
    out8 = (C.c_uint64 * 8)()
    assert L.nk_debug_trace(ctx, out8) != 0 and "NK_TRACE" in msg(ctx)
    L.nk_destroy(ctx)

    tb, st, _ = _load("c2_crossplane", golden_dir)
    eng = _engine(tb, st)
    names = ("px", "py", "pz", "tc", "occ", "mode", "omode", "cfacet", "cx", "cy", "cz", "pid")
    bufs = [C.c_void_p(eng.t[k].data_ptr()) for k in names]                    # never dereferenced: the size check comes first
    n_out = C.c_int64()
    assert L.nk_advance_host(eng.ctx, eng.cap + 1, 1, *bufs, C.byref(n_out), None, None, None) != 0
    assert "capacity" in msg(eng.ctx)
    assert L.nk_bind_particles(eng.ctx, 3, *bufs) != 0 and "capacity" in msg(eng.ctx)      # must be even and >= 2
    assert L.nk_bind_particles(eng.ctx, 2 ** 31, *bufs) != 0 and "2^31" in msg(eng.ctx)    # slots are 32-bit indices
    assert L.nk_bind_particles(eng.ctx, eng.cap, *bufs) == 0                               # and the valid call still binds
    assert L.nk_set_reservoir_mode(eng.ctx, 9, None) != 0 and "unknown" in msg(eng.ctx)
    h2d, d2h = C.c_int64(-1), C.c_int64(-1)
    assert L.nk_last_transfer_bytes(eng.ctx, C.byref(h2d), C.byref(d2h)) == 0 and h2d.value == 0 and d2h.value == 0
    eng.step(2)                                                                # the context is still usable afterwards
    assert eng.timestep() == 2
