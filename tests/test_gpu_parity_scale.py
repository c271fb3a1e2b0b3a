"""GPU-vs-oracle parity at the sizes and on the configurations BASELINE.json names: the CUDA path through the
C ABI against ``oracle/nk_oracle.py`` (keyed Philox draws) on identical inputs.

* README film (configs[1]) at its own size, 1e6 particles, 31^3 x 6 mode table, kernel variant chosen by the library;
* the same film with enough particles per (mode, subvolume) pair that the library picks the table variant (``k_step_tab``, the
  kernel the bench times) on its own;
* parameters_test.txt geometry (configs[0]) at 1e5 particles with the 31^3 table (general kernel path, rough walls);
* a Ge-lattice table with 100 slices (configs[2] layout);
* reservoirs far apart (350 / 250 K), so subvolume temperatures leave the four tau(T) slabs packed into the mode record and the
  full-table fallback runs in k_step, k_step_tab and k_mode_tables.

Integers (census, mode, omega-carrying mode, collision facet, per-SV counts, absorbed counts) bit-exact; positions / occupations /
clocks to 1e-9; T_sv, E_sv, heat flux, kappa to 1e-6 (north_star tolerance).  Every run crosses a convergence step.
The tables come from this repository's host set-up (``bench.workload``), which ``tests/test_host_setup.py`` pins bit for bit
against the reference's own tables; the oracle is pinned against the reference in ``tests/test_oracle_pin.py``.
"""
import numpy as np
import pytest

from oracle import nk_oracle as nko

pytestmark = pytest.mark.gpu

SEED = 31337
RTOL_PARTICLE = 1e-9
RTOL_SV = 1e-6


def _close(name, got, want, rtol, atol=0.0):
    got = np.asarray(got, dtype=float); want = np.asarray(want, dtype=float)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    ok = (np.isnan(got) & np.isnan(want)) | (np.isinf(want) & (got == want)) | (np.abs(got - want) <= atol + rtol * np.abs(want))
    assert ok.all(), f"{name}: {np.count_nonzero(~ok)} of {ok.size} outside rtol {rtol}; worst {np.nanmax(np.abs(got - want)[~ok])}"


def _workload(case, n, mesh, slices=20, eta=0.0, material="si"):
    import bench
    bench.CASE.update(name=case, eta=eta, slices=slices, material=material)
    try:
        return bench.workload(n, mesh)
    finally:
        bench.CASE.update(name="c2", eta=0.0, slices=20, material="si")


def _population(tb, ph, n, T0, res_counter, seed=3, tiled_modes=False):
    """Uniform positions in the bounding box (the geometries here are boxes), modes drawn at random over the active modes (or
    tiled, as Population.initialise_modes does above one particle per mode and subvolume), equilibrium occupation at T0."""
    rs = np.random.RandomState(seed)
    lo, hi = tb["bounds"]
    pos = lo + rs.random_sample((n, 3)) * (hi - lo)
    act = np.vstack(np.where(~ph.inactive_modes_mask)).T
    modes = act[np.arange(n) % act.shape[0]] if tiled_modes else act[rs.randint(0, act.shape[0], n)]
    return nko.make_state(tb, pos, modes, T0, res_counter)


def _engine(tb, st, k0, cap_factor=1.3, sort=False):
    from nanokappa_b200.engine import Engine
    J = tb["omega"].shape[1]
    eng = Engine(0, seed=SEED)
    eng.set_tables(tb, res_counter=st.res_counter)
    eng.allocate(int(st.positions.shape[0] * cap_factor) + 4096)
    eng.load_particles(st.positions, st.modes[:, 0] * J + st.modes[:, 1], st.occupation, ids=st.ids, omodes=st.omega_modes,
                       n_timesteps=st.n_timesteps, collision_facets=st.collision_facets, collision_positions=st.collision_positions)
    eng.set_sv_temperature(st.subvol_temperature)
    eng.set_timestep(k0)
    if sort:
        eng.sort_by_mode()
    return eng


def _compare(label, eng, st, tb, conv=None):
    p = eng.particles()
    order = np.argsort(st.ids)
    assert np.array_equal(p["ids"], st.ids[order]), f"{label}: particle census differs (gpu {p['ids'].shape[0]}, oracle {st.ids.shape[0]})"
    assert np.array_equal(p["modes"], st.modes[order]), f"{label}: modes differ"
    assert np.array_equal(p["omega_modes"], st.omega_modes[order]), f"{label}: omega-carrying modes differ"
    assert np.array_equal(p["collision_facets"], st.collision_facets[order]), f"{label}: collision facets differ"
    _close(f"{label} positions", p["positions"], st.positions[order], RTOL_PARTICLE, atol=1e-9)
    _close(f"{label} occupation", p["occupation"], st.occupation[order], RTOL_PARTICLE)
    _close(f"{label} n_timesteps", p["n_timesteps"], st.n_timesteps[order], RTOL_PARTICLE, atol=1e-9)
    res = eng.results()
    assert np.array_equal(res["subvol_N_p"], st.subvol_N_p), f"{label}: per-SV particle counts differ"
    assert np.array_equal(res["N_leaving"], st.N_leaving), f"{label}: absorbed counts differ"
    _close(f"{label} T_sv", res["subvol_temperature"], st.subvol_temperature, RTOL_SV)
    _close(f"{label} E_sv", res["subvol_energy"], st.subvol_energy, RTOL_SV)
    _close(f"{label} res_counter", eng.res_counter().reshape(st.res_counter.shape), st.res_counter, 1e-12, atol=1e-12)
    if conv:
        scale = np.abs(conv["subvol_heat_flux"]).max()
        _close(f"{label} heat flux", res["subvol_heat_flux"], conv["subvol_heat_flux"], RTOL_SV, atol=RTOL_SV * scale)
        _close(f"{label} res flux", res["res_heat_flux"], conv["res_heat_flux"], RTOL_SV, atol=RTOL_SV * np.abs(conv["res_heat_flux"]).max())
        _close(f"{label} res balance", res["res_energy_balance"], conv["res_energy_balance"], RTOL_SV,
               atol=RTOL_SV * np.abs(conv["res_energy_balance"]).max())
        if tb["sv_slice"]:
            _close(f"{label} kappa", res["kappa"], conv["kappa"], RTOL_SV)
    return res


def _run_both(tb, st, eng, k0, steps, check_at):
    """Advance oracle and device side by side from timestep k0; compare after the steps in `check_at` (absolute step numbers)."""
    rng = nko.KeyedRNG(SEED)
    st.current_timestep = k0
    conv = {}
    seen_conv = False
    with np.errstate(all="ignore"):
        for k in range(k0 + 1, k0 + steps + 1):
            conv.clear()
            nko.run_timestep(tb, st, rng, on_convergence=lambda s: conv.update(
                subvol_heat_flux=s.subvol_heat_flux.copy(), res_heat_flux=s.res_heat_flux.copy(),
                res_energy_balance=s.res_energy_balance.copy(), kappa=s.kappa))
            eng.step(1)
            if k in check_at or conv:
                _compare(f"step {k}", eng, st, tb, dict(conv) if conv else None)
                seen_conv = seen_conv or bool(conv)
    assert seen_conv, "the run did not cross a convergence step"
    assert eng.timestep() == k0 + steps


def test_readme_film_1e6_particles_si31_table_auto_variant():
    """BASELINE configs[1] as is: 1e6 particles, (2e4 A)^3 film, 20 slices, 31^3 x 6 modes; library-chosen kernels."""
    n = 1_000_000
    args, geo, ph, setup, tb = _workload("c2", n, 31)
    st = _population(tb, ph, n, np.full(20, 298.0), setup.res_counter)
    eng = _engine(tb, st, 6)
    _run_both(tb, st, eng, 6, 5, check_at={7, 11})
    assert eng.last_step_variant() == 0          # 5.6 particles per mode: below the table variant's threshold


def test_film_table_variant_selected_by_the_library():
    """The kernel the bench times (k_step_tab + k_mode_tables), reached WITHOUT forcing it: 1e6 particles on an 11^3 table are
    6 particles per (mode, subvolume) pair.  Particles ordered by mode with slot pools, as in the bench."""
    n = 1_000_000
    args, geo, ph, setup, tb = _workload("c2", n, 11)
    st = _population(tb, ph, n, np.full(20, 298.0), setup.res_counter, tiled_modes=True)
    eng = _engine(tb, st, 6, sort=True)
    _run_both(tb, st, eng, 6, 6, check_at={7, 12})
    assert eng.last_step_variant() == 4


def test_parameters_test_geometry_1e5_particles_si31_table():
    """BASELINE configs[0] (parameters_test.txt: box 5e3 x 1e3 x 1e3 A, T/T/R/R/P, 10 slices, linear T) at its own 1e5 particles
    with the 31^3 table; roughness 2 A so that both specular and diffuse reflections occur (the shipped file has eta = 0)."""
    n = 100_000
    args, geo, ph, setup, tb = _workload("c1", n, 31, eta=2.0)
    st = _population(tb, ph, n, np.full(10, 298.0), setup.res_counter)
    eng = _engine(tb, st, 5)
    _run_both(tb, st, eng, 5, 6, check_at={6, 8, 11})
    om_of_mode = tb["omega"][st.modes[:, 0], st.modes[:, 1]]
    assert (om_of_mode != st.omega).any(), "no specular reflection happened"


def test_ge_lattice_100_slices():
    """configs[2] layout: Ge cell, 100 slice subvolumes (beyond the table variant's size limit at 31^3; here 15^3 x 6 modes so
    that the oracle finishes in seconds), both kernel variants."""
    n = 400_000
    args, geo, ph, setup, tb = _workload("c2", n, 15, slices=100, material="ge")
    assert tb["sv_centres"].shape[0] == 100 and abs(tb["volume_unitcell"] - 48.36) < 0.5
    st = _population(tb, ph, n, np.full(100, 298.0), setup.res_counter)
    eng = _engine(tb, st.copy(), 7)
    _run_both(tb, st.copy(), eng, 7, 4, check_at={8, 10})


@pytest.mark.parametrize("variant", ["direct", "tables"])
def test_tau_slab_fallback_far_apart_reservoirs(variant, monkeypatch):
    """Reservoirs at 350 / 250 K with a linear initial profile: most subvolume temperatures lie outside the four tau(T) slabs packed
    into the 64-byte mode record (a 30 K window), so relaxation reads the full tau table (h.tr == -1 in k_step, k_step_tab's tables
    built by k_mode_tables)."""
    monkeypatch.setenv("NK_STEP_TAB", "0" if variant == "direct" else "force")
    n = 200_000
    args, geo, ph, setup, tb = _workload("c2", n, 11)
    tb = dict(tb)
    tb["res_T"] = np.array([350.0, 250.0])
    T0 = np.linspace(347.5, 252.5, 20)
    st = _population(tb, ph, n, T0, setup.res_counter)
    eng = _engine(tb, st, 8)
    _run_both(tb, st, eng, 8, 3, check_at={9, 10})
    T = eng.results()["subvol_temperature"]
    assert T.max() > 340 and T.min() < 260
    assert eng.last_step_variant() == (0 if variant == "direct" else 4)


def test_stl_mesh_find_boundary_and_tiled_rare_path(golden_dir):
    """BASELINE configs[3]: STL-imported mesh (640 triangles, 162 facets), voronoi subvolumes, rough walls.  Operator parity of
    nk_find_boundary on 30k rays (tiles streamed through shared memory) and proof that the step ran the tiled rare path; the
    30-step fixed-draw comparison of this fixture is test_gpu_parity.py::test_step_parity_fixed_draws[c9_stl_voronoi]."""
    import os
    from oracle import gen_golden
    tb, st, _ = gen_golden.load_fixture(os.path.join(golden_dir, "c9_stl_voronoi.npz"))
    assert tb["face_normals"].shape[0] == 640
    eng = _engine(tb, st, 0, cap_factor=2.0)
    r = np.random.default_rng(9)
    lo, hi = tb["bounds"]
    x = lo + r.random((30000, 3)) * (hi - lo)            # inside and (corners of the box) outside the cylinder
    v = r.standard_normal((30000, 3)) * 40
    v[:300, 2] = 0.0; v[300:600, 0] = 0.0
    x = np.vstack([x, st.positions]); v = np.vstack([v, st.group_vel])
    with np.errstate(all="ignore"):
        xc0, tc0, fc0 = nko.find_boundary(tb, x, v)
    xc, tc, fc = eng.find_boundary(x, v)
    assert np.array_equal(fc, fc0), f"{np.count_nonzero(fc != fc0)} facet mismatches"
    _close("tc", tc, tc0, 1e-12)
    _close("xc", xc, xc0, 1e-12, atol=1e-9)
    assert len(np.unique(fc0)) > 100                      # the rays reach most of the 162 facets
    eng.step(2)
    assert eng.last_step_variant() & 8, "a 640-triangle mesh must go through k_rare_tiled"


def test_many_subvolumes_need_the_opt_in_shared_memory():
    """beyond ~230 subvolumes the block-private bins of k_step / k_rare exceed the 48 KB default of dynamic shared
    memory; the launches opt in to the larger size (and nk_set_subvols rejects what cannot fit).  A box with an 8 x 8 x 6 grid
    (384 subvolumes, nearest temperature, rough + periodic walls) against the oracle."""
    import argument_parser as ap
    import contextlib
    import io
    from nanokappa_b200.classes.Geometry import Geometry
    from nanokappa_b200.classes.Phonon import Phonon
    from nanokappa_b200.classes.Population import PopulationSetup
    from nanokappa_b200._lib import NkError
    text = """--mat_folder /nonexistent/ --hdf_file synthetic:5 --poscar_file POSCAR
    --geometry box --dimensions 4e3 4e3 3e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz --subvolumes grid 8 8 6
    --bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 0.5 0.5 -0.1 0.5 0.5 1.1 --bound_cond T T R R P --connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5
    --bound_values 303 297 2 2 --reference_temp local --temp_dist cold --temp_interp nearest --particles total 40000
    --part_dist random_subvol --timestep 1 --iterations 100 --n_mean 10 --results_folder /tmp --conv_crit 0 10 --output screen
    --max_sim_time 0-00:00:00"""
    args = ap.initialise_parser(False).parse_args(text.split())
    args.results_folder = "/tmp"
    with contextlib.redirect_stdout(io.StringIO()):
        geo = Geometry(args)
        ph = Phonon(args, 0)
        np.random.seed(0)
        setup = PopulationSetup(args, geo, ph, seed=0)
    tb = setup.tables(geo, ph)
    S = tb["sv_centres"].shape[0]
    assert S == 384
    st = _population(tb, ph, 40000, np.full(S, 297.0), setup.res_counter)
    eng = _engine(tb, st, 4)
    _run_both(tb, st, eng, 4, 7, check_at={5, 10, 11})
    # and the limit is reported, not a bare launch failure
    from nanokappa_b200.engine import Engine
    big = dict(tb)
    big["sv_centres"] = np.random.default_rng(0).random((1500, 3)) * 1e3
    big["sv_volume"] = np.ones(1500)
    with pytest.raises(NkError, match="n_subvols"):
        Engine(0, seed=1).set_tables(big)
