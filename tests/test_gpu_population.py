"""GPU tests of the reference-shaped Python surface: Geometry / Phonon / Population built from a
parameters file, stepped with run_timestep, output files in the reference's formats, and a statistical
comparison with the oracle (different random streams, so within stated error bars only)."""
import contextlib
import io
import os

import numpy as np
import pytest

import argument_parser as ap
from oracle import gen_golden, nk_oracle as nko

pytestmark = pytest.mark.gpu


def _args(text, folder):
    text = text.replace("kappa-m313131.hdf5", "synthetic:5").replace("--mat_folder test_material/Si/", "--mat_folder /nonexistent/")
    args = ap.initialise_parser(False).parse_args(text.split())
    args.results_folder = str(folder)
    return args


def _population(text, folder, seed=3):
    from nanokappa_b200.classes.Geometry import Geometry
    from nanokappa_b200.classes.Phonon import Phonon
    from nanokappa_b200.classes.Population import Population
    args = _args(text, folder)
    with contextlib.redirect_stdout(io.StringIO()):
        geo = Geometry(args)
        ph = Phonon(args, 0)
        np.random.seed(seed)
        pop = Population(args, geo, ph, device=0, seed=seed)
    return args, geo, ph, pop


def test_population_surface_and_files(tmp_path):
    text = gen_golden.PARAMS_C1.format(eta=2, n=20000)
    args, geo, ph, pop = _population(text, tmp_path)
    n0 = pop.N_p
    assert pop.positions.shape == (n0, 3) and pop.modes.shape == (n0, 2) and pop.occupation.shape == (n0,)
    assert pop.group_vel.shape == (n0, 3) and pop.omega.shape == (n0,) and pop.n_timesteps.shape == (n0,)
    assert set(np.unique(pop.collision_cond)) <= {"T", "P", "R"}
    assert np.array_equal(np.bincount(pop.subvol_id, minlength=10), pop.subvol_N_p)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(120):
            pop.run_timestep(geo, ph)
        pop.write_final_state(geo)
    assert pop.current_timestep == 120 and abs(pop.N_p - n0) < 0.05 * n0
    # convergence.txt: header + row 0 + one row per 10 steps, parsed positionally like the reference does
    pop.view.read_convergence()
    v = pop.view
    assert v.timestep.tolist() == list(range(0, 121, 10))
    assert v.T.shape == (13, 10) and v.sv_phi.shape == (13, 30) and v.sv_k.shape == (13, 10) and v.en_res.shape == (13, 2)
    assert np.allclose(v.T[0], 298.0) and v.T[-1, 0] > v.T[-1, -1] > 297.9            # hot side warms up first
    assert (v.N_p > 0).all() and np.isfinite(v.k[1:]).all()
    # particle_data.txt round trip (--part_dist restart file format)
    data = np.loadtxt(os.path.join(tmp_path, "particle_data.txt"), delimiter=",", comments="#")
    assert data.shape == (pop.N_p, 6)
    assert np.allclose(data[:, 2:5], pop.positions, atol=1e-3) and np.array_equal(data[:, :2].astype(int), pop.modes)
    assert os.path.isfile(os.path.join(tmp_path, "residue.txt")) and os.path.isfile(os.path.join(tmp_path, "subvolumes.txt"))


def test_restart_from_particle_file(tmp_path):
    text = gen_golden.PARAMS_C2.format(n=8000)
    a = tmp_path / "a"; b = tmp_path / "b"; a.mkdir(); b.mkdir()
    args, geo, ph, pop = _population(text, a)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(30):
            pop.run_timestep(geo, ph)
        pop.write_final_state(geo)
    T_before = pop.subvol_temperature.copy()
    text2 = text.replace("--part_dist random_subvol", "--part_dist " + os.path.join(a, "particle_data.txt"))
    args2, geo2, ph2, pop2 = _population(text2, b)
    assert pop2.N_p == pop.N_p
    assert np.allclose(pop2.subvol_temperature, T_before, atol=0.05)      # positions are stored to 1e-3 A, occupation to 7 digits


def test_statistical_agreement_with_oracle(tmp_path):
    """Same physical case, independent random streams: after 150 steps the subvolume temperature
    profiles must agree within the Monte-Carlo noise (sigma_T ~ 0.02 K for 3e4 particles in 10 slices;
    band = 0.12 K), the particle count within 1 %, and the mean heat flux within 25 %."""
    text = gen_golden.PARAMS_C1.format(eta=5, n=30000)
    args, geo, ph, pop = _population(text, tmp_path, seed=11)
    tb = pop.tables
    p = pop._particles()
    st = nko.make_state(tb, p["positions"], p["modes"], np.full(10, 298.0), pop.res_counter, ids=p["ids"])
    rng = nko.KeyedRNG(987)
    flux_o, flux_g = [], []
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        for k in range(150):
            nko.run_timestep(tb, st, rng, on_convergence=lambda s: flux_o.append(s.subvol_heat_flux[:, 0].copy()))
            pop.run_timestep(geo, ph)
            if pop.current_timestep % 10 == 0:
                flux_g.append(pop.subvol_heat_flux[:, 0].copy())
    assert abs(pop.N_p - st.N_p) < 0.01 * st.N_p
    assert np.abs(pop.subvol_temperature - st.subvol_temperature).max() < 0.12
    fo, fg = np.mean(flux_o[5:], axis=0)[:3].mean(), np.mean(flux_g[5:], axis=0)[:3].mean()
    assert fo > 0 and abs(fg - fo) < 0.25 * fo


def test_binary_checkpoint_continues_exactly(tmp_path):
    """save_checkpoint at a convergence-row boundary, continue in a fresh Population: same particles as the
    uninterrupted run (integers and positions identical, occupations to 1e-12: only the order of the sums
    changes because the slots are compacted on load)."""
    text = gen_golden.PARAMS_C1.format(eta=2, n=12000)
    a = tmp_path / "a"; b = tmp_path / "b"
    args, geo, ph, pop = _population(text, a, seed=21)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(20):
            pop.run_timestep(geo, ph)
        pop.save_checkpoint(os.path.join(tmp_path, "ckpt.npz"))
        for _ in range(20):
            pop.run_timestep(geo, ph)
    args2, geo2, ph2, pop2 = _population(text, b, seed=21)
    with contextlib.redirect_stdout(io.StringIO()):
        pop2.load_checkpoint(os.path.join(tmp_path, "ckpt.npz"))
        assert pop2.current_timestep == 20
        for _ in range(20):
            pop2.run_timestep(geo2, ph2)
    p, q = pop._particles(), pop2._particles()
    assert np.array_equal(p["ids"], q["ids"]) and np.array_equal(p["modes"], q["modes"])
    assert np.array_equal(p["collision_facets"], q["collision_facets"])
    assert np.array_equal(p["positions"], q["positions"]) and np.array_equal(p["n_timesteps"], q["n_timesteps"])
    assert np.allclose(p["occupation"], q["occupation"], rtol=1e-12, atol=0)
    assert np.allclose(pop.subvol_temperature, pop2.subvol_temperature, rtol=1e-13, atol=0)
    assert np.array_equal(pop.engine.res_counter(), pop2.engine.res_counter())


def test_converged_film_matches_reference(tmp_path, golden_dir):
    """Converged run, independent random draws: kappa, the temperature profile, the heat flux and the particle count
    of a short cross-plane film must agree with what the REFERENCE ITSELF produced (tests/golden/converged_film.json,
    written by oracle/gen_converged.py executing /root/reference) within the combined block-averaged standard
    errors (5 sigma; the error estimates themselves come from 10 blocks, i.e. are known to ~25 %)."""
    import json
    from oracle import gen_converged as gc
    ref = json.load(open(os.path.join(golden_dir, "converged_film.json")))
    args, geo, ph, pop = _population(ref["params"], tmp_path, seed=5)
    kappa, T, flux, N = [], [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for k in range(1, ref["steps"] + 1):
            pop.run_timestep(geo, ph)
            if k % 10 == 0 and k > ref["discard"]:
                kappa.append(float(pop.kappa)); T.append(pop.subvol_temperature.copy())
                flux.append(pop.subvol_heat_flux[:, 0].copy()); N.append(float(pop.N_p))
    got = gc.summarise(kappa, T, flux, N)
    assert len(kappa) == ref["rows"]
    for name in ("kappa", "T", "flux_x", "N_p"):
        g, r = got[name], ref[name]
        gm, ge, rm, re_ = (np.asarray(x, dtype=float) for x in (g["mean"], g["stderr"], r["mean"], r["stderr"]))
        # 5 sigma of the combined error bars (themselves known to ~25 %), never tighter than a small physical floor
        floor = {"kappa": 3e-3 * abs(rm), "T": 0.02, "flux_x": 0.01 * np.abs(rm).max(), "N_p": 2e-3 * rm}[name]
        band = np.maximum(5.0 * np.sqrt(ge ** 2 + re_ ** 2), floor)
        assert (np.abs(gm - rm) <= band).all(), f"{name}: gpu {gm} +- {ge} vs reference {rm} +- {re_}"
    # and the error bars are small enough for the comparison to mean something: 0.2 % on kappa
    assert 4.0 * np.hypot(got["kappa"]["stderr"], ref["kappa"]["stderr"]) < 2e-3 * ref["kappa"]["mean"] * 2


def test_contains_check_puts_escaped_particles_back(tmp_path):
    """Population.contains_check (Population.py:1712-1722) through nk_outside_slots: particles pushed outside the
    bounding box are found on the device, re-drawn inside the mesh and given a fresh first collision; the others
    are not touched."""
    args, geo, ph, pop = _population(gen_golden.PARAMS_C1.format(eta=5, n=20000), tmp_path, seed=4)
    eng = pop.engine
    before = eng.particles(flush=False)
    t = eng.t
    n_slots, _ = eng.slot_count()
    victims = np.array([3, 777, 15000, n_slots - 1])
    assert (t["mode"][victims] >= 0).all()
    victim_ids = t["pid"][victims].cpu().numpy()
    t["px"][3] = float(geo.bounds[1, 0] + 5.0)
    t["py"][777] = float(geo.bounds[1, 1] + 6.0)
    t["px"][15000] = float(geo.bounds[0, 0] - 7.0)
    t["pz"][n_slots - 1] = float(geo.bounds[0, 2] - 1e-3)
    np.random.seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        pop.contains_check(geo)
    after = eng.particles(flush=False)
    assert np.array_equal(after["ids"], before["ids"])
    moved = np.nonzero((after["positions"] != before["positions"]).any(axis=1))[0]
    assert np.array_equal(np.sort(before["ids"][moved]), np.sort(victim_ids))
    x = after["positions"][moved]
    assert geo.mesh.contains(x).all()
    assert (after["n_timesteps"][moved] > 0).all() and (after["collision_facets"][moved] >= 0).all()
    xc, tc, fc = eng.find_boundary(x, ph.group_vel[after["modes"][moved][:, 0], after["modes"][moved][:, 1], :])
    assert np.array_equal(fc, after["collision_facets"][moved])
    keep = np.setdiff1d(np.arange(before["ids"].shape[0]), moved)
    for f in ("positions", "n_timesteps", "collision_facets", "collision_positions", "occupation"):
        assert np.array_equal(after[f][keep], before[f][keep], equal_nan=True), f
    with contextlib.redirect_stdout(io.StringIO()):
        pop.contains_check(geo)          # nothing left outside: a no-op
    again = eng.particles(flush=False)
    assert np.array_equal(again["positions"], after["positions"])


@pytest.mark.parametrize("name", ["c4_cylinder_voronoi", "c5_box_grid_radial", "c6_cylinder_voronoi_radial", "c7_fixed_rate", "c8_one_to_one"])
def test_population_runs_the_other_configurations(name, tmp_path):
    """The Python surface on the non-slice / radial / debug-emission configurations of the fixtures (host set-up written
    from scratch + CUDA path): 60 steps, physical sanity, convergence rows and the per-connection kappa file."""
    text = gen_golden.CONFIGS[name][0].replace("--particles total 3000", "--particles total 12000")
    np.random.seed(12)                                     # the voronoi centres are drawn at random by Geometry
    args, geo, ph, pop = _population(text, tmp_path, seed=8)
    n0 = pop.N_p
    res_T = np.asarray(pop.res_facet_temperature, dtype=float)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(60):
            pop.run_timestep(geo, ph)
        pop.write_final_state(geo)
    assert pop.current_timestep == 60 and abs(pop.N_p - n0) < 0.1 * n0
    T = pop.subvol_temperature
    # (the cubic RBF field over voronoi centres can run away when the centres are nearly coplanar -- the reference's own
    #  caveat at Population.py:577-586, reproduced step by step: tests/run_radial_instability_check.py; seed 12 is benign)
    assert np.isfinite(T).all() and T.min() >= res_T.min() - 0.5 and T.max() <= res_T.max() + 0.5
    assert T.max() > res_T.min() + 1e-3                                       # heat has entered from the hot side
    x = pop.positions
    assert geo.mesh.contains(x).mean() > 0.999                                # rough walls keep the particles inside
    assert np.array_equal(np.bincount(pop.subvol_id, minlength=pop.n_of_subvols), pop.subvol_N_p)
    Tp = pop.temperatures
    assert np.isfinite(Tp).all() and Tp.shape == (pop.N_p,)
    if "radial" in name:                                                     # the particle field is the RBF through the centres
        assert np.allclose(pop.temperature_interpolator(geo.subvol_center), T, rtol=1e-9)
    pop.view.read_convergence()
    assert pop.view.timestep.tolist() == list(range(0, 61, 10)) and np.isfinite(pop.view.T).all()
    if geo.subvol_type != "slice":
        assert pop.svcon_kappa.shape == (geo.n_of_subvol_con,) and np.isfinite(pop.svcon_kappa).all()


def test_large_populations_dump_the_binary_checkpoint(tmp_path, monkeypatch):
    """Above NK_TEXT_DUMP_MAX particles write_final_state stores particle_data.npz (exact, restartable) instead of the
    reference's 60-bytes-per-particle text file."""
    monkeypatch.setenv("NK_TEXT_DUMP_MAX", "1000")
    args, geo, ph, pop = _population(gen_golden.PARAMS_C2.format(n=5000), tmp_path, seed=2)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(10):
            pop.run_timestep(geo, ph)
        pop.write_final_state(geo)
    assert not os.path.exists(os.path.join(tmp_path, "particle_data.txt"))
    z = np.load(os.path.join(tmp_path, "particle_data.npz"))
    assert z["positions"].shape == (pop.N_p, 3) and int(z["current_timestep"]) == 10
    assert np.array_equal(z["positions"], pop.positions)


def test_command_line_driver_end_to_end(tmp_path):
    """`python nanokappa.py -ff parameters.txt` as a user of the reference runs it (nanokappa.py:71-107 upstream):
    a parameters file in the reference's format, 200 iterations, results folder with the reference's files."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    params = tmp_path / "parameters.txt"
    text = gen_golden.PARAMS_C1.format(eta=2, n=8000).replace("kappa-m313131.hdf5", "synthetic:5") \
        .replace("--mat_folder test_material/Si/", "--mat_folder /nonexistent/").replace("--iterations 1000", "--iterations 200") \
        .replace("--results_folder x", "--results_folder {}".format(tmp_path / "run"))
    params.write_text("\n".join(l for l in text.splitlines() if l.strip()) + "\n")
    out = subprocess.run([sys.executable, os.path.join(root, "nanokappa.py"), "-ff", str(params)], cwd=tmp_path,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Total time" in out.stdout and "Timestep   100" in out.stdout
    folders = [d for d in os.listdir(tmp_path) if d.startswith("run")]
    assert len(folders) == 1
    run = tmp_path / folders[0]
    for f in ("arguments.txt", "convergence.txt", "particle_data.txt", "residue.txt", "subvolumes.txt"):
        assert (run / f).is_file(), f
    rows = [l for l in (run / "convergence.txt").read_text().splitlines() if l and not l.startswith("#")]
    assert len(rows) >= 21                                                   # row 0 + one per 10 steps


def test_maintenance_resort_does_not_change_the_physics(tmp_path, monkeypatch):
    """Large populations are re-ordered by mode every NK_RESORT_EVERY steps (slot layout only).  Sums are merged in fixed
    point and draws are keyed by particle, so the run with re-sorting must repeat the run without it BIT FOR BIT."""
    text = gen_golden.PARAMS_C1.format(eta=2, n=15000)
    out = []
    for every in ("0", "7"):
        monkeypatch.setenv("NK_RESORT_EVERY", every)
        monkeypatch.setenv("NK_RESORT_MIN", "0")
        args, geo, ph, pop = _population(text, tmp_path / every, seed=9)
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(40):
                pop.run_timestep(geo, ph)
        p = pop.engine.particles()
        out.append((pop.engine.results(), p))
    (ra, pa), (rb, pb) = out
    assert np.array_equal(ra["subvol_temperature"], rb["subvol_temperature"]) and np.array_equal(ra["subvol_N_p"], rb["subvol_N_p"])
    assert np.array_equal(ra["subvol_heat_flux"], rb["subvol_heat_flux"])
    for k in ("ids", "modes", "positions", "occupation", "n_timesteps", "collision_facets"):
        assert np.array_equal(pa[k], pb[k], equal_nan=True), k


def test_step_batching_and_seam_methods_equal_run_timestep(tmp_path):
    """Three ways to advance a Population by the same 137 steps must leave identical particles and convergence rows:
    run_timestep one launch sequence per call (library default), run_timestep with the command line's batching (the steps
    between two convergence rows enqueued as one nk_step call, rows read back from asynchronous snapshots, the device up to a
    batch ahead), and the reference's own call sequence drift / fill_reservoirs / add_reservoir_particles / boundary_scattering
    / refresh_temperatures / lifetime_scattering (SURVEY 8b seams)."""
    text = gen_golden.PARAMS_C1.format(eta=2, n=20000).replace("--iterations 1000", "--iterations 137")
    rows, parts = {}, {}
    for label in ("plain", "batched", "seams"):
        folder = tmp_path / label
        args, geo, ph, pop = _population(text, folder, seed=5)
        pop.step_batching = label == "batched"
        with contextlib.redirect_stdout(io.StringIO()):
            if label == "seams":
                for _ in range(137):
                    pop.drift()
                    pop.fill_reservoirs(geo, ph)
                    pop.add_reservoir_particles(geo, ph)
                    pop.boundary_scattering(geo, ph)
                    pop.refresh_temperatures(geo, ph)
                    pop.lifetime_scattering(ph)
                E = pop.calculate_energy(geo, ph)
                assert E.shape == (10,) and np.isfinite(E).all() and pop.calculate_heat_flux(geo, ph).shape == (10, 3)
            else:
                while pop.current_timestep < 137:
                    pop.run_timestep(geo, ph)
            pop.write_final_state(geo)
        assert pop.current_timestep == 137
        parts[label] = pop.engine.particles()
        if label != "seams":
            with open(os.path.join(folder, "convergence.txt")) as fh:
                rows[label] = [ln.split()[1:] for ln in fh if not ln.startswith("#")]      # drop the wall-clock column
    assert len(rows["plain"]) == 14 and rows["plain"] == rows["batched"]
    for label in ("batched", "seams"):
        for f in parts["plain"]:
            assert np.array_equal(parts["plain"][f], parts[label][f], equal_nan=True), f"{label}: {f}"
    with pytest.raises(Exception):
        pop.boundary_scattering(geo, ph)          # outside a drift() ... lifetime_scattering() sequence


def test_contains_operator_on_the_stl_mesh(golden_dir):
    """nk_contains (Mesh.contains_naive, crossing parity over TMA-staged triangle tiles) on the 640-triangle STL cylinder of
    fixture c9, against the analytic answer for a regular prism: inside the 160-gon and between the caps."""
    from nanokappa_b200.engine import Engine
    tb, st, _ = gen_golden.load_fixture(os.path.join(golden_dir, "c9_stl_voronoi.npz"))
    eng = Engine(0, seed=1)
    eng.set_tables(tb, res_counter=st.res_counter)
    r = np.random.default_rng(4)
    lo, hi = tb["bounds"]
    x = lo - 0.05 * (hi - lo) + r.random((200000, 3)) * (hi - lo) * 1.1
    got = eng.contains(x)
    V = tb["face_vertices"].reshape(-1, 3)
    ring = np.unique(np.round(V[np.abs(V[:, 2] - lo[2]) < 1e-6][:, :2], 6), axis=0)
    c = ring.mean(axis=0)
    ring = ring[np.argsort(np.arctan2(ring[:, 1] - c[1], ring[:, 0] - c[0]))]
    ring = ring[np.linalg.norm(ring - c, axis=1) > 1.0]                 # drop the centre vertex of the cap fans
    a, b = ring, np.roll(ring, -1, axis=0)
    edge = b - a
    nrm = np.stack((edge[:, 1], -edge[:, 0]), axis=1); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)     # outward (counter-clockwise ring)
    d = ((x[:, None, :2] - a[None]) * nrm[None]).sum(axis=2).max(axis=1)                                       # > 0: outside the polygon
    dz = np.maximum(lo[2] - x[:, 2], x[:, 2] - hi[2])
    dist = np.maximum(d, dz)
    want = dist < 0
    clear = np.abs(dist) > 1e-6
    assert ring.shape[0] == 160 and 0.3 < want.mean() < 0.8
    assert np.array_equal(got[clear], want[clear]), f"{np.count_nonzero(got[clear] != want[clear])} points misclassified"


def test_population_is_initialised_on_the_device_in_an_arbitrary_mesh(tmp_path, monkeypatch):
    """above NK_DEVICE_INIT_MIN particles the initial positions of ANY geometry are drawn on the device (rejection
    sampling with nk_contains, per-subvolume quotas with nk_classify, Population.py:209-246) -- here the faceted cylinder with
    voronoi subvolumes and rough walls: everybody inside the mesh, quotas as upstream, modes tiled, a run that stays sane."""
    monkeypatch.setenv("NK_DEVICE_INIT_MIN", "100000")
    text = gen_golden.PARAMS_C4.format(eta=3, n=150000)
    np.random.seed(12)
    args, geo, ph, pop = _population(text, tmp_path, seed=8)
    assert pop.N_p == 150000
    eng = pop.engine
    p = eng.particles(flush=False)
    assert eng.contains(p["positions"]).all()
    sv = eng.classify(p["positions"])
    vol = np.asarray(geo.subvol_volume, dtype=float)
    quota = np.ceil(150000 * vol / vol.sum()).astype(int)
    counts = np.bincount(sv, minlength=vol.shape[0])
    assert (counts[:-1] == quota[:-1]).all() and counts[-1] == 150000 - quota[:-1].sum()          # vstack(...)[:N] cuts the last one
    n_act = int((~ph.inactive_modes_mask).sum())
    flat = p["modes"][:, 0] * ph.omega.shape[1] + p["modes"][:, 1]
    assert np.bincount(flat).max() - np.bincount(flat)[np.bincount(flat) > 0].min() <= 1 and np.unique(flat).shape[0] == min(n_act, 150000)
    assert (p["n_timesteps"] > 0).all() and (p["collision_facets"] >= 0).all()
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(30):
            pop.run_timestep(geo, ph)
    assert abs(pop.N_p - 150000) < 0.05 * 150000 and np.isfinite(pop.subvol_temperature).all()
    assert eng.contains(pop.positions).mean() > 0.999


def test_restart_from_the_binary_checkpoint_through_part_dist(tmp_path, monkeypatch):
    """above NK_TEXT_DUMP_MAX particles the end-of-run dump is particle_data.npz, and `--part_dist <that file>` must
    restart from it (the reference's restart workflow, Population.py:284-306, but exact: positions in f64, collision clocks,
    reservoir counters, the open convergence window).  25 steps, dump, restart, 15 more == 40 uninterrupted steps."""
    monkeypatch.setenv("NK_TEXT_DUMP_MAX", "1000")
    text = gen_golden.PARAMS_C2.format(n=8000)
    args, geo, ph, ref = _population(text, tmp_path / "ref", seed=6)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(40):
            ref.run_timestep(geo, ph)
    args, geo, ph, a = _population(text, tmp_path / "a", seed=6)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(25):
            a.run_timestep(geo, ph)
        a.write_final_state(geo)
    ckpt = os.path.join(tmp_path / "a", "particle_data.npz")
    assert os.path.isfile(ckpt)
    args, geo, ph, b = _population(text.replace("--part_dist random_subvol", "--part_dist " + ckpt), tmp_path / "b", seed=6)
    assert b.current_timestep == 25
    with contextlib.redirect_stdout(io.StringIO()):
        while b.current_timestep < 40:
            b.run_timestep(geo, ph)
    pr, pb = ref.engine.particles(), b.engine.particles()
    for f in pr:
        assert np.array_equal(pr[f], pb[f], equal_nan=True), f
    rr, rb = ref.engine.results(), b.engine.results()
    for f in ("subvol_temperature", "subvol_energy", "subvol_heat_flux", "res_energy_balance", "res_heat_flux"):
        assert np.array_equal(rr[f], rb[f]), f
