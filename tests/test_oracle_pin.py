"""Pin the oracle restatement (oracle/nk_oracle.py) against the reference itself.

* fixtures: tests/golden/*.npz were produced by executing the unmodified reference
  (oracle/gen_golden.py); the restatement, fed the same NumPy stream, must reproduce every particle
  array and per-SV vector BIT FOR BIT after 1, 10 and 20 steps.
* live: where /root/reference exists the same comparison runs against the live reference on a
  configuration that is not in the fixtures.
"""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import gen_golden, nk_oracle as nko, ref_harness

FIXTURES = sorted(gen_golden.CONFIGS)


def _eq(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    assert np.array_equal(a, b, equal_nan=True), f"{name}: max abs diff {np.nanmax(np.abs(a.astype(float) - b.astype(float)))}"


@pytest.mark.parametrize("name", FIXTURES)
def test_restatement_matches_reference_fixture(name, golden_dir):
    tb, st, refs = gen_golden.load_fixture(os.path.join(golden_dir, name + ".npz"))
    rng = nko.SequenceRNG()
    np.random.seed(gen_golden.SEED_STEPS)
    conv = {}
    with np.errstate(all="ignore"):
        for k in range(1, max(refs) + 1):
            nko.run_timestep(tb, st, rng, on_convergence=lambda s: conv.update(
                subvol_heat_flux=s.subvol_heat_flux.copy(), res_heat_flux=s.res_heat_flux.copy(),
                res_energy_balance=s.res_energy_balance.copy(),
                subvol_kappa=None if s.subvol_kappa is None else s.subvol_kappa.copy(), kappa=s.kappa))
            if k in refs:
                ref = refs[k]
                for f in gen_golden.REF_FIELDS + ("collision_cond",):
                    _eq(f"step {k} {f}", ref[f], getattr(st, f))
                for f in ("subvol_heat_flux", "res_heat_flux", "res_energy_balance", "subvol_kappa", "kappa"):
                    if "conv_" + f in ref and conv[f] is not None and ref["conv_" + f].size == np.size(conv[f]):
                        _eq(f"step {k} conv {f}", ref["conv_" + f], conv[f])


def test_fixture_covers_every_branch(golden_dir):
    """The fixtures must actually exercise absorption, emission, periodic wrap, specular and diffuse."""
    tb, st, refs = gen_golden.load_fixture(os.path.join(golden_dir, "c1_mixed.npz"))
    last = refs[max(refs)]
    assert last["positions"].shape[0] != st.positions.shape[0] or last["N_leaving"].sum() > 0
    J = tb["omega"].shape[1]
    om_of_mode = tb["omega"][last["modes"][:, 0], last["modes"][:, 1]]
    assert (om_of_mode != last["omega"]).any(), "no specular reflection kept its old omega"
    assert set(np.unique(tb["facet_bc"])) == {nko.BC_T, nko.BC_P, nko.BC_R}


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present on this box")
def test_restatement_matches_live_reference():
    from oracle import extract
    text = gen_golden.PARAMS_C1.format(eta=2, n=2500).replace("slice 10 0", "slice 7 0")
    with contextlib.redirect_stdout(io.StringIO()):
        args, geo, ph, pop = gen_golden.build_reference(text, 5, results="/tmp/nk_pin_live")
    tb = extract.tables_from_reference(geo, ph, pop)
    st = extract.state_from_reference(ph, pop)
    rng = nko.SequenceRNG()
    np.random.seed(99)
    with np.errstate(all="ignore"):
        for k in range(12):
            state = np.random.get_state()
            extract.reference_step(pop, geo, ph)
            after = np.random.get_state()
            np.random.set_state(state)
            nko.run_timestep(tb, st, rng)
            assert np.array_equal(after[1], np.random.get_state()[1]), "random streams diverged"
            for f in gen_golden.REF_FIELDS:
                _eq(f"step {k} {f}", getattr(pop, f), getattr(st, f))


@pytest.mark.parametrize("name", ["c1_mixed", "c2_crossplane"])
def test_restated_thirdparty_formulas(name, golden_dir):
    """interp1d / RegularGridInterpolator / cKDTree restatements against SciPy itself."""
    from scipy.interpolate import RegularGridInterpolator, interp1d
    from scipy.spatial import cKDTree
    tb, st, _ = gen_golden.load_fixture(os.path.join(golden_dir, name + ".npz"))
    r = np.random.default_rng(0)
    lo, hi = tb["bounds"]
    x = r.random((5000, 3)) * (hi - lo) * 1.2 + lo - 0.1 * (hi - lo)
    assert np.array_equal(nko.classify(tb, x), cKDTree(tb["sv_centres"]).query(x)[1])
    T_sv = 298 + 4 * r.random(tb["sv_centres"].shape[0])
    ax = int(tb["slice_axis"])
    f = interp1d(tb["sv_centres"][:, ax], T_sv, kind=tb["temp_interp"], fill_value="extrapolate")
    assert np.array_equal(nko.particle_temperature(tb, T_sv, x), f(x[:, ax]))
    Q, J = tb["omega"].shape
    T = 250 + 100 * r.random(4000)
    modes = np.stack([r.integers(0, Q, 4000), r.integers(0, J, 4000)], axis=1)
    rgi = RegularGridInterpolator((tb["T_grid"], np.arange(Q), np.arange(J)), tb["tau"])
    assert np.array_equal(nko.lifetime_function(tb, T, modes), rgi(np.hstack((T.reshape(-1, 1), modes))))
    E = np.concatenate([tb["energy_array"][[0, -1]], tb["energy_array"][0] + r.random(3000) * np.ptp(tb["energy_array"]),
                        [tb["energy_array"][0] - 1, tb["energy_array"][-1] + 1]])
    tf = interp1d(tb["energy_array"], tb["T_array"], kind="linear", fill_value=(tb["T_array"].min(), tb["T_array"].max()), bounds_error=False)
    assert np.array_equal(nko.temperature_function(tb, E), tf(E))
    ef = interp1d(tb["T_array"], tb["energy_array"], kind="linear", fill_value=(tb["energy_array"].min(), tb["energy_array"].max()), bounds_error=False)
    Tq = np.concatenate([[-5.0, 0.0, 1000.0, 1200.0], 1000 * r.random(2000)])
    assert np.array_equal(nko.crystal_energy_function(tb, Tq), ef(Tq))


def test_philox_known_answers():
    """Random123 Philox4x32-10 known-answer vectors."""
    from oracle.philox import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = tuple(int(v) for v in philox4x32_10(*c, *k))
        assert got == want


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("name", ["c1_mixed", "c2_crossplane", "c5_box_grid_radial"])
def test_convergence_file_format_equals_the_reference(name, tmp_path):
    """convergence.txt is an interface (Visualisation parses it by column position, Population.py:1841-1939): the header and
    the rows written by nanokappa_b200's Population must be character-identical to the reference's for the same state.
    The reference writes its file during construction; our writer is fed the reference's own geometry and values."""
    import types
    from nanokappa_b200.classes.Population import Population
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    ref_dir.mkdir(); our_dir.mkdir()
    with contextlib.redirect_stdout(io.StringIO()):
        args, geo, ph, pop = gen_golden.build_reference(gen_golden.CONFIGS[name][0], 5, results=str(ref_dir))
    theirs = (ref_dir / "convergence.txt").read_text().splitlines()
    stub = types.SimpleNamespace(results_folder_name=str(our_dir), n_of_subvols=pop.n_of_subvols, n_of_reservoirs=pop.n_of_reservoirs)
    for k in ("current_timestep", "t", "total_energy", "res_energy_balance", "res_heat_flux", "N_p", "subvol_temperature",
              "subvol_energy", "subvol_heat_flux", "subvol_N_p", "subvol_kappa", "kappa", "svcon_kappa"):
        if hasattr(pop, k):
            setattr(stub, k, getattr(pop, k))
    Population.open_convergence(stub, geo)
    Population.write_convergence(stub, geo)
    ours = (our_dir / "convergence.txt").read_text().splitlines()
    assert ours[0] == theirs[0], "header differs"
    assert len(ours) == len(theirs) == 2
    strip = lambda row: row.split(" ", 1)[1]           # drop the wall-clock token
    assert strip(ours[1]) == strip(theirs[1]), "row format differs"


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present on this box")
def test_particle_dump_format_equals_the_reference(tmp_path):
    """particle_data.txt is the --part_dist restart format (Population.py:2071-2091): same header lines (but the date) and
    the same rows for the same particles."""
    import types
    from nanokappa_b200.classes.Population import Population
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    ref_dir.mkdir(); our_dir.mkdir()
    with contextlib.redirect_stdout(io.StringIO()):
        args, geo, ph, pop = gen_golden.build_reference(gen_golden.CONFIGS["c1_mixed"][0], 5, results=str(ref_dir))
        pop.write_final_state(geo)
    theirs = (ref_dir / "particle_data.txt").read_text().splitlines()
    stub = types.SimpleNamespace(results_folder_name=str(our_dir), args=args, N_p=pop.N_p, current_timestep=0, view=object(),
                                 _particles=lambda: dict(modes=np.asarray(pop.modes), positions=np.asarray(pop.positions),
                                                         occupation=np.asarray(pop.occupation)))
    Population.write_final_state(stub, geo)
    ours = (our_dir / "particle_data.txt").read_text().splitlines()
    assert len(ours) == len(theirs)
    keep = lambda lines: [l for l in lines if "Date and time" not in l]
    assert keep(ours) == keep(theirs)


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("name", ["c1_mixed", "c5_box_grid_radial"])
def test_postprocessing_chain_equals_the_reference(name, tmp_path):
    """The every-100-steps chain that feeds back into the run (Population.run_timestep :1729-1735): the reference is run for
    201 steps with its own outputs; our Visualisation must parse ITS convergence.txt into the same arrays and rolling
    statistics, and our update_residue / write_subvolume_state, fed those, must write the rows the reference wrote."""
    import types
    from nanokappa_b200.classes.Population import Population
    from nanokappa_b200.classes.Visualisation import Visualisation
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    ref_dir.mkdir(); our_dir.mkdir()
    text = gen_golden.CONFIGS[name][0].replace("--particles total 4000", "--particles total 1500").replace("--particles total 3000", "--particles total 1500")
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        args, geo, ph, pop = gen_golden.build_reference(text, 5, results=str(ref_dir))
        np.random.seed(3)
        for _ in range(201):
            pop.run_timestep(geo, ph)
    rv = pop.view
    with contextlib.redirect_stdout(io.StringIO()):
        ours = Visualisation(args, geo, ph)
        ours.read_convergence()
    slice_type = geo.subvol_type == "slice"
    for k in ("T", "N_p", "en_res", "phi_res", "sv_phi", "mean_T", "std_T", "mean_sv_phi", "std_sv_phi", "mean_en_res", "std_en_res") + \
            (("sv_k", "k", "mean_sv_k", "std_sv_k") if slice_type else ("mean_con_k", "std_con_k", "mean_con_dT", "std_con_dT", "mean_con_phi", "std_con_phi")):
        _eq(k, getattr(rv, k), getattr(ours, k))
    # residue row of step 200 and the subvolume tables written at step 200
    stub = types.SimpleNamespace(view=ours, n_of_subvols=pop.n_of_subvols, n_of_reservoirs=pop.n_of_reservoirs, slice_axis=getattr(pop, "slice_axis", 0),
                                 results_folder_name=str(our_dir), conv_crit=pop.conv_crit, conv_count_min=pop.conv_count_min, conv_count=0,
                                 finish_sim=False, args=args, subvol_volume=np.asarray(pop.subvol_volume))
    Population.initialise_residue(stub, geo)
    # the reference's residue at step 200 compares with the means of step 100: replay both checks
    their_rows = (ref_dir / "residue.txt").read_text().splitlines()
    full = (ref_dir / "convergence.txt").read_text().splitlines()
    for upto in (11, 21):                               # rows of steps 0..100 and 0..200 (header + one row per 10 steps)
        (our_dir / "convergence.txt").write_text("\n".join(full[:upto + 1]) + "\n")
        ours.convergence_file = str(our_dir / "convergence.txt")
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            ours.read_convergence()
            Population.update_residue(stub, geo)
        if upto == 11:
            # run_timestep(200) dumps the state BEFORE it post-processes: the files left by the reference carry the rolling
            # statistics of the step-100 check (Population.py:1729-1733), ours must too
            Population.write_subvolume_state(stub, geo)
    our_rows = (our_dir / "residue.txt").read_text().splitlines()
    assert our_rows[-1] == their_rows[-1] and len(our_rows) == 2
    keep = lambda p: [l for l in p.read_text().splitlines() if "Date and time" not in l]
    assert keep(our_dir / "subvolumes.txt") == keep(ref_dir / "subvolumes.txt")
    if not slice_type:
        assert keep(our_dir / "subvol_connections.txt") == keep(ref_dir / "subvol_connections.txt")
