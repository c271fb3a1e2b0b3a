"""README / BASELINE configs[1] at full size (1e6 particles, box (2e4 A)^3, T/T + periodic walls, 20 slices,
10 000-step run shortened to 200 steps), checked through size-independent properties: census
conservation, containment, non-negative collision clocks, unique ids, determinism for a fixed seed,
reproducibility of the sums across step-kernel variants, and a warming hot side."""
import contextlib
import io

import numpy as np
import pytest
import torch

import argument_parser as ap

pytestmark = pytest.mark.gpu

PARAMS = """
--mat_folder /nonexistent/ --hdf_file synthetic:11 --poscar_file POSCAR
--geometry box --dimensions 20e3 20e3 20e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 20 0 --bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 --bound_cond T T P
--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 0.5 0.5 -0.1 0.5 0.5 1.1 --bound_values 302 298
--reference_temp local --temp_dist cold --temp_interp nearest --particles total 1e6 --part_dist random_subvol
--timestep 1 --iterations 10000 --n_mean 10 --results_folder x --conv_crit 0 10 --output screen --max_sim_time 0-00:00:00
"""


def _run(tmp, seed, steps, env=None, monkeypatch=None, params=None):
    from nanokappa_b200.classes.Geometry import Geometry
    from nanokappa_b200.classes.Phonon import Phonon
    from nanokappa_b200.classes.Population import Population
    if monkeypatch is not None:
        for k in ("NK_STEP_TAB", "NK_RARE_TILED"):
            monkeypatch.delenv(k, raising=False)
        for k, v in (env or {}).items():
            monkeypatch.setenv(k, v)
    args = ap.initialise_parser(False).parse_args((params or PARAMS).split())
    args.results_folder = str(tmp)
    with contextlib.redirect_stdout(io.StringIO()):
        geo = Geometry(args)
        ph = Phonon(args, 0)
        np.random.seed(seed)
        pop = Population(args, geo, ph, device=0, seed=seed)
        pop.engine.sort_by_mode()
        census = [pop.N_p]
        for _ in range(steps // 10):
            pop.engine.step(10)
            census.append(pop.engine.results()["N_p"])
    return geo, pop, census


def test_readme_case_properties(tmp_path, monkeypatch):
    geo, pop, census = _run(tmp_path / "a", 5, 200, {"NK_STEP_TAB": "force"}, monkeypatch)
    eng = pop.engine
    n_slots, n_alive = eng.slot_count()
    res = eng.results()
    assert n_alive == res["N_p"] == int(res["subvol_N_p"].sum())
    assert abs(n_alive - 1_000_000) < 5_000 and n_slots <= eng.cap
    t = eng.t
    live = t["mode"][:n_slots] >= 0
    assert int(live.sum().item()) == n_alive
    ids = t["pid"][:n_slots][live]
    assert torch.unique(ids).numel() == n_alive
    lo = torch.as_tensor(geo.bounds[0] - 1e-6, device=eng.device); hi = torch.as_tensor(geo.bounds[1] + 1e-6, device=eng.device)
    for k, name in enumerate(("px", "py", "pz")):
        v = t[name][:n_slots][live]
        assert bool(((v >= lo[k]) & (v <= hi[k])).all()), f"{name} left the box"
    assert bool((t["tc"][:n_slots][live] >= 0).all()), "a live particle kept an expired collision clock"
    assert bool((t["occ"][:n_slots][live] >= 0).all()) and bool(torch.isfinite(t["occ"][:n_slots][live]).all())
    T = res["subvol_temperature"]
    assert T[0] > 298.05 and abs(T[-1] - 298.0) < 0.05 and (np.diff(T[:6]) < 0).all()      # heat enters from the 302 K side
    assert max(abs(np.diff(census))) < 2_000                                               # emission ~ absorption
    # same seed, other kernel variant: identical integers, sums equal to rounding
    geo2, pop2, census2 = _run(tmp_path / "b", 5, 200, {"NK_STEP_TAB": "0"}, monkeypatch)
    assert census2 == census
    r2 = pop2.engine.results()
    assert np.array_equal(r2["subvol_N_p"], res["subvol_N_p"])
    assert np.allclose(r2["subvol_temperature"], T, rtol=1e-12, atol=0)
    a, b = pop.engine.particles(), pop2.engine.particles()
    assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["modes"], b["modes"])
    assert np.array_equal(a["collision_facets"], b["collision_facets"]) and np.array_equal(a["positions"], b["positions"])
    # same seed, same variant, again: block sums are merged in fixed point (order-independent), draws are keyed by particle,
    # so a run is a pure function of its seed -- temperatures, energies and occupations repeat BIT FOR BIT although slot
    # assignment and block scheduling differ from run to run
    geo3, pop3, census3 = _run(tmp_path / "c", 5, 200, {"NK_STEP_TAB": "force"}, monkeypatch)
    r3 = pop3.engine.results()
    assert census3 == census
    assert np.array_equal(r3["subvol_temperature"], T) and np.array_equal(r3["subvol_energy"], res["subvol_energy"])
    assert np.array_equal(r3["subvol_heat_flux"], res["subvol_heat_flux"])
    c = pop3.engine.particles()
    assert np.array_equal(c["ids"], a["ids"]) and np.array_equal(c["occupation"], a["occupation"])
    assert np.array_equal(c["positions"], a["positions"]) and np.array_equal(c["n_timesteps"], a["n_timesteps"])


# two populations above the 2^20-particle threshold of the pipelined host-buffer call: the cross-plane film (no rough
# facet: `omode` travels sparsely) and a rough-walled bar with linear temperature interpolation (general kernel path)
PIPE_CASES = {
    "film": PARAMS.replace("--particles total 1e6", "--particles total 1.2e6"),
    "rough_bar": PARAMS.replace("--particles total 1e6", "--particles total 1.2e6").replace("--dimensions 20e3 20e3 20e3", "--dimensions 5e3 1e3 1e3")
                       .replace("--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 --bound_cond T T P", "--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 0.5 0.5 -0.1 0.5 0.5 1.1 --bound_cond T T R R P")
                       .replace("--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 0.5 0.5 -0.1 0.5 0.5 1.1 --bound_values 302 298", "--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 --bound_values 302 298 5 5")
                       .replace("--temp_interp nearest", "--temp_interp linear").replace("synthetic:11", "synthetic:7"),
}


@pytest.mark.parametrize("case", sorted(PIPE_CASES))
def test_host_buffer_call_pipelined_equals_simple(case, tmp_path, monkeypatch):
    """nk_advance_host (host SoA in, one timestep, host SoA out): the chunked pipeline (H2D / kernel / D2H overlapped,
    cold arrays returned as a patch of the rewritten slots) must hand back exactly what the plain
    upload-step-download version does, call after call."""
    import ctypes as C
    from nanokappa_b200._lib import check
    out = {}
    for label, env in (("pipelined", {"NK_HOST_PIPELINE": "1"}), ("simple", {"NK_HOST_PIPELINE": "0"}),
                       ("patch_overflow", {"NK_HOST_PIPELINE": "1", "NK_PIPE_PATCH_CAP": "64"}),
                       ("dense_cold", {"NK_HOST_PIPELINE": "1", "NK_HOST_SPARSE": "0"})):
        monkeypatch.delenv("NK_PIPE_PATCH_CAP", raising=False)
        monkeypatch.delenv("NK_HOST_SPARSE", raising=False)
        geo, pop, _ = _run(tmp_path / label, 3, 0, dict(env, NK_STEP_TAB="0"), monkeypatch, params=PIPE_CASES[case])
        eng = pop.engine
        n, _ = eng.slot_count()
        assert n >= 1 << 20, "below the threshold of the pipelined call: the test would compare the simple path with itself"
        names = ("px", "py", "pz", "tc", "occ", "mode", "omode", "cfacet", "cx", "cy", "cz", "pid")
        host = {k: torch.empty(eng.cap, dtype=eng.t[k].dtype, pin_memory=True) for k in names}
        for k in names:
            host[k][:n].copy_(eng.t[k][:n])
        torch.cuda.synchronize()
        S = eng.S
        Tsv = np.zeros(S); Esv = np.zeros(S); Nsv = np.zeros(S, dtype=np.int64)
        hp = lambda k: C.c_void_p(host[k].data_ptr())
        for call in range(6):
            if call == 4:
                # the host arrays are the state: reorder them between calls (ordered by id, so every variant does the
                # same), which makes whatever the device still holds from the previous call wrong for almost every slot
                # -- a hit particle whose cold fields were not uploaded would show up below
                key = torch.where(host["mode"][:n] >= 0, host["pid"][:n], torch.full((n,), 2 ** 63 - 1, dtype=torch.int64))
                perm = torch.argsort(key, stable=True).flip(0)
                for k in names:
                    host[k][:n] = host[k][:n][perm]
            n_out = C.c_int64()
            check(eng.ctx, eng.L.nk_advance_host(eng.ctx, n, 1, *[hp(k) for k in names], C.byref(n_out),
                                                 Tsv.ctypes.data_as(C.c_void_p), Esv.ctypes.data_as(C.c_void_p), Nsv.ctypes.data_as(C.c_void_p)),
                  "nk_advance_host")
            n = n_out.value
        live = (host["mode"][:n] >= 0).numpy()
        order = np.argsort(host["pid"][:n].numpy()[live])
        out[label] = ({k: host[k][:n].numpy()[live][order] for k in names}, Tsv.copy(), Nsv.copy(), n)
    b = out["simple"]
    for label in ("pipelined", "patch_overflow", "dense_cold"):
        a = out[label]
        assert np.array_equal(a[2], b[2]) and int(a[2].sum()) == a[0]["pid"].shape[0]
        for k in a[0]:
            assert np.array_equal(a[0][k], b[0][k], equal_nan=True), f"{k}: {label} host-buffer call differs from the simple one"
        assert np.allclose(a[1], b[1], rtol=1e-12, atol=0)
