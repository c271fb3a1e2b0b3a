"""Not collected by pytest: throughput of the tiled ray-triangle kernels (Mesh.find_boundary, Mesh.py:806-856) on large meshes.

    python tests/run_intersection_bench.py [sides=250,2500] [rays=1e7] [steps=0]

For each `sides` a faceted cylinder with 4 x sides triangles is built by this repository's Geometry (the same primitive the
STL fixture c9 is exported from), `rays` particles are placed inside it with random modes of a 11^3 x 6 table, and
nk_init_collisions (P = N rays against all F triangles, the set-up cost that dominates start-up at N >= 1e8) is
timed with CUDA events, next to the nk_find_boundary operator.  Prints one JSON line per mesh: (ray, triangle) pairs per
second and the fraction of the FP64-pipe bound that the plane test alone sets (11 DADD/DMUL per pair, 64 FP64 lanes per SM).
With steps > 0 it also times that many timesteps (rough walls, voronoi subvolumes: k_rare_tiled)."""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NK_VORONOI_MAX_SAMPLES", "20000")
os.environ.setdefault("NK_VOLUME_MAX_SAMPLES", "2000")      # subvolume volumes by Monte Carlo: irrelevant for the timing

PARAMS = """
--mat_folder /nonexistent/ --hdf_file synthetic:{mesh} --poscar_file POSCAR
--geometry cylinder --dimensions 25000 2500 {sides} --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 10 2
--bound_pos relative 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_cond T T P
--bound_values 304 296
--reference_temp local --temp_dist cold --temp_interp nearest
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 1000
--n_mean 10 --results_folder /tmp --conv_crit 0 10 --output screen --max_sim_time 0-00:00:00
"""


def build(sides, n, mesh=11):
    import argument_parser as ap
    from nanokappa_b200.classes.Geometry import Geometry
    from nanokappa_b200.classes.Phonon import Phonon
    from nanokappa_b200.classes.Population import PopulationSetup
    # side walls are given the LAST boundary condition: periodic would need partner facets, so make them reservoirs-free
    # specular walls without LUTs: 'T T P' is replaced below by plain absorbing caps + a third condition that is never hit
    text = PARAMS.format(mesh=mesh, sides=sides, n=int(n)).replace("--bound_cond T T P", "--bound_cond T T R").replace(
        "--bound_values 304 296", "--bound_values 304 296 0")
    args = ap.initialise_parser(False).parse_args(text.split())
    args.results_folder = "/tmp"
    with contextlib.redirect_stdout(io.StringIO()):
        geo = Geometry(args)
        ph = Phonon(args, 0)
        np.random.seed(0)
        setup = PopulationSetup(args, geo, ph, seed=0)
    return geo, ph, setup, setup.tables(geo, ph)


def main():
    kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    sides_list = [int(s) for s in kv.get("sides", "250,2500").split(",")]
    n = int(float(kv.get("rays", 1e7)))
    steps = int(kv.get("steps", 0))
    reps = int(kv.get("reps", 3))
    from nanokappa_b200.engine import Engine, _dp
    from nanokappa_b200._lib import check
    dev = torch.device("cuda", 0)
    props = torch.cuda.get_device_properties(0)
    for sides in sides_list:
        # the specular LUT of every side facet (F/2 normals x modes) is set-up work that does not matter here: a small table
        geo, ph, setup, tb = build(sides, n, mesh=5 if sides > 400 else 7)
        F = tb["face_normals"].shape[0]
        eng = Engine(0, seed=5)
        eng.set_tables(tb, res_counter=setup.res_counter)
        eng.allocate(int(n * 1.1) + 4096)
        # rays: uniform in the inscribed cylinder, random active modes
        g = torch.Generator(device=dev); g.manual_seed(1)
        lo = tb["bounds"][0]; ext = tb["bounds"][1] - tb["bounds"][0]
        R_in = 0.5 * min(ext[0], ext[1]) * np.cos(np.pi / sides) * 0.999
        r = R_in * torch.sqrt(torch.rand(n, generator=g, dtype=torch.float64, device=dev))
        ph_ = 2 * np.pi * torch.rand(n, generator=g, dtype=torch.float64, device=dev)
        t = eng.t
        t["px"][:n] = lo[0] + 0.5 * ext[0] + r * torch.cos(ph_)
        t["py"][:n] = lo[1] + 0.5 * ext[1] + r * torch.sin(ph_)
        t["pz"][:n] = lo[2] + ext[2] * torch.rand(n, generator=g, dtype=torch.float64, device=dev)
        act = torch.as_tensor(np.nonzero(~ph.inactive_modes_mask.reshape(-1))[0].astype(np.int32), device=dev)
        idx = torch.randint(0, act.numel(), (n,), generator=g, device=dev)
        t["mode"][:n] = act[idx]; t["omode"][:n] = act[idx]; t["mode"][n:] = -1
        t["pid"][:n] = torch.arange(n, device=dev)
        t["occ"][:n] = 0.1
        torch.cuda.synchronize()
        check(eng.ctx, eng.L.nk_set_slot_count(eng.ctx, n), "nk_set_slot_count")
        eng.set_sv_temperature(np.full(tb["sv_centres"].shape[0], 296.0))
        eng.set_timestep(0)
        eng.init_collisions(); torch.cuda.synchronize()           # warm-up
        ms = []
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); eng.init_collisions(); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        miss = int((t["cfacet"][:n] < 0).sum().item())
        # operator seam on device arrays
        x = torch.stack((t["px"][:n], t["py"][:n], t["pz"][:n]), dim=1).contiguous()
        v = torch.as_tensor(tb["group_vel"].reshape(-1, 3), device=dev)[t["mode"][:n].long()].contiguous()
        xc = torch.empty_like(x); tc = torch.empty(n, dtype=torch.float64, device=dev); fc = torch.empty(n, dtype=torch.int32, device=dev)
        check(eng.ctx, eng.L.nk_find_boundary(eng.ctx, n, _dp(x), _dp(v), _dp(xc), _dp(tc), _dp(fc)), "nk_find_boundary"); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        check(eng.ctx, eng.L.nk_find_boundary(eng.ctx, n, _dp(x), _dp(v), _dp(xc), _dp(tc), _dp(fc)), "nk_find_boundary")
        e1.record(); torch.cuda.synchronize()
        ms_op = e0.elapsed_time(e1)
        same = bool((fc == t["cfacet"][:n]).all().item())
        best = min(ms)
        pairs = float(n) * F
        dp_rate = props.multi_processor_count * 64 * 1.965e9          # FP64 instructions per second (one per lane per clock)
        line = {"kernel": "k_init_collisions", "triangles": int(F), "rays": n, "ms": best, "ms_all": ms, "pairs_per_s": pairs / (best * 1e-3),
                "plane_test_bound_ms": pairs * 11 / dp_rate * 1e3, "frac_of_plane_test_bound": (pairs * 11 / dp_rate * 1e3) / best,
                "rays_without_hit": miss, "k_find_boundary_ms": ms_op, "operator_equals_init": same,
                "extrapolated_s_for_1e8_rays": best * 1e-3 * 1e8 / n}
        if steps > 0:
            eng.step(3); torch.cuda.synchronize()
            eng.profile_begin()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); eng.step(steps); e1.record(); torch.cuda.synchronize()
            prof, nprof = eng.profile_end()
            line.update(step_ms=e0.elapsed_time(e1) / steps, k_step_ms=prof["k_step"] / nprof, k_rare_ms=prof["k_rare"] / nprof,
                        rare_variant=eng.last_step_variant(), alive=eng.slot_count()[1])
        print(json.dumps(line), flush=True)
        eng.close(); del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
