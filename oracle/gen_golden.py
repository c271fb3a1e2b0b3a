"""TEST INFRASTRUCTURE -- regenerate ``tests/golden/*.npz`` by EXECUTING the unmodified reference.

Run in the build container (needs ``/root/reference``):

    python -m oracle.gen_golden

Each fixture holds, for one small configuration:
  tb_*   the static tables the reference computed (oracle/extract.py)
  st0_*  the particle state right after ``Population.__init__``
  ref{k}_* the reference's own arrays after k steps driven by NumPy's global generator seeded with
         SEED_STEPS (``reference_step`` = run_timestep without the every-100-step output branch)
The fixtures pin ``oracle/nk_oracle.py`` (sequence RNG, bit-exact) on machines without the
reference, and give the GPU tests reference-built tables (specular LUTs, roulette, mesh planes).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

from . import extract, nk_oracle as nko, ref_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED_INIT = 7
SEED_STEPS = 12345
CHECK_STEPS = (1, 10, 20)

PARAMS_C1 = """
--mat_folder test_material/Si/ --hdf_file kappa-m313131.hdf5 --poscar_file POSCAR
--geometry box --dimensions 5e3 1e3 1e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 10 0
--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_cond T T R R P --connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5
--bound_values 302 298 {eta} {eta}
--reference_temp local --temp_dist cold --temp_interp linear
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 1000
--n_mean 10 --results_folder x --conv_crit 0 10 --colormap jet --output screen --max_sim_time 0-00:00:00
"""

# README cross-plane thin film (README.md:97-110) scaled down: T,T + periodic sides, nearest T
PARAMS_C2 = """
--mat_folder test_material/Si/ --hdf_file kappa-m313131.hdf5 --poscar_file POSCAR
--geometry box --dimensions 20e3 20e3 20e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 20 0
--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5
--bound_cond T T P
--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_values 302 298
--reference_temp local --temp_dist cold --temp_interp nearest
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 1000
--n_mean 10 --results_folder x --conv_crit 0 10 --colormap jet --output screen --max_sim_time 0-00:00:00
"""

# C4-like: faceted cylinder (40 triangles), voronoi subvolumes (nearest T), rough side walls, reservoirs on the caps
PARAMS_C4 = """
--mat_folder test_material/Si/ --hdf_file kappa-m313131.hdf5 --poscar_file POSCAR
--geometry cylinder --dimensions 3000 600 10 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes voronoi 6
--bound_pos relative 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_cond T T R
--bound_values 305 295 {eta}
--reference_temp local --temp_dist cold --temp_interp nearest
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 1000
--n_mean 10 --results_folder x --conv_crit 0 10 --colormap jet --output screen --max_sim_time 0-00:00:00
"""

# box with a 3 x 2 x 1 grid of subvolumes (collapsed z -> 2-D interpolator), cubic RBF temperature, rough walls
PARAMS_C5 = """
--mat_folder test_material/Si/ --hdf_file kappa-m313131.hdf5 --poscar_file POSCAR
--geometry box --dimensions 4e3 2e3 1e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes grid 3 2 1
--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_cond T T R R P --connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5
--bound_values 303 297 {eta} {eta}
--reference_temp local --temp_dist cold --temp_interp radial
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 1000
--n_mean 10 --results_folder x --conv_crit 0 10 --colormap jet --output screen --max_sim_time 0-00:00:00
"""

# C4 proper: an STL-IMPORTED mesh (finely faceted cylinder written by the reference's own Mesh.export_stl, 4 x 160 = 640
# triangles, 162 facets), voronoi subvolumes, rough side walls (eta > 0) -> ray-mesh intersection over many tiles, facet
# tables beyond the shared-memory copies of the rare-path kernel.  `{stl}` is replaced by the generated file.
PARAMS_C9 = """
--mat_folder test_material/Si/ --hdf_file kappa-m313131.hdf5 --poscar_file POSCAR
--geometry {stl} --dimensions 1 1 1 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes voronoi 8
--bound_pos relative 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_cond T T R
--bound_values 304 296 {eta}
--reference_temp local --temp_dist cold --temp_interp nearest
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 1000
--n_mean 10 --results_folder x --conv_crit 0 10 --colormap jet --output screen --max_sim_time 0-00:00:00
"""
STL_SIDES = 160


def write_reference_stl(folder):
    """A faceted cylinder (L 2500 A, R 500 A, STL_SIDES sides) exported as ASCII STL by the reference's Mesh.export_stl
    (Mesh.py:953-975); returns the path.  trimesh is absent here, so Geometry.load_geo_file's tm.load (Geometry.py:82) is
    served by this repository's STL reader (vertices merged like trimesh does)."""
    import sys
    import types
    from nanokappa_b200.classes.Mesh import read_stl
    ref = rh.load_reference()
    args = rh.parse_parameters_text(PARAMS_C4.format(eta=3, n=100).replace("--dimensions 3000 600 10", "--dimensions 2500 500 {}".format(STL_SIDES))
                                    .replace("--subvolumes voronoi 6", "--subvolumes slice 2 2"), "/tmp/nk_golden_results",
                                    overrides=dict(fig_plot=[], output=["screen"]))
    os.makedirs(folder, exist_ok=True)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        geo = ref.Geometry(args)
        geo.mesh.export_stl("nk_c9_cylinder", folder)
    path = os.path.join(folder, "nk_c9_cylinder.stl")

    def _load(name, *a, **k):
        v, f = read_stl(name)
        return types.SimpleNamespace(vertices=v, faces=f)
    sys.modules["trimesh"].__dict__["load"] = _load
    return path


def config_text(name, folder="/tmp/nk_golden_stl"):
    """Parameter text of a configuration with its generated input files in place.  For the STL case the cylinder is
    written by THIS repository's Geometry / Mesh.export_stl (same primitive, same text format as the reference's
    exporter), so boxes without /root/reference can rebuild the input of the fixture."""
    text = CONFIGS[name][0]
    if "{stl}" not in text:
        return text
    import argument_parser as ap
    from nanokappa_b200.classes.Geometry import Geometry
    os.makedirs(folder, exist_ok=True)
    t = PARAMS_C4.format(eta=3, n=100).replace("--dimensions 3000 600 10", "--dimensions 2500 500 {}".format(STL_SIDES)) \
                 .replace("--subvolumes voronoi 6", "--subvolumes slice 2 2").replace("--colormap jet", "")
    args = ap.initialise_parser(False).parse_args(t.split())
    args.results_folder = folder
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        geo = Geometry(args)
        geo.mesh.export_stl("nk_c9_cylinder_host", folder)
    return text.replace("{stl}", os.path.join(folder, "nk_c9_cylinder_host.stl"))


CONFIGS = {
    # name: (parameter text, table mesh n, lattice)
    "c1_specular": (PARAMS_C1.format(eta=0, n=4000), 5),
    "c1_diffuse": (PARAMS_C1.format(eta=10, n=4000), 5),
    "c1_mixed": (PARAMS_C1.format(eta=0.5, n=4000), 5),
    "c2_crossplane": (PARAMS_C2.format(n=6000), 5),
    "c4_cylinder_voronoi": (PARAMS_C4.format(eta=3, n=3000), 5),
    "c5_box_grid_radial": (PARAMS_C5.format(eta=2, n=3000), 5),
    "c6_cylinder_voronoi_radial": (PARAMS_C4.format(eta=3, n=3000).replace("--temp_interp nearest", "--temp_interp radial"), 5),
    # the two debug emission modes of Population.fill_reservoirs on the cross-plane film
    "c7_fixed_rate": (PARAMS_C2.format(n=3000) + "--reservoir_gen fixed_rate\n", 5),
    "c8_one_to_one": (PARAMS_C2.format(n=3000) + "--reservoir_gen one_to_one\n", 5),
    "c9_stl_voronoi": (PARAMS_C9.format(eta=3, n=3000, stl="{stl}"), 5),
}

STATE_FIELDS = ("positions", "modes", "omega", "group_vel", "occupation", "n_timesteps", "collision_facets",
                "collision_positions", "collision_cond", "temperatures", "ids", "subvol_temperature", "res_counter",
                "subvol_id", "energies", "subvol_energy", "subvol_N_p", "subvol_heat_flux", "res_energy_balance",
                "res_heat_flux", "omega_modes")
REF_FIELDS = ("positions", "modes", "omega", "occupation", "n_timesteps", "collision_facets", "collision_positions",
              "subvol_id", "subvol_energy", "subvol_temperature", "subvol_N_p", "res_counter", "N_leaving", "temperatures")


def build_reference(text, n_mesh, results="/tmp/nk_golden_results"):
    from nanokappa_b200 import synthetic
    args = rh.parse_parameters_text(text, results, overrides=dict(fig_plot=[], output=["screen"]))
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        geo = rh.make_geometry(args)
        ph = rh.make_phonon(args, synthetic.make_table(n_mesh))
        pop = rh.make_population(args, geo, ph, seed=SEED_INIT)
    return args, geo, ph, pop


def pack_tables(tb):
    out = {}
    for k, v in tb.items():
        out["tb_" + k] = np.asarray(v)
    return out


def unpack(npz, prefix):
    out = {}
    for k in npz.files:
        if k.startswith(prefix):
            v = npz[k]
            out[k[len(prefix):]] = v.item() if v.shape == () else v
    return out


def load_fixture(path):
    """-> (tb, st0, refs{k: dict})  used by the tests."""
    z = np.load(path, allow_pickle=False)
    tb = unpack(z, "tb_")
    for k in ("temp_interp", "res_gen"):
        if k in tb:
            tb[k] = str(tb[k])
    s0 = unpack(z, "st0_")
    st = nko.State(**{k: s0[k] for k in ("positions", "modes", "omega", "group_vel", "occupation", "n_timesteps",
                                         "collision_facets", "collision_positions", "collision_cond", "temperatures",
                                         "ids", "subvol_temperature", "res_counter")})
    for k in ("subvol_id", "energies", "subvol_energy", "subvol_N_p", "subvol_heat_flux", "res_energy_balance",
              "res_heat_flux", "omega_modes"):
        setattr(st, k, s0[k])
    st.N_p = int(st.subvol_N_p.sum())
    st.N_leaving = np.sum(tb["enter_prob"], axis=(1, 2)).round().astype(int)      # Population.py:344
    st.current_timestep = 0
    refs = {}
    for k in z.files:
        if k.startswith("ref") and "_" in k:
            step, name = k[3:].split("_", 1)
            refs.setdefault(int(step), {})[name] = z[k]
    return tb, st, refs


def generate(name, text, n_mesh):
    if "{stl}" in text:
        text = text.replace("{stl}", write_reference_stl("/tmp/nk_golden_results"))
    args, geo, ph, pop = build_reference(text, n_mesh)
    tb = extract.tables_from_reference(geo, ph, pop)
    st0 = extract.state_from_reference(ph, pop)
    data = pack_tables(tb)
    for f in STATE_FIELDS:
        data["st0_" + f] = np.array(getattr(st0, f), copy=True)
    np.random.seed(SEED_STEPS)
    with np.errstate(all="ignore"):
        for k in range(1, max(CHECK_STEPS) + 1):
            conv = extract.reference_step(pop, geo, ph)
            if k in CHECK_STEPS:
                for f in REF_FIELDS:
                    data[f"ref{k}_{f}"] = np.array(getattr(pop, f), copy=True)
                data[f"ref{k}_collision_cond"] = np.array([nko.BC_CODE[c] for c in pop.collision_cond], dtype=np.int8)
            if conv is not None:
                for f, v in conv.items():
                    data[f"ref{k}_conv_{f}"] = np.array(v, copy=True)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **data)
    print(f"{name}: N0={st0.positions.shape[0]} N{max(CHECK_STEPS)}={pop.positions.shape[0]} -> {path} "
          f"({os.path.getsize(path) / 1e6:.2f} MB)")


def main(argv=None):
    names = (argv or sys.argv[1:]) or list(CONFIGS)
    for name in names:
        text, n_mesh = CONFIGS[name]
        generate(name, text, n_mesh)


if __name__ == "__main__":
    main()
