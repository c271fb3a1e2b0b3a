"""Not collected by pytest: evidence for DESIGN.md section 2 (reference quirks).

`--temp_interp radial` feeds every particle the cubic-RBF temperature field through the subvolume centres.  With six voronoi
centres that are nearly coplanar (np.random.seed(11) below) the field overshoots by tens of kelvin, the particles relax
towards those temperatures and the subvolume temperatures run away within ~15 steps -- the feedback loop the reference's
author describes at Population.py:577-586.  This script steps the CUDA path and the oracle (same keyed draws) side by side
on that geometry: ids identical, T_sv equal to 1e-9 THROUGH the blow-up, i.e. the instability is the algorithm's, and the
GPU reproduces it.

    python tests/run_radial_instability_check.py        (needs a GPU)
"""
import contextlib, io, os, sys, numpy as np
os.environ.setdefault("NK_VORONOI_MAX_SAMPLES", "20000")
os.environ.setdefault("NK_VOLUME_MAX_SAMPLES", "200000")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argument_parser as ap
from oracle import gen_golden, nk_oracle as nko
from nanokappa_b200.classes.Geometry import Geometry
from nanokappa_b200.classes.Phonon import Phonon
from nanokappa_b200.classes.Population import Population
name = sys.argv[1] if len(sys.argv) > 1 else "c6_cylinder_voronoi_radial"
text = gen_golden.CONFIGS[name][0].replace("--particles total 3000", "--particles total 12000")
text = text.replace("kappa-m313131.hdf5", "synthetic:5").replace("--mat_folder test_material/Si/", "--mat_folder /nonexistent/")
args = ap.initialise_parser(False).parse_args(text.split()); args.results_folder = "/tmp/dbg_c6"; os.makedirs("/tmp/dbg_c6", exist_ok=True)
np.random.seed(11)
with contextlib.redirect_stdout(io.StringIO()):
    geo = Geometry(args); ph = Phonon(args, 0); np.random.seed(8); pop = Population(args, geo, ph, device=0, seed=8)
tb = pop.tables
p = pop._particles()
S = tb["sv_centres"].shape[0]
st = nko.make_state(tb, p["positions"], p["modes"], np.asarray(pop.subvol_temperature, dtype=float), pop.res_counter, ids=p["ids"])
print("init occ diff", np.abs(st.occupation - p["occupation"]).max(), "tc diff", np.nanmax(np.abs(st.n_timesteps - p["n_timesteps"])), "T0", pop.subvol_temperature)
rng = nko.KeyedRNG(pop.engine.seed)
eng = pop.engine
with np.errstate(all="ignore"):
    for k in range(1, 81):
        nko.run_timestep(tb, st, rng)
        eng.step(1)
        r = eng.results()
        q = eng.particles()
        order = np.argsort(st.ids)
        same_ids = q["ids"].shape == st.ids.shape and np.array_equal(q["ids"], st.ids[order])
        docc = np.abs(q["occupation"] - st.occupation[order]).max() if same_ids else np.nan
        dpos = np.abs(q["positions"] - st.positions[order]).max() if same_ids else np.nan
        if k % 5 == 0 or not same_ids: print(k, "ids", same_ids, "dT", np.abs(r["subvol_temperature"] - st.subvol_temperature).max(), "docc", docc, "dpos", dpos,
              "Tgpu", np.round(r["subvol_temperature"], 2), "Torc", np.round(st.subvol_temperature, 2), flush=True)
        if k in (1, 2, 5):
            x = q["positions"][:2000]
            print("   particle T gpu-vs-scipy", np.abs(eng.particle_temperature(x) - nko.particle_temperature(tb, r["subvol_temperature"], x)).max())
