"""Not collected by pytest: prints the GPU-side numbers of the converged film next to the reference's
(tests/golden/converged_film.json) as one JSON line -- the record kept under profiles/."""
import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argument_parser as ap                       # noqa: E402
from oracle import gen_converged as gc             # noqa: E402


def main():
    from nanokappa_b200.classes.Geometry import Geometry
    from nanokappa_b200.classes.Phonon import Phonon
    from nanokappa_b200.classes.Population import Population
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "converged_film.json")))
    text = ref["params"].replace("kappa-m313131.hdf5", "synthetic:5").replace("--mat_folder test_material/Si/", "--mat_folder /nonexistent/")
    args = ap.initialise_parser(False).parse_args(text.split())
    out = {}
    for seed in (5, 6, 7):
        args.results_folder = f"/tmp/nk_converged_gpu_{os.getpid()}_{seed}"
        os.makedirs(args.results_folder, exist_ok=True)
        with contextlib.redirect_stdout(io.StringIO()):
            geo = Geometry(args); ph = Phonon(args, 0)
            np.random.seed(seed)
            pop = Population(args, geo, ph, device=0, seed=seed)
            kappa, T, flux, N = [], [], [], []
            for k in range(1, ref["steps"] + 1):
                pop.run_timestep(geo, ph)
                if k % 10 == 0 and k > ref["discard"]:
                    kappa.append(float(pop.kappa)); T.append(pop.subvol_temperature.copy())
                    flux.append(pop.subvol_heat_flux[:, 0].copy()); N.append(float(pop.N_p))
        out[f"gpu_seed{seed}"] = gc.summarise(kappa, T, flux, N)
    out["reference"] = {k: ref[k] for k in ("kappa", "T", "flux_x", "N_p")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
