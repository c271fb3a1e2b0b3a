"""TEST INFRASTRUCTURE -- converged-run reference numbers produced by EXECUTING the reference.

    python -m oracle.gen_converged          (needs /root/reference; writes tests/golden/converged_film.json)

A short cross-plane film (500 A between the 310 K / 290 K reservoirs, periodic sides, 10 slices, 2e4 particles, the
synthetic 5^3 x 6 mode table) reaches its steady state within a few hundred 1-ps steps.  The unmodified reference
(``Population.run_timestep`` minus the every-100-step file output, see ``extract.reference_step``) is run for 1500 steps
with its own ``np.random`` stream; the convergence rows (every 10 steps: kappa, per-slice temperature and heat flux,
particle count) of the last 1000 steps are stored with their block-averaged standard errors.  The GPU test
``tests/test_gpu_population.py::test_converged_film_matches_reference`` runs the same case through the CUDA path with
independent draws and must land within the combined error bars -- the north star's "kappa within statistical error
of the CPU reference on converged runs".
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

from . import extract, gen_golden

PARAMS = """
--mat_folder test_material/Si/ --hdf_file kappa-m313131.hdf5 --poscar_file POSCAR
--geometry box --dimensions 500 400 400 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 10 0
--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5
--bound_cond T T P
--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_values 310 290
--reference_temp local --temp_dist cold --temp_interp nearest
--particles total 20000 --part_dist random_subvol --timestep 1 --iterations 1500
--n_mean 10 --results_folder x --conv_crit 0 10 --colormap jet --output screen --max_sim_time 0-00:00:00
"""
N_MESH = 5
STEPS, DISCARD = 1500, 500
BLOCKS = 10
OUT = os.path.join(gen_golden.GOLDEN, "converged_film.json")


def block_stats(rows, blocks=BLOCKS):
    """mean and standard error of the mean from `blocks` block averages (rows are correlated in time)."""
    rows = np.asarray(rows, dtype=float)
    n = (rows.shape[0] // blocks) * blocks
    b = rows[:n].reshape(blocks, n // blocks, *rows.shape[1:]).mean(axis=1)
    return b.mean(axis=0), b.std(axis=0, ddof=1) / np.sqrt(blocks)


def summarise(kappa, T, flux, N):
    out = {}
    for name, rows in (("kappa", kappa), ("T", T), ("flux_x", flux), ("N_p", N)):
        m, e = block_stats(rows)
        out[name] = {"mean": np.asarray(m).tolist(), "stderr": np.asarray(e).tolist()}
    return out


def main():
    t0 = time.time()
    args, geo, ph, pop = gen_golden.build_reference(PARAMS, N_MESH, results="/tmp/nk_converged_results")
    np.random.seed(2024)
    kappa, T, flux, N = [], [], [], []
    with np.errstate(all="ignore"):
        for k in range(1, STEPS + 1):
            conv = extract.reference_step(pop, geo, ph)
            if conv is not None and k > DISCARD:
                kappa.append(float(pop.kappa))
                T.append(np.array(pop.subvol_temperature, dtype=float))
                flux.append(np.array(conv["subvol_heat_flux"], dtype=float)[:, 0])
                N.append(float(pop.N_p))
    data = summarise(kappa, T, flux, N)
    data.update(steps=STEPS, discard=DISCARD, blocks=BLOCKS, rows=len(kappa), n_mesh=N_MESH, params=PARAMS,
                generated_by="oracle/gen_converged.py executing /root/reference (Population.run_timestep), np.random.seed(2024)",
                seconds=round(time.time() - t0, 1))
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1)
    print(f"kappa = {data['kappa']['mean']:.4f} +- {data['kappa']['stderr']:.4f} W/mK over {len(kappa)} rows; "
          f"N_p = {data['N_p']['mean']:.0f}; {data['seconds']} s -> {OUT}")


if __name__ == "__main__":
    sys.exit(main())
