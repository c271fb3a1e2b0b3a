"""Not collected by pytest: the command-line driver under torchrun on >= 2 GPUs.

    python tests/run_multi_gpu_cli.py [N]     (launches `torchrun --nproc-per-node N nanokappa.py -ff ...` itself)

Runs the README film (2e5 particles, 300 steps, device-side initialisation) once on one GPU and once sharded over N
GPUs with the same parameters file, then compares what rank 0 wrote: same number of convergence rows, particle counts
within 0.5 %, slice temperatures within the Monte-Carlo noise, per-rank particle dumps adding up to N_p."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS = """
--mat_folder /nonexistent/ --hdf_file synthetic:7 --poscar_file POSCAR
--geometry box --dimensions 20e3 20e3 20e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 20 0 --bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 --bound_cond T T P
--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 0.5 0.5 -0.1 0.5 0.5 1.1 --bound_values 302 298
--reference_temp local --temp_dist cold --temp_interp nearest --particles total 4e5 --part_dist random_subvol
--timestep 1 --iterations 300 --n_mean 10 --results_folder {folder} --conv_crit 0 10 --output screen --max_sim_time 0-00:10:00
"""


def rows(folder):
    txt = [l for l in open(os.path.join(folder, "convergence.txt")).read().splitlines() if l and not l.startswith("#")]
    return np.array([[float(x) for x in l.split()[1:]] for l in txt])      # skip the wall-clock token


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    tmp = tempfile.mkdtemp(prefix="nk_cli_")
    out = {}
    for label, cmd in (("single", [sys.executable, os.path.join(ROOT, "nanokappa.py")]),
                       ("sharded", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                                    "--master-port", "29611", os.path.join(ROOT, "nanokappa.py")])):
        pfile = os.path.join(tmp, label + ".txt")
        open(pfile, "w").write("\n".join(l for l in PARAMS.format(folder=os.path.join(tmp, label)).splitlines() if l.strip()) + "\n")
        env = dict(os.environ)
        if label == "single":
            for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
                env.pop(k, None)
        r = subprocess.run(cmd + ["-ff", pfile], cwd=tmp, capture_output=True, text=True, timeout=900, env=env)
        if r.returncode != 0:
            print(label, "FAILED\n", r.stdout[-1500:], r.stderr[-3000:])
            sys.exit(1)
        folder = [os.path.join(tmp, d) for d in os.listdir(tmp) if d.startswith(label + "_")][0]
        out[label] = (folder, rows(folder), r.stdout)
    a, b = out["single"][1], out["sharded"][1]
    ok = a.shape == b.shape
    # columns: timestep, simulation time, total energy, 2 reservoir balances, 2 x 3 reservoir fluxes, N_p, T (20), ...
    NP, T0 = 11, 12
    ok &= bool(np.array_equal(a[:, 0], b[:, 0]))
    ok &= bool(np.abs(a[:, NP] - b[:, NP]).max() <= 0.005 * a[:, NP].max())
    dT = np.abs(a[-1, T0:T0 + 20] - b[-1, T0:T0 + 20]).max()
    ok &= bool(dT < 0.05 and a[-1, T0] > 298.05 and b[-1, T0] > 298.05)
    folder = out["sharded"][0]
    dumps = 0
    for r in range(n):
        f = os.path.join(folder, "particle_data.npz" if r == 0 else f"rank{r}/particle_data.npz")
        ft = f.replace(".npz", ".txt")
        if os.path.isfile(ft):                      # the dump at the end of the run (text up to NK_TEXT_DUMP_MAX particles)
            dumps += np.loadtxt(ft, delimiter=",", comments="#").shape[0]
        elif os.path.isfile(f):
            dumps += np.load(f)["positions"].shape[0]
    ok &= bool(dumps == int(b[-1, NP]))
    print(f"rows {a.shape} vs {b.shape}; N_p end {a[-1, NP]:.0f} vs {b[-1, NP]:.0f}; T[0] end {a[-1, T0]:.3f} vs {b[-1, T0]:.3f}; max |dT| last row {dT:.4f} K; particles in the {n} dumps {dumps}")
    print("MULTI_GPU_CLI", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
