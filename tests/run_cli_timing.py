"""Not collected by pytest: the README case (BASELINE configs[1]: 1e6 particles, (2e4 A)^3 film, 20 slices, 10 000 iterations) through
the command line a user runs -- `python nanokappa.py -ff parameters.txt` -- wall-clock, with and without the batching of the
steps between convergence rows, next to the bare `Engine.step` loop of the same population (bench.py's readme_case_1e6).

    python tests/run_cli_timing.py [iterations=10000] [particles=1e6]
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PARAMS = """--mat_folder /nonexistent_material_folder/
--hdf_file synthetic:31
--poscar_file POSCAR
--geometry box
--dimensions 20e3 20e3 20e3
--scale 1 1 1
--geo_rotation 0 0 0 xyz
--subvolumes slice 20 0
--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5
--bound_cond T T P
--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_values 302 298
--reference_temp local
--temp_dist cold
--temp_interp nearest
--particles total {n}
--part_dist random_subvol
--timestep 1
--iterations {it}
--n_mean 10
--results_folder {out}
--conv_crit 0 10
--output file
--max_sim_time 0-00:00:00
"""


def main():
    kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    it = int(kv.get("iterations", 10000)); n = int(float(kv.get("particles", 1e6)))
    out = {}
    for label, env in (("batched", {"NK_STEP_BATCH": "1"}), ("per_step", {"NK_STEP_BATCH": "0"})):
        with tempfile.TemporaryDirectory() as tmp:
            pfile = os.path.join(tmp, "parameters.txt")
            with open(pfile, "w") as fh:
                fh.write(PARAMS.format(n=n, it=it, out=os.path.join(tmp, "run")))
            t0 = time.perf_counter()
            cmd = [sys.executable, os.path.join(ROOT, "nanokappa.py"), "-ff", pfile]
            if kv.get("profile") and label == "batched":            # where the host time of the loop goes
                cmd = [sys.executable, "-m", "cProfile", "-o", os.path.join(ROOT, "gpurun_out", "cli_profile.pstats"), os.path.join(ROOT, "nanokappa.py"), "-ff", pfile]
            r = subprocess.run(cmd, env=dict(os.environ, **env), capture_output=True, text=True, cwd=tmp)
            wall = time.perf_counter() - t0
            text = r.stdout + r.stderr
            for root, _, files in os.walk(tmp):
                if "output.txt" in files:
                    text += open(os.path.join(root, "output.txt")).read()
            m = re.search(r"Time loop: ([0-9.]+) s for (\d+) timesteps\s+\(([0-9.e+]+) particle", text)
            if r.returncode != 0 or not m:
                print(text[-3000:])
                raise SystemExit(f"{label}: nanokappa.py failed (rc {r.returncode})")
            out[label] = {"wall_s_whole_process": wall, "loop_s": float(m.group(1)), "timesteps": int(m.group(2)),
                          "updates_per_s_loop": float(m.group(3)), "us_per_step_loop": 1e6 * float(m.group(1)) / int(m.group(2))}
    # the bare engine loop of the same population
    import torch
    import bench
    a = type("A", (), {"mesh": 31})()
    out["engine_step_loop"] = bench.readme_case(a, torch.device("cuda", 0), steps=2000, warmup=20)
    out["cli_over_engine"] = out["batched"]["us_per_step_loop"] / (1e3 * out["engine_step_loop"]["ms_per_step"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
