"""Not collected by pytest; runs in the BUILD CONTAINER only (needs /root/reference): how fast the CPU arm's port
(oracle/nk_oracle.py with the SciPy objects and the per-step gc.collect the reference uses) is relative to the UNMODIFIED
reference (loaded by oracle/ref_harness.py) on the same inputs.  `bench.py --impl reference` times the port because the
reference cannot travel to the GPU box; this factor is quoted in its cpu_baseline.sample.

    python tests/run_port_calibration.py [particles=200000] [steps=15] [mesh=11]   ->  profiles/r2_port_calibration.json
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import extract, gen_golden, nk_oracle as nko      # noqa: E402


def main():
    kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    n = int(float(kv.get("particles", 200000))); steps = int(kv.get("steps", 15)); mesh = int(kv.get("mesh", 11))
    text = gen_golden.PARAMS_C2.format(n=n)
    with contextlib.redirect_stdout(io.StringIO()):
        args, geo, ph, pop = gen_golden.build_reference(text, mesh, results="/tmp/nk_port_calibration")
    tb = extract.tables_from_reference(geo, ph, pop)
    st = extract.state_from_reference(ph, pop)
    backend = nko.SciPyBackend(tb, collect_garbage=True)
    rng = nko.SequenceRNG()
    out = {"particles": n, "steps": steps, "mode_table": f"synthetic {mesh}^3 x 6", "case": "README cross-plane film (BASELINE configs[1] geometry)"}
    with np.errstate(all="ignore"):
        np.random.seed(5)
        extract.reference_step(pop, geo, ph)                     # warm-up of both
        np.random.seed(5)
        nko.run_timestep(tb, st, rng, backend=backend)
        np.random.seed(6)
        t0 = time.perf_counter(); upd = 0
        for _ in range(steps):
            extract.reference_step(pop, geo, ph); upd += pop.positions.shape[0]
        t_ref = time.perf_counter() - t0
        out["reference_updates_per_s"] = upd / t_ref
        np.random.seed(6)
        t0 = time.perf_counter(); upd = 0
        for _ in range(steps):
            nko.run_timestep(tb, st, rng, backend=backend); upd += st.N_p
        t_port = time.perf_counter() - t0
        out["port_updates_per_s"] = upd / t_port
    out["port_over_reference"] = out["port_updates_per_s"] / out["reference_updates_per_s"]
    out["same_census_after_run"] = bool(pop.positions.shape[0] == st.positions.shape[0] and np.array_equal(pop.positions, st.positions))
    out["host"] = f"{os.cpu_count()} cores, single thread each"
    path = os.path.join(ROOT, "profiles", "r2_port_calibration.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
