"""Not collected by pytest: how the time per step evolves over a long run as particles are absorbed and emitted, with and
without the per-mode slot pools, and what the maintenance sort costs.  Prints one line per checkpoint.

    python tests/run_order_decay.py [particles] [label=ENV1=v1,ENV2=v2 ...]

Each label runs the same 2000+ steps in a fresh context with the given environment (default set: pools as shipped, no pools)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                     # noqa: E402
from nanokappa_b200.engine import Engine         # noqa: E402


def run(label, env, n, tb, ph, setup, dev):
    keys = ("NK_MODE_POOLS", "NK_POOL_FRAC", "NK_POOL_FIXED", "NK_POOL_MIN")
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = Engine(0, seed=1234)
    eng.set_tables(tb, res_counter=setup.res_counter)
    eng.allocate(int(n * 1.05) + 4096)
    bench.make_ensemble(eng, tb, ph, n, 0, dev)
    eng.sort_by_mode()

    def foreign():
        slots, alive = eng.slot_count()
        md = eng.t["mode"][:slots]
        live = md[md >= 0]
        return int((live[1:] < live[:-1]).sum().item()), slots, alive

    def probe(what, total):
        torch.cuda.synchronize()
        eng.profile_begin()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.step(20); e1.record(); torch.cuda.synchronize()
        prof, nprof = eng.profile_end()
        f, slots, alive = foreign()
        print(f"[{label}] {what:>10s} after {total:5d} steps: {e0.elapsed_time(e1) / 20:.4f} ms/step  k_step {prof['k_step'] / nprof:.4f}  "
              f"k_rare {prof['k_rare'] / nprof:.4f}  slots {slots} alive {alive} order breaks {f}", flush=True)

    done = 0
    eng.step(5); done += 5
    for target in (0, 100, 300, 600, 1000, 2000):
        if target > done:
            eng.step(target - done); done = target
        probe("drifting", done); done += 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.sort_by_mode()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"[{label}] sort_by_mode: {dt * 1e3:.1f} ms", flush=True)
    probe("re-sorted", done)
    eng.close()
    del eng
    torch.cuda.empty_cache()


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 and "=" not in sys.argv[1] else 100000000
    sets = []
    for a in sys.argv[1:]:
        if "=" in a:
            label, _, rest = a.partition("=")
            sets.append((label, dict(kv.split("=") for kv in rest.split(",") if kv)))
    if not sets:
        sets = [("pools", {}), ("no_pools", {"NK_MODE_POOLS": "0"})]
    dev = torch.device("cuda", 0)
    args, geo, ph, setup, tb = bench.workload(n, 31)
    for label, env in sets:
        run(label, env, n, tb, ph, setup, dev)


if __name__ == "__main__":
    main()
