"""Not collected by pytest: how the streaming kernel's time evolves over a long run as emitted particles land in recycled
slots (eroding the mode order made at set-up), and what re-establishing the order costs.  Prints one line per checkpoint."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                     # noqa: E402
from nanokappa_b200.engine import Engine         # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
    dev = torch.device("cuda", 0)
    args, geo, ph, setup, tb = bench.workload(n, 31)
    eng = Engine(0, seed=1234)
    eng.set_tables(tb, res_counter=setup.res_counter)
    eng.allocate(int(n * 1.05) + 4096)
    bench.make_ensemble(eng, tb, ph, n, 0, dev)
    eng.sort_by_mode()

    def probe(label, total):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.step(20); e1.record(); torch.cuda.synchronize()
        slots, alive = eng.slot_count()
        print(f"{label:>12s} after {total:5d} steps: {e0.elapsed_time(e1) / 20:.4f} ms/step  slots {slots} alive {alive}", flush=True)

    done = 0
    eng.step(5); done += 5
    for target in (0, 100, 300, 600, 1000, 2000):
        if target > done:
            eng.step(target - done); done = target
        probe("drifting", done); done += 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.sort_by_mode()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"sort_by_mode: {dt * 1e3:.1f} ms", flush=True)
    probe("re-sorted", done)


if __name__ == "__main__":
    main()
