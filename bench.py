#!/usr/bin/env python
"""bench.py -- particle-timestep updates/s of the Nano-kappa particle loop on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W     CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[1] geometry at configs[4] scale): Si cross-plane thin film, box
(2e4 A)^3, reservoirs 302 K / 298 K on the x faces, periodic y/z walls, 20 slice subvolumes on x,
nearest-subvolume particle temperature, dt = 1 ps, synthetic 31^3 x 6 mode table (the phono3py hdf5
is not shipped), `--particles` per GPU (default 1e8: 8.4 GB of hot particle state, far beyond the
126 MB L2, so every step streams from HBM).  One "step" = one Population.run_timestep.
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes as C
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UPDATE = 84.0   # x,y,z r/w 48 + tc r/w 16 + occupation r/w 16 + mode id read 4

PARAMS = """
--mat_folder {mat} --hdf_file synthetic:{mesh} --poscar_file POSCAR
--geometry box --dimensions 20e3 20e3 20e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 20 0
--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5
--bound_cond T T P
--connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_values 302 298
--reference_temp local --temp_dist cold --temp_interp nearest
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 10000
--n_mean 10 --results_folder bench --conv_crit 0 10 --output screen --max_sim_time 0-00:00:00
"""


PARAMS_C1 = """
--mat_folder {mat} --hdf_file synthetic:{mesh} --poscar_file POSCAR
--geometry box --dimensions 5e3 1e3 1e3 --scale 1 1 1 --geo_rotation 0 0 0 xyz
--subvolumes slice 10 0
--bound_pos relative -0.1 0.5 0.5 1.1 0.5 0.5 0.5 0.5 -0.1 0.5 0.5 1.1
--bound_cond T T R R P --connect_pos relative 0.5 -0.1 0.5 0.5 1.1 0.5
--bound_values 302 298 {eta} {eta}
--reference_temp local --temp_dist cold --temp_interp linear
--particles total {n} --part_dist random_subvol --timestep 1 --iterations 1000
--n_mean 10 --results_folder bench --conv_crit 0 10 --output screen --max_sim_time 0-00:00:00
"""
CASE = {"name": "c2"}


def workload(n_total, mesh):
    """Geometry / Phonon / host set-up tables of the benchmark case (no GPU needed)."""
    import argument_parser as ap
    from nanokappa_b200.classes.Geometry import Geometry
    from nanokappa_b200.classes.Phonon import Phonon
    from nanokappa_b200.classes.Population import PopulationSetup
    if CASE["name"] == "c1":      # parameters_test.txt geometry (configs[0]): rough walls + linear T interpolation -> general kernel path
        text = PARAMS_C1.format(mat="/nonexistent_material_folder/", mesh=mesh, n=int(n_total), eta=CASE.get("eta", 0))
    else:
        text = PARAMS.format(mat="/nonexistent_material_folder/", mesh=mesh, n=int(n_total)).replace("--subvolumes slice 20 0", "--subvolumes slice {} 0".format(int(CASE.get("slices", 20))))
    if CASE.get("material", "si") == "ge":             # Ge cell of test_material/Ge/POSCAR (configs[2]); synthetic table on that lattice
        text = text.replace("synthetic:{}".format(mesh), "synthetic:{}:ge".format(mesh))
    args = ap.initialise_parser(False).parse_args(text.split())
    args.results_folder = "/tmp"
    with contextlib.redirect_stdout(io.StringIO()):
        geo = Geometry(args)
        ph = Phonon(args, 0)
        np.random.seed(0)
        setup = PopulationSetup(args, geo, ph, seed=0)
    return args, geo, ph, setup, setup.tables(geo, ph)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region, polled every few ms through NVML
    (nvidia-ml-py); falls back to `nvidia-smi -lms` when NVML is not importable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []          # (sm_mhz, power_w, reasons_bitmask)
        self.sm_max = None
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.lines = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                              "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._read, daemon=True)
                self.thread.start()
            except Exception:
                self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(sm), pw, int(rs)))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        self.stop_flag = True
        if self.nv is not None:
            if self.thread is not None:
                self.thread.join(timeout=1)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
            nv = self.nv
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            mask = 0
            for _, _, r in self.samples:
                mask |= r
            return {"sm_mhz": float(np.median([x[0] for x in self.samples])), "sm_max_mhz": self.sm_max,
                    "power_w_max": float(max(x[1] for x in self.samples)), "samples": len(self.samples),
                    "reasons": sorted(k for k, v in names.items() if mask & v)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); smax.append(float(p[1])); power.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy restatement + the SciPy objects the reference itself calls)
# --------------------------------------------------------------------------------------------------
def cpu_run(n_particles, steps, warmup, mesh):
    from oracle import nk_oracle as nko
    args, geo, ph, setup, tb = workload(n_particles, mesh)
    rs = np.random.RandomState(1)
    lo, hi = tb["bounds"]
    pos = lo + rs.random_sample((n_particles, 3)) * (hi - lo)
    act = np.vstack(np.where(~ph.inactive_modes_mask)).T
    modes = act[np.arange(n_particles) % act.shape[0]]
    T0 = np.full(tb["sv_centres"].shape[0], float(np.min(tb["res_T"])))
    st = nko.make_state(tb, pos, modes, T0, setup.res_counter)
    backend = nko.SciPyBackend(tb, collect_garbage=True)
    rng = nko.SequenceRNG()
    np.random.seed(2)
    with np.errstate(all="ignore"):
        for _ in range(warmup):
            nko.run_timestep(tb, st, rng, backend=backend)
        updates = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            nko.run_timestep(tb, st, rng, backend=backend)
            updates += st.N_p
        dt = time.perf_counter() - t0
    return updates / dt, dt, updates


def _cpu_replica(job):
    """One independent replica of the CPU sample (separate process): returns (updates, seconds, start, end)."""
    n, steps, warmup, mesh = job
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    t_start = time.time()
    value, dt, updates = cpu_run(n, steps, warmup, mesh)
    return updates, dt, t_start, time.time()


def cpu_run_all_cores(n_particles, steps, warmup, mesh, procs):
    """The reference is single-threaded Python; what a user with a many-core host does is run independent Monte-Carlo
    replicas.  `procs` replicas of the same sample run in parallel processes; throughput = all their timestep updates
    divided by the time the slowest one spent in its timed loop."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_replica, [(n_particles, steps, warmup, mesh)] * procs)
    updates = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return updates / slowest, slowest, updates


PORT_CALIBRATION = os.path.join(ROOT, "profiles", "r2_port_calibration.json")


def port_calibration_note():
    """Port-vs-unmodified-reference factor measured in the build container (tests/run_port_calibration.py)."""
    try:
        c = json.load(open(PORT_CALIBRATION))
        return "port / unmodified reference on the same inputs (build container, {} particles x {} steps): {:.3f}x".format(
            c["particles"], c["steps"], c["port_over_reference"])
    except Exception:
        return "port calibration file missing"


def reference_arm(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # the CPU arm must not touch the GPU or the CUDA library: E(T) table on the host (cached on disk for the replicas)
    os.environ["NK_ENERGY_TABLE"] = "host"
    os.environ.setdefault("NK_TABLE_CACHE", "/tmp/nk_table_cache")
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    n = int(a.ref_particles)
    procs = max(1, min(int(a.ref_procs) if a.ref_procs else (os.cpu_count() or 1), 32))
    single, dt1, _ = cpu_run(n, a.steps, a.warmup, a.mesh)
    if procs > 1:
        value, dt, updates = cpu_run_all_cores(n, a.steps, a.warmup, a.mesh, procs)
    else:
        value, dt = single, dt1
    line = {
        "impl": "reference", "metric": "particle-timestep updates/s", "value": value, "unit": "updates/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(a, n, "cpu"),
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": procs, "kind": "port", "single_core_value": single,
                         "sample": f"{procs} independent replicas (one process per host core; the reference itself is single-threaded) of "
                                   f"{n} particles x {a.steps} timesteps of the same thin-film case (oracle/nk_oracle.py + SciPy cKDTree / "
                                   f"RegularGridInterpolator + gc.collect as the reference calls them); one replica alone: {single:.3e} updates/s; "
                                   f"host has {os.cpu_count()} cores; {port_calibration_note()}"},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def config_dict(a, n_per_gpu, where):
    wl = ("si_thin_film_crossplane: box (2e4 A)^3, T 302/298 K on x faces, periodic sides, {} slice SVs, nearest T, dt 1 ps ".format(int(CASE.get("slices", 20))) +
          "(BASELINE configs[1] geometry at configs[4] scale)") if CASE["name"] == "c2" else \
         (f"parameters_test geometry: box 5e3x1e3x1e3 A, T/T/R/R/P, eta {CASE.get('eta', 0)} A, 10 slice SVs, linear T (BASELINE configs[0] scaled up; diagnostic)")
    return {"workload": wl,
            "particles_per_gpu": int(n_per_gpu), "mode_table": f"synthetic {a.mesh}^3 x 6 on the {CASE.get('material', 'si').capitalize()} cell", "subvolumes": int(CASE.get("slices", 20)) if CASE["name"] == "c2" else 10,
            "particle_order": "tiled modes (as initialised)" if getattr(a, "no_sort", False) else
            "ordered by mode at set-up (nk_sort_by_mode) with per-mode slot pools, so emitted particles land among their own mode; `value` = first steps after set-up, `sustained` = long run incl. re-sorts",
            "l2_policy": "inputs larger than L2 (no flush)" if n_per_gpu * 44 > 2.6e8 else "state fits L2; L2 flushed between timed steps",
            "parallelism": f"particle shards x{a.gpus}, per-step all-reduce of the per-SV vectors ({getattr(a, 'exchange', '?')})" if a.gpus > 1 else "single GPU"}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def make_ensemble(eng, tb, ph, n, rank, dev):
    """Synthetic ensemble created on the device: uniform in the box, modes tiled over the active modes as
    Population.initialise_modes does for >= 1 particle per mode and subvolume, occupation at the cold temperature."""
    import torch
    from nanokappa_b200.engine import _dp
    from nanokappa_b200._lib import check
    S = tb["sv_centres"].shape[0]
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    lo = torch.as_tensor(tb["bounds"][0], device=dev); ext = torch.as_tensor(tb["bounds"][1] - tb["bounds"][0], device=dev)
    t = eng.t
    for k, name in enumerate(("px", "py", "pz")):
        t[name][:n] = lo[k] + torch.rand(n, generator=g, dtype=torch.float64, device=dev) * ext[k]
    act = torch.as_tensor(np.nonzero(~ph.inactive_modes_mask.reshape(-1))[0].astype(np.int32), device=dev)
    idx = (torch.arange(n, device=dev, dtype=torch.int64) + rank * n) % act.numel()
    t["mode"][:n] = act[idx]; t["omode"][:n] = act[idx]; t["mode"][n:] = -1
    t["pid"][:n] = torch.arange(n, device=dev, dtype=torch.int64) + rank * n
    omega_d = torch.as_tensor(tb["omega"].reshape(-1), device=dev)[t["mode"][:n].long()]
    T0 = float(np.min(tb["res_T"]))
    Td = torch.full((n,), T0, dtype=torch.float64, device=dev)
    check(eng.ctx, eng.L.nk_occupation(eng.ctx, n, _dp(Td), _dp(omega_d), _dp(t["occ"])), "nk_occupation")
    del omega_d, Td, idx
    torch.cuda.synchronize()
    check(eng.ctx, eng.L.nk_set_slot_count(eng.ctx, n), "nk_set_slot_count")
    eng.set_sv_temperature(np.full(S, T0))
    eng.set_timestep(0)
    eng.init_collisions()
    eng.synchronize()


def readme_case(a, dev, steps=200, warmup=20):
    """BASELINE configs[1] at its own size (1e6 particles, one GPU): launch-latency regime, reported next to the headline."""
    import torch
    from nanokappa_b200.engine import Engine
    n = 1000000
    args, geo, ph, setup, tb = workload(n, a.mesh)
    eng = Engine(dev.index, seed=4321)
    eng.set_tables(tb, res_counter=setup.res_counter)
    eng.allocate(int(n * 1.05) + 4096)
    make_ensemble(eng, tb, ph, n, 0, dev)
    eng.sort_by_mode()
    eng.step(warmup)
    torch.cuda.synchronize()
    n0 = eng.results()["N_p"]
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(steps)                      # one multi-step call: the host only enqueues (state stays in HBM; it fits the L2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n1 = eng.results()["N_p"]
    eng.close()
    return {"workload": "BASELINE configs[1] as is: 1e6 particles, (2e4 A)^3 film, 20 slices", "value": 0.5 * (n0 + n1) * steps / (ms * 1e-3),
            "unit": "updates/s", "ms_per_step": ms / steps, "steps": steps, "note": "state (84 MB) stays L2-resident between steps; no flush"}


# --------------------------------------------------------------------------------------------------
def gpu_arm(a):
    import torch
    import torch.distributed as dist
    from nanokappa_b200.engine import Engine, _dp
    from nanokappa_b200._lib import check

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    numa = None
    if world > 1:
        from nanokappa_b200.parallel import bind_to_gpu_numa
        if os.environ.get("NK_NUMA_BIND", "1") != "0":
            numa = bind_to_gpu_numa(local)             # before any pinned allocation: host buffers next to the GPU
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = int(a.particles)
    n_total = n * world
    args, geo, ph, setup, tb = workload(n_total, a.mesh)
    S = tb["sv_centres"].shape[0]
    eng = Engine(local, seed=1234)
    eng.set_tables(tb, res_counter=setup.res_counter)
    eng.allocate(int(n * 1.05) + 4096)
    if world > 1:
        check(eng.ctx, eng.L.nk_set_rank(eng.ctx, rank, world), "nk_set_rank")

    make_ensemble(eng, tb, ph, n, rank, dev)
    if not a.no_sort:
        eng.sort_by_mode()          # set-up-time layout choice: neighbours share mode records

    acc_t, fused = None, False
    if world > 1:
        from nanokappa_b200.parallel import ShardedEngine
        sh = ShardedEngine(eng, rank, world)
        acc_t = sh.acc
        if not a.nccl:
            fused = sh.enable_fused_exchange()       # in-kernel exchange over NVLink peer memory, NCCL as fall-back
    a.exchange = "single GPU" if world == 1 else ("fused in-kernel all-reduce over NVLink peer memory" if fused else "NCCL all-reduce between the step halves")

    flush_buf = None
    if n * 44 <= 2.6e8:
        flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def one_step():
        if flush_buf is not None:
            flush_buf.zero_()
        if world > 1 and not fused:
            eng.step_local()
            dist.all_reduce(acc_t)
            eng.step_finalize()
        else:
            eng.step(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()                                   # ranks leave the host set-up at different times
    for _ in range(max(a.warmup, 3)):
        one_step()
    barrier()
    n_alive0 = eng.results()["N_p"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    eng.profile_begin()
    barrier()
    e0.record()
    for _ in range(a.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    prof, nprof = eng.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    n_alive1 = eng.results()["N_p"]
    slots, alive_local = eng.slot_count()
    if world > 1:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    updates = 0.5 * (n_alive0 + n_alive1) * a.steps          # N_p is the global live count (all ranks)
    value = updates / (ms * 1e-3)

    if os.environ.get("NK_TRACE", "0") != "0" and rank == 0:      # diagnostics only: device-side marks of single steps
        for _ in range(4):
            one_step()
            tr = (C.c_uint64 * 8)()
            check(eng.ctx, eng.L.nk_debug_trace(eng.ctx, tr), "nk_debug_trace")
            t0 = tr[0] or tr[2]
            print("[nk trace] us since streaming kernel start: " + " ".join(
                f"{name}={((tr[i] - t0) / 1e3 if tr[i] else float('nan')):.1f}" for i, name in enumerate(
                    ("step_in", "step_out", "rare_in", "rare_items_done", "finalize_in", "finalize_out"))), file=sys.stderr, flush=True)

    # ---- roofline of the streaming kernel: algorithmic bytes = 84 B x live particles of this rank per launch
    peak, peak_src = load_peaks()
    step_ms = prof["k_step"] / max(nprof, 1)
    achieved = BYTES_PER_UPDATE * alive_local / (step_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "kstep_traffic.json")
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_per_update"] * alive_local          # ncu-measured DRAM bytes per update x this launch's updates
        traffic_src = tj["source"]
    roofline = {"bound": "hbm", "kernel": "k_step", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": BYTES_PER_UPDATE * alive_local, "peak_source": peak_src, "bytes_per_update": BYTES_PER_UPDATE, "updates_per_launch": int(alive_local),
                "avg_launch_ms": step_ms,
                "kernel_share_of_step": {k: v / max(sum(prof.values()), 1e-12) for k, v in prof.items()}}

    # ---- sustained: a long run with the maintenance a long run needs (re-sorts), timed as a whole
    sustained = sustained_run(a, eng, world, rank, local, one_step, barrier, peak) if a.sustained_steps > 0 else None

    # ---- end to end through host buffers (pinned): full particle state H2D, one timestep, state + per-SV results D2H
    e2e = e2e_run(a, eng, n, world, rank, one_step, dev, fused)

    if world > 1:
        e2e_t = torch.tensor([e2e["seconds"]], device=dev, dtype=torch.float64)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e["seconds"] = float(e2e_t.item())
    e2e_value = e2e["updates_global"] / e2e["seconds"]
    ceiling = pcie_ceiling(world, dev, e2e["h2d"] / max(n, 1), e2e["d2h"] / max(n, 1), barrier) if not a.no_ceiling else None

    if rank == 0:
        cpu_val, cpu_dt, cpu_upd = (None, None, None)
        cpu = None
        if world == 1 and not a.no_cpu:
            cpu_val, cpu_dt, cpu_upd = cpu_run(int(a.cpu_particles), a.cpu_steps, 1, a.mesh)
            cpu = {"value": cpu_val, "unit": "updates/s", "cores": 1, "kind": "port",
                   "sample": f"{int(a.cpu_particles)} particles x {a.cpu_steps} timesteps of the same case, {cpu_dt:.1f} s "
                             f"(oracle port with the SciPy calls + gc.collect the reference makes; single-threaded like the reference; "
                             f"host has {os.cpu_count()} cores; {port_calibration_note()})"}
        line = {
            "metric": "particle-timestep updates/s", "value": value, "unit": "updates/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(a, n, "gpu"),
            "clocks": clocks, "gpu_launches": int((2 + (1 if prof.get("k_finalize", 0.0) > 0 else 0)) * a.steps),
            "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "timesteps_per_call": 1, "calls": e2e["calls"], "api": e2e["api"], "numa_node_rank0": numa,
                    "pcie_ceiling": ceiling,
                    "frac_of_pcie_ceiling": (e2e_value / ceiling["equal_work_updates_per_s"]) if ceiling else None,
                    "frac_of_ceiling_with_dependent_download": (e2e_value / ceiling["with_dependent_occupation_download_updates_per_s"]) if ceiling else None},
            "roofline": roofline, "cpu_baseline": cpu, "sustained": sustained,
            "particles_alive": int(n_alive1),
        }
        if world == 1 and CASE["name"] == "c2" and not a.no_cpu and n >= 10 ** 7:
            del eng
            torch.cuda.empty_cache()
            line["readme_case_1e6"] = readme_case(a, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sustained_run(a, eng, world, rank, local, one_step, barrier, peak):
    """`--sustained-steps` consecutive timesteps (default 1000, >= 2 s at 1e8 particles) INCLUDING the maintenance pass a long
    run performs (nk_sort_by_mode every NK_RESORT_EVERY = 500 steps, as Population.run_timestep does), timed as one region with
    CUDA events and its own clock record.  This is the number a 10 000-iteration run lives at; `value` above is the first
    `--steps` steps after set-up."""
    import torch
    import torch.distributed as dist
    steps = int(a.sustained_steps)
    every = int(os.environ.get("NK_RESORT_EVERY", 500))
    sampler = ClockSampler(local)
    n0 = eng.results()["N_p"]
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps // 100 + 2)]
    if rank == 0:
        sampler.start()
    eng.profile_begin()
    barrier()
    marks[0].record()
    resorts, sort_ev = 0, []
    for k in range(steps):
        if every > 0 and k > 0 and k % every == 0:
            s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
            s0.record()
            eng.sort_by_mode()
            s1.record()
            sort_ev.append((s0, s1))
            resorts += 1
        one_step()
        if (k + 1) % 100 == 0:
            marks[(k + 1) // 100].record()
    last = torch.cuda.Event(enable_timing=True)
    last.record()
    barrier()
    ms = marks[0].elapsed_time(last)
    prof, nprof = eng.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    n1 = eng.results()["N_p"]
    _, alive_local = eng.slot_count()
    if world > 1:
        tmax = torch.tensor([ms], device=eng.device, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    per100 = [marks[i].elapsed_time(marks[i + 1]) / 100.0 for i in range(steps // 100)]
    kstep_ms = prof["k_step"] / max(nprof, 1)
    return {"value": 0.5 * (n0 + n1) * steps / (ms * 1e-3), "unit": "updates/s", "steps": steps, "ms_per_step": ms / steps,
            "resort_every": every, "resorts": resorts, "resort_ms_each": sum(x.elapsed_time(y) for x, y in sort_ev) / max(resorts, 1),
            "ms_per_step_by_100": per100, "kstep_avg_ms": kstep_ms,
            "kstep_roofline_frac": BYTES_PER_UPDATE * alive_local / (kstep_ms * 1e-3) / 1e9 / peak, "clocks": clocks}


def pcie_ceiling(world, dev, up_bytes, down_bytes, barrier, gb=1.0):
    """What the host<->device link of THIS box gives the e2e number: every rank copies `gb` GB up and `gb` GB
    down at the same time (two streams, pinned buffers, all ranks together, nothing else running), timed with CUDA events.  A
    step moves `up_bytes` up and `down_bytes` down per particle, so a rank cannot exceed min(h2d / up_bytes, d2h / down_bytes)
    updates/s.  `sum_updates_per_s` adds the ranks' limits; `equal_work_updates_per_s` = world x the slowest rank's limit is the
    one that bounds bench.py's e2e (every rank holds the same number of particles and the time is the max over ranks) -- on the
    8-GPU boxes of this pool the two halves of the node differ by 40 % (tests/run_pcie_ceiling.py, profiles/r2_pcie_ceiling_8gpu.json)."""
    import torch
    import torch.distributed as dist
    nelem = int(gb * 1e9 / 8)
    up_h = torch.empty(nelem, dtype=torch.float64, pin_memory=True).fill_(1.0)
    dn_h = torch.empty(nelem, dtype=torch.float64, pin_memory=True)
    up_d = torch.empty(nelem, dtype=torch.float64, device=dev)
    dn_d = torch.ones(nelem, dtype=torch.float64, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    rate = {}
    for phase in ("both", "d2h_alone"):
        for rep in range(2):                               # first repetition = warm-up
            barrier()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if phase == "both":
                with torch.cuda.stream(s_up):
                    e[0].record(); up_d.copy_(up_h, non_blocking=True); e[1].record()
            with torch.cuda.stream(s_dn):
                e[2].record(); dn_h.copy_(dn_d, non_blocking=True); e[3].record()
            torch.cuda.synchronize()
        if phase == "both":
            rate["h2d"] = gb / (e[0].elapsed_time(e[1]) * 1e-3)
            rate["d2h"] = gb / (e[2].elapsed_time(e[3]) * 1e-3)
        else:
            rate["d2h_alone"] = gb / (e[2].elapsed_time(e[3]) * 1e-3)
    mine = torch.tensor([rate["h2d"], rate["d2h"], rate["d2h_alone"]], device=dev, dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    h2d = [float(v[0]) for v in allv]; d2h = [float(v[1]) for v in allv]; d2h_alone = [float(v[2]) for v in allv]
    lim = [min(u * 1e9 / up_bytes, d * 1e9 / down_bytes) for u, d in zip(h2d, d2h)]
    # the occupations (8 B per particle) are final only after the step's temperatures, which depend on every particle: their
    # download cannot overlap the uploads, so a rank needs at least up_bytes / h2d + 8 / d2h_alone per particle
    dep = [1e9 / (up_bytes / u + 8.0 / d) for u, d in zip(h2d, d2h_alone)]
    del up_h, dn_h, up_d, dn_d
    return {"h2d_gbs_per_rank": h2d, "d2h_gbs_per_rank": d2h, "d2h_alone_gbs_per_rank": d2h_alone, "gb_each_way": gb,
            "bytes_up_per_particle": up_bytes, "bytes_down_per_particle": down_bytes,
            "sum_updates_per_s": sum(lim), "equal_work_updates_per_s": world * min(lim),
            "with_dependent_occupation_download_updates_per_s": world * min(dep),
            "how": "concurrent pinned H2D + D2H copies on two streams (then D2H alone), all ranks at once, CUDA events, measured in this run",
            "note": "the dependent-download figure uses the upload rate seen WITH a concurrent download; where the host memory system "
                    "rather than the link limits the copies (several ranks on one box) the call's uploads run partly without one and "
                    "the e2e value can exceed it -- it is a bound only at 1 rank"}


def e2e_run(a, eng, n, world, rank, one_step, dev, fused=False):
    """One public-API call per step with HOST particle arrays: upload the state, advance one timestep,
    download the state and the per-subvolume results.  N=1 goes through the C-ABI host-buffer entry
    point nk_advance_host (also for N>1 when the per-SV sums are exchanged inside the kernel); with the NCCL all-reduce
    between the step halves the same copies are composed around step_local / all-reduce / finalize."""
    import torch
    import torch.distributed as dist
    from nanokappa_b200._lib import check
    t = eng.t
    slots, _ = eng.slot_count()
    names8 = ("px", "py", "pz", "tc", "occ", "cx", "cy", "cz", "pid")
    names4 = ("mode", "omode", "cfacet")
    cap = eng.cap
    host = {k: torch.empty(cap, dtype=t[k].dtype, pin_memory=True) for k in names8 + names4}
    eng.flush_relaxation()
    for k in host:
        host[k][:slots].copy_(t[k][:slots])
    torch.cuda.synchronize()
    calls = max(1, min(a.steps, a.e2e_calls))
    S = eng.S
    Tsv = np.zeros(S); Esv = np.zeros(S); Nsv = np.zeros(S, dtype=np.int64)
    per = 9 * 8 + 3 * 4
    state = {"n": slots, "h2d": 0, "d2h": 0}

    def one_call():
        n_cur = state["n"]
        if world == 1 or fused:
            n_out = C.c_int64()
            hp = lambda k: C.c_void_p(host[k].data_ptr())
            check(eng.ctx, eng.L.nk_advance_host(eng.ctx, n_cur, 1, hp("px"), hp("py"), hp("pz"), hp("tc"), hp("occ"), hp("mode"),
                                                 hp("omode"), hp("cfacet"), hp("cx"), hp("cy"), hp("cz"), hp("pid"), C.byref(n_out),
                                                 Tsv.ctypes.data_as(C.c_void_p), Esv.ctypes.data_as(C.c_void_p),
                                                 Nsv.ctypes.data_as(C.c_void_p)), "nk_advance_host")
            state["n"] = n_out.value
            up, down = C.c_int64(), C.c_int64()
            check(eng.ctx, eng.L.nk_last_transfer_bytes(eng.ctx, C.byref(up), C.byref(down)), "nk_last_transfer_bytes")
            state["h2d"], state["d2h"] = up.value, down.value + 8 * 3 * S
            return int(Nsv.sum())
        for k in host:
            t[k][:n_cur].copy_(host[k][:n_cur], non_blocking=True)
        one_step()
        eng.flush_relaxation()
        state["n"], _ = eng.slot_count()
        for k in host:
            host[k][:state["n"]].copy_(t[k][:state["n"]], non_blocking=True)
        res = eng.results()                  # reports the global N_p
        state["h2d"], state["d2h"] = per * n_cur, per * state["n"] + 8 * 3 * S
        return res["N_p"]

    for _ in range(min(a.warmup, 1)):        # one untimed call: stream / pinned staging set-up, page-in of the host arrays
        one_call()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    updates = 0
    t0 = time.perf_counter()
    for _ in range(calls):
        updates += one_call()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    return {"seconds": sec, "updates_global": updates, "h2d": int(state["h2d"]), "d2h": int(state["d2h"]),
            "calls": calls, "api": "nk_advance_host (C ABI, pinned host SoA)" if (world == 1 or fused) else "Engine host-buffer step (torch pinned copies + step_local/all_reduce/finalize)"}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", 1)))
    p.add_argument("--no-ceiling", dest="no_ceiling", action="store_true", help="skip the PCIe ceiling measurement next to e2e")
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--particles", type=float, default=1e8, help="particles per GPU")
    p.add_argument("--mesh", type=int, default=31, help="q-mesh of the synthetic mode table (31 -> 29791 x 6 modes)")
    p.add_argument("--cpu-particles", type=float, default=2e5, help="bounded CPU sample of the cpu_baseline leg of the GPU arm")
    p.add_argument("--ref-particles", type=float, default=1e6, help="--impl reference: particles per replica (SURVEY 8d: 1e6)")
    p.add_argument("--sustained-steps", type=int, default=1000, help="steps of the sustained measurement (maintenance included); 0 = skip")
    p.add_argument("--material", default="si", choices=["si", "ge"], help="lattice of the synthetic mode table (ge: configs[2])")
    p.add_argument("--cpu-steps", type=int, default=100)
    p.add_argument("--ref-procs", type=int, default=0, help="--impl reference: parallel replicas (0 = one per host core)")
    p.add_argument("--e2e-calls", type=int, default=3)
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--nccl", action="store_true", help="multi-GPU: use the NCCL all-reduce between the step halves instead of the fused exchange")
    p.add_argument("--no-sort", action="store_true", help="keep the tiled mode order of Population.initialise_modes")
    p.add_argument("--case", default="c2", choices=["c2", "c1"], help="c2: README cross-plane film (headline); c1: parameters_test.txt geometry (diagnostic)")
    p.add_argument("--eta", type=float, default=0.0, help="roughness of the R facets in --case c1")
    p.add_argument("--slices", type=int, default=20, help="slice subvolumes of --case c2 (20 = README case; 100 = the configs[2] layout, diagnostic)")
    a = p.parse_args()
    CASE["name"] = a.case; CASE["eta"] = a.eta; CASE["slices"] = a.slices; CASE["material"] = a.material
    if a.impl == "reference":
        reference_arm(a)
    else:
        gpu_arm(a)


if __name__ == "__main__":
    main()
