"""Device-side particle engine: owns the ``nk_ctx`` and the torch tensors of the particle SoA.

This is the thin layer between the reference-shaped Python classes (``nanokappa_b200.classes``) and
the C ABI (``include/nk_b200.h``).  PyTorch is used for what it is good at here -- device memory,
streams, host<->device copies and (in ``parallel``) ``torch.distributed`` -- every number on the hot
path is produced by the CUDA kernels in ``csrc/``.

``tables`` is a plain dict of NumPy arrays / scalars describing one simulation set-up:

mesh      face_normals (F,3) face_k (F) face_lo/face_hi (F,3) face_origins (F,3) face_basis (F,3,3)
          face_facets (F) face_vertices (F,3,3) face_areas (F)                    [Mesh.py:205-243, :314-324]
facets    facet_bc (nf; 0 T,1 P,2 R,3 F) facet_normal facet_centroid (nf,3) facet_area (nf)
          facet_partner facet_res facet_rough (nf; -1 = none) facet_faces_ptr/facet_faces (CSR) bounds (2,3)
                                                                                 [Geometry.py:652-726]
subvols   sv_centres (S,3) sv_volume (S) sv_slice (bool) slice_axis temp_interp ('nearest'|'linear'|'radial') [interp_dims]
                                                                                 [Geometry.py:446-544]
modes     omega (Q,J) group_vel (Q,J,3) tau (NT,Q,J) T_grid (NT) energy_array/T_array (nE)
          hbar kb volume_unitcell n_active eVpsa2_in_Wm2 a_in_m                 [Phonon.py:66-151, :326-401]
run       dt norm_mean particle_density n_dt_to_conv res_facet (R) res_T (R) enter_prob (R,Q,J)
          specularity true_specular spec_out roulette (Fr,Q,J | Fr,Q*J)         [Population.py:146-161, :852-939]
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import NkError, check

INTERP_CODE = {"nearest": 0, "linear": 1, "radial": 2}
RESGEN_CODE = {"constant": 0, "fixed_rate": 1, "one_to_one": 2}


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _dp(t):
    return C.c_void_p(t.data_ptr())


class Engine:
    """One GPU's share of the particle loop."""

    def __init__(self, device=0, seed=0):
        if not torch.cuda.is_available():
            raise NkError("no CUDA device visible: nanokappa_b200 has no CPU path")
        self.L = _lib.lib()
        self.device = torch.device("cuda", device)
        ctx = C.c_void_p()
        rc = self.L.nk_create(device, C.byref(ctx))
        if rc != 0:
            raise NkError("nk_create failed: " + self.L.nk_last_error(None).decode())
        self.ctx = ctx
        self.seed = int(seed)
        self.tb = None
        self.cap = 0
        self.S = self.R = self.M = self.J = 0
        self.t = {}

    def close(self):
        if getattr(self, "ctx", None):
            self.L.nk_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def set_tables(self, tb, res_counter=None, hot_T=None):
        L, ctx = self.L, self.ctx
        self.tb = tb
        Q, J = tb["omega"].shape
        self.M, self.J = Q * J, J
        F = tb["face_normals"].shape[0]
        nf = tb["facet_bc"].shape[0]
        keep = []   # keep converted arrays alive during the calls

        def a(x, conv=_f64):
            y = conv(x)
            keep.append(y)
            return _p(y)

        check(ctx, L.nk_set_mesh(ctx, F, a(tb["face_normals"]), a(tb["face_k"]), a(tb["face_lo"]), a(tb["face_hi"]),
                                 a(tb["face_origins"]), a(tb["face_basis"]), a(tb["face_facets"], _i32),
                                 a(tb["face_vertices"]), a(tb["face_areas"]), nf, a(tb["facet_bc"], _i32),
                                 a(tb["facet_partner"], _i32), a(tb["facet_res"], _i32), a(tb["facet_rough"], _i32),
                                 a(tb["facet_normal"]), a(tb["facet_centroid"]), a(tb["facet_area"]),
                                 a(tb["facet_faces_ptr"], _i32), a(tb["facet_faces"], _i32), a(tb["bounds"])), "nk_set_mesh")
        S = tb["sv_centres"].shape[0]
        self.S = S
        interp = tb["temp_interp"]
        if interp not in INTERP_CODE:
            raise NkError(f"temp_interp '{interp}' is not one of nearest, linear, radial")
        if interp == "linear" and not bool(tb["sv_slice"]):
            interp = "radial"      # Population.py:574-576: linear is for slices only, upstream falls back to the RBF
        check(ctx, L.nk_set_subvols(ctx, S, a(tb["sv_centres"]), a(tb["sv_volume"]), int(bool(tb["sv_slice"])),
                                    int(tb["slice_axis"]), INTERP_CODE[interp]), "nk_set_subvols")
        if interp == "radial":
            from .routines.rbf import cubic_rbf_weights
            dims = np.asarray(tb.get("interp_dims", np.arange(3)), dtype=np.int32)
            shift, scale, W = cubic_rbf_weights(tb["sv_centres"], dims)
            check(ctx, L.nk_set_rbf(ctx, int(dims.shape[0]), a(dims, _i32), a(shift), a(scale), a(W)), "nk_set_rbf")
        NT = tb["T_grid"].shape[0]
        check(ctx, L.nk_set_phonon(ctx, Q, J, NT, a(tb["T_grid"]), a(tb["omega"]), a(tb["group_vel"]), a(tb["tau"]),
                                   float(tb["hbar"]), float(tb["kb"]), float(tb["volume_unitcell"]), int(tb["n_active"]),
                                   tb["energy_array"].shape[0], a(tb["energy_array"]), a(tb["T_array"])), "nk_set_phonon")
        R = tb["res_facet"].shape[0]
        self.R = R
        if hot_T is None:
            hot_T = (float(np.min(tb["res_T"])), float(np.max(tb["res_T"]))) if R > 0 else (295.0, 305.0)
        check(ctx, L.nk_set_population(ctx, float(tb["dt"]), int(bool(tb["norm_mean"])), float(tb["particle_density"]),
                                       int(tb["n_dt_to_conv"]), self.seed, float(tb["eVpsa2_in_Wm2"]), float(tb["a_in_m"]),
                                       float(hot_T[0]), float(hot_T[1])), "nk_set_population")
        if res_counter is None:
            res_counter = np.zeros((R, Q, J))
        check(ctx, L.nk_set_reservoirs(ctx, R, a(tb["res_facet"], _i32), a(tb["res_T"]), a(tb["enter_prob"]),
                                       a(res_counter)), "nk_set_reservoirs")
        res_gen = str(tb.get("res_gen", "constant"))
        if res_gen not in RESGEN_CODE:
            raise NkError(f"reservoir_gen '{res_gen}' is not one of constant, fixed_rate, one_to_one")
        if R > 0:
            n_leaving = np.sum(np.asarray(tb["enter_prob"], dtype=float), axis=(1, 2)).round()       # Population.py:344
            check(ctx, L.nk_set_reservoir_mode(ctx, RESGEN_CODE[res_gen], a(n_leaving)), "nk_set_reservoir_mode")
        Fr = tb["specularity"].shape[0]
        u8 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.uint8))
        check(ctx, L.nk_set_boundary_luts(ctx, Fr, a(tb["specularity"]), a(tb["true_specular"], u8),
                                          a(tb["spec_out"], _i32), a(tb["roulette"])), "nk_set_boundary_luts")

    # ------------------------------------------------------------------------------------------
    def allocate(self, capacity):
        cap = max(int(capacity), 2)
        cap = (cap + 511) // 512 * 512          # whole 512-slot tiles: enables the TMA bulk-copy step kernel
        dev = self.device
        f = lambda: torch.zeros(cap, dtype=torch.float64, device=dev)
        i = lambda fill: torch.full((cap,), fill, dtype=torch.int32, device=dev)
        self.t = dict(px=f(), py=f(), pz=f(), tc=f(), occ=f(), mode=i(-1), omode=i(-1), cfacet=i(-1),
                      cx=f(), cy=f(), cz=f(), pid=torch.zeros(cap, dtype=torch.int64, device=dev))
        t = self.t
        check(self.ctx, self.L.nk_bind_particles(self.ctx, cap, _dp(t["px"]), _dp(t["py"]), _dp(t["pz"]), _dp(t["tc"]),
                                                 _dp(t["occ"]), _dp(t["mode"]), _dp(t["omode"]), _dp(t["cfacet"]),
                                                 _dp(t["cx"]), _dp(t["cy"]), _dp(t["cz"]), _dp(t["pid"])), "nk_bind_particles")
        self.cap = cap
        self.t_back = None

    def load_particles(self, positions, modes_flat, occupation, ids=None, omodes=None,
                       n_timesteps=None, collision_facets=None, collision_positions=None):
        n = int(positions.shape[0])
        if n > self.cap:
            raise NkError("more particles than capacity")
        t, dev = self.t, self.device
        pos = torch.as_tensor(np.asarray(positions, dtype=np.float64), device=dev)
        t["px"][:n] = pos[:, 0]; t["py"][:n] = pos[:, 1]; t["pz"][:n] = pos[:, 2]
        t["occ"][:n] = torch.as_tensor(np.asarray(occupation, dtype=np.float64), device=dev)
        md = torch.as_tensor(np.asarray(modes_flat, dtype=np.int32), device=dev)
        t["mode"][:n] = md
        t["mode"][n:] = -1
        t["omode"][:n] = md if omodes is None else torch.as_tensor(np.asarray(omodes, dtype=np.int32), device=dev)
        t["pid"][:n] = torch.arange(n, device=dev) if ids is None else torch.as_tensor(np.asarray(ids, dtype=np.int64), device=dev)
        if n_timesteps is not None:
            t["tc"][:n] = torch.as_tensor(np.asarray(n_timesteps, dtype=np.float64), device=dev)
            t["cfacet"][:n] = torch.as_tensor(np.asarray(collision_facets, dtype=np.int32), device=dev)
            cp = torch.as_tensor(np.asarray(collision_positions, dtype=np.float64), device=dev)
            t["cx"][:n] = cp[:, 0]; t["cy"][:n] = cp[:, 1]; t["cz"][:n] = cp[:, 2]
        torch.cuda.synchronize(dev)
        check(self.ctx, self.L.nk_set_slot_count(self.ctx, n), "nk_set_slot_count")
        if n_timesteps is None:
            self.init_collisions()

    FIELDS = ("px", "py", "pz", "tc", "occ", "mode", "omode", "cfacet", "cx", "cy", "cz", "pid")

    def sort_by_mode(self, pools=None):
        """Maintenance pass (not part of a timestep): compact the slots and order the live particles by mode
        index with the library's counting sort (``nk_sort_by_mode``: one fused pass over all twelve fields into
        a second set of arrays, which then become the bound ones), so that the 64-byte mode-record gathers of
        neighbouring lanes hit the same cache line.  ``pools`` (default: on when there are >= NK_POOL_MIN = 64
        particles per occupied mode): every mode region keeps a few spare slots and its own free-slot ring, so
        emitted particles land among their own mode and the order survives emission / absorption; ``pools=False``
        gives the plain compaction (live particles in [0, n_live)).  Sums are order independent; particle identity
        is carried by ``pid``."""
        import os
        n, alive = self.slot_count()
        if n == 0:
            return
        if getattr(self, "t_back", None) is None or self.t_back["px"].numel() != self.cap:
            dev = self.device
            self.t_back = {k: torch.empty_like(v) for k, v in self.t.items()}
        frac, fixed = 0.0, 0
        if pools is None:
            pools = os.environ.get("NK_MODE_POOLS", "1") != "0" and alive >= int(os.environ.get("NK_POOL_MIN", 64)) * max(self.M, 1)
        if pools:
            frac = float(os.environ.get("NK_POOL_FRAC", 0.01)); fixed = int(os.environ.get("NK_POOL_FIXED", 2))
        b = self.t_back
        ns, na = C.c_int64(), C.c_int64()
        check(self.ctx, self.L.nk_sort_by_mode(self.ctx, *[_dp(b[k]) for k in self.FIELDS], frac, fixed, C.byref(ns), C.byref(na)),
              "nk_sort_by_mode")
        self.t, self.t_back = self.t_back, self.t

    def set_sv_temperature(self, T):
        T = _f64(T)
        check(self.ctx, self.L.nk_set_sv_temperature(self.ctx, _p(T)), "nk_set_sv_temperature")

    def set_timestep(self, k):
        check(self.ctx, self.L.nk_set_timestep(self.ctx, int(k)), "nk_set_timestep")

    def init_collisions(self):
        check(self.ctx, self.L.nk_init_collisions(self.ctx), "nk_init_collisions")

    def step(self, n=1):
        check(self.ctx, self.L.nk_step(self.ctx, int(n)), "nk_step")

    def step_local(self):
        check(self.ctx, self.L.nk_step_local(self.ctx), "nk_step_local")

    def step_finalize(self):
        check(self.ctx, self.L.nk_step_finalize(self.ctx), "nk_step_finalize")

    def flush_relaxation(self):
        check(self.ctx, self.L.nk_flush_relaxation(self.ctx), "nk_flush_relaxation")

    def profile_begin(self):
        check(self.ctx, self.L.nk_profile_begin(self.ctx), "nk_profile_begin")

    def profile_end(self):
        """-> (dict kernel -> total ms, n_steps) of the steps launched since profile_begin."""
        ms = (C.c_double * 4)()
        n = C.c_int64()
        check(self.ctx, self.L.nk_profile_end(self.ctx, ms, C.byref(n)), "nk_profile_end")
        return dict(k_step=ms[0], k_rare=ms[1], k_finalize=ms[2]), n.value

    def last_step_variant(self):
        """0 direct / 4 table streaming kernel, + 8 if the rare path used triangle tiles (nk_last_step_variant)."""
        return int(self.L.nk_last_step_variant(self.ctx))

    def synchronize(self):
        check(self.ctx, self.L.nk_synchronize(self.ctx), "nk_synchronize")

    def slot_count(self):
        a, b = C.c_int64(), C.c_int64()
        check(self.ctx, self.L.nk_get_slot_count(self.ctx, C.byref(a), C.byref(b)), "nk_get_slot_count")
        return a.value, b.value

    def timestep(self):
        a = C.c_int64()
        check(self.ctx, self.L.nk_get_timestep(self.ctx, C.byref(a)), "nk_get_timestep")
        return a.value

    def res_counter(self):
        out = np.zeros((self.R, self.M))
        if self.R:
            check(self.ctx, self.L.nk_get_res_counter(self.ctx, _p(out)), "nk_get_res_counter")
        return out

    def run_state(self):
        """Reservoir counters, emission deal counters, convergence-window accumulators, N_leaving and the results block
        (what a checkpoint needs besides the particles and T_sv)."""
        R, M = self.R, self.M
        n = self.L.nk_results_len(self.ctx)
        st = dict(res_counter=np.zeros((R, M)), res_fire=np.zeros((R, M), dtype=np.uint8), res_acc=np.zeros(4 * max(R, 1)),
                  n_leaving=np.zeros(max(R, 1)), results_block=np.zeros(n))
        if R:
            check(self.ctx, self.L.nk_get_run_state(self.ctx, _p(st["res_counter"]), _p(st["res_fire"]), _p(st["res_acc"]),
                                                    _p(st["n_leaving"]), _p(st["results_block"])), "nk_get_run_state")
        else:
            check(self.ctx, self.L.nk_get_run_state(self.ctx, None, None, None, None, _p(st["results_block"])), "nk_get_run_state")
        return st

    def set_run_state(self, res_counter=None, res_fire=None, res_acc=None, n_leaving=None, results_block=None):
        keep = []

        def a(x, dt):
            if x is None:
                return None
            y = np.ascontiguousarray(np.asarray(x, dtype=dt))
            keep.append(y)
            return _p(y)
        check(self.ctx, self.L.nk_set_run_state(self.ctx, a(res_counter, np.float64), a(res_fire, np.uint8), a(res_acc, np.float64),
                                                a(n_leaving, np.float64), a(results_block, np.float64)), "nk_set_run_state")

    def checkpoint(self):
        """Everything needed to continue this context's run bit-exactly (besides the static tables and the seed): live
        particles in f64, collision clocks, reservoir counters and deal counters, the accumulators of the open convergence
        window, N_leaving, the results block, T_sv and the step counter.  Flushes the deferred relaxation."""
        p = self.particles(flush=True)
        st = self.run_state()
        T = np.zeros(self.S)
        check(self.ctx, self.L.nk_get_sv_temperature(self.ctx, _p(T)), "nk_get_sv_temperature")
        out = dict(ids=p["ids"], positions=p["positions"], modes=p["modes"], omega_modes=p["omega_modes"], occupation=p["occupation"],
                   n_timesteps=p["n_timesteps"], collision_facets=p["collision_facets"], collision_positions=p["collision_positions"],
                   subvol_temperature=T, current_timestep=np.int64(self.timestep()), seed=np.int64(self.seed))
        out.update(st)
        return out

    def restore(self, z, capacity_factor=1.25):
        """Inverse of ``checkpoint`` on a context whose tables are already set (and, in a sharded run, whose rank is already
        set): nothing is rebuilt, so rank, mailboxes and accumulator buffers stay valid."""
        n = int(z["ids"].shape[0])
        if self.cap < n + 2:
            self.allocate(int(n * capacity_factor) + 1024)
        J = self.J
        modes = np.asarray(z["modes"])
        self.load_particles(z["positions"], modes[:, 0] * J + modes[:, 1], z["occupation"], ids=z["ids"], omodes=z["omega_modes"],
                            n_timesteps=z["n_timesteps"], collision_facets=z["collision_facets"], collision_positions=z["collision_positions"])
        self.set_sv_temperature(z["subvol_temperature"])
        self.set_timestep(int(z["current_timestep"]))
        if self.R:
            self.set_run_state(res_counter=z["res_counter"], res_fire=z["res_fire"], res_acc=z["res_acc"], n_leaving=z["n_leaving"],
                               results_block=z["results_block"])
        else:
            self.set_run_state(results_block=z["results_block"])

    def snapshot_results(self):
        """Enqueue a copy of the results block as it will be after the steps enqueued so far; returns a ticket for
        ``results(ticket)``.  The device is not stalled: more steps can be enqueued before the numbers are looked at."""
        t = C.c_int()
        check(self.ctx, self.L.nk_snapshot_results(self.ctx, C.byref(t)), "nk_snapshot_results")
        return t.value

    def results(self, ticket=None):
        S, R = self.S, self.R
        T = np.zeros(S); E = np.zeros(S); N = np.zeros(S, dtype=np.int64); flux = np.zeros((S, 3)); ksv = np.zeros(S)
        kappa = np.zeros(1); reb = np.zeros(max(R, 1)); rfl = np.zeros((max(R, 1), 3)); nl = np.zeros(max(R, 1), dtype=np.int64)
        etot = np.zeros(1)
        if ticket is None:
            check(self.ctx, self.L.nk_get_results(self.ctx, _p(T), _p(E), _p(N), _p(flux), _p(ksv), _p(kappa), _p(reb), _p(rfl),
                                                  _p(nl), _p(etot)), "nk_get_results")
        else:
            check(self.ctx, self.L.nk_get_snapshot(self.ctx, int(ticket), _p(T), _p(E), _p(N), _p(flux), _p(ksv), _p(kappa), _p(reb),
                                                   _p(rfl), _p(nl), _p(etot)), "nk_get_snapshot")
        return dict(subvol_temperature=T, subvol_energy=E, subvol_N_p=N, subvol_heat_flux=flux, subvol_kappa=ksv,
                    kappa=float(kappa[0]), res_energy_balance=reb[:R], res_heat_flux=rfl[:R], N_leaving=nl[:R],
                    total_energy=float(etot[0]), N_p=int(N.sum()))

    def particles(self, flush=True):
        """Live particles as NumPy arrays in the reference's attribute names (sorted by id)."""
        if flush:
            self.flush_relaxation()
        n, _ = self.slot_count()
        t = self.t
        mode = t["mode"][:n]
        live = (mode >= 0).nonzero().squeeze(1)
        ids = t["pid"][live]
        order = torch.argsort(ids)
        live = live[order]
        g = lambda name: t[name][live].cpu().numpy()
        J = self.J
        md = g("mode").astype(np.int64)
        return dict(ids=g("pid"), positions=np.stack([g("px"), g("py"), g("pz")], axis=1),
                    modes=np.stack([md // J, md % J], axis=1), omega_modes=g("omode").astype(np.int64),
                    occupation=g("occ"), n_timesteps=g("tc"), collision_facets=g("cfacet").astype(np.int64),
                    collision_positions=np.stack([g("cx"), g("cy"), g("cz")], axis=1))

    # ---- operator seams (reference method signatures, NumPy in / NumPy out) ---------------------
    def _dev(self, a, dtype=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(a), device=self.device).to(dtype).contiguous()

    def find_boundary(self, x, v):
        """Mesh.find_boundary(x, v) -> (xc, tc, fc)  (Mesh.py:806-856)."""
        x = self._dev(np.asarray(x, dtype=float).reshape(-1, 3)); v = self._dev(np.asarray(v, dtype=float).reshape(-1, 3))
        n = x.shape[0]
        xc = torch.empty_like(x); tc = torch.empty(n, dtype=torch.float64, device=self.device)
        fc = torch.empty(n, dtype=torch.int32, device=self.device)
        check(self.ctx, self.L.nk_find_boundary(self.ctx, n, _dp(x), _dp(v), _dp(xc), _dp(tc), _dp(fc)), "nk_find_boundary")
        self.synchronize()
        return xc.cpu().numpy(), tc.cpu().numpy(), fc.cpu().numpy().astype(int)

    def contains(self, x):
        """Mesh.contains_naive(x)  (Mesh.py:785-804) -> bool array."""
        x = self._dev(np.asarray(x, dtype=float).reshape(-1, 3))
        out = torch.empty(x.shape[0], dtype=torch.uint8, device=self.device)
        check(self.ctx, self.L.nk_contains(self.ctx, x.shape[0], _dp(x), _dp(out)), "nk_contains")
        self.synchronize()
        return out.cpu().numpy().astype(bool)

    def classify(self, x, counts=False):
        """SubvolClassifier.predict(x)  (Geometry.py:1212)."""
        x = self._dev(np.asarray(x, dtype=float).reshape(-1, 3))
        n = x.shape[0]
        sv = torch.empty(n, dtype=torch.int32, device=self.device)
        cn = torch.zeros(self.S, dtype=torch.int64, device=self.device)
        check(self.ctx, self.L.nk_classify(self.ctx, n, _dp(x), _dp(sv), _dp(cn)), "nk_classify")
        self.synchronize()
        return (sv.cpu().numpy().astype(int), cn.cpu().numpy()) if counts else sv.cpu().numpy().astype(int)

    def calculate_occupation(self, T, omega):
        """Phonon.calculate_occupation  (Phonon.py:338-345)."""
        T, omega = np.broadcast_arrays(np.asarray(T, dtype=float), np.asarray(omega, dtype=float))
        shp = T.shape
        Td = self._dev(T.reshape(-1)); od = self._dev(omega.reshape(-1)); out = torch.empty_like(Td)
        check(self.ctx, self.L.nk_occupation(self.ctx, Td.numel(), _dp(Td), _dp(od), _dp(out)), "nk_occupation")
        self.synchronize()
        return out.cpu().numpy().reshape(shp)

    def lifetime_function(self, Tqj):
        """Phonon.lifetime_function([[T, q, j], ...])  (Phonon.py:326-336)."""
        Tqj = np.asarray(Tqj, dtype=float).reshape(-1, 3)
        Td = self._dev(Tqj[:, 0]); md = self._dev((Tqj[:, 1] * self.J + Tqj[:, 2]).astype(np.int32), torch.int32)
        out = torch.empty_like(Td)
        check(self.ctx, self.L.nk_lifetime(self.ctx, Td.numel(), _dp(Td), _dp(md), _dp(out)), "nk_lifetime")
        self.synchronize()
        return out.cpu().numpy()

    def temperature_function(self, E):
        E = np.asarray(E, dtype=float); Ed = self._dev(E.reshape(-1)); out = torch.empty_like(Ed)
        check(self.ctx, self.L.nk_temperature_of_energy(self.ctx, Ed.numel(), _dp(Ed), _dp(out)), "nk_temperature_of_energy")
        self.synchronize()
        return out.cpu().numpy().reshape(E.shape)

    def crystal_energy_function(self, T):
        T = np.asarray(T, dtype=float); Td = self._dev(T.reshape(-1)); out = torch.empty_like(Td)
        check(self.ctx, self.L.nk_energy_of_temperature(self.ctx, Td.numel(), _dp(Td), _dp(out)), "nk_energy_of_temperature")
        self.synchronize()
        return out.cpu().numpy().reshape(T.shape)

    def particle_temperature(self, x):
        x = self._dev(np.asarray(x, dtype=float).reshape(-1, 3))
        out = torch.empty(x.shape[0], dtype=torch.float64, device=self.device)
        check(self.ctx, self.L.nk_particle_temperature(self.ctx, x.shape[0], _dp(x), _dp(out)), "nk_particle_temperature")
        self.synchronize()
        return out.cpu().numpy()
