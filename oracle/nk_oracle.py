"""TEST INFRASTRUCTURE -- NumPy restatement of Nano-kappa's per-timestep particle loop.

CPU checker for the CUDA path; never imported by ``nanokappa_b200``.  Every function cites the
reference lines (``/root/reference/...``) it restates.  It works on plain arrays:

* ``tb``  -- dict of static tables (mesh planes, SV centres, mode tables, LUTs; see ``TABLE_KEYS``)
* ``st``  -- ``State``: the reference's per-particle SoA (``Population.py:838-850``) + a stable id

Parity status: PINNED.  ``tests/test_oracle_pin.py`` runs this file and the unmodified reference
(``oracle/ref_harness.py``) on the same seeded NumPy stream for several configurations and demands
bit-identical particle arrays and per-SV vectors after every step; ``tests/golden/*.npz`` holds the
same comparison as committed fixtures for boxes without ``/root/reference``.

Two random sources:
* ``SequenceRNG``  consumes ``np.random`` in exactly the reference's call order -> bit parity with
  the reference itself.
* ``KeyedRNG``     Philox keyed by particle id / step / event (``oracle/philox.py``) -> bit parity
  target for the GPU path, which cannot reproduce row-order dependent streams.

Third-party arithmetic restated here (SciPy 1.18.1 / NumPy 2.3.5 as installed): ``interp1d`` linear
with fill values = ``np.interp``; ``interp1d(fill_value='extrapolate')`` linear/nearest;
``RegularGridInterpolator`` linear (only the two T nodes carry weight); ``NearestNDInterpolator`` =
nearest centre by squared distance; ``np.random.choice(p=...)`` = right-searchsorted on the
normalised cumulative sum.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass

import numpy as np

from . import philox

BC_T, BC_P, BC_R, BC_F = 0, 1, 2, 3
BC_CODE = {"T": BC_T, "P": BC_P, "R": BC_R, "F": BC_F}
TOL = 1e-10   # Mesh.py:24

TABLE_KEYS = (
    # mesh (Mesh.py:205-243, :314-324)
    "face_normals", "face_k", "face_lo", "face_hi", "face_origins", "face_basis", "face_facets",
    "face_vertices", "face_areas",
    # facets / boundary conditions (Geometry.py:652-726)
    "facet_bc", "facet_normal", "facet_centroid", "facet_area", "facet_partner", "facet_res", "facet_rough",
    "facet_faces_ptr", "facet_faces", "bounds",
    # subvolumes (Geometry.py:446-544, :1198-1213)
    "sv_centres", "sv_volume", "sv_slice", "slice_axis", "temp_interp",
    # modes (Phonon.py:66-151, :326-401)
    "omega", "group_vel", "tau", "T_grid", "energy_array", "T_array", "hbar", "kb", "volume_unitcell",
    "n_active", "eVpsa2_in_Wm2", "a_in_m",
    # population constants (Population.py:35-125, :146-161, :852-939, :1456-1459)
    "dt", "norm_mean", "particle_density", "res_facet", "res_T", "enter_prob",
    "specularity", "true_specular", "spec_out", "roulette", "n_dt_to_conv",
)


@dataclass
class State:
    positions: np.ndarray
    modes: np.ndarray            # (N,2) int: (q, j) used for v_g and tau
    omega: np.ndarray            # (N,)  carried separately: NOT refreshed on specular hits (:955-971)
    group_vel: np.ndarray
    occupation: np.ndarray
    n_timesteps: np.ndarray
    collision_facets: np.ndarray  # int, -1 = ray escaped (Mesh.py:851)
    collision_positions: np.ndarray
    collision_cond: np.ndarray    # int8 BC code
    temperatures: np.ndarray
    ids: np.ndarray               # int64 stable particle id (test-side bookkeeping)
    subvol_temperature: np.ndarray
    res_counter: np.ndarray       # (R,Q,J)
    subvol_id: np.ndarray = None
    energies: np.ndarray = None
    subvol_energy: np.ndarray = None
    subvol_N_p: np.ndarray = None
    subvol_heat_flux: np.ndarray = None
    subvol_kappa: np.ndarray = None
    kappa: float = 0.0
    N_p: int = 0
    N_leaving: np.ndarray = None
    res_energy_balance: np.ndarray = None
    res_heat_flux: np.ndarray = None
    current_timestep: int = 0
    omega_modes: np.ndarray = None  # (N,) flat mode whose omega the particle carries (bookkeeping)

    def copy(self):
        return copy.deepcopy(self)


# ------------------------------------------------------------------------------------------------
# random sources
# ------------------------------------------------------------------------------------------------
class SequenceRNG:
    """np.random in the reference's own call order."""
    keyed = False

    def emit_dt_uniform(self, n, ids, step):
        return np.random.rand(n)                                   # Population.py:393

    def emit_dice(self, shape, ids, step):
        return np.random.rand(*shape)                              # Population.py:410 (fixed_rate)

    def one_to_one_mode(self, n, ids, step):
        return np.random.rand(n)                                   # Population.py:471

    def one_to_one_dt(self, n, ids, step):
        return np.random.rand(n)                                   # Population.py:483

    def surface(self, n, faces, p, ids, step):
        f = np.random.choice(faces, size=n, p=p)                   # Mesh.py:937
        s = np.random.rand(n, 1)                                    # Mesh.py:941
        r = np.random.rand(n, 1)                                    # Mesh.py:942
        return f, s[:, 0], r[:, 0]

    def rough_dice(self, n, ids, step, event):
        return np.random.rand(n)                                   # Population.py:949

    def diffuse_pick(self, n, ids, step, event):
        return np.random.rand(n)                                   # Population.py:1003


class KeyedRNG:
    """Philox keyed per particle (see oracle/philox.py)."""
    keyed = True

    def __init__(self, seed=0):
        self.seed = int(seed)

    def emit_dt_uniform(self, n, ids, step):
        return philox.uniforms(ids, step, philox.STREAM_EMIT_A, self.seed)[0]

    def emit_dice(self, shape, ids, step):
        return philox.uniforms(ids, step, philox.STREAM_EMIT_C, self.seed)[0].reshape(shape)

    def one_to_one_mode(self, n, ids, step):
        return philox.uniforms(ids, step, philox.STREAM_EMIT_C, self.seed)[0]

    def one_to_one_dt(self, n, ids, step):
        return philox.uniforms(ids, step, philox.STREAM_EMIT_C, self.seed)[1]

    def surface(self, n, faces, p, ids, step):
        u_face = philox.uniforms(ids, step, philox.STREAM_EMIT_A, self.seed)[1]
        s, r = philox.uniforms(ids, step, philox.STREAM_EMIT_B, self.seed)
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        f = np.asarray(faces)[np.searchsorted(cdf, u_face, side="right")]   # np.random.choice semantics
        return f, s, r

    def rough_dice(self, n, ids, step, event):
        return philox.uniforms(ids, step, philox.STREAM_ROUGH0 + np.asarray(event), self.seed)[0]

    def diffuse_pick(self, n, ids, step, event):
        return philox.uniforms(ids, step, philox.STREAM_ROUGH0 + np.asarray(event), self.seed)[1]


# ------------------------------------------------------------------------------------------------
# mode-table functions (Phonon.py)
# ------------------------------------------------------------------------------------------------
def calculate_occupation(tb, T, omega):
    """Bose-Einstein; 0 where T<=0 or omega<=0.  Phonon.py:338-345."""
    T = np.asarray(T, dtype=float)
    omega = np.asarray(omega, dtype=float)
    flag = (T > 0) & (omega > 0)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        occ = np.where(~flag, 0, 1 / (np.exp(omega * tb["hbar"] / (T * tb["kb"])) - 1))
    return occ


def normalise_to_density(tb, x):
    """Phonon.py:392-401."""
    return x / (tb["omega"].shape[0] * tb["volume_unitcell"])


def temperature_function(tb, E):
    """T(E): interp1d(energy_array, T_array, linear, fill=(T_min,T_max)) = np.interp with clamping.
    Phonon.py:387; scipy delegates to np.interp for 1-D float data without extrapolation."""
    E = np.asarray(E, dtype=float)
    xa, ya = tb["energy_array"], tb["T_array"]
    y = np.interp(E, xa, ya)
    y = np.where(E < xa[0], ya[0], y)       # fill_value below = T_min (== ya[0], the table is ascending)
    y = np.where(E > xa[-1], ya[-1], y)
    return y


def crystal_energy_function(tb, T):
    """E(T): Phonon.py:390, fill = (E.min(), E.max())."""
    T = np.asarray(T, dtype=float)
    xa, ya = tb["T_array"], tb["energy_array"]
    y = np.interp(T, xa, ya)
    y = np.where(T < xa[0], ya.min(), y)
    y = np.where(T > xa[-1], ya.max(), y)
    return y


def lifetime_function(tb, T, modes):
    """tau(T,q,j): RegularGridInterpolator((T,q,j), tau) at integer q,j -> linear in T only.
    Phonon.py:326-336; scipy _rgi.py:520-549 (value * weight, lower node first)."""
    Tg = tb["T_grid"]
    T = np.asarray(T, dtype=float)
    if np.any(T < Tg[0]) or np.any(T > Tg[-1]):
        raise ValueError("One of the requested xi is out of bounds in dimension 0")   # same as scipy
    i = np.searchsorted(Tg, T, side="right") - 1
    i = np.clip(i, 0, Tg.shape[0] - 2)
    w = (T - Tg[i]) / (Tg[i + 1] - Tg[i])
    q, j = modes[:, 0], modes[:, 1]
    return tb["tau"][i, q, j] * (1 - w) + tb["tau"][i + 1, q, j] * w


# ------------------------------------------------------------------------------------------------
# geometry seams
# ------------------------------------------------------------------------------------------------
def find_boundary(tb, x, v):
    """Nearest forward ray/triangle hit.  Mesh.py:806-856.  Returns (xc, tc, fc); fc = -1, tc = inf
    when nothing is hit."""
    x = np.asarray(x, dtype=float).reshape(-1, 3)
    v = np.asarray(v, dtype=float).reshape(-1, 3)
    n, k = tb["face_normals"], tb["face_k"]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t = -(np.sum(np.expand_dims(x, 1) * n, axis=2) + k) / np.sum(np.expand_dims(v, 1) * n, axis=2)   # :818
    possible = t >= TOL
    possible &= ~np.isnan(t)
    possible &= ~np.isinf(np.absolute(t))
    in_p, in_f = possible.nonzero()
    c = x[in_p, :] + np.expand_dims(t[in_p, in_f], 1) * v[in_p, :]                                      # :826
    in_bounds = np.logical_and(np.all(c >= tb["face_lo"][in_f, :] - TOL, axis=1),
                               np.all(c <= tb["face_hi"][in_f, :] + TOL, axis=1))                       # :828
    possible[in_p, in_f] = in_bounds
    c = c[in_bounds, :]
    in_p, in_f = possible.nonzero()
    if in_p.shape[0] > 0:
        bar = np.linalg.solve(tb["face_basis"][in_f], (c - tb["face_origins"][in_f])[..., None])[..., 0][:, :2]   # :840
        bar = np.concatenate((bar, 1 - bar.sum(axis=1, keepdims=True)), axis=1)
        possible[in_p, in_f] = np.all(np.logical_and(bar >= 0 - TOL, bar <= 1 + TOL), axis=-1)          # :843
    t = np.where(possible, t, np.inf)
    tc = np.min(t, axis=1)
    fc = tb["face_facets"][np.argmax(t == np.expand_dims(tc, 1), axis=-1)].astype(int)                 # :849
    fc[tc == np.inf] = -1
    with np.errstate(invalid="ignore"):
        xc = x + np.expand_dims(tc, 1) * v                                                              # :854
    return xc, tc, fc


def timesteps_to_boundary(tb, x, v):
    """Population.py:797-830 (the 1e6 stride only bounds memory; results are identical)."""
    if x.shape[0] == 0:
        return np.zeros(0), np.zeros(0, dtype=int), np.zeros((0, 3))
    xc, tc, fc = find_boundary(tb, x, v)
    return tc / tb["dt"], fc, xc


def collision_condition(tb, facets):
    """BC of a facet id; -1 indexes the last facet like NumPy does.  Population.py:657-669, :1487."""
    return tb["facet_bc"][np.asarray(facets, dtype=int)]


def classify(tb, x, kdtree=None):
    """Nearest SV centre.  Geometry.py:1198-1213 (NearestNDInterpolator = cKDTree, p=2).
    Exact ties are tree-order dependent upstream and excluded from parity."""
    x = np.asarray(x, dtype=float).reshape(-1, 3)
    if kdtree is not None:
        return kdtree.query(x)[1].astype(int)
    c = tb["sv_centres"]
    d = x[:, None, :] - c[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    return np.argmin(d2, axis=1).astype(int)


def particle_temperature(tb, T_sv, x):
    """Per-particle T from the SV temperatures.  Population.py:570-590, :694-702.
    slice+linear : interp1d(..., 'linear', fill_value='extrapolate')  (scipy _interpolate.py:491-518)
    slice+nearest: interp1d(..., 'nearest', extrapolate)              (scipy _interpolate.py:361-363, :520-535)
    otherwise    : NearestNDInterpolator(centres, T_sv) == T_sv[classify(x)]"""
    x = np.asarray(x, dtype=float).reshape(-1, 3)
    kind = tb["temp_interp"]
    if tb["sv_slice"] and kind in ("linear", "nearest"):
        ax = int(tb["slice_axis"])
        xs = tb["sv_centres"][:, ax]
        xn = x[:, ax]
        if kind == "linear":
            idx = np.searchsorted(xs, xn).clip(1, len(xs) - 1).astype(int)
            lo, hi = idx - 1, idx
            x_lo, x_hi, y_lo, y_hi = xs[lo], xs[hi], T_sv[lo], T_sv[hi]
            return ((xn - x_lo) / (x_hi - x_lo)) * y_hi + ((x_hi - xn) / (x_hi - x_lo)) * y_lo
        xb = xs / 2.0
        xb = xb[1:] + xb[:-1]
        idx = np.searchsorted(xb, xn, side="left").clip(0, len(xs) - 1)
        return T_sv[idx]
    dims = np.asarray(tb.get("interp_dims", np.arange(3)), dtype=int)
    if kind == "nearest":
        if dims.shape[0] == 3:
            return T_sv[classify(tb, x)]
        from scipy.interpolate import NearestNDInterpolator
        return NearestNDInterpolator(tb["sv_centres"][:, dims], T_sv)(x[:, dims])
    # 'radial', and 'linear' on non-slice subvolumes (Population.py:574-588): the third-party call itself,
    # scipy.interpolate.RBFInterpolator(kernel='cubic') -> degree-1 polynomial tail, epsilon 1, no smoothing
    from scipy.interpolate import RBFInterpolator
    if x.shape[0] == 0:
        return np.zeros(0)
    return RBFInterpolator(tb["sv_centres"][:, dims], T_sv, kernel="cubic")(x[:, dims])


def rbf_weights(centres, dims):
    """Restatement of scipy's cubic RBF system (scipy/interpolate/_rbfinterp_np.py:40-89 and the pythran
    `_build_system`): lhs = [[|c_i - c_j|^3, P], [P^T, 0]] with P = [1, (c - shift)/scale]; returns
    (shift, scale, W) with W = lhs^-1[:, :S], so that coeffs = W @ T_sv and
    T(x) = sum_s coeffs[s] |x - c_s|^3 + coeffs[S] + sum_k coeffs[S+1+k] (x_k - shift_k)/scale_k.
    This is what the host set-up of the CUDA path uploads (nk_set_rbf); tests compare it with RBFInterpolator."""
    y = np.asarray(centres, dtype=float)[:, np.asarray(dims, dtype=int)]
    S, nd = y.shape
    mins, maxs = y.min(axis=0), y.max(axis=0)
    shift = (maxs + mins) / 2
    scale = (maxs - mins) / 2
    scale[scale == 0.0] = 1.0
    yhat = (y - shift) / scale
    lhs = np.zeros((S + nd + 1, S + nd + 1))
    r = np.linalg.norm(y[:, None, :] - y[None, :, :], axis=-1)
    lhs[:S, :S] = r ** 3
    lhs[:S, S] = 1.0
    lhs[:S, S + 1:] = yhat
    lhs[S:, :S] = lhs[:S, S:].T
    rhs = np.zeros((S + nd + 1, S))
    rhs[:S, :] = np.eye(S)
    return shift, scale, np.linalg.solve(lhs, rhs)


def sample_surface_points(tb, f, s, r):
    """Barycentric point on face f.  Mesh.py:939-949."""
    v = tb["face_vertices"][f, :, :]
    s = s.reshape(-1, 1)
    r = r.reshape(-1, 1)
    a = np.zeros((f.shape[0], 3, 1))
    a[:, 0, :] = 1 - (s ** 0.5)
    a[:, 1, :] = (1 - r) * (s ** 0.5)
    a[:, 2, :] = r * (s ** 0.5)
    return np.sum(a * v, axis=1)


def facet_faces(tb, facet):
    p = tb["facet_faces_ptr"]
    return tb["facet_faces"][p[facet]:p[facet + 1]]


# ------------------------------------------------------------------------------------------------
# the timestep
# ------------------------------------------------------------------------------------------------
def drift(tb, st):
    """Population.py:790-795."""
    st.positions += st.group_vel * tb["dt"]
    st.n_timesteps -= 1


def _fill_one_to_one(tb, st, rng):
    """--reservoir_gen one_to_one (Population.py:457-489): every reservoir re-emits as many particles as it absorbed in
    the previous step, modes drawn from the entry-probability roulette, entry time uniform in the step."""
    dt = tb["dt"]
    enter_prob = tb["enter_prob"]
    R, Q, J = enter_prob.shape
    step = st.current_timestep
    pos = np.zeros((0, 3)); modes = np.zeros((0, 2), dtype=int); facet_id = np.zeros(0, dtype=int)
    dt_in = np.zeros(0); ids = np.zeros(0, dtype=np.int64)
    for i in range(R):
        facet = tb["res_facet"][i]
        n = int(st.N_leaving[i])
        if n > 0:
            roulette = np.cumsum(enter_prob[i, :, :])
            roulette /= roulette.max()
            r_ids = philox.one_to_one_id(step, R, i, np.arange(n), Q * J)
            u = rng.one_to_one_mode(n, r_ids, step)
            flat_i = bisect_left_fixed(roulette, u) if rng.keyed else np.searchsorted(roulette, u)
            new_q = np.floor(flat_i / J).astype(int)
            new_j = flat_i - new_q * J
            modes = np.vstack((modes, np.vstack((new_q, new_j)).T))
            dt_in = np.concatenate((dt_in, dt * rng.one_to_one_dt(n, r_ids, step)))
            facet_id = np.concatenate((facet_id, (np.ones(n) * facet).astype(int)))
            faces = facet_faces(tb, facet)
            areas = tb["face_areas"][faces]
            f, s, r = rng.surface(n, faces, areas / areas.sum(), r_ids, step)
            pos = np.vstack((pos, sample_surface_points(tb, f, s, r)))
            ids = np.concatenate((ids, r_ids))
    return pos, modes, facet_id, dt_in, ids


def fill_reservoirs(tb, st, rng):
    """Emission.  Population.py:356-406 (constant), :408-455 (fixed_rate), :457-489 (one_to_one), :491-508.
    Returns the new-particle arrays."""
    dt = tb["dt"]
    enter_prob = tb["enter_prob"]
    R, Q, J = enter_prob.shape
    step = st.current_timestep
    res_gen = str(tb.get("res_gen", "constant"))
    fixed_np = np.floor(enter_prob).astype(int)
    if res_gen == "constant":
        st.res_counter += enter_prob - fixed_np
        in_mask = (st.res_counter >= 1).astype(int)
        st.res_counter -= in_mask
        lead = st.res_counter                      # numerator of the c == 1 entry time
    elif res_gen == "fixed_rate":
        rr, mm = np.meshgrid(np.arange(R), np.arange(Q * J), indexing="ij")
        dice = rng.emit_dice((R, Q, J), philox.emission_id(step, R, rr.ravel(), mm.ravel(), Q * J, 1), step)
        in_mask = (dice <= (enter_prob - fixed_np)).astype(int)
        lead = dice
    elif res_gen == "one_to_one":
        pos, modes, facet_id, dt_in, ids = _fill_one_to_one(tb, st, rng)
        return _new_particle_arrays(tb, pos, modes, facet_id, dt_in, ids)
    else:
        raise ValueError("unknown reservoir generation mode " + res_gen)
    in_np = fixed_np + in_mask
    N_p_facet = in_np.sum(axis=(1, 2))

    pos = np.zeros((0, 3)); modes = np.zeros((0, 2), dtype=int); facet_id = np.zeros(0, dtype=int)
    dt_in = np.zeros(0); ids = np.zeros(0, dtype=np.int64)
    for i in range(R):
        n = N_p_facet[i]
        if n > 0:
            facet = tb["res_facet"][i]
            c = in_np[i].max()
            r_ids = np.zeros(0, dtype=np.int64)
            while c > 0:
                c_modes = np.vstack(np.where(in_np[i] >= c)).T
                c_ids = philox.emission_id(step, R, i, c_modes[:, 0] * J + c_modes[:, 1], Q * J, c)
                p = enter_prob[i, c_modes[:, 0], c_modes[:, 1]]
                if c == 1:
                    c_dt_in = dt * (1 - (lead[i, c_modes[:, 0], c_modes[:, 1]] / p))
                else:
                    u = rng.emit_dt_uniform(c_modes.shape[0], c_ids, step)
                    c_dt_in = dt * (1 - (c - 1 + u) / p)
                c -= 1
                dt_in = np.concatenate((dt_in, c_dt_in))
                modes = np.vstack((modes, c_modes.astype(int)))
                r_ids = np.concatenate((r_ids, c_ids))
            facet_id = np.concatenate((facet_id, (np.ones(n) * facet).astype(int)))
            faces = facet_faces(tb, facet)
            areas = tb["face_areas"][faces]
            f, s, r = rng.surface(n, faces, areas / areas.sum(), r_ids, step)
            pos = np.vstack((pos, sample_surface_points(tb, f, s, r)))
            ids = np.concatenate((ids, r_ids))
    return _new_particle_arrays(tb, pos, modes, facet_id, dt_in, ids)


def _new_particle_arrays(tb, pos, modes, facet_id, dt_in, ids):
    """Population.py:491-508."""
    new = dict(positions=pos, modes=modes, facet_id=facet_id, dt_in=dt_in, ids=ids)
    if modes.shape[0] > 0:
        new["group_vel"] = tb["group_vel"][modes[:, 0], modes[:, 1], :]
        new["omega"] = tb["omega"][modes[:, 0], modes[:, 1]]
    idx = np.where(facet_id.reshape(-1, 1) == tb["res_facet"])[1]
    new["temperatures"] = tb["res_T"][idx]
    if modes.shape[0] > 0:
        new["occupation"] = calculate_occupation(tb, new["temperatures"], new["omega"])
    return new


def add_reservoir_particles(tb, st, new):
    """Population.py:525-552."""
    if new["modes"].shape[0] == 0:
        return
    ts, fc, xc = timesteps_to_boundary(tb, new["positions"], new["group_vel"])
    ts = ts - new["dt_in"] / tb["dt"]
    pos = new["positions"] + new["group_vel"] * new["dt_in"].reshape(-1, 1)
    cond = collision_condition(tb, fc)
    st.positions = np.vstack((st.positions, pos))
    st.modes = np.vstack((st.modes, new["modes"]))
    st.group_vel = np.vstack((st.group_vel, new["group_vel"]))
    st.n_timesteps = np.concatenate((st.n_timesteps, ts))
    st.collision_facets = np.concatenate((st.collision_facets, fc))
    st.collision_positions = np.concatenate((st.collision_positions, xc))
    st.collision_cond = np.concatenate((st.collision_cond, cond))
    st.temperatures = np.concatenate((st.temperatures, new["temperatures"]))
    st.omega = np.concatenate((st.omega, new["omega"]))
    st.occupation = np.concatenate((st.occupation, new["occupation"]))
    st.ids = np.concatenate((st.ids, new["ids"]))
    J = tb["omega"].shape[1]
    st.omega_modes = np.concatenate((st.omega_modes, new["modes"][:, 0] * J + new["modes"][:, 1]))


def _delete(st, mask):
    """Population.py:832-850."""
    keep = ~mask
    for name in ("positions", "group_vel", "omega", "occupation", "temperatures", "n_timesteps", "modes",
                 "collision_facets", "collision_positions", "collision_cond", "ids", "omega_modes"):
        setattr(st, name, getattr(st, name)[keep])


def bisect_left_fixed(arr, keys):
    """Per-key left bisection with fixed initial bounds [0, len) -- NumPy's loop without the
    previous-key shortcut (numpy/_core/src/npysort/binsearch.cpp)."""
    keys = np.asarray(keys, dtype=float)
    lo = np.zeros(keys.shape[0], dtype=np.int64)
    hi = np.full(keys.shape[0], arr.shape[0], dtype=np.int64)
    while np.any(lo < hi):
        act = lo < hi
        mid = (lo + hi) >> 1
        less = np.zeros_like(act)
        less[act] = arr[mid[act]] < keys[act]
        lo = np.where(act & less, mid + 1, lo)
        hi = np.where(act & ~less, mid, hi)
    return lo


def select_reflected_modes(tb, st_T_sv, in_modes, col_fac, col_pos, n_in, omega_in, ids, rng, step, event):
    """Specular / diffuse choice on rough facets.  Population.py:941-1015 ('velocity' model)."""
    J = tb["omega"].shape[1]
    i_rough = tb["facet_rough"][col_fac]
    true_spec = tb["true_specular"][i_rough, in_modes[:, 0], in_modes[:, 1]]
    p = tb["specularity"][i_rough, in_modes[:, 0], in_modes[:, 1]]
    n_p = in_modes.shape[0]
    r = rng.rough_dice(n_p, ids, step, event)
    spec = np.logical_and(true_spec, r <= p)
    diff = ~spec
    out_modes = np.zeros(in_modes.shape, dtype=int)
    n_out = copy.copy(n_in)
    omega_out = copy.copy(omega_in)
    if np.any(spec):
        flat = tb["spec_out"][i_rough[spec], in_modes[spec, 0], in_modes[spec, 1]]
        out_modes[spec, 0] = flat // J
        out_modes[spec, 1] = flat % J
    if np.any(diff):
        # pick_diffuse_modes, Population.py:990-1015: one rand() call per distinct facet, ascending
        d_fac = col_fac[diff]
        d_ids = ids[diff]
        d_ev = np.asarray(event)[diff] if np.ndim(event) else event
        new_modes = np.zeros((d_fac.shape[0], 2))
        for facet in np.unique(d_fac):
            i_p = np.arange(d_fac.shape[0])[d_fac == facet]
            i_f = tb["facet_rough"][facet]
            ev = d_ev[i_p] if np.ndim(d_ev) else d_ev
            u = rng.diffuse_pick(i_p.shape[0], d_ids[i_p], step, ev) * tb["roulette"][i_f, -1]
            if rng.keyed:
                # creation rates can be negative (more specular inflow than outflow), so the cumulative
                # table is not always monotonic and np.searchsorted's answer then depends on the previous
                # key of the batch (it narrows its bounds from it).  The keyed contract is the plain
                # per-key bisection the GPU does; on monotonic tables both are identical.
                flat_i = bisect_left_fixed(tb["roulette"][i_f, :], u)
            else:
                flat_i = np.searchsorted(tb["roulette"][i_f, :], u)
            new_q = np.floor(flat_i / J).astype(int)
            new_modes[i_p, 0] = new_q
            new_modes[i_p, 1] = flat_i - new_q * J
        out_modes[diff, :] = new_modes.astype(int)
        omega_out[diff] = tb["omega"][out_modes[diff, 0], out_modes[diff, 1]]
        T_diff = particle_temperature(tb, st_T_sv, col_pos[diff, :])
        n_out[diff] = calculate_occupation(tb, T_diff, omega_out[diff])
    return out_modes, n_out, omega_out, spec


def boundary_scattering(tb, st, rng, classifier=None):
    """The multi-pass boundary loop.  Population.py:1546-1683, :1463-1544."""
    dt = tb["dt"]
    R = tb["res_facet"].shape[0]
    J = tb["omega"].shape[1]
    step = st.current_timestep
    st.subvol_id = classify(tb, st.positions, classifier)
    idx_all = st.n_timesteps < 0
    calc = np.ones(idx_all.shape)
    calc[idx_all] = 0
    new_ts = copy.copy(st.n_timesteps)
    st.N_leaving = np.zeros(R, dtype=int)
    event = np.zeros(idx_all.shape, dtype=np.int64)   # rough events so far this step (keys the RNG)

    while np.any(calc < 1):
        # I. absorption into reservoirs (:1565-1608)
        cond_res = np.isin(st.collision_cond, (BC_T, BC_F))
        idx_del = np.logical_and(calc < 1, cond_res)
        idx_del = np.logical_and(idx_del, (1 - calc) > new_ts)
        if np.any(idx_del):
            for i in range(R):
                facet = tb["res_facet"][i]
                idx_res = st.collision_facets[idx_del] == facet
                st.N_leaving[i] += int(idx_res.sum())
                om = st.omega[idx_del][idx_res]
                dn = st.occupation[idx_del][idx_res] - calculate_occupation(tb, tb["res_T"][i], om)
                energies = tb["hbar"] * om * dn
                st.res_energy_balance[i] -= energies.sum()
                gv = st.group_vel[idx_del, :][idx_res, :]
                hflux = energies.reshape(-1, 1) * gv / np.sum(gv * tb["facet_normal"][facet, :], axis=1, keepdims=True)
                st.res_heat_flux[i, :] += hflux.sum(axis=0)
            _delete(st, idx_del)
            keep = ~idx_del
            calc, new_ts, event = calc[keep], new_ts[keep], event[keep]

        # II. periodic facets (:1610-1634, :1463-1489)
        idx_per = np.logical_and(calc < 1, st.collision_cond == BC_P)
        idx_per = np.logical_and(idx_per, (1 - calc) > new_ts)
        if np.any(idx_per):
            pos = st.positions[idx_per, :]
            gv = st.group_vel[idx_per, :]
            cf = st.collision_facets[idx_per].astype(int)
            cp = st.collision_positions[idx_per, :]
            cts = calc[idx_per]
            partner = tb["facet_partner"][cf]
            prev = copy.deepcopy(pos)
            first = np.nonzero(cts == 0)[0]
            prev[first, :] -= gv[first, :] * dt
            L = tb["facet_centroid"][partner, :] - tb["facet_centroid"][cf, :]
            new_pos = cp + L
            ts2, fc2, xc2 = timesteps_to_boundary(tb, new_pos, gv)
            cts = cts + np.linalg.norm(cp - prev, axis=1) / np.linalg.norm(gv * dt, axis=1)
            st.positions[idx_per, :] = new_pos
            new_ts[idx_per] = ts2
            st.collision_facets[idx_per] = fc2
            st.collision_positions[idx_per, :] = xc2
            st.collision_cond[idx_per] = collision_condition(tb, fc2)
            calc[idx_per] = cts

        # III. rough facets (:1636-1668, :1491-1544)
        idx_ref = np.logical_and(calc < 1, st.collision_cond == BC_R)
        idx_ref = np.logical_and(idx_ref, (1 - calc) > new_ts)
        if np.any(idx_ref):
            pos = st.positions[idx_ref, :]
            gv = st.group_vel[idx_ref, :]
            cf = st.collision_facets[idx_ref].astype(int)
            cp = st.collision_positions[idx_ref, :]
            cts = calc[idx_ref]
            prev = copy.deepcopy(pos)
            first = cts == 0
            prev[first, :] -= gv[first, :] * dt
            dist = np.linalg.norm(cp - prev, axis=1)
            vel = np.linalg.norm(gv, axis=1)
            new_cts = cts + dist / (vel * dt)
            in_modes = st.modes[idx_ref, :]
            out_modes, n_out, omega_out, spec = select_reflected_modes(
                tb, st.subvol_temperature, in_modes, cf, cp, st.occupation[idx_ref], st.omega[idx_ref],
                st.ids[idx_ref], rng, step, event[idx_ref])
            new_gv = tb["group_vel"][out_modes[:, 0], out_modes[:, 1], :]
            ts2, fc2, xc2 = timesteps_to_boundary(tb, cp, new_gv)
            om_modes = st.omega_modes[idx_ref]
            om_modes = np.where(spec, om_modes, out_modes[:, 0] * J + out_modes[:, 1])
            st.modes[idx_ref, :] = out_modes
            st.positions[idx_ref, :] = cp
            st.group_vel[idx_ref, :] = new_gv
            st.omega[idx_ref] = omega_out
            st.omega_modes[idx_ref] = om_modes
            st.occupation[idx_ref] = n_out
            new_ts[idx_ref] = ts2
            calc[idx_ref] = new_cts
            st.collision_positions[idx_ref, :] = xc2
            st.collision_facets[idx_ref] = fc2
            st.collision_cond[idx_ref] = collision_condition(tb, fc2)
            event[idx_ref] += 1

        # IV. finish the step for particles with no further hit (:1670-1681)
        idx_drift = np.logical_and(calc < 1, (1 - calc) < new_ts)
        if np.any(idx_drift):
            st.positions[idx_drift, :] += st.group_vel[idx_drift, :] * dt * (1 - calc[idx_drift]).reshape(-1, 1)
            new_ts[idx_drift] -= (1 - calc[idx_drift])
            calc[idx_drift] = 1

    st.n_timesteps = copy.copy(new_ts)


def calculate_energy(tb, st):
    """Per-SV energy density, reference_temp 'local'.  Population.py:704-728."""
    S = tb["sv_centres"].shape[0]
    dn = st.occupation - calculate_occupation(tb, st.subvol_temperature[st.subvol_id], st.omega)
    ref = crystal_energy_function(tb, st.subvol_temperature)
    st.energies = tb["hbar"] * st.omega * dn
    e = np.zeros(S)
    for sv in range(S):
        i = np.nonzero(st.subvol_id == sv)[0]
        e[sv] = st.energies[i].sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        if tb["norm_mean"]:
            norm = tb["n_active"] / st.subvol_N_p
            norm = np.where(np.isnan(norm), 0, norm)
        else:
            norm = tb["n_active"] / (tb["particle_density"] * tb["sv_volume"])
        e = e * norm
    e = normalise_to_density(tb, e)
    st.subvol_energy = e + ref


def get_subvol_id(tb, st, classifier=None):
    """Population.py:671-683."""
    S = tb["sv_centres"].shape[0]
    sv = classify(tb, st.positions, classifier)
    st.subvol_N_p = np.array([(sv == i).sum(dtype=int) for i in range(S)])
    st.N_p = int(st.subvol_N_p.sum())
    return sv


def refresh_temperatures(tb, st, classifier=None):
    """Population.py:685-702."""
    st.subvol_id = get_subvol_id(tb, st, classifier)
    calculate_energy(tb, st)
    st.subvol_temperature = temperature_function(tb, st.subvol_energy)
    st.temperatures = particle_temperature(tb, st.subvol_temperature, st.positions)


class SciPyBackend:
    """The third-party objects the reference itself evaluates on the hot path, for the CPU timing legs
    of bench.py: cKDTree behind NearestNDInterpolator (Geometry.py:1210), RegularGridInterpolator for
    tau (Phonon.py:336) and the per-step gc.collect() (Population.py:1769).  Results are identical to
    the NumPy restatements above (tests/test_oracle_pin.py::test_restated_thirdparty_formulas)."""

    def __init__(self, tb, collect_garbage=True):
        from scipy.interpolate import RegularGridInterpolator
        from scipy.spatial import cKDTree
        Q, J = tb["omega"].shape
        self.kdtree = cKDTree(tb["sv_centres"])
        self.rgi = RegularGridInterpolator((tb["T_grid"], np.arange(Q), np.arange(J)), tb["tau"])
        self.collect_garbage = collect_garbage


def lifetime_scattering(tb, st, backend=None):
    """Deterministic relaxation toward Bose-Einstein.  Population.py:1701-1710."""
    if backend is not None:
        tau = backend.rgi(np.hstack((st.temperatures.reshape(-1, 1), st.modes)))
    else:
        tau = lifetime_function(tb, st.temperatures, st.modes)
    n0 = calculate_occupation(tb, st.temperatures, st.omega)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        st.occupation = np.where(tau > 0, n0 + (st.occupation - n0) * np.exp(-tb["dt"] / tau), n0)


def calculate_heat_flux(tb, st):
    """Population.py:730-747."""
    S = tb["sv_centres"].shape[0]
    hf = np.zeros((S, 3))
    for i in range(S):
        ind = np.nonzero(st.subvol_id == i)[0]
        hf[i, :] = np.sum(st.group_vel[ind, :] * st.energies[ind].reshape(-1, 1), axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        if tb["norm_mean"]:
            norm = tb["n_active"] / st.subvol_N_p.reshape(-1, 1)
        else:
            norm = tb["n_active"] / (tb["particle_density"] * tb["sv_volume"].reshape(-1, 1))
        hf = hf * norm
    hf = normalise_to_density(tb, hf)
    return hf * tb["eVpsa2_in_Wm2"]


def calculate_kappa(tb, st):
    """Slice subvolumes only (two reservoirs in facet order).  Population.py:749-771."""
    if not tb["sv_slice"]:
        return
    S = tb["sv_centres"].shape[0]
    ax = int(tb["slice_axis"])
    T = np.zeros(S + 2)
    T[1:-1] = st.subvol_temperature
    T[[0, -1]] = tb["res_T"]
    phi = st.subvol_heat_flux[:, ax]
    L = np.ptp(tb["bounds"][:, ax])
    dx = 2 * L * tb["a_in_m"] / S
    dT = T[2:] - T[:-2]
    DX = L * tb["a_in_m"] * (1 + S) / S
    DT = T[-1] - T[0]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        st.subvol_kappa = -phi * dx / dT
        st.kappa = -np.sum(phi * st.subvol_N_p) * (DX / DT) / st.N_p
    st.subvol_kappa[np.absolute(st.subvol_kappa) == np.inf] = 0


def adjust_reservoir_balance(tb, st):
    """Population.py:1685-1693."""
    R = tb["res_facet"].shape[0]
    if R == 0:
        return
    area = tb["facet_area"][tb["res_facet"]].reshape(-1, 1)
    st.res_heat_flux = st.res_heat_flux * (tb["n_active"] / (tb["particle_density"] * tb["dt"] * tb["n_dt_to_conv"] * area))
    st.res_heat_flux = normalise_to_density(tb, st.res_heat_flux)
    st.res_heat_flux = st.res_heat_flux * tb["eVpsa2_in_Wm2"]
    st.res_energy_balance = st.res_energy_balance * (tb["n_active"] / (tb["particle_density"] * tb["dt"] * tb["n_dt_to_conv"]))
    st.res_energy_balance = normalise_to_density(tb, st.res_energy_balance)


def restart_reservoir_balance(tb, st):
    """Population.py:1695-1699."""
    R = tb["res_facet"].shape[0]
    st.res_heat_flux = np.zeros((R, 3))
    st.res_energy_balance = np.zeros(R)


def run_timestep(tb, st, rng, classifier=None, on_convergence=None, backend=None):
    """One timestep without the every-100-step output branch.  Population.py:1724-1769."""
    if backend is not None:
        classifier = backend.kdtree
    drift(tb, st)
    if tb["res_facet"].shape[0] > 0:
        new = fill_reservoirs(tb, st, rng)
        add_reservoir_particles(tb, st, new)
    boundary_scattering(tb, st, rng, classifier)
    refresh_temperatures(tb, st, classifier)
    lifetime_scattering(tb, st, backend)
    st.current_timestep += 1
    if st.current_timestep % tb["n_dt_to_conv"] == 0:
        st.subvol_heat_flux = calculate_heat_flux(tb, st)
        calculate_kappa(tb, st)
        adjust_reservoir_balance(tb, st)
        if on_convergence is not None:
            on_convergence(st)
        restart_reservoir_balance(tb, st)
    if backend is not None and backend.collect_garbage:
        import gc
        gc.collect()                                                                  # Population.py:1769


def make_state(tb, positions, modes, subvol_temperature, res_counter, ids=None):
    """State of a freshly initialised population: occupation = Bose-Einstein at the subvolume
    temperature, first collisions for everybody (Population.py:270-321)."""
    positions = np.array(positions, dtype=float)
    modes = np.array(modes, dtype=int)
    J = tb["omega"].shape[1]
    R = tb["res_facet"].shape[0]
    n = positions.shape[0]
    sv = classify(tb, positions)
    T_sv = np.array(subvol_temperature, dtype=float)
    omega = tb["omega"][modes[:, 0], modes[:, 1]]
    st = State(positions=positions, modes=modes, omega=omega, group_vel=tb["group_vel"][modes[:, 0], modes[:, 1], :],
               occupation=calculate_occupation(tb, T_sv[sv], omega), n_timesteps=None, collision_facets=None,
               collision_positions=None, collision_cond=None, temperatures=T_sv[sv],
               ids=np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64),
               subvol_temperature=T_sv, res_counter=np.array(res_counter, dtype=float))
    st.omega_modes = modes[:, 0] * J + modes[:, 1]
    st.subvol_id = sv
    st.res_energy_balance = np.zeros(R)
    st.res_heat_flux = np.zeros((R, 3))
    st.N_leaving = np.zeros(R, dtype=int)
    with np.errstate(all="ignore"):
        init_collisions(tb, st)
    return st


def init_collisions(tb, st):
    """First boundary hit of every particle + initial per-SV quantities.  Population.py:308-321."""
    st.n_timesteps, st.collision_facets, st.collision_positions = timesteps_to_boundary(tb, st.positions, st.group_vel)
    st.collision_cond = collision_condition(tb, st.collision_facets)


# ------------------------------------------------------------------------------------------------
# particle shards: the same timestep with the population split over ranks.  Only the
# per-subvolume / per-reservoir sums are exchanged (reduce_fn = all-reduce sum of one f64 vector).
# ------------------------------------------------------------------------------------------------
def emission_owner(fire, copy, mode, world):
    """Rank that injects copy `copy` (0-based) of a table entry of flat mode `mode` that has emitted `fire` particles before
    this step (the device keeps `fire` in 8 bits; nanokappa_b200/csrc/nk_stream.cuh: nk_emit_owner)."""
    return (np.asarray(fire) % 256 + copy + mode) % world


def run_timestep_sharded(tb, st, rng, reduce_fn, rank, world):
    """One timestep of one rank: every rank advances the whole reservoir table (identical on all ranks) and keeps the new
    particles dealt to it (round-robin per table entry, emission_owner), local boundary loop, per-SV sums all-reduced
    before the temperatures are inverted.  With keyed random draws the union of all ranks' particles equals the
    single-rank run."""
    S = tb["sv_centres"].shape[0]
    R = tb["res_facet"].shape[0]
    Q, J = tb["omega"].shape
    M = Q * J
    drift(tb, st)
    if R > 0:
        if getattr(st, "res_fire", None) is None:
            st.res_fire = np.zeros((R, M), dtype=np.int64)
        new = fill_reservoirs(tb, st, rng)
        ids = new["ids"]
        if str(tb.get("res_gen", "constant")) == "one_to_one":
            # the k-th re-emitted particle of a reservoir belongs to rank k % world
            k = (ids - philox.EMIT_ID_BASE) % (M * philox.EMIT_CMAX)
            mine = (k % world) == rank
        else:
            e = ids - philox.EMIT_ID_BASE
            copy = e % philox.EMIT_CMAX
            m = (e // philox.EMIT_CMAX) % M
            r = (e // philox.EMIT_CMAX // M) % R
            mine = emission_owner(st.res_fire[r, m], copy, m, world) == rank
            np.add.at(st.res_fire, (r, m), 1)               # every copy advances the entry's deal counter, on every rank
        n_new = ids.shape[0]
        for k, v in list(new.items()):
            if isinstance(v, np.ndarray) and v.shape[:1] == (n_new,):
                new[k] = v[mine]
        add_reservoir_particles(tb, st, new)
    boundary_scattering(tb, st, rng)
    sv = classify(tb, st.positions)
    st.subvol_id = sv
    dn = st.occupation - calculate_occupation(tb, st.subvol_temperature[sv], st.omega)
    st.energies = tb["hbar"] * st.omega * dn
    vec = np.concatenate([np.bincount(sv, weights=st.energies, minlength=S), np.bincount(sv, minlength=S).astype(float),
                          st.N_leaving.astype(float)])
    vec = reduce_fn(vec)
    e, cnt = vec[:S], vec[S:2 * S]
    st.N_leaving = vec[2 * S:].astype(int)
    st.subvol_N_p = cnt.astype(int)
    st.N_p = int(cnt.sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        norm = tb["n_active"] / cnt if tb["norm_mean"] else tb["n_active"] / (tb["particle_density"] * tb["sv_volume"])
        norm = np.where(np.isnan(norm), 0, norm)
        st.subvol_energy = normalise_to_density(tb, e * norm) + crystal_energy_function(tb, st.subvol_temperature)
    st.subvol_temperature = temperature_function(tb, st.subvol_energy)
    st.temperatures = particle_temperature(tb, st.subvol_temperature, st.positions)
    lifetime_scattering(tb, st)
    st.current_timestep += 1


def shard_state(st, lo, hi):
    """Rows [lo, hi) of a state as an independent State (per-SV vectors are shared values)."""
    out = st.copy()
    for name in ("positions", "group_vel", "omega", "occupation", "temperatures", "n_timesteps", "modes", "collision_facets",
                 "collision_positions", "collision_cond", "ids", "omega_modes", "subvol_id", "energies"):
        v = getattr(out, name)
        if v is not None:
            setattr(out, name, v[lo:hi].copy())
    return out
