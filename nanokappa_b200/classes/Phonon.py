"""Material mode tables (host-side set-up) with the reference's ``Phonon`` surface.

``Phonon(args, mat_index)`` loads a phono3py ``kappa-mNNN.hdf5`` when h5py + phonopy are importable
(Phonon.py:66-187, :515-564) and otherwise one of

* ``--hdf_file something.npz``   a full-Brillouin-zone table saved by ``save_table`` (fields of
  ``nanokappa_b200.synthetic.make_table``), or
* ``--hdf_file synthetic:N``     the analytic N^3 x 6 table of ``nanokappa_b200.synthetic``.

Derived tables follow the reference formulas: omega = 2 pi f (:165-167), v_g rounded to 1e-10
(:102), wavevectors folded into the FBZ (:189-247), tau = 1/(4 pi gamma) with tau = 0 where
gamma <= 0 (:324-329), Bose-Einstein occupation (:338-345), E(T)/T(E) on a 0.1 K grid (:352-390),
density normalisation by Q V_uc (:392-401).  Once an engine is attached the table *functions*
(``calculate_occupation``, ``lifetime_function``, ``temperature_function``,
``crystal_energy_function``) are evaluated by the CUDA kernels.
"""
from __future__ import annotations

import os

import numpy as np

from .Constants import Constants
from .. import synthetic


class Phonon(Constants):
    def __init__(self, arguments, mat_index=0, table=None):
        super().__init__()
        self.args = arguments
        self.mat_index = int(mat_index)
        self.engine = None
        self.get_mat_folder()
        if table is None:
            table = self.load_table()
        self.set_table(table)
        rot_args = getattr(self.args, 'mat_rotation', [])
        if len(rot_args) > 0:
            self.rotate_crystal()
        print('Material initialisation done!')

    # ---- loading -----------------------------------------------------------------------------------
    def get_mat_folder(self):
        mf = getattr(self.args, 'mat_folder', [''])
        folder = os.path.relpath(mf[self.mat_index]) if len(mf) > 0 and mf[self.mat_index] else ''
        if not os.path.isabs(folder):
            folder = os.path.join(os.getcwd(), folder)
        self.mat_folder = folder

    def load_table(self):
        name = self.args.hdf_file[self.mat_index]
        if name.startswith('synthetic:'):
            # synthetic:N[:si|ge] -- analytic N^3 x 6 table on the lattice of the POSCAR when the file exists, else on the
            # built-in Si cell, or on the Ge cell of test_material/Ge/POSCAR (frequencies scaled by 0.58: Ge's optical
            # branches end near 9 THz)
            parts = name.split(':')
            lattice, f_scale = None, 1.0
            poscar = os.path.join(self.mat_folder, self.args.poscar_file[self.mat_index])
            if os.path.isfile(poscar):
                lattice = synthetic.read_poscar_lattice(poscar)
            if len(parts) > 2 and parts[2].lower() == 'ge':
                f_scale = 0.58
                if lattice is None:
                    lattice = synthetic.GE_LATTICE
            return synthetic.make_table(int(parts[1]), lattice=lattice, f_scale=f_scale)
        path = name if os.path.isabs(name) else os.path.join(self.mat_folder, name)
        if path.endswith('.npz'):
            z = np.load(path)
            return {k: z[k] for k in z.files}
        return self.load_hdf_table(path)

    def load_hdf_table(self, path):
        """phono3py hdf5 -> full-BZ table (needs h5py and phonopy, as upstream does)."""
        try:
            import h5py
            from phonopy import Phonopy
            from phonopy.interface.calculator import read_crystal_structure
        except Exception as exc:
            raise Exception('Reading {} needs h5py and phonopy ({}). Convert the file to .npz with '
                            'nanokappa_b200.classes.Phonon.save_table on a machine that has them, or use '
                            '--hdf_file synthetic:N.'.format(path, exc))
        poscar = os.path.join(self.mat_folder, self.args.poscar_file[self.mat_index])
        unitcell, _ = read_crystal_structure(poscar, interface_mode='vasp')
        lattice = np.array(unitcell.cell)
        recip = np.linalg.inv(lattice) * 2 * np.pi
        ph = Phonopy(unitcell, [[1, 0, 0], [0, 1, 0], [0, 0, 1]], primitive_matrix=[[1, 0, 0], [0, 1, 0], [0, 0, 1]])
        rotations = ph.primitive_symmetry.get_reciprocal_operations()
        with h5py.File(path, 'r') as f:
            mesh = np.array(f['mesh']); q = np.array(f['qpoint']); w = np.array(f['weight'])
            freq = np.array(f['frequency']); vel = np.array(f['group_velocity']); T = np.array(f['temperature'])
            gamma = np.array(f['gamma'])
            if self.mat_index in getattr(self.args, 'isotope_scat', []):
                gamma = gamma + np.array(f['gamma_isotope'])
        freq = np.where(freq < 0, 0, freq)
        qf, freq_f = expand_to_full_zone(q, w, freq, 0, rotations, recip, axis=0)
        _, vel_f = expand_to_full_zone(q, w, vel, 1, rotations, recip, axis=0)
        _, gam_f = expand_to_full_zone(q, w, gamma, 0, rotations, recip, axis=1)
        return dict(omega=freq_f * 2 * np.pi, group_vel=vel_f, gamma=gam_f, temperature_array=T, q_points=qf,
                    lattice=lattice, data_mesh=mesh)

    def set_table(self, tab):
        lattice = np.asarray(tab['lattice'], dtype=float)
        self.volume_unitcell = float(abs(np.linalg.det(lattice)))
        self.data_mesh = np.asarray(tab['data_mesh'])
        self.omega = np.array(tab['omega'], dtype=float)
        self.frequency = self.omega / (2 * self.pi)
        self.group_vel = np.around(np.asarray(tab['group_vel'], dtype=float), decimals=10)
        self.temperature_array = np.asarray(tab['temperature_array'], dtype=float)
        gamma = np.asarray(tab['gamma'], dtype=float)
        self.gamma = np.where(gamma > 0, gamma, -1)
        self.q_points = np.array(tab['q_points'], dtype=float)
        self.weights = np.ones(self.q_points.shape[0])
        self.number_of_qpoints = self.q_points.shape[0]
        self.number_of_branches = self.omega.shape[1]
        self.number_of_modes = self.number_of_qpoints * self.number_of_branches
        self.inactive_modes_mask = np.all(self.group_vel == 0, axis=2)
        self.number_of_inactive_modes = self.inactive_modes_mask.sum()
        self.number_of_active_modes = self.number_of_modes - self.number_of_inactive_modes
        self.reciprocal_lattice = np.around(np.linalg.inv(lattice) * 2 * np.pi, decimals=6)
        self.unique_modes = np.stack(np.meshgrid(np.arange(self.number_of_qpoints), np.arange(self.number_of_branches)), axis=-1).reshape(-1, 2).astype(int)
        self.get_wavevectors()
        self.get_norms()
        print('Material info: {:d} q-points; {:d} branches -> {:d} modes in total.'.format(self.number_of_qpoints, self.number_of_branches, self.number_of_modes))
        self.calculate_lifetime()
        self.zero_point = self.calculate_zeropoint()
        self.initialise_temperature_function()
        self.g_counts, self.g_bins = np.histogram(self.omega, bins=100)

    # ---- reciprocal space ----------------------------------------------------------------------------
    def k_to_q(self, k):
        return np.dot(k, np.linalg.inv(self.reciprocal_lattice).T)

    def q_to_k(self, q):
        return np.dot(q, self.reciprocal_lattice.T)

    def find_min_k(self, k, return_disp=False):
        """Equivalent wavevector inside the FBZ (Phonon.py:209-247)."""
        shifts = np.array([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], dtype=float)
        # same neighbour enumeration as np.meshgrid(a, a, a) ravelled: index 13 is the origin cell
        sh = np.array(np.meshgrid([-1, 0, 1], [-1, 0, 1], [-1, 0, 1])).reshape(3, -1).T.astype(float)
        i0 = int(np.nonzero(np.all(sh == 0, axis=1))[0][0])
        q = self.k_to_q(np.asarray(k, dtype=float))
        disp = np.zeros(q.shape)
        active = np.ones(q.shape[0], dtype=bool)
        while active.any():
            cand = q[active][None, :, :] + sh[:, None, :]
            norm = np.linalg.norm(self.q_to_k(cand), axis=-1).T
            imin = np.argmax(norm == norm.min(axis=1, keepdims=True), axis=1)
            if return_disp:
                disp[active] += sh[imin]
            q[active] = cand[imin, np.arange(imin.shape[0])]
            active[active] = imin != i0
        del shifts
        if return_disp:
            return self.q_to_k(q), self.q_to_k(disp)
        return self.q_to_k(q)

    def get_wavevectors(self):
        self.wavevectors = self.find_min_k(self.q_to_k(np.copy(self.q_points)))

    def get_norms(self):
        self.norm_group_vel = np.linalg.norm(self.group_vel, axis=2)
        self.norm_wavevectors = np.linalg.norm(self.wavevectors, axis=1)

    def rotate_crystal(self):
        """--mat_rotation: Euler angles + order per material (Phonon.py:284-314)."""
        import re
        from scipy.spatial.transform import Rotation as rot
        groups, g = [], []
        for i, s in enumerate(self.args.mat_rotation):
            s = str(s)
            g.append(i)
            if re.fullmatch('[A-Z]+|[a-z]+', s):
                groups.append(g)
                g = []
        if groups:
            params = [self.args.mat_rotation[i] for i in groups[self.mat_index]]
            R = rot.from_euler(params[-1], [float(a) for a in params[:-1]], degrees=True)
            self.wavevectors = R.apply(self.wavevectors)
            for j in range(self.number_of_branches):
                self.group_vel[:, j, :] = R.apply(self.group_vel[:, j, :])

    # ---- table functions -----------------------------------------------------------------------------
    def calculate_lifetime(self):
        with np.errstate(divide='ignore', invalid='ignore'):
            self.lifetime = np.where(self.gamma > 0, 1 / (2 * 2 * np.pi * self.gamma), 0)

    def lifetime_function(self, Tqj):
        """tau at rows [T, q, j], linear in T (Phonon.py:326-336)."""
        if self.engine is not None:
            return self.engine.lifetime_function(Tqj)
        Tqj = np.asarray(Tqj, dtype=float).reshape(-1, 3)
        T, q, j = Tqj[:, 0], Tqj[:, 1].astype(int), Tqj[:, 2].astype(int)
        Tg = self.temperature_array
        if np.any(T < Tg[0]) or np.any(T > Tg[-1]):
            raise ValueError('One of the requested xi is out of bounds in dimension 0')
        i = np.clip(np.searchsorted(Tg, T, side='right') - 1, 0, Tg.shape[0] - 2)
        w = (T - Tg[i]) / (Tg[i + 1] - Tg[i])
        return self.lifetime[i, q, j] * (1 - w) + self.lifetime[i + 1, q, j] * w

    def calculate_occupation(self, T, omega):
        if self.engine is not None:
            return self.engine.calculate_occupation(T, omega)
        return self._occupation_host(T, omega)

    def _occupation_host(self, T, omega):
        T = np.asarray(T, dtype=float); omega = np.asarray(omega, dtype=float)
        flag = (T > 0) & (omega > 0)
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            return np.where(~flag, 0, 1 / (np.exp(omega * self.hbar / (T * self.kb)) - 1))

    def calculate_energy(self, T, omega):
        return self.hbar * omega * self._occupation_host(T, omega)

    def normalise_to_density(self, x):
        return x / (self.number_of_qpoints * self.volume_unitcell)

    def calculate_zeropoint(self):
        return self.normalise_to_density(self.hbar * self.omega.sum() / 2)

    def calculate_crystal_energy(self, T):
        """Energy density of the crystal at temperature(s) T (Phonon.py:352-362)."""
        T = np.array(T, dtype=float).reshape((-1, 1, 1))
        e = (self.calculate_energy(T, self.omega) * ~self.inactive_modes_mask).sum(axis=(1, 2))
        return self.normalise_to_density(e) + self.zero_point

    def initialise_temperature_function(self):
        """E(T) on a 0.1 K grid between the table's T_min and T_max (Phonon.py:372-390)."""
        T_min, T_max = self.temperature_array.min(), self.temperature_array.max()
        dT = 0.1
        self.T_array = np.arange(T_min, T_max + dT, dT)
        # optional on-disk cache of the (slow: nT x modes exponentials) table, keyed by its inputs
        cache = os.environ.get('NK_TABLE_CACHE')
        path = None
        if cache:
            import hashlib
            h = hashlib.sha1()
            for arr in (self.omega, self.inactive_modes_mask, self.T_array, np.array([self.volume_unitcell, self.hbar, self.kb])):
                h.update(np.ascontiguousarray(arr).tobytes())
            path = os.path.join(cache, 'nk_energy_table_{}.npy'.format(h.hexdigest()[:16]))
            if os.path.isfile(path):
                self.energy_array = np.load(path)
                return
        self.energy_array = self._energy_table_device() if self.number_of_modes * self.T_array.shape[0] >= 2e8 else None
        if self.energy_array is None:
            chunk = max(1, int(4e6 // max(1, self.number_of_modes)))
            parts = [self.calculate_crystal_energy(self.T_array[i:i + chunk]) for i in range(0, self.T_array.shape[0], chunk)]
            self.energy_array = np.concatenate(parts).reshape(-1)
        if path:
            try:
                os.makedirs(cache, exist_ok=True)
                np.save(path, self.energy_array)
            except OSError:
                pass

    def _energy_table_device(self):
        """Large tables (31^3 x 6 modes x 10 001 temperatures = 1.8e9 exponentials, ~40 s in NumPy) are summed
        by the CUDA library when a GPU is visible; None -> caller falls back to the NumPy evaluation."""
        if os.environ.get('NK_ENERGY_TABLE', 'auto') == 'host':
            return None
        try:
            import ctypes as C
            import torch
            if not torch.cuda.is_available():
                return None
            from .. import _lib
            L = _lib.lib()
            dev = int(os.environ.get('LOCAL_RANK', 0)) % max(torch.cuda.device_count(), 1)
            om = np.ascontiguousarray(self.omega.reshape(-1), dtype=np.float64)
            act = np.ascontiguousarray(~self.inactive_modes_mask.reshape(-1), dtype=np.uint8)
            T = np.ascontiguousarray(self.T_array, dtype=np.float64)
            out = np.zeros(T.shape[0])
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            rc = L.nk_energy_table(dev, om.shape[0], p(om), p(act), T.shape[0], p(T), float(self.hbar), float(self.kb),
                                   float(self.number_of_qpoints * self.volume_unitcell), float(self.zero_point), p(out))
            return out if rc == 0 else None
        except Exception:
            return None

    def temperature_function(self, E):
        if self.engine is not None:
            return self.engine.temperature_function(E)
        E = np.asarray(E, dtype=float)
        y = np.interp(E, self.energy_array, self.T_array)
        y = np.where(E < self.energy_array[0], self.T_array[0], y)
        return np.where(E > self.energy_array[-1], self.T_array[-1], y)

    def crystal_energy_function(self, T):
        if self.engine is not None:
            return self.engine.crystal_energy_function(T)
        T = np.asarray(T, dtype=float)
        y = np.interp(T, self.T_array, self.energy_array)
        y = np.where(T < self.T_array[0], self.energy_array.min(), y)
        return np.where(T > self.T_array[-1], self.energy_array.max(), y)

    def g(self, omega):
        i = np.searchsorted(self.g_bins, omega, side='left')
        return self.g_counts[i - 1]

    def attach_engine(self, engine):
        self.engine = engine


def expand_to_full_zone(qpoints, weights, tensor, rank, rotations, reciprocal_lattice, axis=0):
    """Irreducible wedge -> full Brillouin zone by the crystal's reciprocal point-group operations;
    rank-1 tensors (group velocities) rotate with R_cart = B r B^-1 (reference Phonon.py:515-564)."""
    inv_b = np.linalg.inv(reciprocal_lattice)
    q_out, t_out = [], []
    for i, q in enumerate(qpoints):
        tq = np.take(tensor, i, axis=axis)
        star_q = np.array([np.dot(r, np.mod(q, 1.0)) for r in rotations], dtype=float)
        if rank == 0:
            star_t = np.array([tq for _ in rotations], dtype=float)
        else:
            star_t = np.array([np.dot(np.dot(reciprocal_lattice, np.dot(r, inv_b)), tq.T).T for r in rotations], dtype=float)
        star_q = np.around(np.mod(star_q, 1.0), decimals=6)
        uq, idx = np.unique(star_q, return_index=True, axis=0)
        if weights[i] != len(idx):
            raise Exception('error in FBZ expansion: weight does not match the star of q-point {}'.format(i))
        q_out.append(uq)
        t_out.append(star_t[idx])
    return np.concatenate(q_out, axis=0), np.swapaxes(np.concatenate(t_out, axis=0), 0, axis)


def save_table(path, phonon):
    np.savez_compressed(path, omega=phonon.omega, group_vel=phonon.group_vel,
                        gamma=np.where(phonon.gamma > 0, phonon.gamma, 0.0), temperature_array=phonon.temperature_array,
                        q_points=phonon.q_points, lattice=np.linalg.inv(phonon.reciprocal_lattice / (2 * np.pi)),
                        data_mesh=phonon.data_mesh)
