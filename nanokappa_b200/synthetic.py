"""Synthetic phono3py-like mode tables.

The reference reads ``kappa-mNNN.hdf5`` (``Phonon.py:153-187``); both shipped hdf5 files are absent
from the reference checkout (``.MISSING_LARGE_BLOBS``) and h5py is not installed, so benchmarks and
parity tests run on an analytic table of the same shape and regime: a Gamma-centred ``n^3`` mesh on
the primitive cell of the POSCAR, six isotropic branches folded into the first Brillouin zone,
``gamma ~ omega^2 T``.  The raw arrays use the hdf5 field meanings after FBZ expansion:

* ``q_points (Q,3)``   reduced coordinates in [0,1)
* ``omega (Q,J)``      rad*THz
* ``group_vel (Q,J,3)``  A*THz, rounded to 1e-10 like ``Phonon.py:102``
* ``gamma (NT,Q,J)``   THz on ``temperature_array (NT,)``
"""
from __future__ import annotations

import numpy as np

SI_LATTICE = np.array([[0.0, 2.7343755164098931, 2.7343755164098931],
                       [2.7343755164098931, 0.0, 2.7343755164098931],
                       [2.7343755164098931, 2.7343755164098931, 0.0]])
GE_LATTICE = np.array([[0.0, 2.8916046182323498, 2.8916046182323498],
                       [2.8916046182323498, 0.0, 2.8916046182323498],
                       [2.8916046182323498, 2.8916046182323498, 0.0]])

F_MAX_THZ = np.array([4.0, 4.0, 12.0, 14.0, 15.0, 15.5])


def read_poscar_lattice(path):
    """VASP-5 POSCAR: line 2 scale, lines 3-5 cell vectors as rows (test_material/Si/POSCAR:1-5)."""
    with open(path, "r") as fh:
        lines = fh.read().splitlines()
    scale = float(lines[1].split()[0])
    cell = np.array([[float(t) for t in lines[2 + i].split()[:3]] for i in range(3)])
    return cell * scale


def fold_to_fbz(q, reciprocal_lattice):
    """Shortest equivalent wavevector of each reduced q (same contract as ``Phonon.find_min_k``,
    Phonon.py:209-247: walk over the 27 neighbouring reciprocal cells until the origin cell is the
    minimiser).  Returns cartesian k (N,3)."""
    q = np.array(q, dtype=float)
    shifts = np.array([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], dtype=float)
    i0 = 13
    active = np.ones(q.shape[0], dtype=bool)
    while active.any():
        cand = q[active][None, :, :] + shifts[:, None, :]            # (27, Na, 3)
        k = cand @ reciprocal_lattice.T
        norm = np.linalg.norm(k, axis=-1).T                         # (Na, 27)
        imin = np.argmax(norm == norm.min(axis=1, keepdims=True), axis=1)
        q[active] = cand[imin, np.arange(imin.shape[0])]
        active[active] = imin != i0
    return q @ reciprocal_lattice.T


def make_table(n_mesh=11, lattice=None, nt_step=10.0, t_max=1000.0, gamma_coef=2e-7, f_scale=1.0):
    """Build the raw table.  ``n_mesh=31`` gives the Si/Ge production shape (Q=29791, J=6)."""
    lattice = SI_LATTICE if lattice is None else np.asarray(lattice, dtype=float)
    recip = np.around(np.linalg.inv(lattice) * 2 * np.pi, decimals=6)   # Phonon.py:72,129
    g = np.arange(n_mesh) / n_mesh
    q_points = np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T
    k = fold_to_fbz(q_points, recip)
    knorm = np.linalg.norm(k, axis=1)
    kmax = knorm.max()
    x = knorm / kmax
    with np.errstate(invalid="ignore", divide="ignore"):
        khat = np.where(knorm[:, None] > 0, k / knorm[:, None], 0.0)
    fmax = F_MAX_THZ * f_scale
    nq = q_points.shape[0]
    freq = np.zeros((nq, 6))
    dfdx = np.zeros((nq, 6))
    for j in range(3):
        freq[:, j] = fmax[j] * np.sin(np.pi * x / 2)
        dfdx[:, j] = fmax[j] * (np.pi / 2) * np.cos(np.pi * x / 2)
    for j in range(3, 6):
        freq[:, j] = fmax[j] - 2.0 * x ** 2
        dfdx[:, j] = -4.0 * x
    omega = 2 * np.pi * freq
    vel = (2 * np.pi * dfdx / kmax)[:, :, None] * khat[:, None, :]
    vel = np.around(vel, decimals=10)
    temperature = np.arange(0.0, t_max + nt_step / 2, nt_step)
    gamma = gamma_coef * (omega ** 2)[None, :, :] * temperature[:, None, None] / (2 * np.pi)
    return dict(omega=omega, group_vel=vel, gamma=gamma, temperature_array=temperature,
                q_points=q_points, lattice=lattice, data_mesh=np.array([n_mesh] * 3))
