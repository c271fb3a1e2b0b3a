"""Voronoi subvolume centres by Lloyd relaxation on volume samples (set-up only).

Same contract as the reference's ``routines/subvolumes.distribute`` (subvolumes.py:39-98): start
from ``n_r`` random interior points, repeatedly move every centre to the centroid of the samples
nearest to it, doubling the sample count (1e3 -> 1e6) whenever the largest centre displacement
falls under 1e-8, and stop when that happens at the maximum sample count."""
import os

import numpy as np
from scipy.spatial import cKDTree


def distribute(mesh, n_r, folder=None, view=False, n_s=1000, n_s_max=1000000, criterion=1e-8, max_iter=2000, rng=None):
    rng = np.random if rng is None else rng
    n_s_max = int(float(os.environ.get('NK_VORONOI_MAX_SAMPLES', n_s_max)))     # tests shrink the Lloyd sample budget
    x_r = mesh.sample_volume(n_r, rng=rng)
    x_s = mesh.sample_volume(n_s, rng=rng)
    for it in range(max_iter):
        r = cKDTree(x_r).query(x_s)[1]
        cnt = np.bincount(r, minlength=n_r).astype(float)
        new = np.stack([np.bincount(r, weights=x_s[:, k], minlength=n_r) for k in range(3)], axis=1)
        with np.errstate(invalid='ignore', divide='ignore'):
            new = new / cnt[:, None]
        new = np.where(cnt[:, None] > 0, new, x_r)
        move = np.linalg.norm(new - x_r, axis=1).max()
        x_r = new
        if move < criterion:
            if n_s >= n_s_max:
                break
            n_s = min(int(n_s * 2), n_s_max)
            x_s = mesh.sample_volume(n_s, rng=rng)
    return x_r
