"""Cubic radial-basis temperature field: host set-up for ``nk_set_rbf``.

The reference builds ``scipy.interpolate.RBFInterpolator(centres, T_sv, kernel='cubic')`` every timestep
(Population.py:588, :697-702) and evaluates it at every particle.  The linear system of that interpolator,

    [[ |c_i - c_j|^3 , P ],      P = [1, (c - shift)/scale]   (degree-1 polynomial tail, no smoothing,
     [ P^T           , 0 ]]                                     shift/scale = centre/half-width of the centres' box)

depends on the subvolume centres only; the temperatures enter through the right-hand side.  So it is factorised
once here and handed to the device as ``W = lhs^-1[:, :S]``: each step the closing block of the timestep forms
``coeffs = W @ T_sv`` and the streaming kernel evaluates sum_s coeffs[s] r_s^3 + polynomial per particle.
"""
import numpy as np


def interp_dims(geometry):
    """Coordinates the non-slice interpolator sees: a grid with a collapsed direction drops it (Population.py:697-699)."""
    grid = getattr(geometry, 'grid', None)
    if geometry.subvol_type == 'grid' and grid is not None and np.any(np.asarray(grid) == 1):
        return np.nonzero(np.asarray(grid) != 1)[0].astype(np.int32)
    return np.arange(3, dtype=np.int32)


def cubic_rbf_weights(centres, dims):
    """-> (shift (nd), scale (nd), W (S+nd+1, S))."""
    y = np.ascontiguousarray(np.asarray(centres, dtype=np.float64)[:, np.asarray(dims, dtype=int)])
    S, nd = y.shape
    if S < nd + 1:
        raise Exception('At least {} subvolumes are required for the radial temperature interpolation in {} dimensions.'.format(nd + 1, nd))
    lo, hi = y.min(axis=0), y.max(axis=0)
    shift = (hi + lo) / 2
    scale = (hi - lo) / 2
    scale[scale == 0.0] = 1.0
    n = S + nd + 1
    lhs = np.zeros((n, n))
    d = y[:, None, :] - y[None, :, :]
    r = np.sqrt(np.einsum('ijk,ijk->ij', d, d))
    lhs[:S, :S] = r * r * r
    lhs[:S, S] = 1.0
    lhs[:S, S + 1:] = (y - shift) / scale
    lhs[S:, :S] = lhs[:S, S:].T
    rhs = np.zeros((n, S))
    rhs[:S, :] = np.eye(S)
    try:
        W = np.linalg.solve(lhs, rhs)
    except np.linalg.LinAlgError:
        raise Exception('Singular RBF system: the subvolume centres do not span the interpolated directions.')
    return shift, scale, np.ascontiguousarray(W)
