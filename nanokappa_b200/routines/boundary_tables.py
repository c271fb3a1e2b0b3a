"""Rough-wall look-up tables (host-side set-up, uploaded once to the GPU).

What the reference computes in ``Population.calculate_fbz_specularity`` (Population.py:852-877),
``find_specular_correspondences`` ('velocity' model, :1241-1380, :1456-1459) and
``diffuse_scat_probability`` (:879-939), rebuilt around a k-d tree in velocity space instead of the
sorted-walk over v_x:

* ``specularity[f,q,j]  = exp(-(2 eta_f cos(theta))^2 |k_q|^2)`` masked by ``true_specular``
* ``true_specular[f,q,j]`` the incoming mode has at least one outgoing partner whose velocity equals
  the mirror image of its own within 1e-3 (per component relative to the larger speed, and by
  angle) and whose frequency lies within the sum of the two modes' grid uncertainties |v . dk|
* ``spec_out[f,q,j]``  flat index of the partner used on a specular hit (closest in frequency)
* ``roulette[f,:]``    cumulative diffuse creation rate: outgoing flux max(v.n,0) minus the flux
  fed specularly into each outgoing mode, rounded to 1e-10, cumulated and normalised
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

CRIT = 1e-3


def fbz_specularity(inward_normals, eta, wavevectors, group_vel):
    """(Fr,Q,J) Ziman-type specularity before masking; inward_normals = -facet normals (Fr,3)."""
    n = np.asarray(inward_normals, dtype=float)[:, None, None, :]
    k_norm = np.sum(wavevectors ** 2, axis=1) ** 0.5
    v_norm = np.sum(group_vel ** 2, axis=-1) ** 0.5
    dot = np.sum(group_vel * n, axis=-1)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        cos = dot / v_norm
    eta = np.asarray(eta, dtype=float)[:, None, None]
    spec = np.exp(-(2 * eta * cos) ** 2 * (k_norm[None, :, None] ** 2))
    spec[np.isnan(spec)] = 0
    return spec


def specular_pairs(n, omega, group_vel, delta_omega):
    """All (incoming flat mode, outgoing flat mode) pairs that are mirror images through a wall with
    inward unit normal n."""
    Q, J = omega.shape
    v = group_vel.reshape(-1, 3)
    w = omega.reshape(-1)
    dw = delta_omega.reshape(-1)
    vn = v @ n
    i_in = np.nonzero(vn < 0)[0]
    i_out = np.nonzero(vn > 0)[0]
    if i_in.shape[0] == 0 or i_out.shape[0] == 0:
        return np.zeros((0, 2), dtype=int)
    v_ref = v[i_in] - 2 * n * (v[i_in] @ n)[:, None]
    v_out = v[i_out]
    nr = np.linalg.norm(v_ref, axis=1)
    no = np.linalg.norm(v_out, axis=1)
    tree = cKDTree(v_out)
    radius = np.sqrt(3) * CRIT * nr / (1 - 4 * CRIT) + 1e-12
    pairs = []
    nb = tree.query_ball_point(v_ref, radius)
    for a, lst in enumerate(nb):
        if not lst:
            continue
        b = np.asarray(lst, dtype=int)
        ref = np.fmax(nr[a], no[b])
        d = np.abs(v_ref[a] - v_out[b])
        ok = np.all(d / ref[:, None] < CRIT, axis=1)
        ok &= np.abs(w[i_in[a]] - w[i_out[b]]) < dw[i_in[a]] + dw[i_out[b]]
        if not ok.any():
            continue
        b = b[ok]
        # angle between the mirrored incoming direction and the outgoing one (from the unreflected v, as upstream)
        vi = v[i_in[a]] / np.linalg.norm(v[i_in[a]])
        vi = vi - 2 * n * np.dot(vi, n)
        vo = v_out[b] / no[b][:, None]
        with np.errstate(invalid='ignore'):
            ang = np.arccos(np.sum(vi * vo, axis=1))
        ang[np.isnan(ang)] = np.pi
        b = b[ang < CRIT]
        for bb in np.sort(b):
            pairs.append((i_in[a], i_out[bb]))
    return np.array(pairs, dtype=int).reshape(-1, 2)


def build(rough_normals_outward, eta, phonon, scat_model='velocity'):
    """-> dict(specularity, true_specular, spec_out, roulette, correspondent_modes).
    rough_normals_outward (Fr,3): facet normals of the rough facets (pointing out of the solid)."""
    if scat_model not in ('v', 'vel', 'velocity', 'groupvel', 'group_vel'):
        raise Exception("--bound_scat '{}' is not available on this build (velocity model only)".format(scat_model))
    Q, J = phonon.omega.shape
    Fr = rough_normals_outward.shape[0]
    inward = -np.asarray(rough_normals_outward, dtype=float)
    spec = fbz_specularity(inward, eta, phonon.wavevectors, phonon.group_vel)
    true_spec = np.zeros((Fr, Q, J), dtype=bool)
    spec_out = -np.ones((Fr, Q * J), dtype=np.int64)
    roulette = np.zeros((Fr, Q * J))
    corr_rows = []
    if Fr == 0:
        return dict(specularity=spec, true_specular=true_spec, spec_out=spec_out.reshape(Fr, Q, J), roulette=roulette,
                    correspondent_modes=np.zeros((0, 7)))
    k_grid = phonon.q_to_k(np.absolute(1 / (2 * np.asarray(phonon.data_mesh, dtype=float))))
    delta_omega = np.sum((phonon.group_vel * k_grid) ** 2, axis=2) ** 0.5
    normals = np.round(inward, decimals=10)
    uniq, inv = np.unique(normals, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    w = phonon.omega.reshape(-1)
    v_flat = phonon.group_vel.reshape(-1, 3)
    for i_n, n in enumerate(uniq):
        pairs = specular_pairs(n, phonon.omega, phonon.group_vel, delta_omega)
        facets = np.nonzero(inv == i_n)[0]
        if pairs.shape[0]:
            corr_rows.append(np.hstack((np.tile(n, (pairs.shape[0], 1)), np.stack((pairs[:, 0] // J, pairs[:, 0] % J,
                                                                                   pairs[:, 1] // J, pairs[:, 1] % J), axis=1))))
            # one partner per incoming mode: closest frequency, then lowest index
            dw = np.abs(w[pairs[:, 0]] - w[pairs[:, 1]])
            order = np.lexsort((pairs[:, 1], dw, pairs[:, 0]))
            first = np.ones(order.shape[0], dtype=bool)
            first[1:] = pairs[order[1:], 0] != pairs[order[:-1], 0]
            chosen = pairs[order[first]]
        for f in facets:
            if pairs.shape[0]:
                true_spec[f].reshape(-1)[pairs[:, 0]] = True
                spec_out[f, chosen[:, 0]] = chosen[:, 1]
        # diffuse creation rate: outgoing flux minus what arrives specularly (Population.py:895-939)
        vdn = v_flat @ n
        creation = np.where(vdn > 0, vdn, 0.0)
        destruction = np.where(vdn < 0, -vdn, 0.0)
        for f in facets:
            rate = creation.copy()
            if pairs.shape[0]:
                spec_d = destruction * (spec[f].reshape(-1) * true_spec[f].reshape(-1))
                np.subtract.at(rate, pairs[:, 1], spec_d[pairs[:, 0]])
            rate = np.around(np.where(np.isnan(rate), 0, rate), decimals=10)
            cs = np.cumsum(rate)
            roulette[f] = cs / cs.max()
    spec = true_spec.astype(int) * spec
    corr = np.vstack(corr_rows) if corr_rows else np.zeros((0, 7))
    return dict(specularity=spec, true_specular=true_spec, spec_out=spec_out.reshape(Fr, Q, J), roulette=roulette,
                correspondent_modes=corr)
