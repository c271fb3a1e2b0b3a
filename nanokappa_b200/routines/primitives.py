"""Built-in wire shapes of ``--geometry`` other than box and cylinder (reference: Geometry.py:144-412).

Every one of them is a stack of polygonal rings along z closed by two fan caps, so one builder covers them all:

  zigzag      L R dx dy Ns Nc      Nc sections of length L, radius R; odd rings are displaced by (dx, dy)
  corrugated  L R r Ns Nc          Nc sections of length L; ring radii alternate R, r, R, ...
  castle      L l R r Ns Nc s      Nc sections alternating large (radius R, length L) and small (r, l); s: start large
  star        H R r N              N-pointed star prism of height H (tips at R, notches at r)
  freewire    R0 L0 R1 L1 ... Rn N rings of radius R_i separated by lengths L_i, N sides

Parameter meaning and order follow the reference; vertex / face numbering is this module's own (facets are addressed by
position through --bound_pos, never by number).  ``Mesh`` fixes the winding.
"""
import numpy as np


def _ring(radius, n, z, centre=(0.0, 0.0), phase=0.0):
    a = (np.arange(n) + phase) * 2 * np.pi / n
    r = np.broadcast_to(np.asarray(radius, dtype=float), (n,))
    return np.stack((r * np.cos(a) + centre[0], r * np.sin(a) + centre[1], np.full(n, float(z))), axis=1)


def stack_rings(rings):
    """rings: list of (n, 3) vertex loops with the same n, bottom to top.  -> vertices, faces of the closed surface:
    fan cap on the first ring, two triangles per side quad between consecutive rings, fan cap on the last ring."""
    n = rings[0].shape[0]
    k = len(rings)
    v = np.vstack([r for r in rings] + [rings[0].mean(axis=0, keepdims=True), rings[-1].mean(axis=0, keepdims=True)])
    bottom, top = k * n, k * n + 1
    i = np.arange(n)
    j = (i + 1) % n
    faces = [np.stack((np.full(n, bottom), i, j), axis=1)]
    for s in range(k - 1):
        a, b = s * n, (s + 1) * n
        faces.append(np.stack((a + i, a + j, b + j), axis=1))
        faces.append(np.stack((a + i, b + j, b + i), axis=1))
    faces.append(np.stack((np.full(n, top), (k - 1) * n + i, (k - 1) * n + j), axis=1))
    return v, np.vstack(faces).astype(int)


def generate(shape, dims):
    d = list(dims)
    if shape == 'zigzag':
        L, R, dx, dy, Ns, Nc = float(d[0]), float(d[1]), float(d[2]), float(d[3]), int(d[4]), int(d[5])
        rings = [_ring(R, Ns, i * L, (dx, dy) if i % 2 == 1 else (0.0, 0.0)) for i in range(Nc + 1)]
    elif shape == 'corrugated':
        L, R, r, Ns, Nc = float(d[0]), float(d[1]), float(d[2]), int(d[3]), int(d[4])
        rings = [_ring(r if i % 2 == 1 else R, Ns, i * L) for i in range(Nc + 1)]
    elif shape == 'castle':
        L, l, R, r, Ns, Nc, start_large = float(d[0]), float(d[1]), float(d[2]), float(d[3]), int(d[4]), int(d[5]), bool(d[6])
        if R <= r:
            raise Exception('Outer radius smaller or equal to the inner radius. Check parameters.')
        rings, z, large = [], 0.0, start_large
        if large:
            rings.append(_ring(r, Ns, 0.0))           # annular lid: the wire always starts and ends on the inner radius
        for _ in range(Nc):
            rad, length = (R, L) if large else (r, l)
            rings += [_ring(rad, Ns, z), _ring(rad, Ns, z + length)]
            z += length
            large = not large
        if not large:                                 # last section was a large one
            rings.append(_ring(r, Ns, z))
        keep = [rings[0]]
        for ring in rings[1:]:                        # consecutive small sections share a ring
            if not np.allclose(ring, keep[-1]):
                keep.append(ring)
        rings = keep
    elif shape == 'star':
        H, R, r, N = float(d[0]), float(d[1]), float(d[2]), int(d[3])
        if R <= r:
            raise Exception('Outer radius smaller or equal to the inner radius. Check parameters.')
        radii = np.empty(2 * N)
        radii[0::2], radii[1::2] = r, R               # notch at (i - 1/2) 2 pi / N, tip at i 2 pi / N
        rings = [_ring(radii, 2 * N, z, phase=-0.5) for z in (0.0, H)]
    elif shape == 'freewire':
        R = np.array(d[0:len(d) - 1:2], dtype=float)
        L = np.array(d[1:len(d) - 1:2], dtype=float)
        N = int(d[-1])
        if R.shape[0] < 2 or L.shape[0] < R.shape[0] - 1:
            raise Exception('freewire needs R0 L0 R1 [L1 R2 ...] N.')
        z = np.concatenate(([0.0], np.cumsum(L[:R.shape[0] - 1])))
        rings = [_ring(R[i], N, z[i]) for i in range(R.shape[0])]
    else:
        raise Exception("Unknown geometry '{}'.".format(shape))
    return stack_rings(rings)
