"""TEST INFRASTRUCTURE -- NumPy Philox4x32-10 with the exact keying the CUDA path uses.

The reference draws every uniform from NumPy's global MT19937 in array order
(``Population.py:393, :949, :1003``, ``Mesh.py:937-942``); that order depends on row compaction and
cannot be reproduced on a GPU.  Fixed-draw parity therefore keys every draw by *who* needs it:

    counter = (id_lo, id_hi, step, stream)      key = (seed_lo, seed_hi)

and one Philox call yields two doubles ``u0, u1`` built like NumPy's ``random_sample``
(``(a >> 5) * 2**26 + (b >> 6)) / 2**53``).  Streams:

    0  emission:  u0 = entry-time draw (used when the mode emits c > 1 copies), u1 = face choice
    1  emission:  u0 = s, u1 = r  (barycentric surface sample, Mesh.py:941-947)
    2+e  e-th rough-wall event of the particle in this step: u0 = specular dice, u1 = diffuse pick
    0x40000000  contains_check resample (reserved)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)

STREAM_EMIT_A = 0
STREAM_EMIT_B = 1
STREAM_ROUGH0 = 2          # + index of the boundary event within the step (< 2**16)
STREAM_EMIT_C = 1 << 16    # fixed_rate dice / one_to_one (mode, entry time)

EMIT_ID_BASE = np.int64(1) << np.int64(62)
EMIT_CMAX = 64


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            n0 = hi1 ^ c1 ^ np.uint64(k0)
            n1 = lo1
            n2 = hi0 ^ c3 ^ np.uint64(k1)
            n3 = lo0
            c0, c1, c2, c3 = n0 & MASK, n1, n2 & MASK, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def uniforms(ids, step, stream, seed):
    """Two doubles in [0,1) per id.  ids int64 array; step, stream ints (or arrays); seed int."""
    ids = np.asarray(ids, dtype=np.int64).astype(np.uint64)
    lo = ids & MASK
    hi = ids >> np.uint64(32)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r0, r1, r2, r3 = philox4x32_10(lo, hi, np.asarray(step, dtype=np.uint64), np.asarray(stream, dtype=np.uint64),
                                   seed & 0xFFFFFFFF, seed >> 32)
    u0 = ((r0 >> np.uint32(5)).astype(np.float64) * 67108864.0 + (r1 >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0
    u1 = ((r2 >> np.uint32(5)).astype(np.float64) * 67108864.0 + (r3 >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0
    return u0, u1


def emission_id(step, n_res, r, qj, n_modes, c):
    """Stable 64-bit id of the c-th copy (c >= 1) of mode qj emitted by reservoir r at `step`."""
    step = np.asarray(step, dtype=np.int64)
    return EMIT_ID_BASE + ((step * n_res + np.asarray(r, dtype=np.int64)) * n_modes + np.asarray(qj, dtype=np.int64)) * EMIT_CMAX + (np.asarray(c, dtype=np.int64) - 1)


def one_to_one_id(step, n_res, r, k, n_modes):
    """Id of the k-th particle re-emitted by reservoir r at `step` in the one_to_one mode (k < n_modes * EMIT_CMAX)."""
    step = np.asarray(step, dtype=np.int64)
    return EMIT_ID_BASE + (step * n_res + np.asarray(r, dtype=np.int64)) * (np.int64(n_modes) * EMIT_CMAX) + np.asarray(k, dtype=np.int64)
