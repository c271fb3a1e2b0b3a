"""ctypes binding of ``libnk_b200.so`` (include/nk_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present, importing
``lib()`` / creating a context raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C nanokappa_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NK_LIB", os.path.join(_HERE, "libnk_b200.so"))   # NK_LIB: experiment builds

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_up = C.POINTER(C.c_uint8)
VP = C.c_void_p

# name -> (restype, argtypes).  Kept in one table so the CPU test-suite can check that the library
# exports every symbol the header declares.
SIGNATURES = {
    "nk_create": (C.c_int, [C.c_int, C.POINTER(VP)]),
    "nk_destroy": (None, [VP]),
    "nk_last_error": (C.c_char_p, [VP]),
    "nk_version": (C.c_int, []),
    "nk_set_stream": (C.c_int, [VP, VP]),
    "nk_synchronize": (C.c_int, [VP]),
    "nk_set_mesh": (C.c_int, [VP, C.c_int, VP, VP, VP, VP, VP, VP, VP, VP, VP,
                              C.c_int, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP]),
    "nk_set_subvols": (C.c_int, [VP, C.c_int, VP, VP, C.c_int, C.c_int, C.c_int]),
    "nk_set_rbf": (C.c_int, [VP, C.c_int, VP, VP, VP, VP]),
    "nk_set_reservoir_mode": (C.c_int, [VP, C.c_int, VP]),
    "nk_set_phonon": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, VP, VP, VP, VP, C.c_double, C.c_double, C.c_double,
                                C.c_int64, C.c_int, VP, VP]),
    "nk_set_population": (C.c_int, [VP, C.c_double, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_double, C.c_double,
                                    C.c_double, C.c_double]),
    "nk_set_reservoirs": (C.c_int, [VP, C.c_int, VP, VP, VP, VP]),
    "nk_get_res_counter": (C.c_int, [VP, VP]),
    "nk_set_boundary_luts": (C.c_int, [VP, C.c_int, VP, VP, VP, VP]),
    "nk_bind_particles": (C.c_int, [VP, C.c_int64] + [VP] * 12),
    "nk_set_slot_count": (C.c_int, [VP, C.c_int64]),
    "nk_get_slot_count": (C.c_int, [VP, c_lp, c_lp]),
    "nk_sort_by_mode": (C.c_int, [VP] + [VP] * 12 + [C.c_double, C.c_int, c_lp, c_lp]),
    "nk_results_len": (C.c_int, [VP]),
    "nk_get_run_state": (C.c_int, [VP] + [VP] * 5),
    "nk_set_run_state": (C.c_int, [VP] + [VP] * 5),
    "nk_set_sv_temperature": (C.c_int, [VP, VP]),
    "nk_get_sv_temperature": (C.c_int, [VP, VP]),
    "nk_set_timestep": (C.c_int, [VP, C.c_int64]),
    "nk_get_timestep": (C.c_int, [VP, c_lp]),
    "nk_find_boundary": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP]),
    "nk_contains": (C.c_int, [VP, C.c_int64, VP, VP]),
    "nk_classify": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "nk_occupation": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "nk_lifetime": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "nk_temperature_of_energy": (C.c_int, [VP, C.c_int64, VP, VP]),
    "nk_energy_of_temperature": (C.c_int, [VP, C.c_int64, VP, VP]),
    "nk_particle_temperature": (C.c_int, [VP, C.c_int64, VP, VP]),
    "nk_init_collisions": (C.c_int, [VP]),
    "nk_step": (C.c_int, [VP, C.c_int]),
    "nk_flush_relaxation": (C.c_int, [VP]),
    "nk_profile_begin": (C.c_int, [VP]),
    "nk_profile_end": (C.c_int, [VP, VP, c_lp]),
    "nk_get_results": (C.c_int, [VP] + [VP] * 10),
    "nk_snapshot_results": (C.c_int, [VP, C.POINTER(C.c_int)]),
    "nk_get_snapshot": (C.c_int, [VP, C.c_int] + [VP] * 10),
    "nk_energy_table": (C.c_int, [C.c_int, C.c_int, VP, VP, C.c_int, VP, C.c_double, C.c_double, C.c_double, C.c_double, VP]),
    "nk_advance_host": (C.c_int, [VP, C.c_int64, C.c_int] + [VP] * 12 + [c_lp, VP, VP, VP]),
    "nk_last_transfer_bytes": (C.c_int, [VP, c_lp, c_lp]),
    "nk_debug_trace": (C.c_int, [VP, VP]),
    "nk_last_step_variant": (C.c_int, [VP]),
    "nk_outside_slots": (C.c_int, [VP, C.c_double, VP, C.c_int64, c_lp]),
    "nk_set_rank": (C.c_int, [VP, C.c_int, C.c_int]),
    "nk_acc_buffer": (C.c_int, [VP, C.POINTER(VP), c_lp]),
    "nk_step_local": (C.c_int, [VP]),
    "nk_step_finalize": (C.c_int, [VP]),
    "nk_comm_export": (C.c_int, [VP, VP]),
    "nk_comm_import": (C.c_int, [VP, C.c_int, VP]),
    "nk_comm_enable": (C.c_int, [VP, C.c_int]),
}

_LIB = None


class NkError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(LIB_PATH):
        raise NkError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                      f"(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def check(ctx, rc, what=""):
    if rc != 0:
        msg = lib().nk_last_error(ctx)
        raise NkError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
