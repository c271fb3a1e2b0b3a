"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/nk_b200.h declares, with the ctypes signatures the host layer binds.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "nk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nk_[a-z_0-9]+)\s*\(", src)))


def test_library_is_built():
    from nanokappa_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    from nanokappa_b200 import _lib
    names = _header_functions()
    assert len(names) >= 35
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in nk_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in nanokappa_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_header_argument_counts_match_bindings():
    from nanokappa_b200 import _lib
    src = open(os.path.join(ROOT, "include", "nk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in re.findall(r"\b(nk_[a-z_0-9]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(_lib.SIGNATURES[name][1]), f"{name}: header has {n} args, binding {len(_lib.SIGNATURES[name][1])}"


def test_version_and_no_device_error_path():
    from nanokappa_b200 import _lib
    L = _lib.lib()
    assert L.nk_version() >= 100
    import torch
    if not torch.cuda.is_available():
        ctx = ctypes.c_void_p()
        assert L.nk_create(0, ctypes.byref(ctx)) != 0          # fails loudly, no fallback
        assert b"CUDA" in L.nk_last_error(None)
        from nanokappa_b200.engine import Engine
        with pytest.raises(_lib.NkError):
            Engine(0)
