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


def test_header_is_plain_c(tmp_path):
    """include/nk_b200.h is the drop-in boundary: it must compile as C99 (and as C++) on its own -- plain pointers and
    sizes, no C++ or torch types -- and a C program must link against the library."""
    import shutil
    import subprocess
    from nanokappa_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    src = tmp_path / "use_abi.c"
    src.write_text('#include "nk_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { nk_ctx* c = 0; int rc = nk_create(-1, &c);\n'
                   '  printf("%d %d %s\\n", nk_version(), rc, nk_last_error(0)); return rc == 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-c", str(src), "-o", str(tmp_path / "a.o")])
    if shutil.which("g++") is not None:
        subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-x", "c++", "-c", str(src), "-o", str(tmp_path / "b.o")])
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "use_abi"
    subprocess.check_call(["gcc", str(tmp_path / "a.o"), "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH),
                           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    # without a GPU (or with a bad index) nk_create fails cleanly and explains itself; version is reported either way
    assert out.returncode == 0, out.stderr
    ver, rc, msg = out.stdout.split(" ", 2)
    assert int(ver) >= 100 and int(rc) != 0 and len(msg.strip()) > 0
