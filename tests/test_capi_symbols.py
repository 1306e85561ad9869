"""CPU test: librlitg.so loads and exports every symbol include/rlitg.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rlitg.h")).read()
    return sorted(set(re.findall(r"RLITG_API[^;(]*?\b(rlitg_\w+)\s*\(", src)))


def test_library_exports_all_declared_symbols():
    path = os.path.join(ROOT, "rl-itg_b200", "librlitg.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 8
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.rlitg_version() == 100


def test_argument_validation_without_gpu():
    from rlitg_b200 import _lib
    lib = _lib.lib
    # shape / dtype / null errors are detected before any CUDA call
    assert lib.rlitg_fk_jac(None, 4, 7, 0, 0, None, None, None, None, None) == -1
    assert lib.rlitg_fk_jac(ctypes.c_void_p(16), 4, 6, 0, 0, None, None, None, None, None) == -2
    assert lib.rlitg_fk_jac(ctypes.c_void_p(16), 4, 7, 5, 0, None, None, None, None, None) == -3
    assert lib.rlitg_fk_jac(ctypes.c_void_p(18), 4, 7, 0, 0, None, None, None, None, None) == -4
    assert b"aligned" in lib.rlitg_last_error()
    assert lib.rlitg_fk_jac(ctypes.c_void_p(16), 4, 7, 0, 0, None, None, None, None, None) == -1  # all outputs NULL
    assert lib.rlitg_fk_jac(None, 0, 7, 0, 0, None, None, None, None, None) == 0  # empty batch is a no-op


def test_header_is_strict_c99(tmp_path):
    """include/rlitg.h is a C header (the reference's Python side would bind it through ctypes / cffi, its C++ side
    includes it directly): it must compile as pedantic C99 on its own."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "rlitg.h"\nint main(void) { return rlitg_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
