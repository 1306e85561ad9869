"""ctypes binding of librlitg.so (C ABI: include/rlitg.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librlitg.so")

F32, F64 = 0, 1
Q_SOA, OUT_SOA = 1, 2


class RlitgError(RuntimeError):
    pass


BUILT_IN_THIS_PROCESS = False  # bench.py records whether the library was built on this box or arrived with the checkout
if not os.path.exists(LIB_PATH):
    # not a fallback: the same `make` __graft_entry__.build() runs, for a checkout that arrives without the built library
    import shutil
    import subprocess
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        import fcntl
        with open(os.path.join(_HERE, "csrc", ".build.lock"), "w") as _lock:  # one builder when N ranks start together
            fcntl.flock(_lock, fcntl.LOCK_EX)
            if not os.path.exists(LIB_PATH):
                BUILT_IN_THIS_PROCESS = True
                subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, check=False)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "librlitg.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
        "`make -C rl-itg_b200/csrc`. There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)
lib.rlitg_last_error.restype = C.c_char_p
lib.rlitg_launch_count.restype = C.c_int64

_vp, _i64, _i = C.c_void_p, C.c_int64, C.c_int
lib.rlitg_fk_jac.argtypes = [_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
lib.rlitg_fk.argtypes = [_vp, _i64, _i, _i, _i, _vp, _vp, _vp]
lib.rlitg_allfk.argtypes = [_vp, _i64, _i, _i, _i, _vp, _vp]
lib.rlitg_jacobian.argtypes = [_vp, _i64, _i, _i, _i, _vp, _vp]
lib.rlitg_fk_jac_host.argtypes = [_vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]


def check(rc: int) -> None:
    if rc != 0:
        raise RlitgError("rlitg error %d: %s" % (rc, (lib.rlitg_last_error() or b"").decode()))


def launch_count() -> int:
    return int(lib.rlitg_launch_count())


class on_device:
    """`with on_device(dev):` -- like torch.cuda.device(dev) but free when dev is already current (hot loops)."""
    __slots__ = ("idx", "ctx")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else 0
        self.ctx = None

    def __enter__(self):
        import torch
        if torch.cuda.current_device() != self.idx:
            self.ctx = torch.cuda.device(self.idx)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False
