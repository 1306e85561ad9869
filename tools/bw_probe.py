#!/usr/bin/env python
"""HBM ceilings for write-heavy streams (the FK+Jacobian kernel writes 7.7x what it reads)."""
import torch

def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

for mb in (256, 1024):
    n = mb << 20
    x = torch.empty(n, dtype=torch.uint8, device="cuda"); y = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = t(lambda: x.zero_()); print("memset %4d MB: %.1f us  %.0f GB/s (write only)" % (mb, ms * 1e3, n / ms / 1e6))
    ms = t(lambda: y.copy_(x)); print("copy   %4d MB: %.1f us  %.0f GB/s (read+write)" % (mb, ms * 1e3, 2 * n / ms / 1e6))
    xf = x.view(torch.float32); 
    ms = t(lambda: xf.sum()); print("reduce %4d MB: %.1f us  %.0f GB/s (read only)" % (mb, ms * 1e3, n / ms / 1e6))

# the fk_jac drain path without arithmetic: TMA bulk stores from persistent CTAs, 1 : 7.7 read : write like fk_jac
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlitg_b200 import _lib
lib = _lib.lib
lib.rlitg_write_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
tile = 128 * 54 * 4          # pos + rot + J of 128 configurations
ntiles = 8192
dst = torch.empty(ntiles * tile, dtype=torch.uint8, device="cuda")
src = torch.zeros(ntiles * 128 * 7, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for ctas in (148 * 2, 148 * 3, 148 * 4, 148 * 6):
    for rf in (0, 128 * 7):
        fn = lambda: _lib.check(lib.rlitg_write_probe(dst.data_ptr(), src.data_ptr(), ntiles, tile, rf, ctas, st))
        ms = t(fn, 30)
        tot = ntiles * (tile + rf * 4)
        print("bulk-store probe ctas=%4d read=%d: %.1f us  %.0f GB/s" % (ctas, rf > 0, ms * 1e3, tot / ms / 1e6))
