#!/usr/bin/env python
"""Can the FK + Jacobian kernel write its outputs straight into the caller's pinned host buffers?
Pinned host memory is device-addressable under unified addressing, so the persistent kernel's TMA bulk stores can
target it directly: no staging buffer in HBM, no separate D2H copy, the PCIe write stream starts with the first tile.
This probe drains shared-memory tiles (the fk_jac drain path, no arithmetic: rlitg_write_probe) into a pinned host
buffer and compares the rate with the copy engine's D2H of the same bytes (single copy, and chunked over 1-3 streams).
usage: python tools/zero_copy_probe.py"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlitg_b200 import _lib

lib = _lib.lib
lib.rlitg_write_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]


def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    dev = torch.device("cuda:0")
    B = 1 << 20
    nbytes = B * 54 * 4                      # pos + rot + J of 1M configurations = 226 MB
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    hq = torch.empty(B * 7, dtype=torch.float32).pin_memory()
    dq = torch.empty(B * 7, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    dt = wall(lambda: host.copy_(devbuf, non_blocking=True))
    print("copy engine D2H, one copy of %d MB:            %.3f ms  %.1f GB/s" % (nbytes >> 20, dt * 1e3, nbytes / dt / 1e9))
    for nstreams in (1, 2, 3):
        for nchunks in (8, 16):
            streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
            hv, dv = host.chunk(nchunks), devbuf.chunk(nchunks)

            def chunked():
                for k in range(nchunks):
                    with torch.cuda.stream(streams[k % nstreams]):
                        hv[k].copy_(dv[k], non_blocking=True)
            dt = wall(chunked)
            print("copy engine D2H, %2d chunks over %d stream(s):    %.3f ms  %.1f GB/s" % (nchunks, nstreams, dt * 1e3, nbytes / dt / 1e9))
    side = torch.cuda.Stream(device=dev)

    def duplex():
        with torch.cuda.stream(side):
            dq.copy_(hq, non_blocking=True)
        host.copy_(devbuf, non_blocking=True)
    dt = wall(duplex)
    print("copy engine D2H + 29 MB H2D at once:             %.3f ms  %.1f GB/s out" % (dt * 1e3, nbytes / dt / 1e9))

    for tile in (128 * 54 * 4, 64 * 54 * 4, 32 * 54 * 4):
        ntiles = nbytes // tile
        for ctas in (4, 8, 16, 32, 74, 148, 296):
            fn = lambda: _lib.check(lib.rlitg_write_probe(host.data_ptr(), None, ntiles, tile, 0, ctas, st))
            dt = wall(fn)
            print("bulk stores -> pinned host, tile %5d B, %3d CTAs: %.3f ms  %.1f GB/s" % (tile, ctas, dt * 1e3, ntiles * tile / dt / 1e9))

    # the same with the H2D copy of q running beside it (PCIe is full duplex)
    tile, ntiles = 128 * 54 * 4, nbytes // (128 * 54 * 4)
    for ctas in (16, 148):
        def both():
            with torch.cuda.stream(side):
                dq.copy_(hq, non_blocking=True)
            _lib.check(lib.rlitg_write_probe(host.data_ptr(), None, ntiles, tile, 0, ctas, st))
        dt = wall(both)
        print("bulk stores -> pinned host + 29 MB H2D, %3d CTAs:  %.3f ms  %.1f GB/s out" % (ctas, dt * 1e3, nbytes / dt / 1e9))


if __name__ == "__main__":
    main()
