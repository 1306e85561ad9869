#!/usr/bin/env python
"""Pinned host <-> device copy rates on this box (the ceiling of the `e2e` number: 28 B in, 216 B out per configuration)."""
import torch

dev = torch.device("cuda:0")
n = 226492416
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n // 8, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: h.copy_(d, non_blocking=True))
print("D2H 226 MB: %.2f ms  %.1f GB/s" % (ms, n / ms / 1e6))
ms = timed(lambda: d.copy_(h, non_blocking=True))
print("H2D 226 MB: %.2f ms  %.1f GB/s" % (ms, n / ms / 1e6))


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


ms = timed(both)
print("D2H 226 MB with H2D 28 MB in flight: %.2f ms  -> %.1f M configs/s ceiling for fk_jac_host" % (ms, 1048576 / ms / 1e3))
