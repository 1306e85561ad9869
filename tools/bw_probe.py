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
