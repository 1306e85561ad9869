#!/usr/bin/env python
"""Does the DRAM fetch granularity of L2 misses (cudaLimitMaxL2FetchGranularity, a device-wide hint) change the PIRL env
step at 1M envs, whose target / demo rows are 32-byte gathers?  Prints the limit before / after and the step time.
usage: python tools/l2_fetch_probe.py [bytes]   (run it under ncu --metrics dram__bytes_read.sum for the traffic)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    want = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    torch.zeros(1, device="cuda")
    rt = ctypes.CDLL("libcudart.so.12")
    val = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)   # cudaLimitMaxL2FetchGranularity
    print("L2 fetch granularity before: %d" % val.value)
    if want:
        print("set ->", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(want)))
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        print("L2 fetch granularity after: %d" % val.value)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench_also as BA
    dev = torch.device("cuda:0")
    env, act, ts0, q0, demos, targets_b = BA.pirl_env(torch, dev, 1048576, 16, 0)
    for _ in range(3):
        env.step(act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        env.step(act)
    e1.record()
    torch.cuda.synchronize()
    print("pirl 1M step: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))


if __name__ == "__main__":
    main()
