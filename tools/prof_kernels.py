#!/usr/bin/env python
"""Runs each hot kernel a few times at its benchmark size (for ncu captures and quick timings).
usage: python tools/prof_kernels.py [fk|pirl|torm|all] [--iters K] [--B ...]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="all")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--fkB", type=int, default=1 << 20)
    ap.add_argument("--envB", type=int, default=65536)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--T", type=int, default=200)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    from rlitg_b200 import kinematics as K, pirl_env as P, scene as S, synth, torm as TM
    if a.which in ("probe", "all"):
        import ctypes as C
        from rlitg_b200 import _lib
        lib = _lib.lib
        lib.rlitg_fma_probe.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.rlitg_dfma_probe.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        of, od = torch.zeros(1, device=dev), torch.zeros(1, device=dev, dtype=torch.float64)
        st = torch.cuda.current_stream().cuda_stream
        ms = timeit(lambda: _lib.check(lib.rlitg_fma_probe(148 * 8, 4096, of.data_ptr(), st)), a.iters)
        print("FP32 FMA pipe: %.1f TFLOP/s" % (2.0 * 148 * 8 * 256 * 4096 * 16 / ms / 1e9))
        ms = timeit(lambda: _lib.check(lib.rlitg_dfma_probe(148 * 8, 2048, od.data_ptr(), st)), a.iters)
        print("FP64 FMA pipe: %.2f TFLOP/s" % (2.0 * 148 * 8 * 256 * 2048 * 8 / ms / 1e9))
    if a.which in ("fk", "all"):
        q = torch.tensor(synth.random_configs(a.fkB, seed=1), dtype=torch.float32, device=dev)
        out = {"pos": torch.empty(a.fkB, 3, device=dev), "rot": torch.empty(a.fkB, 3, 3, device=dev),
               "jac": torch.empty(a.fkB, 6, 7, device=dev)}
        ms = timeit(lambda: K.fk_jac(q, out=out), a.iters)
        print("fk_jac B=%d: %.2f us  %.1f GB/s  %.2f G configs/s" % (a.fkB, ms * 1e3, 244 * a.fkB / ms / 1e6, a.fkB / ms / 1e6))
        qs = q.t().contiguous()
        ms = timeit(lambda: K.fk_jac(qs, q_soa=True, out_soa=True), a.iters)
        print("fk_jac SoA B=%d: %.2f us  %.1f GB/s" % (a.fkB, ms * 1e3, 244 * a.fkB / ms / 1e6))
        ms = timeit(lambda: K.allfk(q), a.iters)
        print("allfk B=%d: %.2f us  %.1f GB/s" % (a.fkB, ms * 1e3, 252 * a.fkB / ms / 1e6))
    if a.which in ("pirl", "all"):
        B, Tp = a.envB, 100
        rng = np.random.default_rng(11)
        demos = synth.random_joint_paths(B, Tp, seed=5)
        dq = torch.tensor(demos.reshape(-1, 7), dtype=torch.float32, device=dev)
        pos, quat = K.fk(dq)
        env = P.PirlVecEnv(B, device=dev)
        env.set_problems(torch.cat([pos, quat], 1).reshape(B, Tp, 7), demos, np.full(B, 10 ** 6, dtype=np.int32))
        env.reset()
        act = env.random_action() * 0.1
        ms = timeit(lambda: env.step(act), a.iters)
        print("pirl_env_step B=%d: %.2f us  %.1f GB/s  %.2f G env-steps/s" % (B, ms * 1e3, 536 * B / ms / 1e6, B / ms / 1e6))
        # the kernel without the Python launch path: 50 steps captured in one CUDA graph
        env.step(act); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50):
                env.step(act)
        msg = timeit(g.replay, max(a.iters // 5, 3)) / 50
        print("pirl_env_step B=%d (50 steps per graph): %.2f us per step  %.1f GB/s" % (B, msg * 1e3, 536 * B / msg / 1e6))
    if a.which in ("torm", "all"):
        N, T = a.N, a.T
        traj = torch.tensor(synth.linear_trajectories(N, T, seed=100, noise=0.05), dtype=torch.float32, device=dev)
        targets = synth.random_ee_paths(1, T, seed=7)[0]
        occ, _ = synth.random_box_scene((128, 128, 128), (-0.64, -1.28, -0.20), 0.02, seed=0)
        sc = S.Scene(dev)
        sc.set_occupancy(occ, (-0.64, -1.28, -0.20), 0.02, 1.0)
        tb = TM.TormBatch(targets, TM.TormParameters(min_clearence_=0.3), sc, dev, torch.float32)
        for name, mask in (("full", 2 | 4 | 8 | 16), ("feasible", 1 | 4 | 8 | 16), ("endpose", 2 | 16), ("coll", 8),
                           ("pose_grad", 2), ("eval", 16), ("smooth", 4)):
            ms = timeit(lambda: tb.cost_grad(traj, mask), a.iters)
            print("torm %-9s N=%d T=%d: %.1f us" % (name, N, T, ms * 1e3))


if __name__ == "__main__":
    main()
