#!/usr/bin/env python
"""Where one localOptimizeTSGD iteration spends its time at the benchmark size (1,024 x 200, 128^3 grid).
usage: python tools/prof_tsgd.py [--N 1024] [--T 200]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def ev_time(fn, iters=50, warm=3):
    for _ in range(warm):
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
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--T", type=int, default=200)
    ap.add_argument("--parts-only", action="store_true", help="only the sweep over localOptimizeTSGD(parts=k)")
    a = ap.parse_args()
    from rlitg_b200 import scene as S, synth, torm as TM
    dev = torch.device("cuda:0")
    N, T = a.N, a.T
    traj = torch.tensor(synth.linear_trajectories(N, T, seed=100, noise=0.05), dtype=torch.float32, device=dev)
    targets = synth.random_ee_paths(1, T, seed=7)[0]
    occ, _ = synth.random_box_scene((128, 128, 128), (-0.64, -1.28, -0.20), 0.02, seed=0)
    sc = S.Scene(dev)
    sc.set_occupancy(occ, (-0.64, -1.28, -0.20), 0.02, 1.0)
    tb = TM.TormBatch(targets, TM.TormParameters(min_clearence_=0.3), sc, dev, torch.float32)
    feas = TM.STAGE_SMOOTH | TM.STAGE_COLL | TM.STAGE_EVAL
    pose = TM.STAGE_POSE_GRAD | TM.STAGE_EVAL
    al = torch.full((N,), 0.5, device=dev); lr = torch.full((N,), 0.01, device=dev)
    tw = traj.clone()
    if not a.parts_only:
        r = tb.cost_grad(tw, feas | TM.STAGE_POSE_GRAD)
        inc_s, inc_c, inc_p = r["inc_smooth"].clone(), r["inc_coll"].clone(), r["inc_pose"].clone()
        print("cost feasible   %.1f us" % (1e3 * ev_time(lambda: tb.cost_grad(tw, feas))))
        print("cost end-pose   %.1f us" % (1e3 * ev_time(lambda: tb.cost_grad(tw, pose))))
        print("update feasible %.1f us" % (1e3 * ev_time(lambda: tb.applyIncrements(tw.copy_(traj), True, inc_s, inc_c, None, al, lr, lr_schedule_mode=3))))
        print("update end-pose %.1f us" % (1e3 * ev_time(lambda: tb.applyIncrements(tw.copy_(traj), False, inc_pose=inc_p, alpha=al, learning_rate=lr, lr_schedule_mode=3))))
        def graphed(fn):
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return 1e3 * ev_time(g.replay)
        print("update feasible (graph replay, GPU only) %.1f us" % graphed(lambda: tb.applyIncrements(tw.copy_(traj), True, inc_s, inc_c, None, al, lr, lr_schedule_mode=3)))
        print("update end-pose (graph replay, GPU only) %.1f us" % graphed(lambda: tb.applyIncrements(tw.copy_(traj), False, inc_pose=inc_p, alpha=al, learning_rate=lr, lr_schedule_mode=3)))
        inside = traj * 0.5  # well inside the joint limits: handleJointLimits finds nothing to do
        print("update end-pose, no limit hits (graph)   %.1f us" % graphed(lambda: tb.applyIncrements(tw.copy_(inside), False, inc_pose=inc_p * 0.01, alpha=al, learning_rate=lr, lr_schedule_mode=3)))
        print("update feasible, no limit hits (graph)   %.1f us" % graphed(lambda: tb.applyIncrements(tw.copy_(inside), True, inc_s, inc_c, None, al, lr, lr_schedule_mode=3)))
        print("copy            (graph replay, GPU only) %.1f us" % graphed(lambda: tw.copy_(traj)))
        print("cost feasible   (graph replay, GPU only) %.1f us" % graphed(lambda: tb.cost_grad(tw, feas)))
        print("cost end-pose   (graph replay, GPU only) %.1f us" % graphed(lambda: tb.cost_grad(tw, pose)))
        print("copy only       %.1f us" % (1e3 * ev_time(lambda: tw.copy_(traj))))
        print("collide N*T     %.1f us" % (1e3 * ev_time(lambda: sc.collide(tw.reshape(-1, 7)))))
        import ctypes as C
        from rlitg_b200 import _lib
        any_hit = torch.zeros(N, dtype=torch.uint8, device=dev); st_idx = torch.zeros(N, dtype=torch.int32, device=dev)
        u = torch.rand(N, device=dev)
        off = (C.c_double * 3)(*tb.parameters.base_offset)
        cp = lambda: _lib.check(_lib.lib.rlitg_torm_collision_point(tw.data_ptr(), N, T, 0, sc.handle, off, u.data_ptr(), None,
                                                                    any_hit.data_ptr(), st_idx.data_ptr(),
                                                                    torch.cuda.current_stream().cuda_stream))
        print("collision point (graph replay, GPU only) %.1f us" % graphed(cp))
        print("cost feasible stochastic (graph)          %.1f us" % graphed(lambda: tb.cost_grad(tw, feas, stochastic_idx=st_idx)))
        for mode in (3, 1):
            for stoch in (True, False):
                for use_graph, block in ((False, 1), (True, 1), (True, 8)):
                    for overlap in (True, False):
                        tw.copy_(traj)
                        tb.localOptimizeTSGD(tw, 3, lr_schedule_mode=mode, use_cuda_graph=use_graph, stochastic=stoch, overlap=overlap)
                        iters = 2401 if use_graph else 60   # a graph is captured per call: amortise the capture
                        tw.copy_(traj)
                        torch.cuda.synchronize(); t0 = time.perf_counter()
                        tb.localOptimizeTSGD(tw, iters, lr_schedule_mode=mode, use_cuda_graph=use_graph, stochastic=stoch, overlap=overlap,
                                             graph_iterations=block)
                        torch.cuda.synchronize()
                        print("localOptimizeTSGD mode %d stochastic=%d graph=%d block=%d overlap=%d: %.3f ms per iteration (wall, %d iterations)" %
                              (mode, stoch, use_graph, block, overlap, (time.perf_counter() - t0) * 1e3 / iters, iters))

    # the batch cut into slices that advance on streams of their own (localOptimizeTSGD(parts=k)), inside captured blocks
    for mode in (3, 1):
        for parts in (1, 2, 3, 4):
            for overlap in (True, False):
                tw.copy_(traj)
                tb.localOptimizeTSGD(tw, 3, lr_schedule_mode=mode, use_cuda_graph=False, stochastic=True, parts=parts, overlap=overlap)
                iters = 2401
                tw.copy_(traj)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                tb.localOptimizeTSGD(tw, iters, lr_schedule_mode=mode, use_cuda_graph=True, stochastic=True, graph_iterations=8,
                                     parts=parts, overlap=overlap)
                torch.cuda.synchronize()
                print("localOptimizeTSGD mode %d stochastic=1 graph=1 block=8 parts=%d overlap=%d: %.3f ms per iteration (wall, %d iterations)" %
                      (mode, parts, overlap, (time.perf_counter() - t0) * 1e3 / iters, iters))


if __name__ == "__main__":
    main()
