#!/usr/bin/env python
"""The TORM update step (torm_update_kernel) alone at the benchmark size: feasible / end-pose stage, with and without
joint-limit hits, GPU-only times from CUDA-graph replays.  Also the target of the ncu capture of that kernel
(ncu -k regex:torm_update ... python tools/prof_update.py --iters 2)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--T", type=int, default=200)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    from rlitg_b200 import fetch, scene as S, synth, torm as TM
    dev = torch.device("cuda:0")
    N, T = a.N, a.T
    traj = torch.tensor(synth.linear_trajectories(N, T, seed=100, noise=0.05), dtype=torch.float32, device=dev)
    targets = synth.random_ee_paths(1, T, seed=7)[0]
    occ, _ = synth.random_box_scene((128, 128, 128), (-0.64, -1.28, -0.20), 0.02, seed=0)
    sc = S.Scene(dev)
    sc.set_occupancy(occ, (-0.64, -1.28, -0.20), 0.02, 1.0)
    tb = TM.TormBatch(targets, TM.TormParameters(min_clearence_=0.3), sc, dev, torch.float32)
    r = tb.cost_grad(traj, TM.STAGE_SMOOTH | TM.STAGE_COLL | TM.STAGE_EVAL | TM.STAGE_POSE_GRAD)
    inc_s, inc_c, inc_p = r["inc_smooth"].clone(), r["inc_coll"].clone(), r["inc_pose"].clone()
    al = torch.full((N,), 0.5, device=dev); lr = torch.full((N,), 0.01, device=dev)
    ll = torch.tensor(fetch.LL, device=dev, dtype=torch.float32); ul = torch.tensor(fetch.UL, device=dev, dtype=torch.float32)
    inside = traj.clone()
    inside[:, :, [0, 1, 3, 5]] = torch.minimum(torch.maximum(inside[:, :, [0, 1, 3, 5]], ll[[0, 1, 3, 5]] + 0.3), ul[[0, 1, 3, 5]] - 0.3)
    tw = traj.clone()

    def graphed(fn, prep):
        prep(); fn(); torch.cuda.synchronize()
        gp, gf = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(gp):
            prep()
        with torch.cuda.graph(gf):
            prep(); fn()
        def t(g):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                g.replay()
            e0.record()
            for _ in range(a.iters):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.iters * 1e3
        return t(gf) - t(gp)

    feas = lambda: tb.applyIncrements(tw, True, inc_s, inc_c, None, al, lr, lr_schedule_mode=3)
    endp = lambda: tb.applyIncrements(tw, False, inc_pose=inc_p, alpha=al, learning_rate=lr, lr_schedule_mode=3)
    endp_small = lambda: tb.applyIncrements(tw, False, inc_pose=inc_p * 0.01, alpha=al, learning_rate=lr, lr_schedule_mode=3)
    print("update feasible                 %.1f us" % graphed(feas, lambda: tw.copy_(traj)))
    print("update feasible, no limit hits  %.1f us" % graphed(feas, lambda: tw.copy_(inside)))
    print("update end-pose                 %.1f us" % graphed(endp, lambda: tw.copy_(traj)))
    print("update end-pose, no limit hits  %.1f us" % graphed(endp_small, lambda: tw.copy_(inside)))


if __name__ == "__main__":
    main()
