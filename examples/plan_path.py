#!/usr/bin/env python
"""End-to-end use of the hot path on a reference-format problem, everything on one B200:

    path file (x,y,z;w,x,y,z per line)  ->  start configurations by batched IK  ->  N Linear-initialiser trajectories
    -> iterativeExploration (lock-step TSGD over the N restarts)  ->  best trajectory + its evaluator costs

    python examples/plan_path.py [path_file] [--restarts 256] [--iters 60] [--mode 0|1|2|3]

Default problem: the head of the reference's `square` path (tests/golden/raw/path_square_head.txt) and its IK
solutions of the first pose (tests/golden/raw/square_config.csv), free space.  This mirrors what torm_main does for
one optimiser (torm/src/torm_main.cpp:227-262) with the restarts as the batch axis."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from rlitg_b200 import io, torm as TM, torm_init as TI
    raw = os.path.join(ROOT, "tests", "golden", "raw")
    ap = argparse.ArgumentParser()
    ap.add_argument("path_file", nargs="?", default=os.path.join(raw, "path_square_head.txt"))
    ap.add_argument("--start-configs", default=os.path.join(raw, "square_config.csv"))
    ap.add_argument("--restarts", type=int, default=256)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--mode", type=int, default=3, help="lr_schedule_mode_ of TormOptimizer (0..3)")
    a = ap.parse_args()
    # under torchrun every rank plans its own shard of the restarts; the only exchange is the arg-min of the best cost
    # and a broadcast of the winning trajectory
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    say = print if rank == 0 else (lambda *x, **k: None)

    poses = io.load_path(a.path_file)                       # [T,7] x y z qx qy qz qw
    T = poses.shape[0]
    starts = io.load_configs(a.start_configs)               # IK solutions of pose 0 shipped with the reference
    say("problem: %d target poses, %d start configurations on file" % (T, starts.shape[0]))

    # goal configurations: batched IK of the last pose from random seeds (TRAC-IK's role in the reference)
    t0 = time.perf_counter()
    goals, ok = TI.dls_ik_candidates(poses[-1], 4 * a.restarts, seed=rank, device=dev)
    goals = goals[ok][: a.restarts]
    say("IK: %d / %d seeds converged (%.1f ms)" % (int(ok.sum()), ok.numel(), 1e3 * (time.perf_counter() - t0)))
    N = goals.shape[0]

    pr = TM.TormParameters(exploration_iter_=a.iters)
    init = TI.TrajectoryInitializer(poses, pr, None, dev)
    rng = np.random.default_rng(rank)
    trajs = []
    for n in range(N):                                      # Linear initialiser per restart (one start each)
        st = starts[rng.integers(0, starts.shape[0])]
        tr, _, _ = init.getJointInterpolatedTrajectory(st, goals[n:n + 1])
        trajs.append(tr)
    trajs = torch.stack(trajs)

    tb = TM.TormBatch(poses, pr, None, dev)
    c0 = tb.getWeightedPoseCost(trajs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = tb.iterativeExploration(trajs, lr_schedule_mode=a.mode)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    k = out["argmin"]
    if world > 1:
        from rlitg_b200 import dist as D
        best, gid, win = D.broadcast_best(out["best_traj"], out["best_cost"], offset=rank * N)
        say("%d ranks x %d restarts: global best cost %.6f from restart %d (rank %d)" % (world, N, best, gid, gid // N))
        assert win.shape == out["best_traj"][0].shape
    e = tb.evaluate(out["best_traj"][k:k + 1].contiguous())
    ev = [float(e["cost_pose"][0]), float(e["cost_rot"][0]), float(e["cost_vel"][0])]
    say("TSGD: %d restarts x %d iterations in %.1f ms; feasible %d / %d" % (N, a.iters, 1e3 * dt, int(out["success"].sum()), N))
    say("initial weighted pose cost: best %.5f  median %.5f" % (float(c0.min()), float(c0.median())))
    say("best restart %d (this rank): weighted pose cost %.6f  (position %.6f m, rotation %.6f rad, velocity %.4f)" %
          (k, float(out["best_cost"][k]), ev[0], ev[1], ev[2]))
    ok_any = bool(out["success"].any())
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0 if ok_any else 1


if __name__ == "__main__":
    sys.exit(main())
