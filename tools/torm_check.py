#!/usr/bin/env python
"""Parity statistics of kernel 3 (TORM cost + gradient) against the CPU oracle, FP32 and FP64, on the shapes of
tests/test_torm_gpu.py plus the benchmark shape (1,024 x 200, 128^3 grid); prints the error distribution of the FP32
collision increments (the quantity SURVEY 8d bounds at 1e-4) and the kernel timings."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch


def main():
    from oracle import oracle as O
    from rlitg_b200 import fetch, scene as S, synth, torm as TM
    from test_torm_gpu import make_problem, voxel_boundary_rows, rel_rows
    dev = torch.device("cuda:0")
    O.set_num_threads(len(os.sched_getaffinity(0)))
    mask = O.STAGE_POSE_GRAD | O.STAGE_SMOOTH | O.STAGE_COLL | O.STAGE_EVAL
    for (T, N, dt, cl, quirk, pinv) in [(100, 6, 0.1, 0.3, True, True), (37, 5, 0.2, 0.5, False, True),
                                        (64, 4, 1.0, 0.3, True, False), (300, 3, 0.1, 0.3, True, True),
                                        (553, 2, 0.1, 0.3, True, True), (7, 3, 0.1, 0.3, True, True)]:
        traj, targets, osc, gsc = make_problem(N, T, 7 + T, dev)
        op = O.torm_params(T, dt=dt, clearance=cl, affine_axis_quirk=quirk, use_pinv=pinv)
        Aband, _, _ = O.smooth_setup(T, dt, (1.0, 0.0, 0.0), 0.0)
        want = O.torm_cost_grad(traj, targets, osc, op, Aband, mask)
        pr = TM.TormParameters(time_duration_=dt, min_clearence_=cl, affine_axis_quirk=quirk, use_pseudo_inverse_=pinv)
        skip = voxel_boundary_rows(traj, (-0.3, -0.96, 0.0), 0.025)
        for name, dtp in (("f64", torch.float64), ("f32", torch.float32)):
            tb = TM.TormBatch(targets, pr, gsc, dev, dtp)
            got = tb.cost_grad(torch.tensor(traj, dtype=dtp, device=dev), mask, want_active=True)
            torch.cuda.synchronize()
            gc = got["inc_coll"].double().cpu().numpy()
            rr = rel_rows(gc, want["inc_coll"], 1e-6)
            rk = rr[~skip]
            cc = got["costs"].double().cpu().numpy()
            ce = np.abs(cc[:, 5] - want["costs"][:, 5]).max() / np.abs(want["costs"][:, 5]).max()
            se = rel_rows(got["inc_smooth"].double().cpu().numpy(), want["inc_smooth"], 1e-9).max()
            ae = np.abs(got["active"].cpu().numpy() - want["active"]).max()
            print("T %3d N %d dt %.1f pinv %d quirk %d %s | coll inc: med %.1e >1e-4 %.4f max %.1e (all rows max %.1e, skipped %d) "
                  "| coll cost rel %.1e | smooth %.1e | active diff %.1e" %
                  (T, N, dt, pinv, quirk, name, np.median(rk), (rk > 1e-4).mean(), rk.max(), rr.max(), int(skip.sum()), ce, se, ae))
        gsc.close()
    # ---- benchmark shape: all 1,024 trajectories against the oracle (FP32), then timings
    N, T = 1024, 200
    traj = synth.linear_trajectories(N, T, seed=100, noise=0.05).astype(np.float32).astype(np.float64)
    targets = synth.random_ee_paths(1, T, seed=7)[0]
    occ, _ = synth.random_box_scene((128, 128, 128), (-0.64, -1.28, -0.20), 0.02, seed=0)
    link, center, radius = fetch.default_spheres()
    t0 = time.time()
    osc = O.Scene(O.build_field(occ, 0.02, 1.0), (-0.64, -1.28, -0.20), 0.02, link, center, radius)
    op = O.torm_params(T, dt=0.1, clearance=0.3)
    Aband, _, _ = O.smooth_setup(T, 0.1, (1.0, 0.0, 0.0), 0.0)
    want = O.torm_cost_grad(traj, targets, osc, op, Aband, mask)
    print("oracle 1024 x 200: %.1f s" % (time.time() - t0))
    gsc = S.Scene(dev)
    gsc.set_occupancy(occ, (-0.64, -1.28, -0.20), 0.02, 1.0)
    tb = TM.TormBatch(targets, TM.TormParameters(min_clearence_=0.3), gsc, dev, torch.float32)
    tt = torch.tensor(traj, dtype=torch.float32, device=dev)
    got = tb.cost_grad(tt, mask, want_active=True)
    torch.cuda.synchronize()
    skip = np.zeros((N, T - 1), dtype=bool)
    rr = rel_rows(got["inc_coll"].double().cpu().numpy(), want["inc_coll"], 1e-6)
    cc = got["costs"].double().cpu().numpy()
    print("bench shape FP32 | coll inc: med %.1e >1e-4 %.4f >1e-3 %.5f max %.1e (no exclusions) | coll cost rel %.1e | active %.2f vs %.2f"
          % (np.median(rr), (rr > 1e-4).mean(), (rr > 1e-3).mean(), rr.max(),
             np.abs(cc[:, 5] - want["costs"][:, 5]).max() / np.abs(want["costs"][:, 5]).max(),
             float(got["active"].mean()), float(want["active"].mean())))
    for k in (0, 1, 2, 3, 4, 6, 8, 9):
        print("   cost[%d] rel err %.1e" % (k, np.abs(cc[:, k] - want["costs"][:, k]).max() / np.abs(want["costs"][:, k]).max()))

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    st = torch.randint(0, T - 1, (N,), dtype=torch.int32, device=dev)
    for name, m, kw in (("full", 2 | 4 | 8 | 16, {}), ("feasible", 1 | 4 | 8 | 16, {}), ("endpose", 2 | 16, {}), ("coll", 8, {}),
                        ("coll+smooth", 12, {}), ("pose_grad", 2, {}), ("eval", 16, {}), ("smooth", 4, {}),
                        ("coll stochastic", 8, {"stochastic_idx": st}), ("feasible stochastic", 1 | 4 | 8 | 16, {"stochastic_idx": st})):
        print("torm %-20s %.1f us" % (name, timeit(lambda: tb.cost_grad(tt, m, **kw))))


if __name__ == "__main__":
    main()
