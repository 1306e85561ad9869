#!/usr/bin/env python
"""Which term makes the FP32 collision increments lose digits?  FP32 kernel vs FP64 oracle on the same inputs while
switching the pseudo-inverse, the affine-axis quirk and dt (1/dt^2 scales the stencil acceleration)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    from oracle import oracle as O
    from rlitg_b200 import fetch, scene as S, synth, torm as TM
    dev = torch.device("cuda:0")
    N, T = 12, 100
    for seed in (107, 7919 * 5 + 107):
        traj = synth.linear_trajectories(N, T, seed=seed, noise=0.03).astype(np.float32).astype(np.float64)
        targets = np.round(synth.random_ee_paths(1, T, seed=seed + 1)[0], 6)
        occ, _ = synth.random_box_scene((64, 96, 64), (-0.3, -0.96, 0.0), 0.025, seed=seed, n_clutter=3)
        link, center, radius = fetch.default_spheres()
        osc = O.Scene(O.build_field(occ, 0.025, 1.0), (-0.3, -0.96, 0.0), 0.025, link, center, radius)
        gsc = S.Scene(dev)
        gsc.set_occupancy(occ, (-0.3, -0.96, 0.0), 0.025, 1.0)
        for dt in (0.1, 1.0):
            for pinv in (True, False):
                for quirk in (True, False):
                    op = O.torm_params(T, dt=dt, clearance=0.3, affine_axis_quirk=quirk, use_pinv=pinv)
                    want = O.torm_cost_grad(traj, targets, osc, op, None, O.STAGE_COLL)["inc_coll"]
                    pr = TM.TormParameters(time_duration_=dt, min_clearence_=0.3, affine_axis_quirk=quirk, use_pseudo_inverse_=pinv)
                    out = {}
                    for name, dtp in (("f32", torch.float32), ("f64", torch.float64)):
                        tb = TM.TormBatch(targets, pr, gsc, dev, dtp)
                        got = tb.calculateCollisionIncrements(torch.tensor(traj, dtype=dtp, device=dev)).double().cpu().numpy()
                        err = np.abs(got - want).max(-1) / (np.abs(want).max(-1) + 1e-6)
                        out[name] = (np.median(err), (err > 1e-4).mean(), err.max())
                    print("seed %6d dt %.1f pinv %d quirk %d | FP32 med %.1e frac>1e-4 %.3f max %.1e | FP64 max %.1e" %
                          (seed, dt, pinv, quirk, *out["f32"], out["f64"][2]))
        gsc.close()


if __name__ == "__main__":
    main()
