#!/usr/bin/env python
"""Smallest run of every kernel, for compute-sanitizer (one tool per gpurun call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rlitg_b200 import fetch, kinematics as K, pirl_env as P, scene as S, synth, torm as TM, torm_init as TI

dev = torch.device("cuda:0")
for dt in (torch.float32, torch.float64):
    for B in (1, 131, 128 * 1024 + 5):  # tile-per-CTA, partial tile, persistent + tail
        q = torch.tensor(synth.random_configs(B, seed=1), dtype=dt, device=dev)
        K.fk_jac(q, want_quat=True); K.fk_jac(q.t().contiguous(), q_soa=True, out_soa=True); K.allfk(q)
    q8 = torch.tensor(synth.random_configs(200, seed=2, D=8), dtype=dt, device=dev)
    K.fk_jac(q8); K.allfk(q8)
    occ, _ = synth.random_box_scene((40, 48, 40), (-0.2, -0.7, 0.0), 0.035, seed=1, n_clutter=2)
    sc = S.Scene(dev); sc.set_occupancy(occ, (-0.2, -0.7, 0.0), 0.035, 1.0)
    sc.collide(torch.tensor(synth.random_configs(300, seed=3), dtype=dt, device=dev))
    for T, N in ((5, 2), (37, 3), (300, 2)):
        traj = torch.tensor(synth.linear_trajectories(N, T, seed=4), dtype=dt, device=dev)
        tb = TM.TormBatch(synth.random_ee_paths(1, T, seed=5)[0], TM.TormParameters(min_clearence_=0.3), sc, dev, dt)
        r = tb.cost_grad(traj, 2 | 4 | 8 | 16, want_active=True)
        tb.cost_grad(traj, 8, stochastic_idx=np.zeros(N, np.int32))
        tb.applyIncrements(traj, True, r["inc_smooth"], r["inc_coll"], None, 0.5, 0.01)
        tb.applyIncrements(traj, False, inc_pose=r["inc_pose"])
    B, Tp = 200, 12
    demos = synth.random_joint_paths(B, Tp, seed=6)
    pos, quat = K.fk(torch.tensor(demos.reshape(-1, 7), dtype=dt, device=dev))
    env = P.PirlVecEnv(B, scene=sc, device=dev, dtype=dt, want_errs=True, zDim=3, t=2)
    env.set_problems(torch.cat([pos, quat], 1).reshape(B, Tp, 7), demos, np.full(B, Tp - 1, np.int32))
    env.reset()
    for _ in range(3):
        env.step(env.random_action())
    P.pinv_step(q8, torch.zeros(200, 6, dtype=dt, device=dev) + 0.01)
    ti = TI.TrajectoryInitializer(synth.random_ee_paths(1, 25, seed=7)[0], TM.TormParameters(min_clearence_=0.3), sc, dev, dt)
    ti.getNewTrajectory(demos[0, 0], [synth.random_configs(6, seed=k) for k in range(len(ti.simplified_points_))])
    sc.close()
torch.cuda.synchronize()
print("sanitize run ok")
