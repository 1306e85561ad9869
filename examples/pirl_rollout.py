#!/usr/bin/env python
"""PIRL side of the hot path, end to end on one B200:

    seeded problems (target paths + demonstration configs)  ->  PirlVecEnv (B environments, one kernel per step)
    driven by a stock torch MLP policy (what pirl/script/algos/sac feeds PirlEnv_Fetch one env at a time)
    ->  env-steps/s with the policy in the loop, eager and as one CUDA graph per step
    ->  behaviour-cloning dataset of the demonstrations (pirl_gen_dataset) in one launch

    python examples/pirl_rollout.py [--envs 16384] [--steps 99]
The learner itself (SAC / BC training) is out of scope and would be stock PyTorch on these tensors."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from rlitg_b200 import bc_dataset, kinematics as K, pirl_env as P, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=99)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, T = a.envs, a.steps + 1

    demos = synth.random_joint_paths(B, T, seed=1)                       # [B,T,7] demonstration configurations
    pos, quat = K.fk(torch.tensor(demos.reshape(-1, 7), dtype=torch.float32, device=dev))
    targets = torch.cat([pos, quat], dim=1).reshape(B, T, 7)               # their end-effector poses = the target paths
    env = P.PirlVecEnv(B, rewardType="Task+NullSpaceImitation", obsType="Conf+LinkPoses", t=1, zDim=0, device=dev)
    env.set_problems(targets, demos, np.full(B, T - 1, dtype=np.int32))

    torch.manual_seed(0)
    policy = torch.nn.Sequential(torch.nn.Linear(env.obs_dim, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256),
                                 torch.nn.ReLU(), torch.nn.Linear(256, 7), torch.nn.Tanh()).to(dev)

    def act(obs):
        return 0.26 * policy(obs)                                          # action limits of pirl/launch/pirl.yaml:13-14

    with torch.no_grad():
        obs = env.reset()
        env.step(act(obs))                                                 # warm-up (cuBLAS handles, kernel attributes)
        obs = env.reset()
        ret = torch.zeros(B, device=dev)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(a.steps):
            obs, rew, done, info = env.step(act(obs))
            ret += rew
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("eager : %d envs x %d steps in %.1f ms -> %.2f G env-steps/s (policy in the loop), mean return %.2f"
              % (B, a.steps, 1e3 * dt, B * a.steps / dt / 1e9, float(ret.mean())))

        # one CUDA graph per step: policy forward + env step
        obs = env.reset()
        abuf = torch.zeros(B, 7, device=dev)
        abuf.copy_(act(env.obs)); env.step(abuf)                           # warm-up outside capture
        env.reset()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            abuf.copy_(act(env.obs))
            env.step(abuf)
        env.reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(a.steps):
            g.replay()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("graph : %d envs x %d steps in %.1f ms -> %.2f G env-steps/s" % (B, a.steps, 1e3 * dt, B * a.steps / dt / 1e9))

    # behaviour-cloning pairs of the first 2,000 demonstrations: X [M,133], Y [M,7]
    nd = min(B, 2000)
    tp = targets[:nd].cpu().numpy()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    X, Y = bc_dataset.pirl_gen_dataset(list(demos[:nd]), list(tp), device=dev)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("BC dataset: %d demonstrations -> X %s, Y %s in %.1f ms" % (nd, tuple(X.shape), tuple(Y.shape), 1e3 * dt))
    return 0


if __name__ == "__main__":
    sys.exit(main())
