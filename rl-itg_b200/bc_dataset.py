"""Batched behaviour-cloning dataset builder: `pirl_gen_dataset` of
pirl/script/imitation_learning/data_loader.py:66-141 for many demonstrations at once.

For every demonstration (forwards and reversed), every gap in `gaplist` and every node i the reference appends
    X = [q_i (7) | 8 link poses as (p, R[:,0], R[:,1]) (72) | t future target errors relative to the EE (9 t)]
    Y = wrapped(q[min(i + gap, n-1)] - q_i)
with future target k = target_poses[min(i + k*gap, n-1)], k = 1..t (t = 6: 133-d observations), after one
`allfk_srv` round trip per node.  Here all rows of all demonstrations go through ONE `rlitg_pirl_obs` launch: each
row is presented as a tiny path problem of t+1 poses (row 0 unused, rows 1..t = its gathered future targets) at
timestep 0, which makes the env's own target indexing min(timestep + 1 + k, max_timestep) pick exactly those poses.
The index bookkeeping (gathers, the wrapped difference) is torch plumbing; the arithmetic is the env kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, fetch
from ._lib import check, lib
from .pirl_env import make_cfg, _code


def refine_continuous_joints(configs: np.ndarray) -> np.ndarray:
    """data_loader.py:8-20 on an [n,7] array (continuous joints folded into (-pi, pi] once)."""
    q = np.array(configs, dtype=np.float64, copy=True)
    two_pi = 2 * np.pi
    for j in (2, 4, 6):
        v = q[:, j]
        buf = v - np.trunc(v / two_pi) * two_pi
        hi, lo = v > np.pi, v < -np.pi
        buf_hi = np.where(buf > np.pi, buf - two_pi, buf)
        buf_lo = np.where(buf < -np.pi, buf + two_pi, buf)
        q[:, j] = np.where(hi, buf_hi, np.where(lo, buf_lo, v))
    return q


def _row_indices(lengths, t, gaplist, reverse):
    """Global row numbers (into the concatenated demonstrations) of the current node, its t future targets and its
    action target, in the reference's append order: reverse pass -> demonstration -> gap -> node."""
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
    cur, fut, act = [], [], []
    ks = np.arange(1, t + 1, dtype=np.int64)
    for rev in ((0, 1) if reverse else (0,)):
        for s, n in zip(starts, lengths):
            i = np.arange(n, dtype=np.int64)
            for gap in gaplist:
                f = np.minimum(i[:, None] + ks[None, :] * gap, n - 1)   # :118-123
                a = np.minimum(i + gap, n - 1)                          # :128-130
                if rev:  # the lists were reversed in place (:95-98): local index l -> n-1-l
                    cur.append(s + (n - 1 - i)); fut.append(s + (n - 1 - f)); act.append(s + (n - 1 - a))
                else:
                    cur.append(s + i); fut.append(s + f); act.append(s + a)
    return np.concatenate(cur), np.concatenate(fut), np.concatenate(act)


def pirl_gen_dataset(configs_list, target_poses_list, t=6, gaplist=(1, 2, 3), reverse=True, device="cuda:0",
                     dtype=torch.float32):
    """configs_list[k] [n_k,7], target_poses_list[k] [n_k,7] (x y z qx qy qz qw) -> DATA_X [M, 79 + 9 t],
    DATA_Y [M,7] on `device`, M = (2 if reverse else 1) * len(gaplist) * sum(n_k), rows in the reference's order."""
    device = torch.device(device)
    lengths = np.array([len(c) for c in configs_list], dtype=np.int64)
    assert all(len(p) == n for p, n in zip(target_poses_list, lengths)), "configs / target poses length mismatch"
    Q = torch.as_tensor(np.concatenate([refine_continuous_joints(c) for c in configs_list]), dtype=dtype, device=device)
    P = torch.as_tensor(np.concatenate([np.asarray(p, dtype=np.float64) for p in target_poses_list]), dtype=dtype,
                        device=device)
    cur, fut, act = _row_indices(lengths, int(t), tuple(gaplist), reverse)
    cur_t = torch.as_tensor(cur, device=device)
    fut_t = torch.as_tensor(fut, device=device)
    act_t = torch.as_tensor(act, device=device)
    M = cur.shape[0]
    q = Q[cur_t].contiguous()
    paths = torch.empty(M, t + 1, 7, dtype=dtype, device=device)
    paths[:, 1:] = P[fut_t]
    paths[:, 0] = paths[:, 1]  # never read (timestep 0 looks at rows 1..t)
    ts = torch.zeros(M, dtype=torch.int32, device=device)
    mts = torch.full((M,), t, dtype=torch.int32, device=device)
    cfg = make_cfg(obsType="Conf+LinkPoses", t=int(t), zDim=0, mode="test")
    X = torch.empty(M, int(lib.rlitg_pirl_obs_dim(C.byref(cfg))), dtype=dtype, device=device)
    with _lib.on_device(device):
        check(lib.rlitg_pirl_obs(C.byref(cfg), M, _code(q), q.data_ptr(), ts.data_ptr(), mts.data_ptr(),
                                 paths.data_ptr(), t + 1, None, X.data_ptr(),
                                 torch.cuda.current_stream(device).cuda_stream))
    d = Q[act_t] - q                                           # :131-138: single wrap on the continuous joints
    cj = torch.tensor(fetch.CONTINUOUS, dtype=torch.bool, device=device)
    d = torch.where(cj & (d > np.pi), d - 2 * np.pi, torch.where(cj & (d < -np.pi), d + 2 * np.pi, d))
    return X, d
