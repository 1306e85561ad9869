"""Seeded synthetic inputs shaped like the reference's problems (host side, numpy).  Used by tests and bench.py.

Distributions follow the reference's generators: end-effector key poses as
torm/src/utils/pirl_problem_generator.cpp:28-33,102-111 samples them (interpolated lerp + slerp like
torm/src/utils/interpolator6D.cpp:87-141), box scenes like torm/script/make_scene.py:39-74, Linear-initialiser
like trajectories (torm/src/torm_optimizer.cpp:784-807 + noise)."""
from __future__ import annotations

import numpy as np

from . import fetch


def random_configs(B, seed=0, D=7):
    rng = np.random.default_rng(seed)
    q = rng.uniform(fetch.LL, fetch.UL, size=(B, 7))
    if D == 8:
        q = np.concatenate([rng.uniform(*fetch.TORSO_RANGE, size=(B, 1)), q], axis=1)
    return q


def _quat_from_rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], axis=-1)  # x y z w


def _slerp(q0, q1, t):
    d = np.sum(q0 * q1, axis=-1, keepdims=True)
    q1 = np.where(d < 0, -q1, q1)
    d = np.clip(np.abs(d), -1, 1)
    th = np.arccos(d)
    s = np.sin(th)
    small = s < 1e-6
    w0 = np.where(small, 1 - t, np.sin((1 - t) * th) / np.where(small, 1, s))
    w1 = np.where(small, t, np.sin(t * th) / np.where(small, 1, s))
    q = w0 * q0 + w1 * q1
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def random_ee_paths(B, T=100, n_key=5, seed=0, digits=None):
    """[B,T,7] target paths (x y z qx qy qz qw) in the chain-base frame; ``digits`` rounds like the 6-digit CSVs."""
    rng = np.random.default_rng(seed)
    key_p = np.stack([rng.uniform(0.6, 1.0, (B, n_key)), rng.uniform(-0.7, 0.7, (B, n_key)),
                      rng.uniform(0.4, 1.1, (B, n_key))], axis=-1) - fetch.BASE_OFFSET
    roll = rng.uniform(-np.pi, np.pi, (B, n_key))
    pitch = np.clip(rng.normal(0.298, 0.572, (B, n_key)), -2.61799387799, 2.61799387799)
    yaw = np.clip(rng.normal(0.0, 1.3, (B, n_key)), -2.09439510239, 2.09439510239)
    key_q = _quat_from_rpy(roll, pitch, yaw)
    s = np.linspace(0, n_key - 1, T)
    i0 = np.minimum(s.astype(int), n_key - 2)
    t = (s - i0)[None, :, None]
    p = (1 - t) * key_p[:, i0] + t * key_p[:, i0 + 1]
    q = _slerp(key_q[:, i0], key_q[:, i0 + 1], t)
    out = np.concatenate([p, q], axis=-1)
    return np.round(out, digits) if digits is not None else out


def random_joint_paths(B, T=100, seed=0, step=0.08):
    """[B,T,7] smooth joint-space demonstrations inside the limits (random walk with momentum)."""
    rng = np.random.default_rng(seed)
    lo, hi = fetch.LL + 0.3, fetch.UL - 0.3
    q = np.empty((B, T, 7))
    q[:, 0] = rng.uniform(lo, hi, size=(B, 7))
    v = rng.uniform(-step, step, size=(B, 7))
    for i in range(1, T):
        v = 0.9 * v + 0.1 * rng.uniform(-step, step, size=(B, 7))
        nxt = q[:, i - 1] + v
        v = np.where((nxt < lo) | (nxt > hi), -v, v)
        q[:, i] = np.clip(q[:, i - 1] + v, lo, hi)
    return q


def linear_trajectories(N, T=200, seed=0, noise=0.05):
    """[N,T,7]: per trajectory random start and goal, joint-linear interpolation + N(0, noise) on the free rows."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(fetch.LL, fetch.UL, size=(N, 1, 7))
    b = rng.uniform(fetch.LL, fetch.UL, size=(N, 1, 7))
    s = np.linspace(0, 1, T)[None, :, None]
    traj = (1 - s) * a + s * b
    traj[:, 1:] += rng.normal(0, noise, size=(N, T - 1, 7))
    return traj


def random_box_scene(shape_zyx=(128, 128, 128), origin=(-0.64, -1.28, -0.20), res=0.02, seed=0, n_clutter=None):
    """Occupancy uint8 [nz,ny,nx]: a table box in front of the robot plus 0-3 clutter boxes on it
    (sizes and placement ranges of torm/script/make_scene.py:39-74).  Returns (occ, boxes)."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape_zyx
    occ = np.zeros(shape_zyx, dtype=np.uint8)
    boxes = []
    tz = rng.uniform(0.55, 0.75)
    boxes.append(((0.8, 0.0, tz / 2), (rng.uniform(0.5, 0.7), rng.uniform(1.2, 1.9), tz)))
    if n_clutter is None:
        n_clutter = int(rng.integers(0, 4))
    for _ in range(n_clutter):
        sx, sy, sz = rng.uniform(0.05, 0.2), rng.uniform(0.05, 0.2), rng.uniform(0.1, 0.4)
        boxes.append(((rng.uniform(0.6, 1.0), rng.uniform(-0.6, 0.6), tz + sz / 2), (sx, sy, sz)))
    xs = origin[0] + res * np.arange(nx)
    ys = origin[1] + res * np.arange(ny)
    zs = origin[2] + res * np.arange(nz)
    for (cx, cy, cz), (sx, sy, sz) in boxes:
        mx = np.abs(xs - cx) <= sx / 2
        my = np.abs(ys - cy) <= sy / 2
        mz = np.abs(zs - cz) <= sz / 2
        occ[np.ix_(mz, my, mx)] = 1
    return occ, boxes
