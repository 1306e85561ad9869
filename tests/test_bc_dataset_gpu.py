"""Batched BC dataset builder (rl-itg_b200/bc_dataset.py) against a literal CPU replay of
pirl/script/imitation_learning/data_loader.py:66-141 built on the oracle's all-link FK.
Tolerances: FP64 1e-10; FP32 observation 1e-5 relative (floor 2e-6), actions 1e-6."""
import numpy as np
import pytest


def _as_rotation_matrix(w, x, y, z):
    # numpy-quaternion as_rotation_matrix (normalising form)
    n = w * w + x * x + y * y + z * z
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def _calc_pos_and_rot(src):  # data_loader.py:22-32
    M = _as_rotation_matrix(src[6], src[3], src[4], src[5])
    return [src[0], src[1], src[2], M[0, 0], M[1, 0], M[2, 0], M[0, 1], M[1, 1], M[2, 1]]


def _calc_pos_and_rot_error(target, src):  # data_loader.py:34-48
    tM = _as_rotation_matrix(target[6], target[3], target[4], target[5])
    eM = _as_rotation_matrix(src[6], src[3], src[4], src[5])
    R = tM @ np.linalg.inv(eM)
    return [target[0] - src[0], target[1] - src[1], target[2] - src[2], R[0, 0], R[1, 0], R[2, 0], R[0, 1], R[1, 1], R[2, 1]]


def reference_loop(configs_list, target_list, t, gaplist, reverse):
    from oracle import oracle as O
    from rlitg_b200.bc_dataset import refine_continuous_joints
    X, Y = [], []
    for rev in ((0, 1) if reverse else (0,)):
        for configs, targets in zip(configs_list, target_list):
            configs = [list(r) for r in refine_continuous_joints(configs)]
            targets = [list(r) for r in np.asarray(targets, dtype=np.float64)]
            poses = [list(r.reshape(-1)) for r in O.allfk(np.array(configs))]
            if rev:
                configs.reverse(); targets.reverse(); poses.reverse()
            n = len(configs)
            for gap in gaplist:
                for i in range(n):
                    xi = list(configs[i])
                    for s in range(8):
                        xi += _calc_pos_and_rot(poses[i][7 * s:7 * s + 7])
                    for f in range(i + gap, i + t * gap + 1, gap):
                        xi += _calc_pos_and_rot_error(targets[min(f, n - 1)], poses[i][49:56])
                    ti = min(i + gap, n - 1)
                    yi = []
                    for k in range(7):
                        d = configs[ti][k] - configs[i][k]
                        if k in (2, 4, 6):
                            if d > np.pi:
                                d -= 2 * np.pi
                            elif d < -np.pi:
                                d += 2 * np.pi
                        yi.append(d)
                    X.append(xi); Y.append(yi)
    return np.array(X), np.array(Y)


def make_demos(seed):
    from oracle import oracle as O
    from rlitg_b200 import synth
    rng = np.random.default_rng(seed)
    configs, targets = [], []
    for k, n in enumerate((4, 5, 7, 11, 3)):  # the reference's demos have 4..7 waypoints (data_loader.py:56-63)
        q = synth.random_joint_paths(1, n, seed=seed + k, step=0.3)[0]
        if k == 1:
            q[:, 4] += 5.0   # continuous joint several radians out -> refine_continuous_joints
        if k == 2:
            q[:, 2] = np.linspace(2.9, 3.6, n)  # crosses +pi after refinement -> wrapped action
        configs.append(q)
        targets.append(np.round(O.fk_quat(q + rng.normal(0, 0.01, size=q.shape)), 6))
    return configs, targets


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("t,gaplist,reverse", [(6, (1, 2, 3), True), (1, (1,), False), (3, (2, 5), True)])
def test_bc_dataset_matches_reference_loop(cuda_device, dtype, t, gaplist, reverse):
    import torch
    from rlitg_b200 import bc_dataset
    configs, targets = make_demos(3)
    X, Y = bc_dataset.pirl_gen_dataset(configs, targets, t=t, gaplist=gaplist, reverse=reverse, device=cuda_device,
                                       dtype=getattr(torch, dtype))
    Xr, Yr = reference_loop(configs, targets, t, gaplist, reverse)
    assert X.shape == Xr.shape == (Xr.shape[0], 79 + 9 * t) and Y.shape == Yr.shape
    if t == 6 and reverse:
        assert X.shape[0] == 2 * 3 * sum(len(c) for c in configs) and X.shape[1] == 133  # "R^133" (:104)
    X, Y = X.cpu().numpy().astype(np.float64), Y.cpu().numpy().astype(np.float64)
    if dtype == "float64":
        np.testing.assert_allclose(X, Xr, rtol=0, atol=1e-10)
        np.testing.assert_allclose(Y, Yr, rtol=0, atol=1e-12)
    else:
        assert np.all(np.abs(X - Xr) <= 1e-5 * np.abs(Xr) + 2e-6), np.abs(X - Xr).max()
        np.testing.assert_allclose(Y, Yr, rtol=0, atol=1e-6)
    assert np.abs(Yr).max() < np.pi + 1e-9 and np.any(np.abs(Yr[:, 2]) > 0)


def test_row_order_and_refinement_cpu():
    """Index bookkeeping without a GPU: order rev -> demo -> gap -> node, clamping at the last node, reversal."""
    from rlitg_b200.bc_dataset import _row_indices, refine_continuous_joints
    cur, fut, act = _row_indices(np.array([3, 2]), 2, (1, 2), True)
    assert cur.tolist()[:6] == [0, 1, 2, 0, 1, 2] and cur.tolist()[6:10] == [3, 4, 3, 4]
    assert fut[:3].tolist() == [[1, 2], [2, 2], [2, 2]] and fut[3:6].tolist() == [[2, 2], [2, 2], [2, 2]]
    assert act[:6].tolist() == [1, 2, 2, 2, 2, 2]
    half = cur.shape[0] // 2
    assert cur[half:half + 3].tolist() == [2, 1, 0] and fut[half].tolist() == [1, 0] and act[half + 2] == 0
    q = np.zeros((4, 7)); q[:, 2] = [3.5, -3.5, 7.0, -10.0]; q[:, 0] = 9.0
    r = refine_continuous_joints(q)
    np.testing.assert_allclose(r[:, 2], [3.5 - 2 * np.pi, -3.5 + 2 * np.pi, 7.0 - 2 * np.pi, -10.0 + 4 * np.pi - 0.0], atol=1e-12)
    assert np.all(r[:, 0] == 9.0)
