"""The reference's own initial trajectories (torm/launch/config/initial_traj/<prob>-<method>-0-0.csv, outputs of its
Linear and Greedy initialisers; fixtures in tests/golden/init_traj.npz) pin the host-side rules of torm_init.py:

  Linear  (`torm_jli`): every row is fillInLinearInterpolation(start, goal)              (torm_optimizer.cpp:784-807, 1021-1109)
  Greedy  (`torm`):     linear between the simplified points, every 10th pose + the last (torm_main.cpp:135-146, :939-1018)
  and the rows at those points are IK solutions of the target poses there (FK + KDL::diff of the oracle vanish).
CPU only; CSV precision (6 decimals) sets the tolerances."""
import numpy as np
import pytest
import torch

PROBLEMS = ("hello", "zig", "rotation", "square")


def _fill(start, goal, num):
    try:
        from rlitg_b200 import torm_init as TI
    except OSError:
        pytest.skip("librlitg.so not built")
    return TI.fill_in_linear_interpolation(torch.tensor(start), torch.tensor(goal[None]), num)[0].numpy(), TI


@pytest.mark.parametrize("prob", PROBLEMS)
def test_linear_initialiser_output_is_the_fill_rule(init_traj, prob):
    tr = init_traj[prob + "_torm_jli_test"]
    T = tr.shape[0]
    seg, _ = _fill(tr[0], tr[-1], T - 1)
    assert seg.shape == tr.shape and np.abs(seg - tr).max() < 2e-5


@pytest.mark.parametrize("prob", PROBLEMS)
def test_greedy_initialiser_output_is_piecewise_linear_at_the_simplified_points(init_traj, prob):
    tr = init_traj[prob + "_torm_test"]
    T = tr.shape[0]
    _, TI = _fill(tr[0], tr[-1], T - 1)
    sp = TI.simplified_points(T, 10)
    assert sp[-1] == T - 1 and sp[0] == 10
    s0 = 0
    for g in sp:
        seg, _ = _fill(tr[s0], tr[g], g - s0)
        d = seg - tr[s0:g + 1]
        d[:, [2, 4, 6]] = (d[:, [2, 4, 6]] + np.pi) % (2 * np.pi) - np.pi   # continuous joints may be stored a turn apart
        assert np.abs(d).max() < 2e-5, (prob, s0, g)
        s0 = g
    # and NOT linear across a simplified point (otherwise the check above would be vacuous)
    seg, _ = _fill(tr[0], tr[sp[1]], sp[1])
    assert np.abs(seg - tr[:sp[1] + 1]).max() > 1e-3


@pytest.mark.parametrize("prob", PROBLEMS)
def test_initialiser_rows_at_simplified_points_solve_the_targets(init_traj, prob):
    from oracle import oracle as O
    path = init_traj[prob + "_path"]
    for method, rows in (("torm_test", None), ("torm_jli_test", [path.shape[0] - 1])):
        tr = init_traj[prob + "_" + method]
        if rows is None:
            rows = list(range(10, tr.shape[0], 10))
            if rows[-1] != tr.shape[0] - 1:
                rows.append(tr.shape[0] - 1)
        tw = O.pose_diff(tr[[0] + rows], path[[0] + rows])
        assert np.linalg.norm(tw[:, :3], axis=1).max() < 2e-5               # TRAC-IK tolerance 1e-5 + CSV rounding
        # `rotation` stores unnormalised 6-digit quaternions: KDL::Rotation::Quaternion is then not orthonormal
        assert np.linalg.norm(tw[:, 3:], axis=1).max() < (2e-3 if prob == "rotation" else 3e-5)
