"""CPU tests: the oracle (oracle/rlitg_oracle.c) against the reference's own data fixtures
and against self-consistency properties for the quantities no reference artefact pins."""
import numpy as np
import pytest

from oracle import oracle as O
from conftest import random_configs


def rot_err_angle(R, Rt):
    """angle of R^T Rt by the well-conditioned atan2 form"""
    rel = np.einsum("bij,bik->bjk", R, Rt)
    a = np.stack([rel[:, 2, 1] - rel[:, 1, 2], rel[:, 0, 2] - rel[:, 2, 0], rel[:, 1, 0] - rel[:, 0, 1]], 1)
    r = np.linalg.norm(a, axis=1)
    f = (np.trace(rel, axis1=1, axis2=2) - 1) / 2
    return np.arctan2(r / 2, f)


def test_hello_known_answer(fk_pairs):
    # pirl/launch/cfg/target_path/hello.yaml:2-3: FK(init_config) = (0.8, 0.45, 0.25), identity rotation
    out = O.fk_quat(fk_pairs["hello_init_q"][None])[0]
    raw = fk_pairs["hello_init_pose_raw"]
    assert np.allclose(raw, [0.8, 0.45, 0.25, 0, 0, 0, 1])
    assert np.abs(out[:3] - raw[:3]).max() < 1e-9
    assert np.abs(out[3:6]).max() < 1e-9 and abs(out[6] - 1) < 1e-12


@pytest.mark.parametrize("key,ptol,rtol", [("named", 2e-5, 1e-3), ("rand", 2e-5, 5e-5), ("obs", 2e-5, 2.5e-3)])
def test_fk_fixture_pairs(fk_pairs, key, ptol, rtol):
    # start_conf/*.csv rows are IK solutions (TRAC-IK/KDL) of the path poses: CSV precision 6 digits; the rotation
    # tolerance is the IK solver's own residual (zig / random_obs rows reach 1e-3 .. 2e-3 rad)
    q, pose = fk_pairs[key + "_q"], fk_pairs[key + "_pose"]
    pos, R = O.fk(q)
    assert np.linalg.norm(pos - pose[:, :3], axis=1).max() < ptol
    assert rot_err_angle(R, O.quat_to_rot(pose[:, 3:])).max() < rtol


def test_allfk_consistent_with_fk():
    q = random_configs(257)
    a = O.allfk(q)
    f = O.fk_quat(q)
    assert np.abs(a[:, 7] - f).max() < 1e-14
    # segment 0 (shoulder_pan_link) sits at the shoulder_pan_joint origin
    assert np.abs(a[:, 0, :3] - [0.119525, 0, 0.34858]).max() < 1e-15


@pytest.mark.parametrize("D", [7, 8])
def test_jacobian_vs_finite_difference(D):
    q = random_configs(64, D, seed=3)
    J = O.jacobian(q)
    h = 1e-6
    for j in range(D):
        dq = np.zeros(D); dq[j] = h
        pp, Rp = O.fk(q + dq)
        pm, Rm = O.fk(q - dq)
        lin = (pp - pm) / (2 * h)
        assert np.abs(lin - J[:, :3, j]).max() < 1e-7
        # angular: (Rp Rm^T) ~ I + [w] 2h
        W = np.einsum("bij,bkj->bik", Rp, Rm)
        w = np.stack([W[:, 2, 1] - W[:, 1, 2], W[:, 0, 2] - W[:, 2, 0], W[:, 1, 0] - W[:, 0, 1]], 1) / (4 * h)
        assert np.abs(w - J[:, 3:, j]).max() < 1e-7


def test_torso_column_and_offset():
    q8 = random_configs(32, 8, seed=5)
    p8, R8 = O.fk(q8)
    p7, R7 = O.fk(q8[:, 1:])
    off = np.array([-0.086875, 0.0, 0.37743])
    assert np.abs(p8 - (p7 + off + np.outer(q8[:, 0], [0, 0, 1]))).max() < 1e-14
    J8 = O.jacobian(q8)
    assert np.abs(J8[:, :, 0] - [0, 0, 1, 0, 0, 0]).max() == 0
    assert np.abs(J8[:, :, 1:] - O.jacobian(q8[:, 1:])).max() < 1e-14


def test_pinv_properties():
    q = random_configs(128, seed=7)
    J = O.jacobian(q)
    rng = np.random.default_rng(1)
    v = rng.normal(size=(128, 6))
    dq = O.pinv_step(q, v)
    ref = np.stack([np.linalg.pinv(J[i], rcond=0) @ v[i] for i in range(128)])
    good = O.manipulability(q) > 1e-3
    assert good.sum() > 100
    assert np.abs(dq[good] - ref[good]).max() < 1e-9
    # J J+ v = v for full row rank
    assert np.abs(np.einsum("bij,bj->bi", J[good], dq[good]) - v[good]).max() < 1e-9
    m = np.array([np.sqrt(abs(np.linalg.det(J[i] @ J[i].T))) for i in range(128)])
    assert np.allclose(m, O.manipulability(q), rtol=1e-9, atol=1e-12)


def test_pinv_truncation_at_singularity():
    # elbow straight + wrist aligned: rank-deficient J; truncated SVD (eps 1e-5) must stay bounded
    q = np.zeros((1, 7))
    v = np.array([[0.01, 0.02, -0.01, 0.1, -0.05, 0.02]])
    dq = O.pinv_step(q, v)
    assert np.isfinite(dq).all() and np.abs(dq).max() < 10
    J = O.jacobian(q)[0]
    ref = np.linalg.pinv(J, rcond=1e-5 / np.linalg.svd(J, compute_uv=False)[0]) @ v[0]
    assert np.abs(dq[0] - ref).max() < 1e-9


def test_nullspace_norm():
    q = random_configs(64, seed=11)
    err = np.random.default_rng(2).normal(size=(64, 7))
    out = O.nullspace_norm(err, q)
    J = O.jacobian(q)
    for i in range(64):
        N = np.eye(7) - J[i].T @ np.linalg.inv(J[i] @ J[i].T) @ J[i]
        assert abs(out[i] - np.linalg.norm(N @ err[i])) < 1e-8 * max(1, out[i])


def test_quat_roundtrip_and_rot_angle():
    rng = np.random.default_rng(0)
    qt = rng.normal(size=(200, 4)); qt /= np.linalg.norm(qt, axis=1, keepdims=True)
    R = O.quat_to_rot(qt)
    back = O.rot_to_quat(R)
    sign = np.sign(np.sum(back * qt, axis=1, keepdims=True))
    assert np.abs(back * sign - qt).max() < 1e-12
    ang, ax = O.rot_angle(np.eye(3))
    assert ang == 0 and np.allclose(ax, [0, 0, 1])
    ang, ax = O.rot_angle(np.diag([1.0, -1.0, -1.0]))
    assert abs(ang - np.pi) < 1e-15 and np.allclose(np.abs(ax), [1, 0, 0])


def test_smoothness_known_answer():
    # torm/src/torm_optimizer.cpp:235 `// 96.77624`: max coeff of A_free^-1 for the `square` path
    # (T = 320), dt = 1.0 (torm/src/pirl_datagen.cpp:129), velocity weight only
    Aband, scale, _ = O.smooth_setup(320, 1.0, (1.0, 0.0, 0.0), 0.0)
    assert abs(scale - 96.77624) < 2e-5
    # symmetric, 13-banded, interior rows sum to zero (before ridge)
    G = 330
    A = np.zeros((G, G))
    for i in range(G):
        for k in range(13):
            j = i + k - 6
            if 0 <= j < G:
                A[i, j] = Aband[i, k]
    assert np.abs(A - A.T).max() < 1e-9
    assert np.abs(A[20:300].sum(axis=1)).max() < 1e-9


def test_smoothness_inverse_scaling():
    T = 40
    Aband, scale, Ainv = O.smooth_setup(T, 0.1, (1.0, 0.0, 0.0), 0.0, want_inverse=True)
    assert abs(Ainv.max() - 1.0) < 1e-12  # after scaling max(A^-1) = 1
    G = T + 10
    A = np.zeros((G, G))
    for i in range(G):
        for k in range(13):
            j = i + k - 6
            if 0 <= j < G:
                A[i, j] = Aband[i, k]
    Af = A[6:6 + T - 1, 6:6 + T - 1]
    assert np.abs(Af @ Ainv - np.eye(T - 1)).max() < 1e-8


ANCHORS = {  # SURVEY.md 8c regression anchors (survey-computed, not reference-published): weighted pose cost
    "hello": dict(torm_test=0.008976, torm_jli_test=0.301658, pirl_rl_test_0=0.028859, pirl_bc_test=0.025048),
    "zig": dict(torm_test=0.020445, torm_jli_test=0.373526, pirl_rl_test_0=0.029367, pirl_bc_test=0.026999),
    "rotation": dict(torm_test=0.011282, torm_jli_test=0.072687, pirl_rl_test_0=0.038744, pirl_bc_test=0.030672),
    "square": dict(torm_test=0.003403, torm_jli_test=0.258495, pirl_rl_test_0=0.030088),
}


@pytest.mark.parametrize("prob", list(ANCHORS))
def test_traj_evaluator_anchors(init_traj, prob):
    path = init_traj[prob + "_path"]
    p = O.torm_params(path.shape[0], dt=0.1)
    for method, want in ANCHORS[prob].items():
        ev = O.traj_eval(init_traj["%s_%s" % (prob, method)][None], path, p)[0]
        assert abs(ev[0] + 0.17 * ev[1] - want) < 2e-6, (prob, method, ev)
    if prob == "hello":
        ev = O.traj_eval(init_traj["hello_torm_test"][None], path, p)[0]
        assert np.allclose(ev[:5], [0.007168, 0.010636, 0.17986, 0.2681, 4.037], rtol=2e-4)
