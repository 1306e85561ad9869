"""GPU parity on the edge cases of the path: smallest / largest shapes, single units, boundary indices, alternate
parameter sets of the reference's mains, spheres leaving the grid."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(cuda_device, link, center, radius, shape=(40, 56, 40), origin=(0.1, -0.7, 0.1), res=0.03, seed=2):
    from oracle import oracle as O
    from rlitg_b200 import scene as S, synth
    occ, _ = synth.random_box_scene(shape, origin, res, seed=seed, n_clutter=3)
    osc = O.Scene(O.build_field(occ, res, 1.0), origin, res, link, center, radius)
    gsc = S.Scene(cuda_device, spheres=(link, center, radius))
    gsc.set_occupancy(occ, origin, res, 1.0)
    return osc, gsc


@pytest.mark.parametrize("T,N", [(2, 3), (3, 1), (4, 2), (5, 1), (6, 33), (33, 1), (65, 2), (130, 3), (257, 2), (553, 1)])
def test_torm_smallest_and_chunk_boundaries(cuda_device, T, N):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, synth, torm as TM
    link, center, radius = fetch.default_spheres()
    osc, gsc = _scene(cuda_device, link, center, radius)
    traj = synth.linear_trajectories(N, T, seed=T, noise=0.02)
    targets = np.round(synth.random_ee_paths(1, T, seed=T + 1)[0], 6)
    # T < 5: the initialisers' short segments (getNewTrajectory scores 2..10-row segments with the end-pose and collision
    # terms only, torm/src/torm_optimizer.cpp:977-984); the evaluator and the smoothness band need five rows.
    # T = 65 / 130 / 257 / 553: two, three, five and sixteen chunks = cluster sizes 2, 3, 5, 8 (the last one looping)
    mask = O.STAGE_POSE_GRAD | O.STAGE_COLL | ((O.STAGE_EVAL | O.STAGE_SMOOTH) if T >= 5 else 0)
    op = O.torm_params(T, dt=0.2, clearance=0.5)
    Aband = O.smooth_setup(T, 0.2, (1.0, 0.0, 0.0), 0.0)[0] if T >= 5 else None
    want = O.torm_cost_grad(traj, targets, osc, op, Aband, mask)
    if T >= 5:
        tb = TM.TormBatch(targets, TM.TormParameters(time_duration_=0.2, min_clearence_=0.5), gsc, cuda_device, torch.float64)
    else:
        tb = TM.TormBatch(targets, TM.TormParameters(time_duration_=0.2, min_clearence_=0.5), gsc, cuda_device, torch.float64,
                          smoothness=False)
    got = tb.cost_grad(torch.tensor(traj, dtype=torch.float64, device=cuda_device), mask)
    c = got["costs"].cpu().numpy()
    ok = np.isfinite(want["costs"])
    assert np.abs(c[ok] - want["costs"][ok]).max() < 1e-8
    for k in ("inc_smooth", "inc_coll", "inc_pose") if T >= 5 else ("inc_coll", "inc_pose"):
        sc = max(1.0, np.abs(want[k]).max())
        assert np.abs(got[k].cpu().numpy() - want[k]).max() < 1e-7 * sc, k
    gsc.close()


@pytest.mark.parametrize("S", [1, 64])
def test_torm_sphere_table_extremes_and_out_of_grid(cuda_device, S):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import synth, torm as TM
    rng = np.random.default_rng(S)
    link = np.sort(rng.integers(0, 8, size=S)).astype(np.int32)
    center = np.float32(rng.uniform(-0.1, 0.1, size=(S, 3))).astype(np.float64)
    radius = np.float32(rng.uniform(0.02, 0.08, size=S)).astype(np.float64)
    osc, gsc = _scene(cuda_device, link, center, radius, shape=(24, 24, 24), origin=(0.5, -0.3, 0.4), res=0.03)  # small grid: most spheres fall outside
    T, N = 30, 3
    traj = synth.linear_trajectories(N, T, seed=5, noise=0.02)
    targets = synth.random_ee_paths(1, T, seed=6)[0]
    op = O.torm_params(T, dt=0.1, clearance=0.3)
    want = O.torm_cost_grad(traj, targets, osc, op, None, O.STAGE_COLL)
    tb = TM.TormBatch(targets, TM.TormParameters(min_clearence_=0.3), gsc, cuda_device, torch.float64)
    got = tb.cost_grad(torch.tensor(traj, dtype=torch.float64, device=cuda_device), TM.STAGE_COLL, want_active=True)
    assert np.abs(got["costs"].cpu().numpy()[:, 5] - want["costs"][:, 5]).max() < 1e-10
    assert np.abs(got["inc_coll"].cpu().numpy() - want["inc_coll"]).max() < 1e-7 * max(1.0, np.abs(want["inc_coll"]).max())
    assert np.abs(got["active"].cpu().numpy() - want["active"]).max() < 1e-6
    gsc.close()


def test_torm_reference_parameter_sets(cuda_device):
    """min_clearence_ 0.001 (torm_main), 1.0 (torm_evaluation); J^T instead of the pseudo-inverse; stochastic
    index at both ends of the free range."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, synth, torm as TM
    link, center, radius = fetch.default_spheres()
    osc, gsc = _scene(cuda_device, link, center, radius)
    T, N = 50, 4
    traj = synth.linear_trajectories(N, T, seed=9, noise=0.02)
    targets = synth.random_ee_paths(1, T, seed=10)[0]
    tt = torch.tensor(traj, dtype=torch.float64, device=cuda_device)
    for clearance, use_pinv, st in ((0.001, True, None), (1.0, False, None), (0.3, True, [0, T - 2, 0, T - 2])):
        op = O.torm_params(T, dt=0.1, clearance=clearance, use_pinv=use_pinv)
        want = O.torm_cost_grad(traj, targets, osc, op, None, O.STAGE_COLL, st)
        tb = TM.TormBatch(targets, TM.TormParameters(min_clearence_=clearance, use_pseudo_inverse_=use_pinv), gsc,
                          cuda_device, torch.float64)
        got = tb.cost_grad(tt, TM.STAGE_COLL, st)
        assert np.abs(got["costs"].cpu().numpy()[:, 5] - want["costs"][:, 5]).max() < 1e-10
        assert np.abs(got["inc_coll"].cpu().numpy() - want["inc_coll"]).max() < 1e-7 * max(1.0, np.abs(want["inc_coll"]).max())
    gsc.close()


def test_pirl_single_env_two_pose_path(cuda_device):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import pirl_env as P, synth
    demos = synth.random_joint_paths(1, 2, seed=1)
    tg = O.fk_quat(demos.reshape(-1, 7)).reshape(1, 2, 7)
    env = P.PirlVecEnv(1, device=cuda_device, dtype=torch.float64)
    env.set_problems(tg, demos, [1])
    obs0 = env.reset().cpu().numpy().copy()
    ocfg = O.pirl_cfg()
    assert np.abs(obs0 - O.pirl_obs(ocfg, demos[:, 0], [0], [1], tg)).max() < 1e-12
    act = np.full((1, 7), 0.01, dtype=np.float32)
    obs, rew, done, info = env.step(torch.tensor(act, device=cuda_device))
    w = O.pirl_env_step(ocfg, demos[:, 0], act, [0], [1], tg, demos)
    assert np.abs(obs.cpu().numpy() - w["obs"]).max() < 1e-10 and bool(done[0].item()) and bool(info["is_timeout"][0].item())


def test_fk_limits_and_wrapped_angles(cuda_device):
    """joint values at the limits, at +-pi and several turns away (the sincos range reduction), D = 7 and 8"""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, kinematics as K
    base = np.stack([fetch.LL, fetch.UL, np.zeros(7), np.full(7, np.pi), np.full(7, -np.pi), np.full(7, 9.4247),
                     np.full(7, -12.5663), np.full(7, 1e-8)])
    for D in (7, 8):
        q = base if D == 7 else np.concatenate([np.linspace(0, 0.38615, len(base))[:, None], base], axis=1)
        pos, R, J = O.fk_jac(q)
        for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 4e-6)):
            r = K.fk_jac(torch.tensor(q, dtype=dtype, device=cuda_device))
            assert np.abs(r["pos"].double().cpu().numpy() - pos).max() < tol
            assert np.abs(r["rot"].double().cpu().numpy() - R).max() < tol
            assert np.abs(r["jac"].double().cpu().numpy() - J).max() < tol
