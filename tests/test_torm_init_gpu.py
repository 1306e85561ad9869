"""GPU tests for the batched Greedy / Linear initialisers: the selection must equal a literal
CPU replay of TormOptimizer::getNewTrajectory / getJointInterpolatedTrajectory built from oracle pieces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cpu_fill(st, gt, num, cj):
    """fillInLinearInterpolation torm/src/torm_optimizer.cpp:784-807"""
    out = np.empty((num + 1, 7)); out[0] = st
    for i in range(7):
        diff = gt[i] - st[i]
        if cj[i] and abs(diff) > np.pi:
            diff = 2 * np.pi - abs(diff) if diff < 0 else -2 * np.pi + abs(diff)
        theta = diff / num
        for t in range(1, num + 1):
            out[t, i] = st[i] + t * theta
    return out


def cpu_greedy(start, cands, targets, pts, osc, clearance, pose_w=1.0, col_w=0.05):
    from oracle import oracle as O
    cj = [0, 0, 1, 0, 1, 0, 1]
    T = targets.shape[0]
    traj = np.empty((T, 7)); traj[0] = start
    sp0, chosen = 0, []
    for i, sp in enumerate(pts):
        best, bi, bseg = np.inf, -1, None
        for j, c in enumerate(cands[i]):
            seg = cpu_fill(traj[sp0], c, sp - sp0, cj)
            p = O.torm_params(seg.shape[0], clearance=clearance)
            r = O.torm_cost_grad(seg[None], targets[sp0:sp + 1], osc, p, None, O.STAGE_POSE | O.STAGE_COLL)["costs"][0]
            e = (r[8] + 0.17 * r[9]) / (seg.shape[0] - 1)
            if best < e:
                continue
            cost = pose_w * e + col_w * r[5]
            if j == 0 or best > cost:
                best, bi, bseg = cost, j, seg
        traj[sp0 + 1:sp + 1] = bseg[1:]
        chosen.append(bi); sp0 = sp
    return traj, chosen


def test_greedy_and_linear_match_cpu_replay(cuda_device):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, scene as S, synth, torm as TM, torm_init as TI
    T, Cn = 47, 12
    rng = np.random.default_rng(5)
    demo = synth.random_joint_paths(1, T, seed=3)[0]
    targets = np.round(O.fk_quat(demo), 6)
    occ, _ = synth.random_box_scene((64, 96, 64), (-0.3, -0.96, 0.0), 0.025, seed=2, n_clutter=3)
    link, center, radius = fetch.default_spheres()
    osc = O.Scene(O.build_field(occ, 0.025, 1.0), (-0.3, -0.96, 0.0), 0.025, link, center, radius)
    gsc = S.Scene(cuda_device); gsc.set_occupancy(occ, (-0.3, -0.96, 0.0), 0.025, 1.0)
    pts = TI.simplified_points(T)
    assert pts == [10, 20, 30, 40, 46]
    # candidates: the demonstration's own configuration plus perturbed / wrapped variants
    cands = []
    for sp in pts:
        c = demo[sp] + rng.normal(0, 0.4, size=(Cn, 7)); c[3] = demo[sp]; c[5, 2] += 2 * np.pi - 0.3
        cands.append(np.clip(c, fetch.LL - 3.2, fetch.UL + 3.2))
    pr = TM.TormParameters(min_clearence_=0.3)
    ti = TI.TrajectoryInitializer(targets, pr, gsc, cuda_device, torch.float64)
    traj, chosen, best = ti.getNewTrajectory(demo[0], cands)
    want, wchosen = cpu_greedy(demo[0], cands, targets, pts, osc, 0.3)
    assert chosen.cpu().tolist() == wchosen
    assert np.abs(traj.cpu().numpy() - want).max() < 1e-12
    # FP32 build picks the same candidates on this problem
    ti32 = TI.TrajectoryInitializer(targets, pr, gsc, cuda_device, torch.float32)
    t32, c32, _ = ti32.getNewTrajectory(demo[0], cands)
    assert c32.cpu().tolist() == wchosen and np.abs(t32.double().cpu().numpy() - want).max() < 5e-6
    # Linear: 30 goal candidates for the last pose
    goals = demo[-1] + rng.normal(0, 0.3, size=(30, 7)); goals[7] = demo[-1]
    lt, li, lb = ti.getJointInterpolatedTrajectory(demo[0], goals)
    wl, wi = cpu_greedy(demo[0], [goals], targets, [T - 1], osc, 0.3)
    assert int(li.item()) == wi[0] and np.abs(lt.cpu().numpy() - wl).max() < 1e-12
    gsc.close()


@pytest.mark.parametrize("T,gap", [(43, 10), (23, 1), (12, 10)])
def test_greedy_short_segments(cuda_device, T, gap):
    """Segments of 2..4 rows: the reference's own paths end on them (hello: 553 poses -> last segment 550..552;
    random_16: 382 poses -> 380..381) and its pirl configurations use gap = 1, where EVERY segment has two rows.  Only the
    evaluator and the smoothness band need five rows; the end-pose and collision terms are per waypoint."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, scene as S, synth, torm as TM, torm_init as TI
    Cn = 7
    rng = np.random.default_rng(T)
    demo = synth.random_joint_paths(1, T, seed=T)[0]
    targets = np.round(O.fk_quat(demo), 6)
    occ, _ = synth.random_box_scene((64, 96, 64), (-0.3, -0.96, 0.0), 0.025, seed=4, n_clutter=3)
    link, center, radius = fetch.default_spheres()
    osc = O.Scene(O.build_field(occ, 0.025, 1.0), (-0.3, -0.96, 0.0), 0.025, link, center, radius)
    gsc = S.Scene(cuda_device); gsc.set_occupancy(occ, (-0.3, -0.96, 0.0), 0.025, 1.0)
    pts = TI.simplified_points(T, gap)
    assert min(b - a for a, b in zip([0] + pts[:-1], pts)) <= 3   # at least one segment of <= 4 rows
    cands = []
    for sp in pts:
        c = demo[sp] + rng.normal(0, 0.3, size=(Cn, 7)); c[2] = demo[sp]
        cands.append(np.clip(c, fetch.LL - 3.2, fetch.UL + 3.2))
    pr = TM.TormParameters(min_clearence_=0.3)
    ti = TI.TrajectoryInitializer(targets, pr, gsc, cuda_device, torch.float64, gap=gap)
    traj, chosen, best = ti.getNewTrajectory(demo[0], cands)
    want, wchosen = cpu_greedy(demo[0], cands, targets, pts, osc, 0.3)
    assert chosen.cpu().tolist() == wchosen
    assert np.abs(traj.cpu().numpy() - want).max() < 1e-12
    gsc.close()


def test_dls_ik_candidates(cuda_device):
    """New work (not a parity target): candidates must reproduce the target pose through FK."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import synth, torm_init as TI
    q_true = synth.random_configs(1, seed=8)[0] * 0.6
    target = O.fk_quat(q_true[None])[0]
    q, ok = TI.dls_ik_candidates(target, 256, seed=1, device=cuda_device, dtype=torch.float64)
    assert ok.float().mean().item() > 0.3
    sol = q[ok].cpu().numpy()
    tw = O.pose_diff(sol, np.tile(target, (sol.shape[0], 1)))
    assert np.abs(tw[:, :3]).max() < 2e-4 and np.abs(tw[:, 3:]).max() < 2e-3
    assert np.unique(np.round(sol, 1), axis=0).shape[0] > 3  # distinct IK branches, like random restarts give


def test_set_pirl_trajectory_matches_scalar_replay(cuda_device):
    """setPIRLTrajectory (torm_optimizer.cpp:898-936) for B starts at once vs a CPU loop: oracle observation -> the same
    policy -> setAction (test mode) per simplified point, then the literal fillInLinearInterpolation."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, synth, torm_init as TI
    T, gap, B = 47, 10, 5
    poses = np.round(synth.random_ee_paths(1, T, seed=3)[0], 6)
    rng = np.random.default_rng(1)
    starts = synth.random_configs(B, seed=8) * 0.5
    W = torch.tensor(rng.normal(0, 0.05, size=(88, 7)), dtype=torch.float64, device=cuda_device)

    def policy(obs):   # any deterministic map: a fixed linear layer with tanh squashing to the action limits
        return 0.26 * torch.tanh(obs.double() @ W)

    got = TI.set_pirl_trajectory(poses, starts, policy, gap=gap, device=cuda_device, dtype=torch.float64).cpu().numpy()
    sp = TI.simplified_points(T, gap)
    path = np.concatenate([poses[:1], poses[sp]])[None]
    cfg = O.pirl_cfg("Task", "Conf+LinkPoses", 1, 0, train_mode=False)
    Wc = W.cpu().numpy()
    for b in range(B):
        q = starts[b:b + 1].copy()
        ts = np.zeros(1, dtype=np.int32); mts = np.full(1, len(sp), dtype=np.int32)
        obs = O.pirl_obs(cfg, q, ts, mts, path)
        gen = []
        for k in range(len(sp)):
            # the policy sees float32 observations (torm/src/pirl_interpolator.cpp:182) and emits float32 actions
            act = (0.26 * np.tanh(obs.astype(np.float32).astype(np.float64) @ Wc)).astype(np.float32)
            r = O.pirl_env_step(cfg, q, act, ts, mts, path, None)
            q, ts, obs = r["q"], r["timestep"], r["obs"]
            gen.append(q[0].copy())
        want = np.empty((T, 7)); want[0] = starts[b]
        s0 = 0
        for i, g in enumerate(sp):
            want[g] = gen[i]
            num = g - s0
            for j in range(7):
                diff = want[g, j] - want[s0, j]
                if fetch.CONTINUOUS[j] and abs(diff) > np.pi:
                    diff = 2 * np.pi - abs(diff) if diff < 0 else -2 * np.pi + abs(diff)
                theta = diff / num
                for tt, row in enumerate(range(s0 + 1, g + 1), start=1):
                    want[row, j] = want[s0, j] + tt * theta
            s0 = g
        assert np.abs(got[b] - want).max() < 1e-9, (b, np.abs(got[b] - want).max())
    assert np.abs(got[:, 1:] - got[:, :-1]).max() > 1e-3


def test_example_plan_path_runs(cuda_device):
    """examples/plan_path.py: reference-format path file -> batched IK -> Linear initialiser -> iterativeExploration."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "plan_path.py"), "--restarts", "32", "--iters", "25"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "best restart" in r.stdout
