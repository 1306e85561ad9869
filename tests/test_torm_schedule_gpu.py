"""TormBatch.localOptimizeTSGD for every lr_schedule_mode_ (0 fixed weights, 1 adaptive + per-trajectory stage, 2 cosine,
3 fixed) against a literal scalar replay of TormOptimizer::localOptimizeTSGD (torm/src/torm_optimizer.cpp:381-598) run
one trajectory at a time on the CPU oracle.  FP64 build; the replay and the batch must take the same decisions
(stage, alpha_, learning_rate_, best-trajectory records) and end on the same trajectories to 1e-7."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_problem(N, T, seed, cuda_device, grid=(64, 96, 64), origin=(-0.3, -0.96, 0.0), res=0.025):
    from oracle import oracle as O
    from rlitg_b200 import fetch, scene as S, synth
    traj = synth.linear_trajectories(N, T, seed=seed, noise=0.03)
    targets = np.round(synth.random_ee_paths(1, T, seed=seed + 1)[0], 6)
    occ, _ = synth.random_box_scene(grid, origin, res, seed=seed, n_clutter=3)
    link, center, radius = fetch.default_spheres()
    osc = O.Scene(O.build_field(occ, res, 1.0), origin, res, link, center, radius)
    gsc = S.Scene(cuda_device)
    gsc.set_occupancy(occ, origin, res, 1.0)
    return traj, targets, osc, gsc


def replay_one(traj, targets, osc, op, Aband, Ainv, pr, maxiter, mode, T, rand=None):
    """The reference loop for ONE trajectory with scalar state, on the oracle.  rand [maxiter, 2]: the uniform draws of
    use_stochastic_descent_ (None = every free waypoint contributes to the collision increment)."""
    from oracle import oracle as O
    from rlitg_b200 import fetch
    traj = traj.copy()
    lr0 = pr.learning_rate_
    lr, lr_max, lr_min = lr0, lr0 * 1000.0, lr0 / 100.0
    alpha = 0.5
    infea_col = infea_vel = False
    best_cost, best_traj = np.inf, traj.copy()
    log = []
    limit = 0.1 if mode == 0 else pr.joint_update_limit_
    for it in range(maxiter):
        for stage in (0, 1):
            if stage == 0:
                feasible_stage = not (mode == 1 and lr < 1e-2 and not infea_col and not infea_vel)
            else:
                feasible_stage = (mode == 1 and lr == lr_max and infea_col)
            mask = (O.STAGE_SMOOTH | O.STAGE_COLL | O.STAGE_EVAL) if feasible_stage else (O.STAGE_POSE_GRAD | O.STAGE_EVAL)
            # isCurrentTrajectoryCollisionFree :695-717 runs on the same trajectory before the increments: collision_point_
            hit = O.sphere_collide(traj, osc, fetch.BASE_OFFSET)
            first = int(np.argmax(hit)) if hit.any() else -1
            st = None
            if rand is not None:  # :1502-1515: collision_point_, else (int)(u (free_end - free_start) + free_start)
                F = T - 1
                st = first - 1 if first >= 0 else min(max(int(np.float32(rand[it, stage]) * np.float32(F - 1)), 0), F - 1)
            r = O.torm_cost_grad(traj[None], targets, osc, op, Aband, mask, None if st is None or st < 0 else [st])
            if st is not None and st < 0 and r["inc_coll"] is not None:
                r["inc_coll"][:] = 0.0  # collision_point_ = start row: index -1, no increment is taken (appendix B item 4)
            cur = r["costs"][0, 6]
            infea_col = first >= 0
            infea_vel = bool(r["costs"][0, 7] != 0)
            if mode in (1, 2):
                if not infea_col and not infea_vel:
                    alpha = 0.5
                elif infea_col and not infea_vel:
                    alpha = 0.5 if alpha < 0.5 else min(alpha + 0.1, 1.0)
                elif not infea_col and infea_vel:
                    alpha = 0.5 if alpha > 0.5 else max(alpha - 0.1, 0.0)
                else:
                    alpha = 0.5
            elif mode == 3:
                alpha = 0.5
            if mode == 1:
                if not infea_col and not infea_vel:
                    lr = lr0 if lr > lr0 else max(lr / 2.0, lr_min)
                else:
                    lr = min(lr * 2.0, lr_max)
            elif mode == 2:
                period = maxiter // 2
                ph = float(it % period) / period
                lr = lr_min + 0.5 * (lr_max - lr_min) * (1 + np.cos(ph * np.pi))
            elif mode == 3:
                lr = lr0
            if best_cost > cur and not infea_col and not infea_vel:
                best_cost, best_traj = cur, traj.copy()
            log.append((int(feasible_stage), alpha, lr, infea_col, infea_vel))
            traj = O.torm_update(traj[None], Ainv, 0 if feasible_stage else 1, mode, r["inc_smooth"], r["inc_coll"],
                                 r["inc_pose"], alpha, lr, sw=pr.smoothness_update_weight_ / T,
                                 ow=pr.obstacle_update_weight_, limit=limit)[0]
    return traj, best_traj, best_cost, log


@pytest.mark.parametrize("stochastic", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_lockstep_schedule_matches_scalar_replay(cuda_device, mode, stochastic):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import torm as TM
    N, T, maxiter = 6, 30, 6
    traj, targets, osc, gsc = make_problem(N, T, 21, cuda_device)
    rng = np.random.default_rng(2)
    # a spread of situations: clean tracking starts, one far too fast (velocity violation), one through the table
    base = O.fk_quat(traj[0])
    targets = np.round(base, 6)
    for n in range(N):
        traj[n] = traj[0] + (rng.normal(0, 0.01 * (n + 1), size=(T, 7)) if n else 0.0)
        traj[n, 0] = traj[0, 0]
    traj[4, 1:, 0] += np.linspace(0, 1.4, T - 1) * np.where(np.arange(T - 1) % 2, 1, -1) * 0.2   # zig-zag: too fast
    traj[5, :, 1] += 0.9   # leans into the table from the first row on: collision_point_ = the start row (index -1)
    pr = TM.TormParameters(min_clearence_=0.3, learning_rate_=0.02, use_stochastic_descent_=stochastic)
    op = O.torm_params(T, dt=pr.time_duration_, clearance=0.3)
    Aband, _, Ainv = O.smooth_setup(T, pr.time_duration_, (1.0, 0.0, 0.0), 0.0, want_inverse=True)
    tb = TM.TormBatch(targets, pr, gsc, cuda_device, torch.float64)
    t = torch.tensor(traj, dtype=torch.float64, device=cuda_device)
    best_traj, best_cost, state = tb.localOptimizeTSGD(t, maxiter, lr_schedule_mode=mode, return_state=True, seed=5)
    final = t.cpu().numpy()
    rand = state["rand"].cpu().numpy() if stochastic else None
    assert (rand is not None) == stochastic
    seen_stage_mix = False
    for n in range(N):
        ftraj, btraj, bcost, log = replay_one(traj[n], targets, osc, op, Aband, Ainv, pr, maxiter, mode, T,
                                              None if rand is None else rand[:, :, n])
        assert np.abs(final[n] - ftraj).max() < 1e-7, (mode, n, np.abs(final[n] - ftraj).max())
        assert abs(float(state["alpha"][n]) - log[-1][1]) < 1e-12 and abs(float(state["learning_rate"][n]) - log[-1][2]) < 1e-12 * max(1, log[-1][2])
        assert bool(state["infea_col"][n]) == log[-1][3] and bool(state["infea_vel"][n]) == log[-1][4]
        if np.isfinite(bcost):
            assert abs(float(best_cost[n]) - bcost) < 1e-9 * max(1.0, bcost)
            assert np.abs(best_traj[n].cpu().numpy() - btraj).max() < 1e-7
        else:
            assert not np.isfinite(float(best_cost[n]))
        stages = [l[0] for l in log]
        seen_stage_mix |= stages[0::2] != [1] * maxiter or stages[1::2] != [0] * maxiter
    if mode == 1 and os.environ.get("RLITG_TEST_SEED", "0") == "0":  # a property of the default data set
        assert seen_stage_mix, "mode 1 should make some trajectory leave the plain feasible / end-pose alternation"
    gsc.close()


def test_iterative_exploration_restarts_as_batch(cuda_device, init_traj):
    """iterativeExploration with the restarts as the batch axis on the reference's `zig` problem: the per-optimiser best
    over two epochs equals the better of two separate TSGD runs, the early-return threshold stops after one epoch."""
    import torch
    from rlitg_b200 import torm as TM
    path = init_traj["zig_path"]
    a, b = init_traj["zig_torm_jli_test"], init_traj["zig_pirl_bc_test"]
    rng = np.random.default_rng(0)
    first = np.stack([a, b, a + np.r_[np.zeros((1, 7)), rng.normal(0, 0.02, size=(a.shape[0] - 1, 7))]])
    second = np.stack([b, a, b])
    pr = TM.TormParameters(min_clearence_=0.001, exploration_iter_=8)
    tb = TM.TormBatch(path, pr, None, cuda_device, torch.float64)
    g = lambda x: torch.tensor(x, dtype=torch.float64, device=cuda_device)
    out = tb.iterativeExploration(g(first), new_trajectories=lambda eph: g(second), max_epochs=2, lr_schedule_mode=3)
    assert out["epochs"] == 2 and bool(out["success"].all())
    _, c1 = tb.localOptimizeTSGD(g(first), 8, lr_schedule_mode=3)
    _, c2 = tb.localOptimizeTSGD(g(second), 8, lr_schedule_mode=3)
    want = torch.minimum(c1, c2)
    assert torch.allclose(out["best_cost"], want, rtol=1e-12, atol=0)
    assert out["argmin"] == int(torch.argmin(want))
    got_cost = tb.getWeightedPoseCost(out["best_traj"])
    assert torch.allclose(got_cost, out["best_cost"], rtol=1e-9)   # the recorded rows are the ones with that cost
    early = tb.iterativeExploration(g(first), new_trajectories=lambda eph: g(second), max_epochs=5, lr_schedule_mode=3,
                                    return_threshold=1.0)
    assert early["epochs"] == 1
    only_first = tb.iterativeExploration(g(first), max_epochs=5, lr_schedule_mode=3)
    assert only_first["epochs"] == 1 and torch.allclose(only_first["best_cost"], c1, rtol=1e-12)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("dtype,block", [("float32", 8), ("float64", 8), ("float32", 1), ("float32", 4)])
def test_graph_replay_equals_eager(cuda_device, mode, dtype, block):
    """localOptimizeTSGD replays captured blocks of `block` iterations (no host synchronisation inside; 11 iterations =
    one eager + full blocks + an eager remainder): bit-identical to the eager loop, for every schedule mode, including
    the per-iteration cosine rate of mode 2 and the per-iteration draws of the stochastic waypoint."""
    import torch
    from rlitg_b200 import torm as TM
    N, T, maxiter = 24, 40, 11
    traj, targets, osc, gsc = make_problem(N, T, 33, cuda_device)
    pr = TM.TormParameters(min_clearence_=0.3, learning_rate_=0.02)
    tb = TM.TormBatch(targets, pr, gsc, cuda_device, getattr(torch, dtype))
    a = torch.tensor(traj, dtype=getattr(torch, dtype), device=cuda_device)
    b = a.clone()
    bt_e, bc_e, st_e = tb.localOptimizeTSGD(a, maxiter, lr_schedule_mode=mode, return_state=True, use_cuda_graph=False)
    bt_e, bc_e = bt_e.clone(), bc_e.clone()
    st_e = {k: v.clone() for k, v in st_e.items()}
    bt_g, bc_g, st_g = tb.localOptimizeTSGD(b, maxiter, lr_schedule_mode=mode, return_state=True, use_cuda_graph=True,
                                            graph_iterations=block)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert torch.equal(bt_e, bt_g) and torch.equal(bc_e, bc_g)
    for k in st_e:
        assert torch.equal(st_e[k], st_g[k]), k
    assert float((a - torch.tensor(traj, dtype=a.dtype, device=cuda_device)).abs().max()) > 1e-3
    gsc.close()


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("dtype,parts,graph", [("float32", 2, True), ("float64", 3, True), ("float32", 2, False)])
def test_batch_slices_equal_whole(cuda_device, mode, dtype, parts, graph):
    """localOptimizeTSGD(parts=k) cuts the batch into k contiguous slices that run the loop on streams of their own (the
    latency-bound launches of one slice under the cost kernels of another): trajectories are independent, so every
    output is bit-identical to the single-slice loop -- also with a last slice that is shorter than the others."""
    import torch
    from rlitg_b200 import torm as TM
    N, T, maxiter = 40, 40, 11
    traj, targets, osc, gsc = make_problem(N, T, 34, cuda_device)
    pr = TM.TormParameters(min_clearence_=0.3, learning_rate_=0.02)
    tb = TM.TormBatch(targets, pr, gsc, cuda_device, getattr(torch, dtype))
    a = torch.tensor(traj, dtype=getattr(torch, dtype), device=cuda_device)
    b = a.clone()
    bt_e, bc_e, st_e = tb.localOptimizeTSGD(a, maxiter, lr_schedule_mode=mode, return_state=True, use_cuda_graph=False)
    bt_e, bc_e = bt_e.clone(), bc_e.clone()
    st_e = {k: v.clone() for k, v in st_e.items()}
    bt_g, bc_g, st_g = tb.localOptimizeTSGD(b, maxiter, lr_schedule_mode=mode, return_state=True, use_cuda_graph=graph,
                                            graph_iterations=4, parts=parts)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert torch.equal(bt_e, bt_g) and torch.equal(bc_e, bc_g)
    for k in st_e:
        assert torch.equal(st_e[k], st_g[k]), k
    gsc.close()
