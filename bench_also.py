"""Secondary workloads of bench.py: the TORM cost + gradient iteration (BASELINE.json configs[3], the second half of
the metric), the PIRL env step (configs[2]) and the strong split of configs[4].  Every leg returns a compact record
(value, roofline, cpu_baseline, e2e) that bench.py nests under the keys the driver keeps, plus details for the line."""
import ctypes as C
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GRID = ((128, 128, 128), (-0.64, -1.28, -0.20), 0.02)


def time_loop(torch, fn, steps, warmup, flush=None, blocks=1):
    """ms per call, CUDA events on the current stream.  blocks > 1: the `steps`-call block is timed `blocks` times and the
    median block is reported (short blocks move by several per cent from one to the next).  With `flush`, every call is
    timed on its own and an L2 flush (a write larger than L2) runs between calls outside the timed spans."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        res = []
        for _ in range(blocks):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / steps)
        return sorted(res)[len(res) // 2]
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def pipe_peaks(torch, dev):
    """FP32 / FP64 FMA-pipe peaks measured on this GPU (rlitg_fma_probe / rlitg_dfma_probe), TFLOP/s"""
    from rlitg_b200 import _lib
    lib = _lib.lib
    lib.rlitg_fma_probe.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.rlitg_dfma_probe.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    of, od = torch.zeros(1, device=dev), torch.zeros(1, device=dev, dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    ctas = 148 * 8
    f32 = f64 = 0.0
    for _ in range(4):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        _lib.check(lib.rlitg_fma_probe(ctas, 4096, of.data_ptr(), st))
        e1.record()
        _lib.check(lib.rlitg_dfma_probe(ctas, 2048, od.data_ptr(), st))
        e2.record()
        torch.cuda.synchronize()
        f32 = max(f32, 2.0 * ctas * 256 * 4096 * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        f64 = max(f64, 2.0 * ctas * 256 * 2048 * 8 / (e1.elapsed_time(e2) * 1e-3) / 1e12)
    return f32, f64


def profile_traffic(name):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/traffic.json), or None"""
    import json
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(name)
    except Exception:  # noqa: BLE001
        return None


# ----------------------------------------------------------------------------------------------------------- TORM
def torm_problem(torch, dev, N, T, seed):
    from rlitg_b200 import scene as S, synth, torm as TM
    traj = torch.tensor(synth.linear_trajectories(N, T, seed=seed, noise=0.05), dtype=torch.float32, device=dev)
    targets = synth.random_ee_paths(1, T, seed=7)[0]
    occ, _ = synth.random_box_scene(GRID[0], GRID[1], GRID[2], seed=0)
    sc = S.Scene(dev)
    sc.set_occupancy(occ, GRID[1], GRID[2], 1.0)
    tb = TM.TormBatch(targets, TM.TormParameters(min_clearence_=0.3), sc, dev, torch.float32)
    return traj, targets, occ, sc, tb


def run_torm(torch, dev, rank, world, max_over_ranks, barrier, steps, warmup, fma_peak, cpu=True, N=1024, T=200):
    from rlitg_b200 import fetch, torm as TM
    traj, targets, occ, sc, tb = torm_problem(torch, dev, N, T, 100 + rank)
    full = TM.STAGE_POSE_GRAD | TM.STAGE_SMOOTH | TM.STAGE_COLL | TM.STAGE_EVAL
    feas = TM.STAGE_POSE | TM.STAGE_SMOOTH | TM.STAGE_COLL | TM.STAGE_EVAL   # feasible stage of localOptimizeTSGD
    pose = TM.STAGE_POSE_GRAD | TM.STAGE_EVAL                                 # end-pose stage
    r = tb.cost_grad(traj, full, want_active=True)
    torch.cuda.synchronize()
    A = float(r["active"].mean().item())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    barrier()
    ms_full = max_over_ranks(time_loop(torch, lambda: tb.cost_grad(traj, full), steps, warmup, blocks=3))
    ms_cold = max_over_ranks(time_loop(torch, lambda: tb.cost_grad(traj, full), min(steps, 10), 2, flush))
    ms_feas = max_over_ranks(time_loop(torch, lambda: tb.cost_grad(traj, feas), steps, warmup))
    ms_pose = max_over_ranks(time_loop(torch, lambda: tb.cost_grad(traj, pose), steps, warmup))
    ms_coll = max_over_ranks(time_loop(torch, lambda: tb.cost_grad(traj, TM.STAGE_COLL | TM.STAGE_SMOOTH), steps, warmup))
    st = torch.randint(0, T - 1, (N,), dtype=torch.int32, device=dev)
    ms_feas_st = max_over_ranks(time_loop(torch, lambda: tb.cost_grad(traj, feas, stochastic_idx=st), steps, warmup))
    del flush
    flop_wp = 2060.0 + 430.0 * A  # SURVEY 8d per-waypoint figure
    wps = world * N * (T - 1)
    tf = flop_wp * N * (T - 1) / (ms_full * 1e-3) / 1e12
    roof = {"bound": "fp32_fma", "achieved": tf, "peak": fma_peak, "unit": "TFLOP/s", "frac": tf / fma_peak if fma_peak else None,
            "traffic": profile_traffic("torm_coll_kernel"), "flop_per_waypoint": flop_wp,
            "kernel": "torm_coll_kernel<float,fk64> || torm_pose_eval_kernel<float> (two streams, one call)",
            "peak_source": "measured in this run (rlitg_fma_probe)"}
    # ---- end to end: host trajectories through rlitg_torm_cost_grad_host (H2D traj, kernels, D2H increments + costs)
    npin = lambda *s: torch.empty(*s, dtype=torch.float32).pin_memory()
    h_traj = npin(N, T, 7); h_traj.copy_(traj.cpu())
    out = {"inc_smooth": npin(N, T - 1, 7).numpy(), "inc_coll": npin(N, T - 1, 7).numpy(), "inc_pose": npin(N, T - 1, 7).numpy(),
           "costs": npin(N, TM.NCOST).numpy()}
    hv = h_traj.numpy()
    for _ in range(3):
        tb.cost_grad_host(hv, full, out=out)
    barrier()
    n_e2e = max(5, min(steps, 20))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        tb.cost_grad_host(hv, full, out=out)
    dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": wps * n_e2e / dt, "unit": "waypoints/s", "ms_per_iteration": 1e3 * dt / n_e2e,
           "h2d_bytes_per_step": N * T * 7 * 4, "d2h_bytes_per_step": 3 * N * (T - 1) * 7 * 4 + N * TM.NCOST * 4,
           "api": "rlitg_torm_cost_grad_host (pinned host buffers)"}
    # ---- CPU baseline: the FP64 oracle on all host threads, bounded sample of the same trajectories
    cb = None
    if cpu and rank == 0 and world == 1:
        from oracle import oracle as O
        O.set_num_threads(len(os.sched_getaffinity(0)))
        n_s = 128
        link, center, radius = fetch.default_spheres()
        osc = O.Scene(O.build_field(occ, GRID[2], 1.0), GRID[1], GRID[2], link, center, radius)
        op = O.torm_params(T, dt=0.1, clearance=0.3)
        Aband, _, _ = O.smooth_setup(T, 0.1, (1.0, 0.0, 0.0), 0.0)
        tj = traj[:n_s].double().cpu().numpy()
        O.torm_cost_grad(tj[:8], targets, osc, op, Aband, full)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 6.0:
            O.torm_cost_grad(tj, targets, osc, op, Aband, full); reps += 1
        dtc = time.perf_counter() - t0
        cb = {"value": reps * n_s * (T - 1) / dtc, "unit": "waypoints/s", "cores": O.num_threads(), "kind": "port",
              "sample": "%d x %d of the %d trajectories x %d waypoints, all stages, FP64 oracle, %.1f s" % (reps, n_s, N, T - 1, dtc)}
    # ---- one whole TSGD iteration: both stages, collision-point tests, bookkeeping, update steps
    tsgd = {}
    try:
        tw = traj.clone()
        for name, kw in (("stochastic", dict(stochastic=True)), ("all_waypoints", dict(stochastic=False))):
            for gname, g in (("eager", False), ("cuda_graph", True)):
                tw.copy_(traj)
                tb.localOptimizeTSGD(tw, 3, lr_schedule_mode=3, use_cuda_graph=g, **kw)   # warm-up
                tw.copy_(traj)
                n_it = 1201 if g else 30   # the graph (8 iterations per replay) is captured inside the call: amortised over 150 replays
                torch.cuda.synchronize(); barrier(); t0 = time.perf_counter()
                tb.localOptimizeTSGD(tw, n_it, lr_schedule_mode=3, use_cuda_graph=g, **kw)
                torch.cuda.synchronize()
                tsgd["ms_per_iteration_%s_%s" % (name, gname)] = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_it)
        tsgd["note"] = ("localOptimizeTSGD(30 iterations eager / 1,201 as CUDA-graph replays of 8 iterations each incl. the capture, lr_schedule_mode 3), "
                        "wall clock per iteration: two cost calls, two collision-point tests, two bookkeeping launches, two update "
                        "steps; stochastic = the reference's default (one waypoint per trajectory takes the collision increment)")
    except Exception as e:  # noqa: BLE001
        tsgd = {"error": repr(e)}
    sc.close()
    rec = {"metric": "torm_cost_grad_waypoints_per_s", "value": wps / (ms_full * 1e-3), "unit": "waypoints/s",
           "ms_per_iteration": ms_full, "target_ms": 1.0, "roofline": roof, "cpu_baseline": cb, "e2e": e2e}
    detail = {"config": {"workload": "torm_cost_grad_1024x200_grid128", "trajectories_per_gpu": N, "T": T, "spheres": 24,
                         "grid": "128^3 @ 0.02 m", "clearance": 0.3, "stages": "pose+pinv, smoothness, collision, evaluator",
                         "l2": "iterations back to back (working set ~50 MB, L2 resident as in the optimiser loop); cold = L2 flushed"},
              "mean_active_spheres": A, "ms_per_iteration_cold": ms_cold, "ms_feasible_stage": ms_feas,
              "ms_feasible_stage_stochastic": ms_feas_st, "ms_endpose_stage": ms_pose, "ms_collision_smoothness_kernel": ms_coll,
              "tsgd_iteration": tsgd}
    return rec, detail


# ----------------------------------------------------------------------------------------------------------- PIRL
def pirl_env(torch, dev, B, Tp, seed):
    from rlitg_b200 import kinematics as K, pirl_env as P, synth
    rng = np.random.default_rng(11 + seed)
    demos = synth.random_joint_paths(B, Tp, seed=5 + seed)
    dq = torch.tensor(demos.reshape(-1, 7) + rng.normal(0, 0.01, size=(B * Tp, 7)), dtype=torch.float32, device=dev)
    pos, quat = K.fk(dq)
    targets_b = torch.cat([pos, quat], dim=1).reshape(B, Tp, 7)
    env = P.PirlVecEnv(B, device=dev, dtype=torch.float32)
    env.set_problems(targets_b, demos, np.full(B, Tp - 1, dtype=np.int32))
    ts0 = torch.tensor(rng.integers(0, Tp - 2, size=B), dtype=torch.int32, device=dev)
    act = env.random_action()
    env.reset()
    q0 = torch.tensor(demos[np.arange(B), ts0.cpu().numpy()], dtype=torch.float32, device=dev)
    env.timestep.copy_(ts0); env.cur_conf.copy_(q0)
    env.max_timestep.fill_(10 ** 6)   # keep stepping: the timing loop must not run into done / timeout logic changes
    return env, act, ts0, q0, demos, targets_b


def run_pirl(torch, dev, rank, world, max_over_ranks, barrier, steps, warmup, hbm_peak, cpu=True, sizes=((16384, 100), (1048576, 16))):
    BYTES_ENV = 536.0  # read q 28 + a 28 + two target poses 56 + demo 28 + index 4; write obs 352 + q' 28 + reward 4 + flags 8
    rec, detail = None, {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B, Tp in sizes:
        env, act, ts0, q0, demos, targets_b = pirl_env(torch, dev, B, Tp, rank)
        step = lambda: env.step(act)
        barrier()
        ms = max_over_ranks(time_loop(torch, step, steps, warmup, blocks=3))
        env.timestep.copy_(ts0)
        ms_cold = max_over_ranks(time_loop(torch, step, min(steps, 10), 2, flush))
        us_graph = None
        try:
            g, _ = env.capture_step_graph(act)
            us_graph = 1e3 * max_over_ranks(time_loop(torch, g.replay, steps, warmup, blocks=3))
        except Exception:  # noqa: BLE001
            pass
        # the kernel without the launch path: 50 steps captured in ONE graph (what a captured roll-out loop pays per step)
        us_gpu = None
        try:
            env.step(act); torch.cuda.synchronize()
            g50 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g50):
                for _ in range(50):
                    env.step(act)
            us_gpu = 1e3 * max_over_ranks(time_loop(torch, g50.replay, max(steps // 5, 3), 2, blocks=3)) / 50
            del g50
        except Exception:  # noqa: BLE001
            pass
        gbs = BYTES_ENV * B / (ms * 1e-3) / 1e9
        d = {"envs_per_gpu": B, "path_poses": Tp, "us_per_step": ms * 1e3, "us_per_step_cold": ms_cold * 1e3,
             "us_per_step_cuda_graph": us_graph, "us_per_step_50_steps_per_graph": us_gpu,
             "env_steps_per_s": world * B / (ms * 1e-3), "gb_per_s": gbs, "hbm_frac": gbs / hbm_peak}
        detail["envs_%d" % B] = d
        if B == 16384:
            # ---- end to end at the BASELINE config: action H2D (pinned), step, obs / reward / done D2H (pinned)
            h_act = act.cpu().pin_memory()
            d_act = torch.empty_like(act)
            h_obs = torch.empty(B, env.obs_dim).pin_memory(); h_rew = torch.empty(B).pin_memory()
            h_done = torch.empty(B, dtype=torch.uint8).pin_memory()

            def e2e_step():
                d_act.copy_(h_act, non_blocking=True)
                o, r_, dn, _ = env.step(d_act)
                h_obs.copy_(o, non_blocking=True); h_rew.copy_(r_, non_blocking=True); h_done.copy_(dn, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            for _ in range(5):
                e2e_step()
            barrier()
            n_e2e = 200
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_step()
            dt = max_over_ranks(time.perf_counter() - t0)
            e2e = {"value": world * B * n_e2e / dt, "unit": "env-steps/s", "us_per_step": 1e6 * dt / n_e2e,
                   "h2d_bytes_per_step": B * 7 * 4, "d2h_bytes_per_step": B * (env.obs_dim * 4 + 4 + 1),
                   "api": "PirlVecEnv.step (rlitg_pirl_env_step), action from / obs+reward+done to pinned host buffers, synchronised every step"}
            cb = None
            if cpu and rank == 0 and world == 1:
                from oracle import oracle as O
                O.set_num_threads(len(os.sched_getaffinity(0)))
                cfg = O.pirl_cfg()
                qn = q0.double().cpu().numpy(); an = act.cpu().numpy(); tsn = ts0.cpu().numpy().astype(np.int32)
                mts = np.full(B, 10 ** 6, np.int32); tg = targets_b.double().cpu().numpy()
                O.pirl_env_step(cfg, qn[:64], an[:64], tsn[:64], mts[:64], tg[:64], demos[:64])
                t0 = time.perf_counter(); reps = 0
                while time.perf_counter() - t0 < 4.0:
                    O.pirl_env_step(cfg, qn, an, tsn, mts, tg, demos); reps += 1
                dtc = time.perf_counter() - t0
                cb = {"value": reps * B / dtc, "unit": "env-steps/s", "cores": O.num_threads(), "kind": "port",
                      "sample": "%d x %d envs, one step each, FP64 oracle (OpenMP), %.1f s" % (reps, B, dtc)}
            # value = the device-side rate: 50 steps per CUDA-graph replay (what a captured roll-out loop pays per step; every
            # step does all its work, the action tensor is the same).  One launch per step from Python is bound by the
            # launch path at this batch size (us_per_step) and is kept next to it.
            best_us = us_gpu if us_gpu else ms * 1e3
            rec = {"metric": "pirl_env_steps_per_s", "value": world * B / (best_us * 1e-6), "unit": "env-steps/s",
                   "value_launch": "50 steps per CUDA-graph replay" if us_gpu else "one launch per step from Python",
                   "value_python_launch": world * B / (ms * 1e-3), "us_per_step": ms * 1e3,
                   "us_per_step_cuda_graph": us_graph, "us_per_step_50_steps_per_graph": us_gpu, "cpu_baseline": cb, "e2e": e2e}
        del env, demos, targets_b
    big = detail.get("envs_1048576")
    if rec is not None:
        src = big if big else detail["envs_16384"]
        rec["roofline"] = {"bound": "hbm", "achieved": src["gb_per_s"], "peak": hbm_peak, "unit": "GB/s", "frac": src["hbm_frac"],
                           "traffic": profile_traffic("pirl_env_kernel_1M"), "bytes_per_env": BYTES_ENV,
                           "at": "%d envs (the HBM-bound size; 16,384 envs = %.1f MB is launch / latency bound: %.1f us per step)"
                                 % (src["envs_per_gpu"], 16384 * BYTES_ENV / 1e6, detail["envs_16384"]["us_per_step"]),
                           "kernel": "pirl_env_kernel<float,step,linkobs,row8,fast>"}
    detail["config"] = {"workload": "pirl_env_step", "obs": "Conf+LinkPoses t=1 z=0 (88 floats)", "reward": "Task+NullSpaceImitation",
                        "l2": "16k-env batch is L2 resident (8.8 MB); 1M envs (562 MB) is not; cold = L2 flushed"}
    return rec, detail


# ----------------------------------------------------------------------------------------- config 5: the strong split
def run_config5(torch, dev, rank, world, max_over_ranks, barrier, steps, warmup):
    """BASELINE.json configs[4]: 8M configs / 65,536 envs / 8,192 trajectories split evenly over the ranks (total work fixed).
    Whole-job throughput per workload; the driver's N = 1, 2, 4, 8 lines give the strong-scaling curve."""
    from rlitg_b200 import kinematics as K, synth, torm as TM
    out = {"split": "8M configs / 65,536 envs / 8,192 trajectories over %d rank(s)" % world}
    Bk = (1 << 23) // world
    q = torch.tensor(synth.random_configs(Bk, seed=40 + rank), dtype=torch.float32, device=dev)
    o = {"pos": torch.empty(Bk, 3, device=dev), "rot": torch.empty(Bk, 3, 3, device=dev), "jac": torch.empty(Bk, 6, 7, device=dev)}
    barrier()
    ms = max_over_ranks(time_loop(torch, lambda: K.fk_jac(q, out=o), steps, warmup, blocks=3))
    out["fk_jacobian_8M"] = {"configs_per_rank": Bk, "ms": ms, "configs_per_s": (1 << 23) / (ms * 1e-3)}
    del q, o
    Be = 65536 // world
    env, act, ts0, q0, demos, tg = pirl_env(torch, dev, Be, 100, 100 + rank)
    barrier()
    ms = max_over_ranks(time_loop(torch, lambda: env.step(act), steps, warmup, blocks=3))
    # the same 50 steps captured in one CUDA graph: at 8,192 envs per rank a launch from Python costs more than the step
    us_g = None
    try:
        torch.cuda.synchronize()
        g50 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g50):
            for _ in range(50):
                env.step(act)
        us_g = 1e3 * max_over_ranks(time_loop(torch, g50.replay, max(steps // 5, 3), 2, blocks=3)) / 50
        del g50
    except Exception:  # noqa: BLE001
        pass
    best_us = us_g if us_g else ms * 1e3
    out["pirl_65536"] = {"envs_per_rank": Be, "us": best_us, "us_python_launch": ms * 1e3, "us_50_steps_per_graph": us_g,
                         "env_steps_per_s": 65536 / (best_us * 1e-6), "env_steps_per_s_python_launch": 65536 / (ms * 1e-3)}
    del env
    Nt = 8192 // world
    traj, targets, occ, sc, tb = torm_problem(torch, dev, Nt, 200, 300 + rank)
    full = TM.STAGE_POSE_GRAD | TM.STAGE_SMOOTH | TM.STAGE_COLL | TM.STAGE_EVAL
    barrier()
    ms = max_over_ranks(time_loop(torch, lambda: tb.cost_grad(traj, full), max(3, steps // 4), 2, blocks=3))
    out["torm_8192"] = {"trajectories_per_rank": Nt, "ms": ms, "waypoints_per_s": 8192 * 199 / (ms * 1e-3)}
    sc.close()
    return out


# ----------------------------------------------------------------------- FK / Jacobian in the other layouts and sizes
def run_fk_layouts(torch, dev, rank, world, max_over_ranks, barrier, steps, warmup):
    from rlitg_b200 import fetch, kinematics as K, synth
    fk = {}
    for name, Bk, D, soa in (("soa_1M", 1 << 20, 7, True), ("aos_1M_D8_torso", 1 << 20, 8, False)):
        g = torch.Generator(device="cpu"); g.manual_seed(3 + rank)
        lo = torch.tensor(([0.0] if D == 8 else []) + list(fetch.LL), dtype=torch.float32)
        hi = torch.tensor(([0.38615] if D == 8 else []) + list(fetch.UL), dtype=torch.float32)
        qk = (lo + (hi - lo) * torch.rand(Bk, D, generator=g)).to(dev)
        if soa:
            qk = qk.t().contiguous()
            outk = {"pos": torch.empty(3, Bk, device=dev), "rot": torch.empty(9, Bk, device=dev), "jac": torch.empty(6 * D, Bk, device=dev)}
            fn = lambda: K.fk_jac(qk, q_soa=True, out_soa=True, out=outk)
        else:
            outk = {"pos": torch.empty(Bk, 3, device=dev), "rot": torch.empty(Bk, 3, 3, device=dev), "jac": torch.empty(Bk, 6, D, device=dev)}
            fn = lambda: K.fk_jac(qk, out=outk)
        barrier()
        msk = max_over_ranks(time_loop(torch, fn, steps, warmup))
        bpc = 4 * (D + 12 + 6 * D)
        fk[name] = {"us_per_step": msk * 1e3, "configs_per_s": world * Bk / (msk * 1e-3), "gb_per_s": bpc * Bk / (msk * 1e-3) / 1e9}
        del qk, outk
    qa = torch.tensor(synth.random_configs(1 << 20, seed=9 + rank), dtype=torch.float32, device=dev)
    oa = torch.empty(1 << 20, 8, 7, device=dev)
    barrier()
    msk = max_over_ranks(time_loop(torch, lambda: K.allfk(qa, out=oa), steps, warmup))
    fk["allfk_1M"] = {"us_per_step": msk * 1e3, "configs_per_s": world * (1 << 20) / (msk * 1e-3), "gb_per_s": 252 * (1 << 20) / (msk * 1e-3) / 1e9}
    return fk


def cpu_batch1():
    """BASELINE configs[0]: the reference's own batch-1 CPU path -- one 100-pose path, one thread, per waypoint FK +
    Jacobian + pose error (comparable with its recorded 0.003 ms per KDL FK call, pirl/script/module_topic_speed_test.py:72)"""
    from oracle import oracle as O
    from rlitg_b200 import synth
    O.set_num_threads(1)
    path = synth.random_ee_paths(1, 100, seed=0)[0]
    qs = synth.random_configs(99, seed=0)
    t0 = time.perf_counter(); reps = 0
    while time.perf_counter() - t0 < 1.0:
        O.fk_jac(qs); O.pose_diff(qs, path[1:]); reps += 1
    dt1 = time.perf_counter() - t0
    O.set_num_threads(len(os.sched_getaffinity(0)))
    return {"us_per_waypoint": 1e6 * dt1 / (reps * 99), "cores": 1, "kind": "port",
            "workload": "99 x (FK + Jacobian + pose error), one thread, FP64 oracle",
            "reference_recorded": "0.003 ms per KDL FK call, 1.1 ms through the ROS service (authors' machine)"}
