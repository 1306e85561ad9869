"""GPU parity: fused PIRL env step (kernel 2), null-space norm, pinv step, pose diff vs the CPU oracle.

Stated tolerances.  FP64 build: 1e-10 on everything.  FP32: observation 1e-5 relative (floor 2e-6); dist_err
2e-6; rot_err compared through cos(r/2) to 5e-7 (r = 2 acos|w| is ill-conditioned near 0: SURVEY section 7);
reward 2e-4 + the propagated rot_err / null-space sensitivities; null-space objective 1e-6 + 1.5 eps32 cond(J) |e|
(FP64: 1e-12 + 0.3 eps64 cond(J)^2 |e|); flags identical except where an error sits within tolerance of its threshold
(counted)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def build_case(B, T, seed, with_scene, cuda_device):
    from oracle import oracle as O
    from rlitg_b200 import fetch, synth
    rng = np.random.default_rng(seed)
    demos = synth.random_joint_paths(B, T, seed=seed)
    noisy = demos + rng.normal(0, 0.01, size=demos.shape)
    targets = np.round(O.fk_quat(noisy.reshape(-1, 7)).reshape(B, T, 7), 6)  # 6-digit poses like the demo CSVs
    mts = rng.integers(12, T, size=B).astype(np.int32)
    ts = np.array([rng.integers(0, m) for m in mts], dtype=np.int32)
    ts[: B // 16] = mts[: B // 16] - 1  # some envs time out on this step
    q = demos[np.arange(B), ts] + rng.normal(0, 0.02, size=(B, 7))
    q = np.clip(q, fetch.LL + 1e-3, fetch.UL - 1e-3)
    act = rng.uniform(-0.26, 0.26, size=(B, 7)).astype(np.float32)
    act[B // 2:: 7] *= 6.0  # some joint-limit violations / wraps
    q[B // 3:: 11, 2] = 3.1  # continuous joints near the seam
    osc = gsc = None
    if with_scene:
        from rlitg_b200 import scene as S
        occ, _ = synth.random_box_scene((64, 96, 64), (-0.2, -0.96, 0.0), 0.02, seed=seed, n_clutter=3)
        link, center, radius = fetch.default_spheres()
        osc = O.Scene(O.build_field(occ, 0.02, 1.0), (-0.2, -0.96, 0.0), 0.02, link, center, radius)
        gsc = S.Scene(cuda_device)
        gsc.set_occupancy(occ, (-0.2, -0.96, 0.0), 0.02, 1.0)
    return dict(demos=demos, targets=targets, mts=mts, ts=ts, q=q, act=act, osc=osc, gsc=gsc)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("reward_type,obs_type,t,z,with_scene,refine,B", [
    ("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0, False, False, 3001),  # the trained setting (88-d obs): the fast-path kernel
    ("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0, False, False, 16384),  # BASELINE config 3: the benchmarked batch
    ("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0, True, False, 3001),   # same with a scene: the row is re-assembled after a collision
    ("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 32, True, False, 3001),  # obs scenes: 120-d obs + collision
    ("Task+Imitation", "Conf+EE", 3, 0, True, False, 3001),
    ("Imitation", "onlyConf", 2, 4, False, False, 3001),
    ("Task", "onlyLinkPoses", 1, 0, False, True, 3001),
])
def test_env_step_parity(cuda_device, dtype, reward_type, obs_type, t, z, with_scene, refine, B):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import pirl_env as P
    T = 40
    c = build_case(B, T, 5, with_scene, cuda_device)
    rng = np.random.default_rng(9)
    latent = rng.normal(size=(B, z)) if z else None
    for mode in ("train", "test"):
        ocfg = O.pirl_cfg(reward_type, obs_type, t, z, train_mode=(mode == "train"), refine_flag=refine)
        want = O.pirl_env_step(ocfg, c["q"], c["act"], c["ts"], c["mts"], c["targets"], c["demos"], latent, c["osc"])
        env = P.PirlVecEnv(B, reward_type, obs_type, t, z, refine, c["gsc"], cuda_device, getattr(torch, dtype),
                           want_errs=True)
        env.set_problems(c["targets"], c["demos"], c["mts"], latent)
        env.reset(start_conf=c["q"], mode=mode)
        env.timestep.copy_(torch.tensor(c["ts"], device=cuda_device))
        obs, rew, done, info = env.step(torch.tensor(c["act"], device=cuda_device))
        torch.cuda.synchronize()
        f64 = dtype == "float64"
        g = dict(obs=obs.double().cpu().numpy(), rew=rew.double().cpu().numpy(), done=done.cpu().numpy(),
                 flags=env.flags.cpu().numpy(), errs=env.errs.double().cpu().numpy(),
                 q=env.cur_conf.double().cpu().numpy(), ts=env.timestep.cpu().numpy(), cnt=env.counters.cpu().numpy())
        assert np.array_equal(g["ts"], want["timestep"])
        assert want["obs"].shape == g["obs"].shape == (B, env.obs_dim)
        # flags can legitimately differ when an error is within tolerance of a threshold, or (FP32) when a sphere
        # centre sits on a voxel boundary; such envs are counted and excluded
        tol = 1e-10 if f64 else 3e-6
        e = want["errs"]
        near = (np.abs(e[:, 0] - 0.2) < tol) | (np.abs(e[:, 0] - 0.04) < tol) | (np.abs(e[:, 0] - 0.08) < tol) | \
               (np.abs(e[:, 1] - 0.104) < 30 * tol) | (np.abs(e[:, 2] - 0.087) < tol)
        flag_mism = (g["flags"] != want["flags"]).any(axis=1)
        bad = flag_mism & ~near
        assert bad.sum() <= (0 if f64 else 3 * (1 + B // 4096)), (int(bad.sum()), np.nonzero(bad)[0][:5])
        ok = ~flag_mism
        assert ok.mean() > 0.995
        if mode == "train":
            assert want["flags"][:, 1].sum() > 10  # joint-limit rejections exercised
            if with_scene:
                assert want["flags"][:, 0].sum() > 10  # collisions exercised
        assert np.abs(g["q"][ok] - want["q"][ok]).max() < (1e-12 if f64 else 1e-6)  # two FP32 roundings at |q| ~ pi
        rel, floor = (1e-10, 1e-12) if f64 else (1e-5, 2e-6)
        assert (np.abs(g["obs"][ok] - want["obs"][ok]) / (floor + rel * np.abs(want["obs"][ok]))).max() <= 1.0
        assert np.abs(g["errs"][ok, 0] - e[ok, 0]).max() < (1e-12 if f64 else 2e-6)
        assert np.abs(np.cos(g["errs"][ok, 1] / 2) - np.cos(e[ok, 1] / 2)).max() < (1e-12 if f64 else 5e-7)
        assert np.abs(g["errs"][ok, 2] - e[ok, 2]).max() < (1e-12 if f64 else 1e-6)
        dnull = np.zeros(B)
        if not refine:
            # null-space objective |(I - J^+ J) e|: the FP32 build projects with a QR of J^T (error ~ cond(J) eps32 |e|),
            # the FP64 build with a Cholesky solve of J J^T (~ cond^2 eps64 |e|); measured constants 0.6 and 0.08
            # (tools/ns_cond.py), asserted with margin.  |e|_2 <= |e|_1 = 7 * imitate_err.
            sv = np.linalg.svd(O.jacobian(want["q"]), compute_uv=False)
            cond = sv[:, 0] / sv[:, -1]
            enorm = 7.0 * e[:, 2]
            ntol = (1e-12 + 0.3 * 1.1e-16 * cond ** 2 * enorm) if f64 else (1e-6 + 1.5 * 6e-8 * cond * enorm)
            dnull = np.abs(g["errs"][:, 3] - e[:, 3])
            assert (dnull[ok] <= ntol[ok]).all(), (dnull[ok] / ntol[ok]).max()
            assert np.median(dnull[ok]) < (1e-12 if f64 else 2e-6)
        # reward: recompute the sensitivity of the rot-error and null-space terms (|d/dx| <= 10 and <= 16)
        drot = np.abs(g["errs"][:, 1] - e[:, 1])
        rtol = (1e-9 if f64 else 2e-4) + 10.0 * drot + 16.0 * dnull
        assert (np.abs(g["rew"][ok] - want["reward"][ok]) <= rtol[ok]).all()
        assert np.array_equal(g["done"][ok], want["done"][ok])
        assert np.array_equal(g["cnt"][ok], want["counters"][ok])


def test_reset_observation_and_scalar_wrapper(cuda_device):
    from oracle import oracle as O
    from rlitg_b200 import pirl_env as P
    c = build_case(8, 30, 2, False, cuda_device)
    env = P.PirlEnvFetch(device=cuda_device)
    assert env.observation_space_shape == (88,) and env.action_space_shape == (7,)
    T = int(c["mts"][0]) + 1
    obs = env.reset(c["targets"][0, :T], c["demos"][0, :T])
    ocfg = O.pirl_cfg()
    want = O.pirl_obs(ocfg, c["demos"][:1, 0], [0], [T - 1], c["targets"][:1, :T])
    assert np.abs(obs - want[0]).max() < 1e-12
    assert env.max_timestep == T - 1 and env.timestep == 0
    q = c["demos"][0, 0].copy()
    ts = 0
    for k in range(5):
        a = env.random_action().astype(np.float32)
        o, r, d, info = env.step(a)
        w = O.pirl_env_step(ocfg, q[None], a[None], [ts], [T - 1], c["targets"][:1, :T], c["demos"][:1, :T])
        q, ts = w["q"][0], int(w["timestep"][0])
        assert np.abs(o - w["obs"][0]).max() < 1e-10 and abs(r - w["reward"][0]) < 1e-9 and d == bool(w["done"][0])
        assert set(info) == {"task_suc_count", "dist_suc_count", "rot_suc_count", "imitate_suc_count", "is_col",
                             "is_jlimit", "is_early_stop", "is_timeout"}
        assert env.timestep == ts


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 2e-5)])
def test_nullspace_pinv_posediff(cuda_device, dtype, tol):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import pirl_env as P, synth
    B = 5000
    q = synth.random_configs(B, seed=21)
    rng = np.random.default_rng(3)
    err = rng.normal(0, 0.3, size=(B, 7))
    tw = rng.normal(0, 0.05, size=(B, 6))
    tgt = np.round(O.fk_quat(q + rng.normal(0, 0.05, size=q.shape)), 6)
    dt = getattr(torch, dtype)
    tq = torch.tensor(q, dtype=dt, device=cuda_device)
    man = O.manipulability(q)
    good = man > 1e-2  # pinv conditioning: exclusions counted (SURVEY 8d tolerances)
    assert good.mean() > 0.85
    ns = P.nullspace_norm(torch.tensor(err, dtype=dt, device=cuda_device), tq).double().cpu().numpy()
    sv = np.linalg.svd(O.jacobian(q), compute_uv=False)
    cond = sv[:, 0] / sv[:, -1]
    enorm = np.linalg.norm(err, axis=1)
    ntol = (1e-12 + 0.3 * 1.1e-16 * cond ** 2 * enorm) if dtype == "float64" else (1e-6 + 1.5 * 6e-8 * cond * enorm)
    dn = np.abs(ns - O.nullspace_norm(err, q))
    assert (dn <= ntol).all(), (dn / ntol).max()           # every configuration, conditioning-aware
    assert dn[good].max() < tol * 25                        # and a flat bound away from singularities
    dq = P.pinv_step(tq, torch.tensor(tw, dtype=dt, device=cuda_device)).double().cpu().numpy()
    want = O.pinv_step(q, tw)
    scale = np.abs(want).max(axis=1, keepdims=True) + 1e-3
    assert (np.abs(dq - want) / scale)[good].max() < (1e-8 if dtype == "float64" else 2e-4)
    pd = P.pose_diff(tq, torch.tensor(tgt, dtype=dt, device=cuda_device)).double().cpu().numpy()
    assert np.abs(pd - O.pose_diff(q, tgt)).max() < (1e-10 if dtype == "float64" else 3e-6)
    # D = 8 path
    q8 = synth.random_configs(512, seed=5, D=8)
    tw8 = rng.normal(0, 0.05, size=(512, 6))
    dq8 = P.pinv_step(torch.tensor(q8, dtype=dt, device=cuda_device),
                      torch.tensor(tw8, dtype=dt, device=cuda_device)).double().cpu().numpy()
    w8 = O.pinv_step(q8, tw8)
    g8 = O.manipulability(q8) > 1e-2
    assert (np.abs(dq8 - w8) / (np.abs(w8).max(axis=1, keepdims=True) + 1e-3))[g8].max() < (1e-8 if dtype == "float64" else 2e-4)


def test_pinv_singular_configuration(cuda_device):
    """Rank-deficient J (arm stretched out): KDL's truncated SVD must be reproduced, not a blown-up solve."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import pirl_env as P
    q = np.zeros((4, 7)); q[1, 1] = 0.3; q[2, 3] = 1e-7; q[3, 5] = 1e-9
    tw = np.tile(np.array([[0.01, 0.02, -0.01, 0.1, -0.05, 0.02]]), (4, 1))
    want = O.pinv_step(q, tw)
    for dtype, tol in (("float64", 1e-8), ("float32", 1e-3)):
        dt = getattr(torch, dtype)
        got = P.pinv_step(torch.tensor(q, dtype=dt, device=cuda_device),
                          torch.tensor(tw, dtype=dt, device=cuda_device)).double().cpu().numpy()
        assert np.isfinite(got).all() and np.abs(got - want).max() < tol, (dtype, got, want)


def test_policy_rollout_matches_cpu_loop(cuda_device):
    """batched PirlInterpolator::interpolate with a stand-in torch policy, FP64, vs a CPU loop built
    from the oracle's env step in test mode."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import pirl_env as P
    B, T = 64, 25
    c = build_case(B, T, 13, False, cuda_device)
    torch.manual_seed(0)
    W = (torch.randn(88, 7, dtype=torch.float64) * 0.02)
    policy = lambda obs: torch.tanh(obs.double() @ W.to(obs.device)) * 0.26
    env = P.PirlVecEnv(B, device=cuda_device, dtype=torch.float64)
    mts = np.full(B, T - 1, dtype=np.int32)
    env.set_problems(c["targets"], c["demos"], mts)
    got = P.rollout_policy(env, policy, n_steps=T - 1).cpu().numpy()
    ocfg = O.pirl_cfg(train_mode=False)
    q = c["demos"][:, 0].copy(); ts = np.zeros(B, dtype=np.int32)
    obs = O.pirl_obs(ocfg, q, ts, mts, c["targets"])
    for k in range(T - 1):
        act = (np.tanh(obs.astype(np.float32).astype(np.float64) @ W.numpy()) * 0.26).astype(np.float32)
        r = O.pirl_env_step(ocfg, q, act, ts, mts, c["targets"], c["demos"])
        q, ts, obs = r["q"], r["timestep"], r["obs"]
        assert np.abs(got[:, k] - q).max() < 1e-6, k  # float32 observations feed the policy on both sides


def test_step_cuda_graph_replay(cuda_device):
    """The fused step is capturable (no allocation / sync inside): graph replays equal eager steps."""
    import torch
    from rlitg_b200 import pirl_env as P
    B, T = 512, 30
    c = build_case(B, T, 17, False, cuda_device)
    mts = np.full(B, T - 1, dtype=np.int32)
    envs = []
    for _ in range(2):
        e = P.PirlVecEnv(B, device=cuda_device, dtype=torch.float32)
        e.set_problems(c["targets"], c["demos"], mts)
        e.reset()
        envs.append(e)
    act = torch.zeros(B, 7, device=cuda_device)
    g, out = envs[0].capture_step_graph(act)
    envs[0].reset()  # capture ran the step twice (warm-up + capture): start both envs from the same state
    gen = torch.Generator(device="cpu"); gen.manual_seed(0)
    for k in range(5):
        a = (torch.rand(B, 7, generator=gen) * 0.52 - 0.26).to(cuda_device)
        act.copy_(a)
        g.replay()
        o2, r2, d2, _ = envs[1].step(a)
        assert torch.equal(out[0], o2) and torch.equal(out[1], r2) and torch.equal(out[2], d2)
        assert torch.equal(envs[0].timestep, envs[1].timestep)


def test_stepping_finished_episodes_stays_in_bounds(cuda_device):
    """The reference raises IndexError when a finished episode is stepped; the kernel clamps the path index instead
    (memory-safety guard).  Run far past the end of every path on a batch whose last path ends the buffer."""
    import torch
    from rlitg_b200 import pirl_env as P
    B, T = 4096, 14
    c = build_case(B, T, 19, False, cuda_device)
    env = P.PirlVecEnv(B, device=cuda_device, dtype=torch.float32)
    env.set_problems(c["targets"], c["demos"], np.full(B, T - 1, dtype=np.int32))
    env.reset()
    act = torch.zeros(B, 7, device=cuda_device)
    for _ in range(5 * T):
        obs, rew, done, info = env.step(act)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert int(env.timestep.min().item()) == 5 * T


@pytest.mark.parametrize("D", [7, 8])
def test_manipulability_parity(cuda_device, D):
    """checkSingularity's sqrt(det(J J^T)) (torm/src/torm_optimizer.cpp:758)"""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import pirl_env as P, synth
    q = synth.random_configs(4000, seed=31, D=D)
    q[0, -7:] = 0.0  # a singular configuration
    want = O.manipulability(q)
    for dtype, rel in ((torch.float64, 1e-9), (torch.float32, 2e-4)):
        got = P.manipulability(torch.tensor(q, dtype=dtype, device=cuda_device)).double().cpu().numpy()
        assert np.abs(got - want).max() <= rel * want.max() + (1e-12 if dtype == torch.float64 else 1e-6)
    assert want[0] < 1e-6


@pytest.mark.parametrize("setting", [("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0), ("Task+Imitation", "Conf+EE", 3, 0)])
def test_table_layouts_agree(cuda_device, setting):
    """The step reads the target / demo tables in three layouts (RlitgPirlCfg.row_floats): packed [B,Tmax,7], 32-byte rows
    [B,Tmax,8] and the 128-byte step records [B,Tmax,32] PirlVecEnv builds.  Same values in, so the outputs agree to FP32
    rounding (the three template instantiations contract their multiply-adds differently: 2e-6); episodes that end on this
    step exercise the records' end-of-path branch."""
    import copy
    import ctypes as C
    import torch
    from rlitg_b200 import _lib, pirl_env as P
    reward_type, obs_type, t, z = setting
    B, T = 2000, 30
    c = build_case(B, T, 11, False, cuda_device)
    env = P.PirlVecEnv(B, reward_type, obs_type, t, z, False, None, cuda_device, torch.float32, want_errs=True)
    env.set_problems(c["targets"], c["demos"], c["mts"])
    assert env.cfg.row_floats == 32
    q0 = torch.tensor(c["q"], dtype=torch.float32, device=cuda_device)
    ts0 = torch.tensor(c["ts"], device=cuda_device)
    act = torch.tensor(c["act"], device=cuda_device)
    tp = env.target_poses.contiguous()      # [B,T,7] packed copies of what the records hold
    dc = env.demo_configs.contiguous()

    def run(row_floats):
        q, ts = q0.clone(), ts0.clone()
        cnt = torch.zeros(B, 4, dtype=torch.int32, device=cuda_device)
        obs = torch.empty(B, env.obs_dim, device=cuda_device); rew = torch.empty(B, device=cuda_device)
        done = torch.empty(B, dtype=torch.uint8, device=cuda_device); flags = torch.empty(B, 8, dtype=torch.uint8, device=cuda_device)
        errs = torch.empty(B, 4, device=cuda_device)
        cfg = copy.copy(env.cfg)
        if row_floats == 32:
            tgt, dem = env._records, None
        elif row_floats == 8:
            tgt = torch.zeros(B, T, 8, device=cuda_device); tgt[..., :7] = tp
            dem = torch.zeros(B, T, 8, device=cuda_device); dem[..., :7] = dc
        else:
            tgt, dem = tp, dc
        cfg.row_floats = row_floats
        with _lib.on_device(cuda_device):
            _lib.check(_lib.lib.rlitg_pirl_env_step(C.byref(cfg), B, _lib.F32, q.data_ptr(), act.data_ptr(), ts.data_ptr(),
                                                    env.max_timestep.data_ptr(), tgt.data_ptr(),
                                                    None if dem is None else dem.data_ptr(), T, None, None, None,
                                                    obs.data_ptr(), rew.data_ptr(), done.data_ptr(), flags.data_ptr(),
                                                    cnt.data_ptr(), errs.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return [x.cpu().numpy() for x in (q, ts, obs, rew, done, flags, cnt, errs)]

    ref = run(32)
    assert (c["ts"] + 1 >= c["mts"]).sum() > 50   # end-of-path steps are in the batch
    for rf in (8, 7):
        got = run(rf)
        for k, (a, b) in enumerate(zip(ref, got)):
            if a.dtype.kind == "f":
                assert np.abs(a.astype(np.float64) - b).max() < (2e-4 if k == 3 else 2e-6), (rf, k)   # k = 3: the reward
            else:
                assert (a != b).sum() <= 2, (rf, k)   # a flag / counter may flip when an error sits on its threshold
