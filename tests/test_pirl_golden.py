"""The PIRL env step against episodes recorded from the reference's OWN env class.

tests/golden/pirl_steps.npz is written by tests/golden/make_pirl_golden.py, which imports and steps
`PirlEnv_Fetch` (/root/reference/pirl/script/envs/PirlEnv.py) under stubs of its ROS services.  Executing the reference's
code needs a human's authorisation in this environment and both attempts (rounds 1 and 2) were refused by the sandbox
policy, so the fixture does not exist yet: these tests SKIP until it does (the ask is in the script's header and in
DESIGN.md section 2).  Once the fixture is there
  * the CPU test pins the oracle (`orc_pirl_env_step`, `orc_pirl_obs`) on the reference: observation, reward, done,
    the four flags, the four success counters and the configuration of every step of all 24 episodes, 1e-10;
  * the GPU test runs the same episodes through rlitg_pirl_env_step (FP64 build 1e-9, FP32 build as the tolerances of
    tests/test_pirl_gpu.py) -- and the parity of row A9 moves from "oracle self-consistency" to "reference-pinned".
The replay logic below is exercised without the fixture by tests/test_pirl_twin.py::test_golden_replay_plumbing, which
writes a surrogate file of the same layout from the NumPy twin."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from rlitg_b200 import fetch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "pirl_steps.npz")


def load(path=GOLDEN):
    if not os.path.exists(path):
        pytest.skip("tests/golden/pirl_steps.npz is absent: running the reference's PirlEnv.py (make_pirl_golden.py) "
                    "still needs a human's authorisation")
    return np.load(path)


def scene_of(g):
    shape = tuple(int(v) for v in g["scene_shape"])
    occ = np.unpackbits(g["scene_occ_bits"])[: int(np.prod(shape))].reshape(shape).astype(np.uint8)
    origin, res = tuple(float(v) for v in g["scene_origin"]), float(g["scene_res"])
    return occ, origin, res


def episodes(g):
    """(key prefix, rewardType, obsType, t, zDim, mode, with_scene, max_timestep) per episode"""
    for ep, m in enumerate(g["meta"]):
        yield "ep%d_" % ep, str(m[0]), str(m[1]), int(m[2]), int(m[3]), str(m[4]), bool(int(m[5])), int(m[6])


def replay_oracle(g, tol=1e-10):
    """Steps the oracle through every recorded episode with the state fed back; returns the number of steps checked."""
    occ, origin, res = scene_of(g)
    osc = O.Scene(O.build_field(occ, res, 1.0), origin, res, g["sphere_link"], g["sphere_center"], g["sphere_radius"])
    base_off = tuple(float(v) for v in g["base_off"])
    n = 0
    for key, rtype, otype, t, z, mode, with_scene, mts in episodes(g):
        cfg = O.pirl_cfg(rtype, otype, t, z, train_mode=(mode == "train"))
        demos, targets = g[key + "demos"], g[key + "targets"]
        assert targets.shape[0] - 1 == mts
        latent = np.zeros((1, z)) if z else None   # sceneType "" -> envLatent = [0.0] * z_dim (PirlEnv.py:416)
        q = demos[:1].copy()                       # reset: cur_conf = demo_configs[0] (PirlEnv.py:366-371)
        ts = np.zeros(1, np.int32)
        cnt = np.zeros((1, 4), np.int32)
        m = np.array([mts], np.int32)
        obs0 = O.pirl_obs(cfg, q, ts, m, targets[None], latent)
        assert obs0.shape[1] == g[key + "obs0"].shape[0]
        assert np.abs(obs0[0] - g[key + "obs0"]).max() < tol, key
        for k, a in enumerate(g[key + "action"]):
            w = O.pirl_env_step(cfg, q, a[None], ts, m, targets[None], demos[None], latent=latent,
                                scene=osc if with_scene else None, base_off=base_off, counters=cnt)
            assert np.abs(w["q"][0] - g[key + "conf"][k]).max() < 1e-12, (key, k)
            assert np.abs(w["obs"][0] - g[key + "obs"][k]).max() < tol, (key, k)
            assert abs(w["reward"][0] - g[key + "reward"][k]) < max(tol, 1e-9), (key, k, w["reward"][0], g[key + "reward"][k])
            assert bool(w["done"][0]) == bool(g[key + "done"][k]), (key, k)
            assert [bool(v) for v in w["flags"][0][:4]] == [bool(v) for v in g[key + "flags"][k]], (key, k)
            assert list(w["counters"][0]) == list(g[key + "counters"][k]), (key, k)
            q, ts, cnt = w["q"], w["timestep"], w["counters"]
            n += 1
    return n


def test_reference_constants():
    """limits / action bounds / observation widths as the reference's __init__ set them (PirlEnv.py:118-174,76-90)"""
    g = load()
    for si in range(4):
        assert np.array_equal(g["set%d_ll" % si], fetch.LL) and np.array_equal(g["set%d_ul" % si], fetch.UL)
        assert np.array_equal(g["set%d_act_ul" % si], fetch.ACT_UL)
    for key, rtype, otype, t, z, mode, with_scene, mts in episodes(g):
        si = int(key[2:-1]) // 6
        assert int(g["set%d_obs_dim" % si]) == O.pirl_obs_dim(O.pirl_cfg(rtype, otype, t, z))


def test_oracle_matches_reference_episodes():
    g = load()
    assert replay_oracle(g) > 200


def replay_gpu(g, device, dtype_name):
    import torch
    from rlitg_b200 import pirl_env as P, scene as S
    dtype = getattr(torch, dtype_name)
    f64 = dtype_name == "float64"
    gsc = None
    if any(e[6] for e in episodes(g)):
        occ, origin, res = scene_of(g)
        gsc = S.Scene(device)
        gsc.set_spheres(g["sphere_link"], g["sphere_center"], g["sphere_radius"])
        gsc.set_occupancy(occ, origin, res, 1.0)
    n = n_excl = 0
    for key, rtype, otype, t, z, mode, with_scene, mts in episodes(g):
        demos, targets = g[key + "demos"], g[key + "targets"]
        env = P.PirlVecEnv(1, rtype, otype, t, z, False, gsc if with_scene else None, device, dtype, want_errs=True)
        env.set_problems(targets[None], demos[None], np.array([mts], np.int32))
        obs0 = env.reset(mode=mode).double().cpu().numpy()[0]
        rel, floor = (1e-9, 1e-11) if f64 else (1e-5, 2e-6)
        assert (np.abs(obs0 - g[key + "obs0"]) / (floor + rel * np.abs(g[key + "obs0"]))).max() <= 1.0, key
        for k, a in enumerate(g[key + "action"]):
            # the recorded state is fed in at every step: a flag that flips within tolerance of its threshold then costs
            # one excluded step, not the rest of the episode
            if k > 0:
                env.cur_conf.copy_(torch.tensor(g[key + "conf"][k - 1][None], device=device, dtype=dtype))
                env.counters.copy_(torch.tensor(g[key + "counters"][k - 1][None], device=device))
            obs, rew, done, info = env.step(torch.tensor(a[None], device=device))
            torch.cuda.synchronize()
            fl = env.flags.cpu().numpy()[0][:4]
            e = env.errs.double().cpu().numpy()[0]
            tol = 1e-9 if f64 else 3e-6
            near = (abs(e[0] - 0.2) < tol or abs(e[0] - 0.04) < tol or abs(e[0] - 0.08) < tol or abs(e[1] - 0.104) < 30 * tol or
                    abs(e[2] - 0.087) < tol)
            n += 1
            if [bool(v) for v in fl] != [bool(v) for v in g[key + "flags"][k]]:
                assert near or (not f64 and with_scene), (key, k)   # FP32: a sphere centre on a voxel boundary
                n_excl += 1
                continue
            assert np.abs(env.cur_conf.double().cpu().numpy()[0] - g[key + "conf"][k]).max() < (1e-12 if f64 else 1e-6), (key, k)
            o = obs.double().cpu().numpy()[0]
            assert (np.abs(o - g[key + "obs"][k]) / (floor + rel * np.abs(g[key + "obs"][k]))).max() <= 1.0, (key, k)
            if near:
                n_excl += 1
                continue
            assert abs(float(rew[0]) - g[key + "reward"][k]) < (1e-8 if f64 else 2e-3), (key, k)
            assert bool(done[0]) == bool(g[key + "done"][k]), (key, k)
            assert list(env.counters.cpu().numpy()[0]) == list(g[key + "counters"][k]), (key, k)
    if gsc is not None:
        gsc.close()
    assert n_excl <= max(2, n // 50), (n_excl, n)
    return n


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_kernel_matches_reference_episodes(cuda_device, dtype):
    g = load()
    assert replay_gpu(g, cuda_device, dtype) > 200
