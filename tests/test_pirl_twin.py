"""An independent NumPy twin of the PIRL env step against the C oracle (CPU only).

The oracle's `orc_pirl_env_step` is pinned on reference artefacts only through FK (tests/test_oracle_golden.py); the
reference's own env class cannot be executed here without a human's authorisation (tests/golden/make_pirl_golden.py,
DESIGN.md section 2).  Until then this twin is the second, structurally different statement of
pirl/script/envs/PirlEnv.py:184-349,495-542 + pirl/src/kinematic_solver.cpp:332-351 that the oracle must agree with:
  * scalar Python in the reference's own shape (one env, one step at a time, lists and numpy vectors),
  * link poses from 4 x 4 homogeneous transforms (tests/test_oracle_twin.py), not from the oracle's chain walk,
  * the end-effector quaternion from scipy's Rotation (not KDL's trace-branch GetQuaternion),
  * the null-space objective as || (I - pinv(J) J) e || with numpy's SVD pseudo-inverse (not J^T (J J^T)^-1 J).
Episodes are stepped for several steps with the state fed back, with actions that hit joint limits and wrap the
continuous joints."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import oracle as O
from rlitg_b200 import fetch, synth
from test_oracle_twin import twin_fk_jac

REWARD = {"Task+NullSpaceImitation": (True, False, True, 10.0, 1.0, 1.0, 0.0), "Task+Imitation": (True, True, False, 10.0, 1.0, 1.0, 0.0),
          "Task": (True, False, False, 10.0, 1.0, 1.0, 0.0), "Imitation": (False, True, False, 0.0, 0.0, 0.0, 10.0)}


def calc_reward(err, p):
    return p[0] * np.exp(p[1] * err) - p[2] * (p[3] * err) ** 2


def quat_to_matrix(x, y, z, w):
    """numpy-quaternion as_rotation_matrix: normalised by |q|^2"""
    n = w * w + x * x + y * y + z * z
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


class TwinEnv:
    """One environment, scalar state, in the shape of PirlEnv_Fetch (no collision service: scene-free episodes)."""

    def __init__(self, reward_type, obs_type, t, z_dim, mode):
        self.task, self.imit, self.null, self.task_suc_r, self.dist_suc_r, self.rot_suc_r, self.imit_suc_r = REWARD[reward_type]
        self.obs_type, self.t, self.z_dim, self.mode = obs_type, t, z_dim, mode
        self.ll, self.ul = fetch.LL.copy(), fetch.UL.copy()
        self.c_joints = [2, 4, 6]

    def reset(self, targets, demos, q0, latent):
        self.targets, self.demos = targets, demos
        self.max_timestep = targets.shape[0] - 1
        self.cur = q0.copy()
        self.timestep = 0
        self.latent = list(latent)
        self.counts = [0, 0, 0, 0]

    def link_frames(self):
        return twin_fk_jac(self.cur, 7)

    def obs(self, frames):
        o = []
        pose9 = lambda F: [F[0, 3], F[1, 3], F[2, 3], F[0, 0], F[1, 0], F[2, 0], F[0, 1], F[1, 1], F[2, 1]]
        if self.obs_type == "onlyConf":
            o += self.cur.tolist()
        elif self.obs_type == "Conf+LinkPoses":
            o += self.cur.tolist()
            for F in frames:
                o += pose9(F)
        elif self.obs_type == "Conf+EE":
            o += self.cur.tolist() + pose9(frames[-1])
        else:
            for F in frames:
                o += pose9(F)
        ee = frames[-1]
        for i in range(self.t):
            k = min(self.timestep + i + 1, self.max_timestep)
            tg = self.targets[k]
            tM = quat_to_matrix(tg[3], tg[4], tg[5], tg[6])
            Rct = tM @ np.linalg.inv(ee[:3, :3])
            o += [tg[0] - ee[0, 3], tg[1] - ee[1, 3], tg[2] - ee[2, 3], Rct[0, 0], Rct[1, 0], Rct[2, 0], Rct[0, 1], Rct[1, 1], Rct[2, 1]]
        return np.array(o + self.latent)

    def step(self, a):
        self.timestep += 1
        jlimit = False
        o2 = self.cur + a.astype(np.float64)
        for j in self.c_joints:
            if o2[j] > np.pi:
                o2[j] -= 2 * np.pi
            elif o2[j] < -np.pi:
                o2[j] += 2 * np.pi
        clipped = np.clip(o2, self.ll, self.ul)
        inside = np.allclose(o2, clipped)
        if self.mode == "train":
            if inside:
                self.cur = clipped
            else:
                jlimit = True
        else:
            self.cur = clipped
        frames, J = self.link_frames()
        ob = self.obs(frames)
        ee = frames[-1]
        tg = self.targets[self.timestep]
        d_err = float(np.linalg.norm(tg[:3] - ee[:3, 3]))
        ex, ey, ez, ew = Rotation.from_matrix(ee[:3, :3]).as_quat()
        w = tg[6] * ew + tg[3] * ex + tg[4] * ey + tg[5] * ez   # (tq * conj(eq)).w
        r_err = 0.0 if abs(w) > 1.0 else 2 * np.arccos(abs(w))
        dist_rwd, rot_rwd = calc_reward(d_err, [2.0, -65.6, 30.0, 1.0]), calc_reward(r_err, [2.0, -5.0, 0.0, 0.0])
        err = self.demos[self.timestep] - self.cur
        for c in self.c_joints:
            if err[c] > np.pi:
                err[c] -= 2 * np.pi
            elif err[c] < -np.pi:
                err[c] += 2 * np.pi
        imit_err = float(np.abs(err).sum() / 7)
        imit_rwd = calc_reward(imit_err, [1.0, -3.8, 0.05, 1.0])
        null_obj = float(np.linalg.norm((np.eye(7) - np.linalg.pinv(J) @ J) @ err))
        null_rwd = calc_reward(null_obj, [1.0, -15, 0.5, 1.0])
        timeout = self.timestep == self.max_timestep
        early = d_err > 0.2
        ds, rs, isuc = d_err < 0.04, r_err < 0.104, imit_err < 0.087
        self.counts[0] += int(ds and rs); self.counts[1] += int(ds); self.counts[2] += int(rs); self.counts[3] += int(isuc)
        rew = -1.0 * jlimit - 3.0 * early
        if self.task:
            near = d_err < 0.08
            rew += dist_rwd + self.dist_suc_r * ds + (rot_rwd if near else 0) + (self.rot_suc_r * rs if near else 0) + self.task_suc_r * (ds and rs)
        if self.imit:
            rew += imit_rwd + self.imit_suc_r * isuc
        if self.null:
            rew += null_rwd
        done = early or timeout
        return ob, rew, done, dict(jlimit=jlimit, outside=not inside, early=early, timeout=timeout, ds=ds, rs=rs, isuc=isuc, d_err=d_err, r_err=r_err,
                                   imit_err=imit_err, null_obj=null_obj)


@pytest.mark.parametrize("reward_type,obs_type,t,z,mode", [
    ("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0, "train"),   # the trained setting (pirl/script/train_main.py:15-16,42)
    ("Task+Imitation", "Conf+EE", 3, 0, "train"),
    ("Task", "onlyConf", 2, 4, "test"),
    ("Imitation", "onlyLinkPoses", 1, 0, "test"),
    ("Task+NullSpaceImitation", "Conf+LinkPoses", 2, 3, "test"),
])
def test_oracle_env_step_matches_numpy_twin(reward_type, obs_type, t, z, mode):
    T, episodes, steps = 14, 6, 9
    rng = np.random.default_rng(len(reward_type) * 131 + t)
    cfg = O.pirl_cfg(reward_type, obs_type, t, z, train_mode=(mode == "train"))
    n_jl = n_es = n_wrap = 0
    for ep in range(episodes):
        demos = synth.random_joint_paths(1, T, seed=1000 + 17 * ep + t)[0]
        targets = np.round(O.fk_quat(demos + rng.normal(0, 0.004, size=demos.shape)), 6)
        latent = rng.normal(size=z)
        env = TwinEnv(reward_type, obs_type, t, z, mode)
        env.reset(targets, demos, demos[0], latent)
        q = demos[:1].copy()
        ts = np.zeros(1, np.int32)
        cnt = np.zeros((1, 4), np.int32)
        for k in range(steps):
            a = demos[k + 1] - env.cur
            for j in (2, 4, 6):
                a[j] = (a[j] + np.pi) % (2 * np.pi) - np.pi
            a = a + rng.normal(0, [0.002, 0.01, 0.03][ep % 3], size=7)
            if rng.random() < 0.25:
                a[rng.integers(0, 7)] += rng.choice([-1.0, 1.0]) * rng.uniform(0.3, 3.0)   # limit hits, +-2 pi wraps
            a = a.astype(np.float32)
            before = env.cur.copy()
            ob, rew, done, info = env.step(a)
            w = O.pirl_env_step(cfg, q, a[None], ts, np.array([T - 1], np.int32), targets[None], demos[None],
                                latent=latent[None] if z else None, counters=cnt)
            n_jl += info["outside"]; n_es += info["early"]; n_wrap += int(np.any(np.abs(before + a) > np.pi))
            assert np.abs(w["q"][0] - env.cur).max() < 1e-12
            assert w["obs"].shape[1] == ob.shape[0] and np.abs(w["obs"][0] - ob).max() < 1e-9
            assert abs(w["reward"][0] - rew) < 1e-8, (ep, k, w["reward"][0], rew)
            assert bool(w["done"][0]) == bool(done)
            f = w["flags"][0]
            assert (bool(f[1]), bool(f[2]), bool(f[3]), bool(f[4]), bool(f[5])) == (info["jlimit"], info["early"], info["timeout"], info["ds"], info["rs"])
            assert list(w["counters"][0][:3]) == env.counts[:3]
            assert abs(w["errs"][0, 0] - info["d_err"]) < 1e-12 and abs(w["errs"][0, 1] - info["r_err"]) < 1e-7
            if cfg.imitate_reward_flag or cfg.null_reward_flag:
                assert abs(w["errs"][0, 2] - info["imit_err"]) < 1e-12 and w["counters"][0][3] == env.counts[3]
            if cfg.null_reward_flag:
                assert abs(w["errs"][0, 3] - info["null_obj"]) < 1e-9
            q, ts, cnt = w["q"], w["timestep"], w["counters"]
            if done:
                break
    assert n_jl + n_wrap > 0   # the episodes exercise the joint-limit rejection (train) / clipping (test) or the wrap


def write_surrogate(path):
    """A file in the layout of tests/golden/pirl_steps.npz written from the NumPy twin (scene-free episodes; the twin has
    no collision service).  Returns the number of episodes."""
    settings = [("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0, "train"), ("Task+Imitation", "Conf+EE", 3, 0, "train"),
                ("Task", "onlyConf", 2, 4, "train"), ("Imitation", "onlyLinkPoses", 1, 0, "test")]
    T = 12
    rng = np.random.default_rng(77)
    link, center, radius = fetch.default_spheres()
    out = {"scene_occ_bits": np.packbits(np.zeros(8 * 8 * 8, np.uint8)), "scene_shape": np.array((8, 8, 8)),
           "scene_origin": np.array((5.0, 5.0, 5.0)), "scene_res": np.array(0.02), "base_off": np.array(fetch.BASE_OFFSET),
           "sphere_link": link, "sphere_center": center, "sphere_radius": radius}
    meta, ep = [], 0
    for si, (rtype, otype, t, z, mode) in enumerate(settings):
        for k in range(2):
            demos = synth.random_joint_paths(1, T, seed=500 + 10 * si + k)[0]
            targets = np.round(O.fk_quat(demos + rng.normal(0, 0.004, size=demos.shape)), 6)
            env = TwinEnv(rtype, otype, t, z, mode)
            env.reset(targets, demos, demos[0], np.zeros(z))
            rec = {n: [] for n in ("action", "obs", "reward", "done", "flags", "counters", "conf")}
            obs0 = env.obs(env.link_frames()[0])
            for step in range(env.max_timestep):
                a = demos[step + 1] - env.cur
                for j in (2, 4, 6):
                    a[j] = (a[j] + np.pi) % (2 * np.pi) - np.pi
                a = (a + rng.normal(0, 0.004, size=7)).astype(np.float32)
                if step == 3:
                    a[1] += 2.5   # a joint-limit hit
                o, r, d, info = env.step(a)
                rec["action"].append(a); rec["obs"].append(o); rec["reward"].append(r); rec["done"].append(bool(d))
                rec["flags"].append([False, info["jlimit"], info["early"], info["timeout"]])
                rec["counters"].append(list(env.counts)); rec["conf"].append(env.cur.copy())
                if d:
                    break
            key = "ep%d_" % ep
            out[key + "demos"], out[key + "targets"], out[key + "obs0"] = demos, targets, obs0
            out[key + "action"] = np.array(rec["action"], np.float32)
            out[key + "obs"], out[key + "reward"] = np.array(rec["obs"]), np.array(rec["reward"])
            out[key + "done"], out[key + "flags"] = np.array(rec["done"], np.uint8), np.array(rec["flags"], np.uint8)
            out[key + "counters"], out[key + "conf"] = np.array(rec["counters"], np.int32), np.array(rec["conf"])
            meta.append([rtype, otype, str(t), str(z), mode, "0", str(env.max_timestep)])
            ep += 1
    out["meta"] = np.array(meta)
    np.savez_compressed(path, **out)
    return ep


def test_golden_replay_plumbing(tmp_path):
    """tests/test_pirl_golden.py replays tests/golden/pirl_steps.npz (episodes recorded from the reference's own env
    class -- absent until a human authorises running it).  Its replay logic is exercised here on the surrogate file."""
    from test_pirl_golden import replay_oracle
    path = str(tmp_path / "surrogate_pirl_steps.npz")
    ep = write_surrogate(path)
    assert replay_oracle(np.load(path), tol=1e-9) >= ep * 4


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_golden_replay_plumbing_gpu(cuda_device, tmp_path, dtype):
    """the GPU half of the same replay (rlitg_pirl_env_step, FP64 and FP32 builds) on the surrogate file"""
    from test_pirl_golden import replay_gpu
    path = str(tmp_path / "surrogate_pirl_steps.npz")
    ep = write_surrogate(path)
    assert replay_gpu(np.load(path), cuda_device, dtype) >= ep * 4
