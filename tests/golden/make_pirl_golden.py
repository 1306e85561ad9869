#!/usr/bin/env python
"""Builds tests/golden/pirl_steps.npz by EXECUTING the reference's own env class, `PirlEnv_Fetch`
(/root/reference/pirl/script/envs/PirlEnv.py, imported from where it lies, never copied), on seeded episodes.

STATUS: NOT RUN YET.  Executing code of the external reference repository (MinsungYoon/RL-ITG, mounted at
/root/reference) needs a human's explicit authorisation in this environment; the attempts in round 1 and in round 2
(`python tests/golden/make_pirl_golden.py`, as written below) were refused by the sandbox policy -- "needs human
authorization for executing the external reference repo's code" -- and were not worked around.  THE ASK, for whoever
operates this repository: approve importing and executing /root/reference/pirl/script/envs/PirlEnv.py (from the
MinsungYoon/RL-ITG reference) in the dev container through this script; a plain yes is enough.  Until someone approves
and runs
    python tests/golden/make_pirl_golden.py
in the dev container (where /root/reference is mounted), tests/golden/pirl_steps.npz does not exist and the PIRL
step stays "parity unpinned" (DESIGN.md section 2).  The GPU box has no /root/reference, so the vectors would be
committed together with this script.

What runs from the reference: `__init__` (constants :118-174, limits via rospy.get_param <- the reference's own
pirl/launch/pirl.yaml), `reset` / `setup_problem` (CSV parsing, refine_continuous_joints), `step`, `_set_action`,
`_get_obs`, `calc_pos_and_rot`, `calc_pos_and_rot_error`, `_compute_task_reward`, `_compute_imitate_reward`,
`_compute_null_space_imitate_reward`, `calc_reward` -- i.e. every line SURVEY row A9 lists.

What is stubbed (third-party modules that do not exist in this image):
  rospy            logwarn (no-op), get_param (the reference's pirl.yaml; /use_rviz = False), ServiceProxy ->
                   the ROS services answered by the CPU oracle: /fk_solver, /allLinkfk_solver, /jaco_reward
                   (oracle FK is itself pinned on the reference's 68,000 (q -> pose) pairs, test_oracle_golden.py)
                   and /collision_check = the oracle's sphere-vs-grid test on a synthetic box scene (the real service
                   is an FCL mesh test; the env only consumes its boolean)
  quaternion       numpy-quaternion: the `np.quaternion(w, x, y, z)` value type, its Hamilton product and
                   `as_rotation_matrix` restated from the published package (normalised / unnormalised branches)
  utils, vae_model, pirl_msgs.srv   names only (visualiser, VAE and .srv types are not touched by these episodes)
Only `prepare_problemSet` (a directory scan of /data/torm_data) is overridden, to point at the episode's CSV files.
"""
import os
import sys
import tempfile
import types

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from rlitg_b200 import fetch, synth  # noqa: E402

BASE_OFF = (-0.086875, 0.0, 0.37743)
SCENE_SHAPE, SCENE_ORIGIN, SCENE_RES, SCENE_SEED = (64, 96, 64), (-0.2, -0.96, 0.0), 0.02, 5


class _Resp:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Quat:
    """numpy-quaternion value type: components w, x, y, z; `*` is the Hamilton product; norm() is w^2+x^2+y^2+z^2."""

    def __init__(self, w, x, y, z):
        self.w, self.x, self.y, self.z = float(w), float(x), float(y), float(z)

    def __mul__(self, o):
        a, b = self, o
        return _Quat(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
                     a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                     a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
                     a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w)

    def norm(self):
        return self.w * self.w + self.x * self.x + self.y * self.y + self.z * self.z


def _as_rotation_matrix(q):
    n = q.norm()
    if n == 0.0:
        raise ZeroDivisionError("Input to `as_rotation_matrix({0})` has zero norm".format(q))
    if abs(n - 1.0) < np.finfo(float).eps:
        s = 2.0
    else:
        s = 2.0 / n
    if s == 2.0:
        return np.array([[1 - 2 * (q.y ** 2 + q.z ** 2), 2 * (q.x * q.y - q.z * q.w), 2 * (q.x * q.z + q.y * q.w)],
                         [2 * (q.x * q.y + q.z * q.w), 1 - 2 * (q.x ** 2 + q.z ** 2), 2 * (q.y * q.z - q.x * q.w)],
                         [2 * (q.x * q.z - q.y * q.w), 2 * (q.y * q.z + q.x * q.w), 1 - 2 * (q.x ** 2 + q.y ** 2)]])
    return np.array([[1 - 2 * (q.y ** 2 + q.z ** 2) / n, 2 * (q.x * q.y - q.z * q.w) / n, 2 * (q.x * q.z + q.y * q.w) / n],
                     [2 * (q.x * q.y + q.z * q.w) / n, 1 - 2 * (q.x ** 2 + q.z ** 2) / n, 2 * (q.y * q.z - q.x * q.w) / n],
                     [2 * (q.x * q.z - q.y * q.w) / n, 2 * (q.y * q.z + q.x * q.w) / n, 1 - 2 * (q.x ** 2 + q.y ** 2) / n]])


def install_stubs(services):
    with open(os.path.join(REF, "pirl/launch/pirl.yaml")) as f:
        params = yaml.safe_load(f)

    def get_param(name):
        if name == "/use_rviz":
            return False
        node = params
        for part in name.strip("/").split("/"):
            node = node[part]
        return node

    rospy = types.ModuleType("rospy")
    rospy.logwarn = lambda *a, **k: None
    rospy.get_param = get_param
    rospy.ServiceProxy = lambda name, srv_type: services[name]
    sys.modules["rospy"] = rospy

    quat = types.ModuleType("quaternion")
    quat.quaternion = _Quat
    quat.as_rotation_matrix = _as_rotation_matrix
    sys.modules["quaternion"] = quat
    np.quaternion = _Quat  # the package registers its dtype on numpy

    utils = types.ModuleType("utils")
    utils.PathVisualizer = object
    sys.modules["utils"] = utils
    vae = types.ModuleType("vae_model")
    vae.VanillaVAE = object
    sys.modules["vae_model"] = vae
    msgs = types.ModuleType("pirl_msgs")
    srv = types.ModuleType("pirl_msgs.srv")
    for n in ("collision", "scene_set", "scene_reset", "fk", "ik", "rndvalconf", "allfk", "jaco"):
        setattr(srv, n, n)
    msgs.srv = srv
    sys.modules["pirl_msgs"] = msgs
    sys.modules["pirl_msgs.srv"] = srv


def make_scene():
    occ, _ = synth.random_box_scene(SCENE_SHAPE, SCENE_ORIGIN, SCENE_RES, seed=SCENE_SEED, n_clutter=3)
    link, center, radius = fetch.default_spheres()
    return occ, O.Scene(O.build_field(occ, SCENE_RES, 1.0), SCENE_ORIGIN, SCENE_RES, link, center, radius)


def write_csv(path, rows):
    with open(path, "w") as f:
        for r in rows:
            f.write(",".join(repr(float(v)) for v in r) + "\n")


SETTINGS = [
    # rewardType, obsType, t, zDim, mode, scene
    ("Task+NullSpaceImitation", "Conf+LinkPoses", 1, 0, "train", True),   # the trained setting (train_main.py:15-16,42)
    ("Task+Imitation", "Conf+EE", 3, 0, "train", False),
    ("Task", "onlyConf", 2, 4, "train", True),
    ("Imitation", "onlyLinkPoses", 1, 0, "test", False),
]
EPISODES_PER_SETTING = 6
T = 24


def main():
    occ, oscene = make_scene()
    state = {"scene": None}
    services = {
        "/collision_check": lambda conf: _Resp(collision_result=bool(
            state["scene"] is not None and O.sphere_collide(np.array([conf]), state["scene"], BASE_OFF)[0])),
        "/scene_set": lambda *a: _Resp(result=True),
        "/scene_reset": lambda *a: _Resp(result=True),
        "/ik_solver": lambda *a: _Resp(ik_result=[]),
        "/fk_solver": lambda conf: _Resp(fk_result=tuple(O.fk_quat(np.array([conf]))[0].tolist()), result=True),
        "/allLinkfk_solver": lambda conf: _Resp(allfk_result=tuple(O.allfk(np.array([conf]))[0].reshape(-1).tolist()),
                                                result=True),
        "/jaco_reward": lambda err, conf: _Resp(objective_value=float(
            O.nullspace_norm(np.array([err]), np.array([conf]))[0])),
    }
    install_stubs(services)
    sys.path.insert(0, os.path.join(REF, "pirl/script/envs"))
    import PirlEnv as ref_env  # the reference module, executed in place

    tmp = tempfile.mkdtemp(prefix="pirl_golden_")

    class Env(ref_env.PirlEnv_Fetch):
        def prepare_problemSet(self):  # the only override: no /data/torm_data scan
            self.n_free_prob = self.n_obs_prob = 0

    out = {"scene_occ_bits": np.packbits(occ.reshape(-1)), "scene_shape": np.array(SCENE_SHAPE),
           "scene_origin": np.array(SCENE_ORIGIN), "scene_res": np.array(SCENE_RES), "base_off": np.array(BASE_OFF)}
    link, center, radius = fetch.default_spheres()
    out["sphere_link"], out["sphere_center"], out["sphere_radius"] = link, center, radius
    rng = np.random.default_rng(20240607)
    meta = []
    ep = 0
    n_flags = np.zeros(4, dtype=int)
    for si, (rtype, otype, t, z, mode, with_scene) in enumerate(SETTINGS):
        env = Env(sceneType="", rewardType=rtype, obsType=otype, t=t, zDim=z, refineFlag=False)
        state["scene"] = oscene if with_scene else None
        for k in range(EPISODES_PER_SETTING):
            demos = synth.random_joint_paths(1, T, seed=1000 * si + k)[0]
            if k == 4:
                demos[:, 4] += 4.0  # a continuous joint several radians out: refine_continuous_joints (:577-590)
            if k == 5:
                demos[:, 1] = np.linspace(1.40, 1.515, T)  # rides the shoulder_lift upper limit -> clipping
            targets = np.round(O.fk_quat(demos + rng.normal(0, 0.004, size=demos.shape)), 6)
            prefix = os.path.join(tmp, "ep%d" % ep)
            write_csv(prefix + "_config.csv", demos)
            write_csv(prefix + "_targetpose.csv", targets)
            obs0 = env.reset(env_num=-1, prob_path=prefix, mode=mode, set_from_start=True)
            noise = [0.001, 0.004, 0.012, 0.03, 0.004, 0.01][k]
            acts, obs, rew, done, flags, cnt, conf = [], [], [], [], [], [], []
            for step in range(env.max_timestep):
                a = env.demo_configs[step + 1] - env.cur_conf
                for j in env.c_joints:  # shortest way round on continuous joints
                    a[j] = (a[j] + np.pi) % (2 * np.pi) - np.pi
                a = a + rng.normal(0, noise, size=7)
                if rng.random() < 0.1:
                    a[rng.integers(0, 7)] += rng.choice([-1.0, 1.0]) * rng.uniform(0.3, 2.5)  # limit hits, wraps
                a = a.astype(np.float32)
                o, r, d, info = env.step(a)
                acts.append(a); obs.append(o); rew.append(r); done.append(bool(d))
                flags.append([info["is_col"], info["is_jlimit"], info["is_early_stop"], info["is_timeout"]])
                cnt.append([info["task_suc_count"], info["dist_suc_count"], info["rot_suc_count"],
                            info["imitate_suc_count"]])
                conf.append(np.array(env.cur_conf, dtype=np.float64))
                if d:
                    break
            key = "ep%d_" % ep
            out[key + "demos"] = np.array(env.demo_configs, dtype=np.float64)   # after refine_continuous_joints
            out[key + "demos_raw"] = demos
            out[key + "targets"] = np.array(env.target_poses, dtype=np.float64)
            out[key + "obs0"] = np.asarray(obs0, dtype=np.float64)
            out[key + "action"] = np.array(acts, dtype=np.float32)
            out[key + "obs"] = np.array(obs, dtype=np.float64)
            out[key + "reward"] = np.array(rew, dtype=np.float64)
            out[key + "done"] = np.array(done, dtype=np.uint8)
            out[key + "flags"] = np.array(flags, dtype=np.uint8)
            out[key + "counters"] = np.array(cnt, dtype=np.int32)
            out[key + "conf"] = np.array(conf)
            n_flags += np.array(flags, dtype=int).sum(0)
            meta.append([rtype, otype, str(t), str(z), mode, str(int(with_scene)), str(env.max_timestep)])
            ep += 1
        # constants of __init__ as the reference set them (checked against the restated tables in the tests)
        out["set%d_ll" % si], out["set%d_ul" % si] = env.ll, env.ul
        out["set%d_act_ul" % si] = env.act_ul
        out["set%d_obs_dim" % si] = np.array(env.observation_space_shape[0])
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "pirl_steps.npz"), **out)
    steps = sum(out["ep%d_reward" % e].shape[0] for e in range(ep))
    print("episodes", ep, "steps", steps, "flags col/jlimit/early/timeout", n_flags.tolist())


if __name__ == "__main__":
    main()
