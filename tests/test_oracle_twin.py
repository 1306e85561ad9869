"""An independent NumPy twin of the kinematics (4 x 4 homogeneous transforms, Rodrigues rotations, textbook geometric
Jacobian) against the C oracle at 1e-12, and -- when the reference checkout is present (the dev container) -- the chain
constants of both against the reference's own URDF / YAML data files.  CPU only."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from oracle import oracle as O
from rlitg_b200 import fetch

REF = "/root/reference"
# (name, origin xyz, axis, kind) in chain order: fetch_description/robots/fetch.urdf
CHAIN = [
    ("torso_lift_joint", (-0.086875, 0.0, 0.37743), (0, 0, 1), "prismatic"),
    ("shoulder_pan_joint", (0.119525, 0.0, 0.34858), (0, 0, 1), "revolute"),
    ("shoulder_lift_joint", (0.117, 0.0, 0.06), (0, 1, 0), "revolute"),
    ("upperarm_roll_joint", (0.219, 0.0, 0.0), (1, 0, 0), "continuous"),
    ("elbow_flex_joint", (0.133, 0.0, 0.0), (0, 1, 0), "revolute"),
    ("forearm_roll_joint", (0.197, 0.0, 0.0), (1, 0, 0), "continuous"),
    ("wrist_flex_joint", (0.1245, 0.0, 0.0), (0, 1, 0), "revolute"),
    ("wrist_roll_joint", (0.1385, 0.0, 0.0), (1, 0, 0), "continuous"),
    ("gripper_axis", (0.16645, 0.0, 0.0), None, "fixed"),
]


def _tf(xyz, R=None):
    T = np.eye(4)
    T[:3, 3] = xyz
    if R is not None:
        T[:3, :3] = R
    return T


def _rodrigues(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def twin_fk_jac(q, D):
    """-> frames after every segment (list of 4x4), geometric Jacobian 6 x D at the last frame's origin."""
    chain = CHAIN if D == 8 else CHAIN[1:]
    T = np.eye(4)
    frames, joint_origin, joint_axis, kinds = [], [], [], []
    k = 0
    for name, xyz, axis, kind in chain:
        T = T @ _tf(xyz)
        if kind != "fixed":
            joint_origin.append(T[:3, 3].copy())
            joint_axis.append(T[:3, :3] @ np.asarray(axis, dtype=np.float64))
            kinds.append(kind)
            if kind == "prismatic":
                T = T @ _tf(np.asarray(axis, dtype=np.float64) * q[k])
            else:
                T = T @ _tf((0, 0, 0), _rodrigues(axis, q[k]))
            k += 1
        frames.append(T.copy())
    p = T[:3, 3]
    J = np.zeros((6, D))
    for j in range(D):
        if kinds[j] == "prismatic":
            J[:3, j] = joint_axis[j]
        else:
            J[:3, j] = np.cross(joint_axis[j], p - joint_origin[j])
            J[3:, j] = joint_axis[j]
    return frames, J


@pytest.mark.parametrize("D", [7, 8])
def test_numpy_twin_matches_oracle(D):
    rng = np.random.default_rng(5)
    lo = np.concatenate([[0.0], fetch.LL]) if D == 8 else fetch.LL
    hi = np.concatenate([[0.38615], fetch.UL]) if D == 8 else fetch.UL
    q = rng.uniform(lo, hi, size=(200, D))
    pos, rot, J = O.fk_jac(q)
    allfk = O.allfk(q)
    for b in range(q.shape[0]):
        frames, Jt = twin_fk_jac(q[b], D)
        assert np.abs(frames[-1][:3, 3] - pos[b]).max() < 1e-12
        assert np.abs(frames[-1][:3, :3] - rot[b]).max() < 1e-12
        assert np.abs(Jt - J[b]).max() < 1e-12
        assert len(frames) == D + 1
        for s, F in enumerate(frames):
            assert np.abs(F[:3, 3] - allfk[b, s, :3]).max() < 1e-12
            R = O.quat_to_rot(allfk[b, s, 3:])
            assert np.abs(np.asarray(R).reshape(3, 3) - F[:3, :3]).max() < 1e-12


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_chain_constants_against_the_reference_data_files():
    """The twin's table (and through the test above the oracle's and the kernels') against the reference's URDF, and
    the limits in rl-itg_b200/fetch.py against its YAML files."""
    import yaml
    root = ET.parse(os.path.join(REF, "fetch_description/robots/fetch.urdf")).getroot()
    joints = {j.get("name"): j for j in root.findall("joint")}
    for name, xyz, axis, kind in CHAIN:
        j = joints[name]
        assert j.get("type") == kind
        o = j.find("origin")
        assert np.allclose([float(v) for v in o.get("xyz").split()], xyz, atol=1e-12, rtol=0)  # the URDF writes 0.06 as 0.0599999999999999
        assert np.abs([float(v) for v in o.get("rpy").split()]).max() < 1e-15   # all joint rpy are 0 (torso -6e-17)
        if axis is not None:
            assert [float(v) for v in j.find("axis").get("xyz").split()] == list(axis)
    assert float(joints["torso_lift_joint"].find("limit").get("upper")) == fetch.TORSO_RANGE[1]
    # the chain is connected in this order
    for (child_joint, _, _, _), (parent_joint, _, _, _) in zip(CHAIN[1:], CHAIN[:-1]):
        assert joints[child_joint].find("parent").get("link") == joints[parent_joint].find("child").get("link")
    with open(os.path.join(REF, "pirl/launch/pirl.yaml")) as f:
        y = yaml.safe_load(f)["robot"]
    assert y["ll"] == fetch.LL.tolist() and y["ul"] == fetch.UL.tolist()
    assert y["vel_ll"] == fetch.ACT_LL.tolist() and y["vel_ul"] == fetch.ACT_UL.tolist()
    assert y["continuous_joints"] == np.nonzero(fetch.CONTINUOUS)[0].tolist()
    assert y["fk_base_position"] == fetch.BASE_OFFSET.tolist()
    assert y["joint_names"] == [n for n, _, _, k in CHAIN[1:] if k != "fixed"]
    with open(os.path.join(REF, "fetch_moveit_assistant/config/joint_limits.yaml")) as f:
        jl = yaml.safe_load(f)["joint_limits"]
    assert [jl[n]["max_velocity"] for n in y["joint_names"]] == fetch.VEL_LIMIT.tolist()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_parameter_defaults_against_the_reference_main():
    """TormParameters defaults (rl-itg_b200/torm.py) against the assignments of initParameters in the reference's
    torm/src/torm_main.cpp:50-92, read as text (nothing of the reference is executed), and the PIRL env constants of
    make_cfg against the literals of pirl/script/envs/PirlEnv.py:122-172."""
    import re
    import ctypes as C
    pytest.importorskip("torch")
    try:
        from rlitg_b200.torm import TormParameters
        from rlitg_b200.pirl_env import make_cfg
    except OSError:
        pytest.skip("librlitg.so not built")
    src = open(os.path.join(REF, "torm/src/torm_main.cpp")).read()
    body = src[src.index("void initParameters"):src.index("// random_obs_63_1")]
    found = {}
    for name, val in re.findall(r"params_\.(\w+)\s*=\s*([^;]+);", body):
        found.setdefault(name, []).append(val.strip())
    pr = TormParameters()
    checked = 0
    for name, vals in found.items():
        val = vals[-1]  # the non-debug branch comes last
        if name == "smoothness_update_weight_":
            assert val == "30.0/endPose_size" and pr.smoothness_update_weight_ == 30.0
        elif val in ("true", "false"):
            assert getattr(pr, name) == (val == "true"), name
        else:
            assert float(getattr(pr, name)) == float(val), name
        checked += 1
    assert checked >= 22
    env = open(os.path.join(REF, "pirl/script/envs/PirlEnv.py")).read()
    lit = lambda n: float(re.search(r"self\.%s\s*=\s*([-0-9.e]+)" % n, env).group(1))
    lst = lambda n: [float(v) for v in re.search(r"self\.%s\s*=\s*\[([^\]]+)\]" % n, env).group(1).split(",")]
    c = make_cfg()
    assert (c.thr_early_stop, c.thr_dist, c.thr_rot, c.thr_imitate) == (lit("threshold_early_stop_dist"), lit("threshold_dist_err"),
                                                                       lit("threshold_rot_err"), lit("threshold_imitate_err"))
    assert c.zero_task_dist == 2 * lit("threshold_dist_err")
    assert list(c.dist_p) == lst("dist_reward_params") and list(c.rot_p) == lst("rot_reward_params")
    assert list(c.imit_p) == lst("imit_reward_params") and list(c.null_p) == lst("null_space_imit_reward_params")
    assert (c.w_col, c.w_jlimit, c.w_early_stop) == (lit("w_col"), lit("w_jlimit"), lit("w_early_stop"))
