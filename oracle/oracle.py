"""ctypes front-end of the CPU ORACLE (oracle/rlitg_oracle.c).

TEST INFRASTRUCTURE ONLY.  Nothing under ``rl-itg_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do.
All arrays are float64 / C-contiguous; shapes are documented in rlitg_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

NCOST = 10
STAGE_POSE, STAGE_POSE_GRAD, STAGE_SMOOTH, STAGE_COLL, STAGE_EVAL = 1, 2, 4, 8, 16
INACTIVE = np.float32(1e30)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, no third-party dependencies)."""
    src = os.path.join(_HERE, "rlitg_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "rlitg_oracle.h")))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_rot_angle_kdl.restype = C.c_double
    return _lib


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OrcScene(C.Structure):
    _fields_ = [("S", C.c_int), ("link", C.POINTER(C.c_int32)), ("center", C.POINTER(C.c_double)),
                ("radius", C.POINTER(C.c_double)), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("origin", C.c_double * 3), ("res", C.c_double), ("field", C.POINTER(C.c_float))]


class OrcTormParams(C.Structure):
    _fields_ = [("T", C.c_int), ("D", C.c_int), ("dt", C.c_double), ("rot_scale", C.c_double),
                ("jac_weight", C.c_double), ("clearance", C.c_double), ("pinv_ridge", C.c_double),
                ("vik_eps", C.c_double), ("use_pinv", C.c_int), ("affine_axis_quirk", C.c_int),
                ("base_off", C.c_double * 3), ("cjoint", C.c_int * 8), ("vel_limit", C.c_double * 8)]


class OrcPirlCfg(C.Structure):
    _fields_ = [("ll", C.c_double * 7), ("ul", C.c_double * 7), ("cjoint", C.c_int * 7),
                ("n_future", C.c_int), ("z_dim", C.c_int), ("obs_type", C.c_int), ("train_mode", C.c_int),
                ("refine_flag", C.c_int), ("task_reward_flag", C.c_int), ("imitate_reward_flag", C.c_int),
                ("null_reward_flag", C.c_int), ("task_suc_r", C.c_double), ("dist_suc_r", C.c_double),
                ("rot_suc_r", C.c_double), ("imit_suc_r", C.c_double), ("thr_early_stop", C.c_double),
                ("thr_dist", C.c_double), ("thr_rot", C.c_double), ("thr_imitate", C.c_double),
                ("zero_task_dist", C.c_double), ("dist_p", C.c_double * 4), ("rot_p", C.c_double * 4),
                ("imit_p", C.c_double * 4), ("null_p", C.c_double * 4), ("w_col", C.c_double),
                ("w_jlimit", C.c_double), ("w_early_stop", C.c_double)]


# ---------------------------------------------------------------------------------- kinematics
def fk(q):
    q = _f64(q); B, D = q.shape
    pos = np.empty((B, 3)); rot = np.empty((B, 9))
    assert lib().orc_fk(_p(q), C.c_int64(B), D, _p(pos), _p(rot)) == 0
    return pos, rot.reshape(B, 3, 3)


def fk_quat(q):
    q = _f64(q); B, D = q.shape
    out = np.empty((B, 7))
    assert lib().orc_fk_quat(_p(q), C.c_int64(B), D, _p(out)) == 0
    return out


def allfk(q):
    q = _f64(q); B, D = q.shape
    out = np.empty((B, D + 1, 7))
    assert lib().orc_allfk(_p(q), C.c_int64(B), D, _p(out)) == 0
    return out


def jacobian(q):
    q = _f64(q); B, D = q.shape
    J = np.empty((B, 6, D))
    assert lib().orc_jacobian(_p(q), C.c_int64(B), D, _p(J)) == 0
    return J


def fk_jac(q):
    q = _f64(q); B, D = q.shape
    pos = np.empty((B, 3)); rot = np.empty((B, 9)); J = np.empty((B, 6, D))
    assert lib().orc_fk_jac(_p(q), C.c_int64(B), D, _p(pos), _p(rot), _p(J)) == 0
    return pos, rot.reshape(B, 3, 3), J


def pose_diff(q, target):
    q = _f64(q); target = _f64(target); B, D = q.shape
    tw = np.empty((B, 6))
    assert lib().orc_pose_diff(_p(q), _p(target), C.c_int64(B), D, _p(tw)) == 0
    return tw


def pinv_step(q, twist, eps=1e-5):
    q = _f64(q); twist = _f64(twist); B, D = q.shape
    dq = np.empty((B, D))
    assert lib().orc_pinv_step(_p(q), _p(twist), C.c_int64(B), D, C.c_double(eps), _p(dq)) == 0
    return dq


def nullspace_norm(err, q):
    err = _f64(err); q = _f64(q); B = q.shape[0]
    out = np.empty(B)
    assert lib().orc_nullspace_norm(_p(err), _p(q), C.c_int64(B), _p(out)) == 0
    return out


def manipulability(q):
    q = _f64(q); B, D = q.shape
    out = np.empty(B)
    assert lib().orc_manipulability(_p(q), C.c_int64(B), D, _p(out)) == 0
    return out


def quat_to_rot(quat_xyzw):
    qq = _f64(quat_xyzw).reshape(-1, 4)
    R = np.empty((qq.shape[0], 9))
    for i in range(qq.shape[0]):
        lib().orc_quat_to_rot_kdl(_p(qq[i]), _p(R[i]))
    return R.reshape(-1, 3, 3)


def rot_to_quat(R):
    RR = _f64(R).reshape(-1, 9)
    out = np.empty((RR.shape[0], 4))
    for i in range(RR.shape[0]):
        lib().orc_rot_to_quat_kdl(_p(RR[i]), _p(out[i]))
    return out


def rot_angle(R):
    RR = _f64(R).reshape(9)
    ax = np.empty(3)
    return float(lib().orc_rot_angle_kdl(_p(RR), _p(ax))), ax


# ---------------------------------------------------------------------------------- scene
def build_field(occ, res, max_prop=1.0):
    """occ: uint8 [nz, ny, nx] -> float32 records [nz, ny, nx, 4] = (d, gx, gy, gz)."""
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    nz, ny, nx = occ.shape
    field = np.empty((nz, ny, nx, 4), dtype=np.float32)
    assert lib().orc_build_field(_p(occ, C.c_uint8), nx, ny, nz, C.c_double(res), C.c_double(max_prop),
                                 _p(field, C.c_float)) == 0
    return field


class Scene:
    """Keeps the numpy buffers alive behind an OrcScene struct."""

    def __init__(self, field, origin, res, sphere_link, sphere_center, sphere_radius):
        self.field = np.ascontiguousarray(field, dtype=np.float32)
        self.link = np.ascontiguousarray(sphere_link, dtype=np.int32)
        self.center = _f64(sphere_center)
        self.radius = _f64(sphere_radius)
        nz, ny, nx, _ = self.field.shape
        s = OrcScene()
        s.S = int(self.link.shape[0])
        s.link = _p(self.link, C.c_int32); s.center = _p(self.center); s.radius = _p(self.radius)
        s.nx, s.ny, s.nz = nx, ny, nz
        s.origin = (C.c_double * 3)(*[float(v) for v in origin])
        s.res = float(res)
        s.field = _p(self.field, C.c_float)
        self.c = s


def sphere_collide(q, scene: Scene, base_off):
    q = _f64(q); B = q.shape[0]
    out = np.empty(B, dtype=np.uint8)
    bo = _f64(base_off)
    assert lib().orc_sphere_collide(_p(q), C.c_int64(B), C.byref(scene.c), _p(bo), _p(out, C.c_uint8)) == 0
    return out


# ---------------------------------------------------------------------------------- TORM
def torm_params(T, dt=0.1, rot_scale=0.17, jac_weight=1.0, clearance=0.3, pinv_ridge=1e-4, vik_eps=1e-5,
                use_pinv=True, affine_axis_quirk=True, base_off=(-0.086875, 0.0, 0.37743),
                cjoint=(0, 0, 1, 0, 1, 0, 1), vel_limit=(1.256, 1.454, 1.571, 1.521, 1.571, 2.268, 2.268), D=7):
    p = OrcTormParams()
    p.T, p.D, p.dt, p.rot_scale, p.jac_weight = T, D, dt, rot_scale, jac_weight
    p.clearance, p.pinv_ridge, p.vik_eps = clearance, pinv_ridge, vik_eps
    p.use_pinv, p.affine_axis_quirk = int(use_pinv), int(affine_axis_quirk)
    p.base_off = (C.c_double * 3)(*base_off)
    p.cjoint = (C.c_int * 8)(*(list(cjoint) + [0] * (8 - len(cjoint))))
    p.vel_limit = (C.c_double * 8)(*(list(vel_limit) + [0.0] * (8 - len(vel_limit))))
    return p


def smooth_setup(T, dt, w3=(1.0, 0.0, 0.0), ridge=0.0, want_inverse=False):
    G = T + 10
    Aband = np.empty((G, 13))
    scale = C.c_double(0)
    Ainv = np.empty((T - 1, T - 1)) if want_inverse else None
    w = _f64(w3)
    rc = lib().orc_smooth_setup(T, C.c_double(dt), _p(w), C.c_double(ridge), _p(Aband), C.byref(scale), _p(Ainv))
    assert rc == 0
    return Aband, scale.value, Ainv


def torm_cost_grad(traj, targets, scene, p, Aband, stage_mask, stochastic_idx=None):
    traj = _f64(traj); targets = _f64(targets)
    N, T, D = traj.shape
    assert T == p.T and D == p.D
    F = T - 1
    inc_s = np.zeros((N, F, D)); inc_c = np.zeros((N, F, D)); inc_p = np.zeros((N, F, D))
    costs = np.zeros((N, NCOST)); active = np.zeros(N)
    st = None if stochastic_idx is None else np.ascontiguousarray(stochastic_idx, dtype=np.int32)
    Ab = None if Aband is None else _f64(Aband)
    rc = lib().orc_torm_cost_grad(_p(traj), _p(targets), C.byref(scene.c) if scene is not None else None,
                                  C.byref(p), _p(Ab), int(stage_mask), _p(st, C.c_int32), C.c_int64(N),
                                  _p(inc_s), _p(inc_c), _p(inc_p), _p(costs), _p(active))
    assert rc == 0
    return dict(inc_smooth=inc_s, inc_coll=inc_c, inc_pose=inc_p, costs=costs, active=active)


def traj_eval(traj, targets, p):
    traj = _f64(traj); targets = _f64(targets)
    N = traj.shape[0]
    out = np.empty((N, 6))
    assert lib().orc_traj_eval(_p(traj), _p(targets), C.byref(p), C.c_int64(N), _p(out)) == 0
    return out


def torm_update(traj, Ainv, stage, mode, inc_smooth, inc_coll, inc_pose, alpha, lr, sw=0.15, ow=1.0, limit=0.17,
                cjoint=(0, 0, 1, 0, 1, 0, 1), jmin=None, jmax=None):
    """-> updated copy of traj [N,T,D]"""
    traj = _f64(traj).copy(); N, T, D = traj.shape
    Ainv = _f64(Ainv)
    f = lambda a: None if a is None else _f64(a)
    inc_smooth, inc_coll, inc_pose = f(inc_smooth), f(inc_coll), f(inc_pose)
    alpha = f(np.broadcast_to(alpha, (N,))); lr = f(np.broadcast_to(lr, (N,)))
    cj = np.ascontiguousarray(cjoint, dtype=np.int32)
    jmin = _f64([-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592] if jmin is None else jmin)
    jmax = _f64([1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592] if jmax is None else jmax)
    rc = lib().orc_torm_update(_p(traj), C.c_int64(N), T, D, _p(Ainv), int(stage), int(mode), _p(inc_smooth),
                               _p(inc_coll), _p(inc_pose), _p(alpha), _p(lr), C.c_double(sw), C.c_double(ow),
                               C.c_double(limit), _p(cj, C.c_int32), _p(jmin), _p(jmax))
    assert rc == 0
    return traj


# ---------------------------------------------------------------------------------- PIRL
def pirl_cfg(reward_type="Task+NullSpaceImitation", obs_type="Conf+LinkPoses", n_future=1, z_dim=0,
             train_mode=True, refine_flag=False):
    c = OrcPirlCfg()
    c.ll = (C.c_double * 7)(-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592)
    c.ul = (C.c_double * 7)(1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592)
    c.cjoint = (C.c_int * 7)(0, 0, 1, 0, 1, 0, 1)
    c.n_future, c.z_dim = n_future, z_dim
    c.obs_type = {"onlyConf": 0, "Conf+LinkPoses": 1, "Conf+EE": 2, "onlyLinkPoses": 3}[obs_type]
    c.train_mode, c.refine_flag = int(train_mode), int(refine_flag)
    flags = {"Task+NullSpaceImitation": (1, 0, 1, 10.0, 1.0, 1.0, 0.0), "Task+Imitation": (1, 1, 0, 10.0, 1.0, 1.0, 0.0),
             "Task": (1, 0, 0, 10.0, 1.0, 1.0, 0.0), "Imitation": (0, 1, 0, 0.0, 0.0, 0.0, 10.0)}[reward_type]
    c.task_reward_flag, c.imitate_reward_flag, c.null_reward_flag = flags[0], flags[1], flags[2]
    c.task_suc_r, c.dist_suc_r, c.rot_suc_r, c.imit_suc_r = flags[3:]
    c.thr_early_stop, c.thr_dist, c.thr_rot, c.thr_imitate = 0.2, 0.04, 0.104, 0.087
    c.zero_task_dist = 0.04 * 2
    c.dist_p = (C.c_double * 4)(2.0, -65.6, 30.0, 1.0)
    c.rot_p = (C.c_double * 4)(2.0, -5.0, 0.0, 0.0)
    c.imit_p = (C.c_double * 4)(1.0, -3.8, 0.05, 1.0)
    c.null_p = (C.c_double * 4)(1.0, -15, 0.5, 1.0)
    c.w_col, c.w_jlimit, c.w_early_stop = 10.0, 1.0, 3.0
    return c


def pirl_obs_dim(cfg):
    return int(lib().orc_pirl_obs_dim(C.byref(cfg)))


def pirl_env_step(cfg, q, action, timestep, max_timestep, targets, demos, latent=None, scene=None,
                  base_off=(-0.086875, 0.0, 0.37743), counters=None):
    """Returns dict with updated q / timestep (copies) and obs, reward, done, flags, counters, errs."""
    q = _f64(q).copy(); B = q.shape[0]
    action = np.ascontiguousarray(action, dtype=np.float32)
    ts = np.ascontiguousarray(timestep, dtype=np.int32).copy()
    mts = np.ascontiguousarray(max_timestep, dtype=np.int32)
    targets = _f64(targets); Tmax = targets.shape[1]
    demos = None if demos is None else _f64(demos)
    lat = None if latent is None else _f64(latent)
    O = pirl_obs_dim(cfg)
    obs = np.empty((B, O)); rew = np.empty(B); done = np.empty(B, dtype=np.uint8)
    flags = np.empty((B, 8), dtype=np.uint8); errs = np.empty((B, 4))
    cnt = np.zeros((B, 4), dtype=np.int32) if counters is None else np.ascontiguousarray(counters, dtype=np.int32).copy()
    bo = _f64(base_off)
    rc = lib().orc_pirl_env_step(C.byref(cfg), C.c_int64(B), _p(q), _p(action, C.c_float), _p(ts, C.c_int32),
                                 _p(mts, C.c_int32), _p(targets), _p(demos), Tmax, _p(lat),
                                 C.byref(scene.c) if scene is not None else None, _p(bo), _p(obs), _p(rew),
                                 _p(done, C.c_uint8), _p(flags, C.c_uint8), _p(cnt, C.c_int32), _p(errs))
    assert rc == 0
    return dict(q=q, timestep=ts, obs=obs, reward=rew, done=done, flags=flags, counters=cnt, errs=errs)


def pirl_obs(cfg, q, timestep, max_timestep, targets, latent=None):
    q = _f64(q); B = q.shape[0]
    ts = np.ascontiguousarray(timestep, dtype=np.int32); mts = np.ascontiguousarray(max_timestep, dtype=np.int32)
    targets = _f64(targets); Tmax = targets.shape[1]
    lat = None if latent is None else _f64(latent)
    obs = np.empty((B, pirl_obs_dim(cfg)))
    assert lib().orc_pirl_obs(C.byref(cfg), C.c_int64(B), _p(q), _p(ts, C.c_int32), _p(mts, C.c_int32),
                              _p(targets), Tmax, _p(lat), _p(obs)) == 0
    return obs


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
