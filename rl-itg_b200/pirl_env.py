"""PIRL path-wise IK environment on the GPU: B environments per step in one kernel launch.

Mirrors PirlEnv_Fetch (pirl/script/envs/PirlEnv.py): same constructor arguments (sceneType is replaced by an
optional Scene), same attributes the learners read (observation_space_shape, action_space_shape, act_ll, act_ul,
timestep, max_timestep, random_action(); pirl/script/algos/sac/sac.py:41-46,285-334) and the same
``reset`` / ``step(a) -> obs, reward, done, info`` contract, with a leading batch axis.  ``PirlEnvFetch`` is the
batch-of-one wrapper with the reference's scalar numpy API.

What stays on the host, as in the reference: choosing problems and reading their files (PirlEnv.py:394-487).
Problems are handed over as arrays: target_poses [B,Tmax,7] (x y z qx qy qz qw), demo_configs [B,Tmax,7],
max_timestep [B] (= len - 1 of each path; rows past it are padding)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, fetch
from ._lib import check, lib


class RlitgPirlCfg(C.Structure):
    _fields_ = [("ll", C.c_double * 7), ("ul", C.c_double * 7), ("cjoint", C.c_int32 * 7),
                ("n_future", C.c_int32), ("z_dim", C.c_int32), ("obs_type", C.c_int32), ("train_mode", C.c_int32),
                ("refine_flag", C.c_int32), ("task_reward_flag", C.c_int32), ("imitate_reward_flag", C.c_int32),
                ("null_reward_flag", C.c_int32), ("task_suc_r", C.c_double), ("dist_suc_r", C.c_double),
                ("rot_suc_r", C.c_double), ("imit_suc_r", C.c_double), ("thr_early_stop", C.c_double),
                ("thr_dist", C.c_double), ("thr_rot", C.c_double), ("thr_imitate", C.c_double),
                ("zero_task_dist", C.c_double), ("dist_p", C.c_double * 4), ("rot_p", C.c_double * 4),
                ("imit_p", C.c_double * 4), ("null_p", C.c_double * 4), ("w_col", C.c_double),
                ("w_jlimit", C.c_double), ("w_early_stop", C.c_double), ("row_floats", C.c_int32), ("record_has_demo", C.c_int32)]


_vp, _i64, _i = C.c_void_p, C.c_int64, C.c_int
lib.rlitg_pirl_obs_dim.argtypes = [C.POINTER(RlitgPirlCfg)]
lib.rlitg_pirl_env_step.argtypes = [C.POINTER(RlitgPirlCfg), _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.rlitg_pirl_obs.argtypes = [C.POINTER(RlitgPirlCfg), _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]
lib.rlitg_nullspace_norm.argtypes = [_vp, _vp, _i64, _i, _vp, _vp]
lib.rlitg_pinv_step.argtypes = [_vp, _vp, _i64, _i, _i, C.c_double, _vp, _vp]
lib.rlitg_pose_diff.argtypes = [_vp, _vp, _i64, _i, _i, _vp, _vp]
lib.rlitg_manipulability.argtypes = [_vp, _i64, _i, _i, _vp, _vp]

OBS_TYPES = {"onlyConf": 0, "Conf+LinkPoses": 1, "Conf+EE": 2, "onlyLinkPoses": 3}
# (task, imitate, null-space) flags and (task, dist, rot, imit) success rewards: PirlEnv.py:130-161
REWARD_TYPES = {"Task+NullSpaceImitation": (1, 0, 1, 10.0, 1.0, 1.0, 0.0), "Task+Imitation": (1, 1, 0, 10.0, 1.0, 1.0, 0.0),
                "Task": (1, 0, 0, 10.0, 1.0, 1.0, 0.0), "Imitation": (0, 1, 0, 0.0, 0.0, 0.0, 10.0)}
FLAG_NAMES = ("is_col", "is_jlimit", "is_early_stop", "is_timeout", "is_dist_suc", "is_rot_suc", "is_imitate_suc",
              "is_task_suc")


def make_cfg(rewardType="Task+NullSpaceImitation", obsType="Conf+LinkPoses", t=1, zDim=0, mode="train",
             refineFlag=False) -> RlitgPirlCfg:
    """Constants of PirlEnv_Fetch.__init__ (pirl/script/envs/PirlEnv.py:118-174) and pirl/launch/pirl.yaml."""
    c = RlitgPirlCfg()
    c.ll = (C.c_double * 7)(*fetch.LL)
    c.ul = (C.c_double * 7)(*fetch.UL)
    c.cjoint = (C.c_int32 * 7)(*[int(v) for v in fetch.CONTINUOUS])
    c.n_future, c.z_dim, c.obs_type = int(t), int(zDim), OBS_TYPES[obsType]
    c.train_mode, c.refine_flag = int(mode == "train"), int(bool(refineFlag))
    f = REWARD_TYPES[rewardType]
    c.task_reward_flag, c.imitate_reward_flag, c.null_reward_flag = f[0], f[1], f[2]
    c.task_suc_r, c.dist_suc_r, c.rot_suc_r, c.imit_suc_r = f[3:]
    c.thr_early_stop, c.thr_dist, c.thr_rot, c.thr_imitate = 0.2, 0.04, 0.104, 0.087
    c.zero_task_dist = 0.04 * 2
    c.dist_p = (C.c_double * 4)(2.0, -65.6, 30.0, 1.0)
    c.rot_p = (C.c_double * 4)(2.0, -5.0, 0.0, 0.0)
    c.imit_p = (C.c_double * 4)(1.0, -3.8, 0.05, 1.0)
    c.null_p = (C.c_double * 4)(1.0, -15, 0.5, 1.0)
    c.w_col, c.w_jlimit, c.w_early_stop = 10.0, 1.0, 3.0
    return c


def _code(t):
    return _lib.F32 if t.dtype == torch.float32 else _lib.F64


def _stream():
    return torch.cuda.current_stream().cuda_stream


try:  # the raw handle of the current stream without building a torch.cuda.Stream object (hot path of PirlVecEnv.step)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


lib.rlitg_pirl_bind.argtypes = [C.POINTER(RlitgPirlCfg), _i64, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp,
                                _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_void_p)]
lib.rlitg_pirl_bind.restype = C.c_int
lib.rlitg_pirl_step_bound.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.rlitg_pirl_step_bound.restype = C.c_int
lib.rlitg_pirl_bound_destroy.argtypes = [C.c_void_p]
lib.rlitg_pirl_bound_destroy.restype = None
_step_bound = lib.rlitg_pirl_step_bound


def nullspace_norm(err, q):
    """jaco_srv(err, q).objective_value for a batch (pirl/src/kinematic_solver.cpp:332-351)."""
    err, q = err.contiguous(), q.contiguous()
    out = torch.empty(q.shape[0], device=q.device, dtype=q.dtype)
    with _lib.on_device(q.device):
        check(lib.rlitg_nullspace_norm(err.data_ptr(), q.data_ptr(), q.shape[0], _code(q), out.data_ptr(), _stream()))
    return out


def pinv_step(q, twist, eps=1e-5):
    """TormIKSolver::vikSolver for a batch (torm/src/torm_ik_solver.cpp:169-171)."""
    q, twist = q.contiguous(), twist.contiguous()
    dq = torch.empty_like(q)
    with _lib.on_device(q.device):
        check(lib.rlitg_pinv_step(q.data_ptr(), twist.data_ptr(), q.shape[0], q.shape[1], _code(q), float(eps),
                                  dq.data_ptr(), _stream()))
    return dq


def manipulability(q):
    """sqrt(det(J J^T)) per configuration: the measure TormOptimizer::checkSingularity compares with
    singularity_lower_bound_ (torm/src/torm_optimizer.cpp:748-765)."""
    q = q.contiguous()
    out = torch.empty(q.shape[0], device=q.device, dtype=q.dtype)
    with _lib.on_device(q.device):
        check(lib.rlitg_manipulability(q.data_ptr(), q.shape[0], q.shape[1], _code(q), out.data_ptr(), _stream()))
    return out


def pose_diff(q, target):
    """KDL::diff(FK(q), target) for a batch -> twist [B,6]."""
    q, target = q.contiguous(), target.contiguous()
    tw = torch.empty(q.shape[0], 6, device=q.device, dtype=q.dtype)
    with _lib.on_device(q.device):
        check(lib.rlitg_pose_diff(q.data_ptr(), target.data_ptr(), q.shape[0], q.shape[1], _code(q), tw.data_ptr(),
                                  _stream()))
    return tw


def refine_continuous_joints(conf):
    """PirlEnv_Fetch.refine_continuous_joints (PirlEnv.py:577-590) on a [..., 7] tensor: continuous joints beyond +-pi are
    brought back by whole turns (int() truncation, then one more turn if still outside)."""
    cj = torch.tensor(fetch.CONTINUOUS, device=conf.device, dtype=torch.bool)
    two_pi = 2 * np.pi
    wrapped = conf - torch.trunc(conf / two_pi) * two_pi
    wrapped = torch.where((conf > np.pi) & (wrapped > np.pi), wrapped - two_pi, wrapped)
    wrapped = torch.where((conf < -np.pi) & (wrapped < -np.pi), wrapped + two_pi, wrapped)
    return torch.where((conf.abs() > np.pi) & cj, wrapped, conf)


class PirlVecEnv:
    """B path-following IK environments stepped by one kernel (rlitg_pirl_env_step)."""

    def __init__(self, num_envs, rewardType="Task+NullSpaceImitation", obsType="Conf+LinkPoses", t=1, zDim=0,
                 refineFlag=False, scene=None, device="cuda:0", dtype=torch.float32, want_errs=False):
        self.B = int(num_envs)
        self.device = torch.device(device)
        self._dev_index = self.device.index if self.device.index is not None else 0
        self.dtype = dtype
        self.scene = scene
        self.mode = "train"
        self._args = dict(rewardType=rewardType, obsType=obsType, t=t, zDim=zDim, refineFlag=refineFlag)
        self.cfg = make_cfg(mode=self.mode, **self._args)
        self.n_dof = 7
        self.ll, self.ul = fetch.LL.copy(), fetch.UL.copy()
        self.act_ll, self.act_ul = fetch.ACT_LL.copy(), fetch.ACT_UL.copy()
        self.c_joints = [2, 4, 6]
        self.action_space_shape = (self.n_dof,)
        self.obs_dim = int(lib.rlitg_pirl_obs_dim(C.byref(self.cfg)))
        self.observation_space_shape = (self.obs_dim,)
        kw = dict(device=self.device)
        B = self.B
        self.cur_conf = torch.zeros(B, 7, dtype=dtype, **kw)
        self.timestep = torch.zeros(B, dtype=torch.int32, **kw)
        self.max_timestep = torch.ones(B, dtype=torch.int32, **kw)
        self.counters = torch.zeros(B, 4, dtype=torch.int32, **kw)
        self.obs = torch.empty(B, self.obs_dim, dtype=dtype, **kw)
        self.reward = torch.empty(B, dtype=dtype, **kw)
        self.done = torch.empty(B, dtype=torch.uint8, **kw)
        self.flags = torch.empty(B, 8, dtype=torch.uint8, **kw)
        self.errs = torch.empty(B, 4, dtype=dtype, **kw) if want_errs else None
        self.target_poses = self.demo_configs = self.latent = None
        self._records = None
        self.Tmax = 0
        self._pre = self._post = self._info = self._bound = None
        self._base_off = np.ascontiguousarray(scene.base_offset if scene is not None else fetch.BASE_OFFSET, dtype=np.float64)

    # -- problems ------------------------------------------------------------------------------------------
    def set_problems(self, target_poses, demo_configs, max_timestep, latent=None):
        """target_poses [B,Tmax,7], demo_configs [B,Tmax,7] or None, max_timestep [B]; latent [B,zDim]."""
        conv = lambda a: None if a is None else torch.as_tensor(a).to(self.device, self.dtype).contiguous()

        tp, dc = conv(target_poses), conv(demo_configs)
        assert tp.shape[0] == self.B and tp.shape[2] == 7
        if dc is not None and not self._args["refineFlag"]:
            dc = refine_continuous_joints(dc)   # the reference refines EVERY demo row in reset (PirlEnv.py:362-365)
        # 128-byte STEP RECORDS (cfg.row_floats = 32): entry (b,t) = [target(t) | demo(t) | target(min(t+1, Tmax-1)) | pad],
        # 8 values each.  A per-thread gather costs a whole 128-byte line on the B200 whatever its size
        # (tools/probes/gather_probe.cu), so all the table rows one env step reads sit in one line; target_poses /
        # demo_configs stay 7-wide views of the records (slot 2 is derived: call set_problems again after editing them)
        Tm = tp.shape[1]
        rec = torch.zeros(self.B, Tm, 32, device=self.device, dtype=self.dtype)
        rec[..., 0:7] = tp
        if dc is not None:
            rec[..., 8:15] = dc
        rec[:, :-1, 16:23] = tp[:, 1:]
        rec[:, -1, 16:23] = tp[:, -1]
        self._records = rec
        self.target_poses = rec[..., 0:7]
        self.demo_configs = None if dc is None else rec[..., 8:15]
        self.cfg.row_floats = 32
        self.cfg.record_has_demo = 0 if dc is None else 1
        self.Tmax = int(self.target_poses.shape[1])
        self.max_timestep = torch.as_tensor(max_timestep).to(self.device, torch.int32).contiguous()
        self.latent = conv(latent)
        if self.cfg.z_dim > 0 and self.latent is None:
            self.latent = torch.zeros(self.B, self.cfg.z_dim, device=self.device, dtype=self.dtype)
        self._pre = None

    def reset(self, start_conf=None, mode="train"):
        """PirlEnv_Fetch.reset (PirlEnv.py:352-392): cur_conf = demo_configs[0] (continuous joints refined to
        (-pi, pi]) unless start_conf is given; counters and timestep cleared; returns the observation."""
        if mode != self.mode:
            self.mode = mode
            rf, hd = self.cfg.row_floats, self.cfg.record_has_demo
            self.cfg = make_cfg(mode=mode, **self._args)
            self.cfg.row_floats, self.cfg.record_has_demo = rf, hd
        if start_conf is None:
            q0 = self.demo_configs[:, 0, :].clone()   # already refined to (-pi, pi] by set_problems
        else:
            q0 = torch.as_tensor(start_conf).to(self.device, self.dtype)
        self.cur_conf.copy_(q0.reshape(self.B, 7))  # in place: captured CUDA graphs keep pointing at this buffer
        self.timestep.zero_()
        self.counters.zero_()
        self._pre = None
        return self._observe()

    def _observe(self):
        with _lib.on_device(self.device):
            check(lib.rlitg_pirl_obs(C.byref(self.cfg), self.B, _code(self.cur_conf), self.cur_conf.data_ptr(),
                                     self.timestep.data_ptr(), self.max_timestep.data_ptr(),
                                     self._records.data_ptr(), self.Tmax,
                                     None if self.latent is None else self.latent.data_ptr(), self.obs.data_ptr(),
                                     _stream()))
        return self.obs

    def _bind(self):
        """argument tuple of rlitg_pirl_env_step for the current buffers (pointers as python ints: hot path)"""
        p = lambda t: None if t is None else t.data_ptr()
        self._info = {"task_suc_count": self.counters[:, 0], "dist_suc_count": self.counters[:, 1],
                      "rot_suc_count": self.counters[:, 2], "imitate_suc_count": self.counters[:, 3],
                      "is_col": self.flags[:, 0], "is_jlimit": self.flags[:, 1], "is_early_stop": self.flags[:, 2],
                      "is_timeout": self.flags[:, 3]}
        self._pre = (C.byref(self.cfg), self.B, _code(self.cur_conf), self.cur_conf.data_ptr())
        self._post = (self.timestep.data_ptr(), self.max_timestep.data_ptr(), self._records.data_ptr(),
                      None, self.Tmax, p(self.latent),
                      None if self.scene is None else self.scene.handle, self._base_off.ctypes.data,
                      self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(), self.flags.data_ptr(),
                      self.counters.data_ptr(), p(self.errs))
        # the same argument list held by a C handle: the per-step call then marshals three arguments instead of twenty
        self._unbind()
        h = C.c_void_p()
        check(lib.rlitg_pirl_bind(*self._pre, *self._post, C.byref(h)))
        self._bound = h

    def _unbind(self):
        if getattr(self, "_bound", None):
            lib.rlitg_pirl_bound_destroy(self._bound)
        self._bound = None

    def __del__(self):
        try:
            self._unbind()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def step(self, action):
        """action: [B,7] tensor (any float dtype; converted to float32 like the policy output).
        Returns obs [B,O], reward [B], done [B] (uint8), info dict of tensors (keys of PirlEnv.py:273-282).
        The returned tensors are the env's own buffers (overwritten by the next step)."""
        a = action
        if not (type(a) is torch.Tensor and a.dtype == torch.float32 and a.is_cuda and a.is_contiguous()):
            a = torch.as_tensor(action, device=self.device).to(torch.float32).contiguous()
        if self._pre is None:
            self._bind()
        # at 16k envs the kernel takes ~5 us: the launch path is kept short (the argument tuples are bound once; no
        # stream / device objects are built when this env's device is already current)
        idx = self._dev_index
        if torch.cuda.current_device() == idx:
            rc = _step_bound(self._bound, a.data_ptr(), _raw_stream(idx))
        else:
            with torch.cuda.device(idx):
                rc = _step_bound(self._bound, a.data_ptr(), _raw_stream(idx))
        if rc:
            check(rc)
        return self.obs, self.reward, self.done, self._info

    def capture_step_graph(self, action_buffer):
        """CUDA-graph the step for launch-bound batch sizes: returns (graph, outputs).  `action_buffer` is a
        persistent float32 [B,7] CUDA tensor the caller overwrites before each `graph.replay()`."""
        assert action_buffer.dtype == torch.float32 and action_buffer.is_cuda and action_buffer.is_contiguous()
        self.step(action_buffer)  # warm-up outside capture (sets kernel attributes)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.step(action_buffer)
        return g, out

    def random_action(self):
        lo = torch.tensor(self.act_ll, device=self.device, dtype=torch.float32)
        hi = torch.tensor(self.act_ul, device=self.device, dtype=torch.float32)
        return lo + (hi - lo) * torch.rand(self.B, 7, device=self.device)


class PirlEnvFetch:
    """Batch-of-one wrapper with the reference's scalar numpy API (PirlEnv_Fetch): reset(...) -> obs,
    step(a) -> obs, reward, done, info (python scalars / numpy), float64 like the reference."""

    def __init__(self, rewardType="Task+NullSpaceImitation", obsType="Conf+LinkPoses", t=1, zDim=0, refineFlag=False,
                 scene=None, device="cuda:0"):
        self.vec = PirlVecEnv(1, rewardType, obsType, t, zDim, refineFlag, scene, device, torch.float64)
        for k in ("observation_space_shape", "action_space_shape", "act_ll", "act_ul", "n_dof", "ll", "ul", "c_joints"):
            setattr(self, k, getattr(self.vec, k))

    @property
    def timestep(self):
        return int(self.vec.timestep[0].item())

    @property
    def max_timestep(self):
        return int(self.vec.max_timestep[0].item())

    @property
    def cur_conf(self):
        return self.vec.cur_conf[0].cpu().numpy()

    def reset(self, target_poses, demo_configs, latent=None, mode="train", start_conf=None):
        tp = np.asarray(target_poses, dtype=np.float64)[None]
        dc = None if demo_configs is None else np.asarray(demo_configs, dtype=np.float64)[None]
        self.vec.set_problems(tp, dc, [tp.shape[1] - 1], None if latent is None else np.asarray(latent)[None])
        return self.vec.reset(None if start_conf is None else np.asarray(start_conf)[None], mode)[0].cpu().numpy()

    def step(self, action):
        obs, rew, done, info = self.vec.step(np.asarray(action, dtype=np.float32)[None])
        info = {k: (bool(v[0].item()) if k.startswith("is_") else int(v[0].item())) for k, v in info.items()}
        return obs[0].cpu().numpy(), float(rew[0].item()), bool(done[0].item()), info

    def random_action(self):
        return np.random.uniform(self.act_ll, self.act_ul)


def rollout_policy(env: PirlVecEnv, policy, start_conf=None, n_steps=None, deterministic=True):
    """Batched PirlInterpolator::interpolate (torm/src/pirl_interpolator.cpp:102-140): for every
    environment, observation -> policy -> setAction (add, single +-2pi wrap, clamp to the limits; no rejection),
    one waypoint per step.  `policy(obs [B,O] float32) -> action [B,7]` is any torch callable (the reference's is a
    TorchScript MLP).  Returns the generated configurations [B, n_steps, 7]."""
    obs = env.reset(start_conf, mode="test")
    n_steps = int(env.max_timestep.max().item()) if n_steps is None else n_steps
    out = torch.empty(env.B, n_steps, 7, device=env.device, dtype=env.dtype)
    with torch.no_grad():
        for k in range(n_steps):
            act = policy(obs.float())
            obs, _, _, _ = env.step(act)
            out[:, k] = env.cur_conf
    return out
