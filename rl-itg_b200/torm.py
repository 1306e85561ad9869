"""TORM path-following cost / gradient terms on the GPU, batched over N candidate trajectories.

Python facade with the reference's member names (TormOptimizer, torm/include/torm/torm_optimizer.h:295-312;
traj_evaluator, torm/include/torm/traj_evaluator.h:12-41; TormCost, torm/include/torm/torm_cost.h:49-77), each
taking / returning tensors with a leading batch axis N.  All of them run rlitg_torm_cost_grad (one kernel launch,
one CTA per trajectory) with the matching stage mask.

Trajectories are [N,T,7] with row 0 the (fixed) start configuration -- the reference's full_trajectory_; its padded
group_trajectory_ ([T+10,7], five clamped rows each side, torm/src/torm_trajectory.cpp:59-97) is implicit.
Increments are [N,T-1,7] over the free waypoints, like the reference's *_increments_ matrices."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, fetch
from ._lib import check, lib

STAGE_POSE, STAGE_POSE_GRAD, STAGE_SMOOTH, STAGE_COLL, STAGE_EVAL = 1, 2, 4, 8, 16
STAGE_KEEP_COSTS, STAGE_COLL_ROW_ONLY = 32, 64
NCOST = 10
COST_NAMES = ("pose", "rot", "vel", "acc", "jerk", "collision", "weighted", "n_vel_violation", "endpose_pos_sum",
              "endpose_rot_sum")


class RlitgTormParams(C.Structure):
    _fields_ = [("T", C.c_int32), ("D", C.c_int32), ("dt", C.c_double), ("rot_scale", C.c_double),
                ("jac_weight", C.c_double), ("clearance", C.c_double), ("pinv_ridge", C.c_double),
                ("vik_eps", C.c_double), ("use_pinv", C.c_int32), ("affine_axis_quirk", C.c_int32),
                ("base_off", C.c_double * 3), ("cjoint", C.c_int32 * 8), ("vel_limit", C.c_double * 8)]


_vp = C.c_void_p
lib.rlitg_smooth_create.argtypes = [C.c_int, C.c_double, _vp, C.c_double, C.POINTER(_vp)]
lib.rlitg_smooth_destroy.argtypes = [_vp]
lib.rlitg_smooth_get.argtypes = [_vp, _vp, C.POINTER(C.c_double), _vp]
lib.rlitg_torm_cost_grad.argtypes = [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.POINTER(RlitgTormParams), C.c_int, _vp,
                                     _vp, _vp, _vp, _vp, _vp, _vp]
lib.rlitg_torm_cost_grad_staged.argtypes = [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.POINTER(RlitgTormParams), C.c_int,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.rlitg_torm_target_frames.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp]
lib.rlitg_torm_cost_grad_frames.argtypes = [_vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.POINTER(RlitgTormParams), C.c_int,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]


@dataclass
class TormParameters:
    """torm/include/torm/torm_parameters.h:53-85; defaults = torm/src/torm_main.cpp:50-92."""
    planning_time_limit_: float = 1.0
    smoothness_update_weight_: float = 30.0  # divided by T by the mains
    obstacle_update_weight_: float = 1.0
    jacobian_update_weight_: float = 1.0
    learning_rate_: float = 0.01
    pose_cost_weight_: float = 1.0
    collision_cost_weight_: float = 0.05
    rotation_scale_factor_: float = 0.17
    smoothness_cost_velocity_: float = 1.0
    smoothness_cost_acceleration_: float = 0.0
    smoothness_cost_jerk_: float = 0.0
    use_stochastic_descent_: bool = True
    use_velocity_check_: bool = True
    use_singularity_check_: bool = False
    use_collision_check_: bool = True
    ridge_factor_: float = 0.0
    use_pseudo_inverse_: bool = True
    pseudo_inverse_ridge_factor_: float = 1e-4
    joint_update_limit_: float = 0.17
    min_clearence_: float = 0.001
    exploration_iter_: int = 100
    traj_generation_iter_: int = 50
    time_duration_: float = 0.1
    singularity_lower_bound_: float = 0.005
    # build-side switches
    affine_axis_quirk: bool = True
    vik_eps: float = 1e-5
    base_offset: tuple = tuple(fetch.BASE_OFFSET)
    continuous: tuple = tuple(int(v) for v in fetch.CONTINUOUS)
    vel_limit: tuple = tuple(fetch.VEL_LIMIT)

    def to_c(self, T: int) -> RlitgTormParams:
        p = RlitgTormParams()
        p.T, p.D = int(T), 7
        p.dt, p.rot_scale, p.jac_weight = self.time_duration_, self.rotation_scale_factor_, self.jacobian_update_weight_
        p.clearance, p.pinv_ridge, p.vik_eps = self.min_clearence_, self.pseudo_inverse_ridge_factor_, self.vik_eps
        p.use_pinv, p.affine_axis_quirk = int(self.use_pseudo_inverse_), int(self.affine_axis_quirk)
        p.base_off = (C.c_double * 3)(*self.base_offset)
        p.cjoint = (C.c_int32 * 8)(*(list(self.continuous) + [0]))
        p.vel_limit = (C.c_double * 8)(*(list(self.vel_limit) + [0.0]))
        return p


class TormCost:
    """Smoothness operator (TormCost + the scaling in TormOptimizer::initialize), device resident."""

    def __init__(self, T, dt, weights=(1.0, 0.0, 0.0), ridge=0.0, device="cuda:0"):
        self.T, self.device = int(T), torch.device(device)
        self._h = _vp()
        w = np.ascontiguousarray(weights, dtype=np.float64)
        with _lib.on_device(self.device):
            check(lib.rlitg_smooth_create(self.T, float(dt), w.ctypes.data, float(ridge), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def get(self, want_inverse=False):
        G, F = self.T + 10, self.T - 1
        band = np.empty((G, 13)); scale = C.c_double(0)
        inv = np.empty((F, F)) if want_inverse else None
        check(lib.rlitg_smooth_get(self._h, band.ctypes.data, C.byref(scale), None if inv is None else inv.ctypes.data))
        return band, scale.value, inv

    def getMaxQuadCostInvValue(self):
        return self.get()[1]

    def close(self):
        if self._h:
            lib.rlitg_smooth_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class TormBatch:
    """N candidate trajectories against one target path: the cost / increment members of TormOptimizer."""

    def __init__(self, end_poses, parameters: TormParameters | None = None, scene=None, device="cuda:0",
                 dtype=torch.float32, smoothness=None):
        """end_poses: [T,7] target path (x y z qx qy qz qw, chain-base frame) -- endPoses_desired_.
        smoothness: build the TormCost operator (default: when T >= 5; the initialisers' short segments are scored with
        the end-pose and collision terms only and need none)."""
        self.device, self.dtype = torch.device(device), dtype
        self.parameters = parameters or TormParameters()
        self.scene = scene
        self.endPoses_desired_ = torch.as_tensor(end_poses).to(self.device, dtype).contiguous()
        self.T = int(self.endPoses_desired_.shape[0])
        self.num_vars_free_ = self.T - 1
        pr = self.parameters
        if smoothness is None:
            smoothness = self.T >= 5
        self.joint_costs_ = TormCost(self.T, pr.time_duration_, (pr.smoothness_cost_velocity_,
                                     pr.smoothness_cost_acceleration_, pr.smoothness_cost_jerk_), pr.ridge_factor_,
                                     self.device) if smoothness else None
        self._cp = pr.to_c(self.T)
        self._buf = {}
        self._call_cache = {}
        # the path as frames (targets + the evaluator's inter-target poses), built once and shared by every call
        self._frames = torch.empty(2 * self.T - 1, 12, device=self.device, dtype=dtype)
        with _lib.on_device(self.device):
            check(lib.rlitg_torm_target_frames(self.endPoses_desired_.data_ptr(), self.T,
                                               _lib.F32 if dtype == torch.float32 else _lib.F64, self._frames.data_ptr(),
                                               torch.cuda.current_stream(self.device).cuda_stream))

    def _out(self, name, shape, dtype=None):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, device=self.device, dtype=dtype or self.dtype)
            self._buf[name] = t
            self._call_cache.clear()
        return t

    def cost_grad(self, traj, stage_mask, stochastic_idx=None, want_active=False, traj_stage=None):
        """One call of rlitg_torm_cost_grad.  Returns dict(costs [N,10], inc_smooth, inc_coll, inc_pose
        [N,T-1,7] or None, active [N] or None).  Output tensors are reused between calls.
        traj_stage: optional int32 [N] device tensor (0 feasible / 1 end-pose per trajectory): each trajectory only gets
        the increments its stage needs (rlitg_torm_cost_grad_staged); rows of the other stage are left untouched."""
        if not traj.is_contiguous():
            traj = traj.contiguous()
        st = None
        if stochastic_idx is not None:
            st = stochastic_idx if (torch.is_tensor(stochastic_idx) and stochastic_idx.dtype == torch.int32
                                    and stochastic_idx.is_cuda) else \
                torch.as_tensor(stochastic_idx).to(self.device, torch.int32).contiguous()
        key = (traj.data_ptr(), tuple(traj.shape), traj.dtype, int(stage_mask), None if st is None else st.data_ptr(),
               bool(want_active), None if traj_stage is None else traj_stage.data_ptr())
        hit = self._call_cache.get(key)
        if hit is None:
            N, T, D = traj.shape
            if T != self.T or D != 7:
                raise ValueError("traj must be [N,%d,7]" % self.T)
            if traj.dtype != self.dtype:
                raise TypeError("traj dtype %s does not match the batch dtype %s" % (traj.dtype, self.dtype))
            F = T - 1
            inc_s = self._out("inc_smooth", (N, F, 7)) if stage_mask & STAGE_SMOOTH else None
            inc_c = self._out("inc_coll", (N, F, 7)) if stage_mask & STAGE_COLL else None
            inc_p = self._out("inc_pose", (N, F, 7)) if stage_mask & STAGE_POSE_GRAD else None
            costs = self._out("costs", (N, NCOST))
            active = self._out("active", (N,), torch.float32) if want_active else None
            p = lambda t: None if t is None else t.data_ptr()
            code = _lib.F32 if traj.dtype == torch.float32 else _lib.F64
            args = (traj.data_ptr(), self.endPoses_desired_.data_ptr(), self._frames.data_ptr(), N, code,
                    None if self.scene is None else self.scene.handle,
                    None if self.joint_costs_ is None else self.joint_costs_.handle, C.byref(self._cp),
                    int(stage_mask), p(traj_stage), p(st), p(inc_s), p(inc_c), p(inc_p), costs.data_ptr(), p(active))
            hit = (args, dict(costs=costs, inc_smooth=inc_s, inc_coll=inc_c, inc_pose=inc_p, active=active))
            if len(self._call_cache) > 64:
                self._call_cache.clear()
            self._call_cache[key] = hit
        with _lib.on_device(self.device):
            rc = lib.rlitg_torm_cost_grad_frames(*hit[0], torch.cuda.current_stream().cuda_stream)
        if rc:
            check(rc)
        return hit[1]

    # ---- reference member names (leading batch axis) ------------------------------------------------------
    def getEndPoseCost(self, traj, grad=False):
        """torm/src/torm_optimizer.cpp:1399-1425: returns [N] = sum ||dp|| + rotation_scale_factor_ sum ||drot||;
        with grad=True also fills endPose_increments_ (returned as second value)."""
        r = self.cost_grad(traj, STAGE_POSE_GRAD if grad else STAGE_POSE)
        c = r["costs"][:, 8] + self.parameters.rotation_scale_factor_ * r["costs"][:, 9]
        return (c, r["inc_pose"]) if grad else c

    def calculateEndPoseIncrements(self, traj):
        return self.cost_grad(traj, STAGE_POSE_GRAD)["inc_pose"]

    def calculateSmoothnessIncrements(self, traj):
        return self.cost_grad(traj, STAGE_SMOOTH)["inc_smooth"]

    def calculateCollisionIncrements(self, traj, stochastic_idx=None):
        """performForwardKinematics + calculateCollisionIncrements (:1302-1382, :1484-1553)."""
        return self.cost_grad(traj, STAGE_COLL, stochastic_idx)["inc_coll"]

    def getCollisionCost(self, traj):
        """:1448-1463 over rows start_idx-1 .. end_idx (the T real rows): [N]."""
        return self.cost_grad(traj, STAGE_COLL)["costs"][:, 5]

    def evaluate(self, traj):
        """traj_evaluator::getCost (torm/src/utils/traj_evaluator.cpp:104-132): dict of [N] tensors."""
        c = self.cost_grad(traj, STAGE_EVAL)["costs"]
        return {"cost_pose": c[:, 0], "cost_rot": c[:, 1], "cost_vel": c[:, 2], "cost_acc": c[:, 3],
                "cost_jerk": c[:, 4], "weighted": c[:, 6], "n_vel_violation": c[:, 7]}

    # ---- binary feasibility, one flag per trajectory --------------------------------------------
    def isCurrentTrajectoryCollisionFree(self, traj):
        """torm/src/torm_optimizer.cpp:695-717 over all T rows: True = no waypoint collides (sphere-vs-grid stand-in for
        the FCL test; True for every trajectory when use_collision_check_ is off or there is no scene)."""
        N = traj.shape[0]
        if self.scene is None or not self.parameters.use_collision_check_:
            return torch.ones(N, dtype=torch.bool, device=self.device)
        return self.scene.collide(traj.reshape(-1, 7)).reshape(N, self.T).sum(dim=1) == 0

    def checkJointVelocityLimit(self, traj):
        """:720-745: True = every consecutive pair of rows stays within vel_limit * time_duration_ (continuous joints by
        shortest angular distance); always True when use_velocity_check_ is off."""
        if not self.parameters.use_velocity_check_:
            return torch.ones(traj.shape[0], dtype=torch.bool, device=self.device)
        return self.cost_grad(traj, STAGE_EVAL)["costs"][:, 7] == 0

    def checkSingularity(self, traj):
        """:748-765: True = sqrt(det(J J^T)) >= singularity_lower_bound_ on every free row; always True when
        use_singularity_check_ is off."""
        N = traj.shape[0]
        if not self.parameters.use_singularity_check_:
            return torch.ones(N, dtype=torch.bool, device=self.device)
        from .pirl_env import manipulability
        yosh = manipulability(traj[:, 1:].reshape(-1, 7).contiguous()).reshape(N, self.T - 1)
        return ~(yosh.abs() < self.parameters.singularity_lower_bound_).any(dim=1)

    def getWeightedPoseCost(self, traj):
        return self.cost_grad(traj, STAGE_EVAL)["costs"][:, 6]


# ------------------------------------------------------------------------------------------------ update step
class RlitgTormUpdate(C.Structure):
    _fields_ = [("stage", C.c_int32), ("lr_schedule_mode", C.c_int32), ("smoothness_update_weight", C.c_double),
                ("obstacle_update_weight", C.c_double), ("joint_update_limit", C.c_double),
                ("joint_min", C.c_double * 7), ("joint_max", C.c_double * 7)]


lib.rlitg_torm_update.argtypes = [_vp, C.c_int64, C.c_int, _vp, C.POINTER(RlitgTormParams), C.POINTER(RlitgTormUpdate),
                                  _vp, _vp, _vp, _vp, _vp, _vp]


class RlitgTormSchedule(C.Structure):
    _fields_ = [("lr_schedule_mode", C.c_int32), ("stage", C.c_int32), ("use_collision_check", C.c_int32),
                ("use_velocity_check", C.c_int32), ("learning_rate", C.c_double), ("learning_rate_min", C.c_double),
                ("learning_rate_max", C.c_double), ("hits_per_trajectory", C.c_int32), ("reserved_", C.c_int32)]


lib.rlitg_torm_stage_decide.argtypes = [C.c_int64, C.c_int, _vp, _vp, _vp, C.POINTER(RlitgTormSchedule), _vp, _vp]
lib.rlitg_torm_bookkeep.argtypes = [_vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    C.POINTER(RlitgTormSchedule), _vp]
lib.rlitg_torm_update_staged.argtypes = [_vp, C.c_int64, C.c_int, _vp, C.POINTER(RlitgTormParams),
                                         C.POINTER(RlitgTormUpdate), _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.rlitg_torm_collision_point.argtypes = [_vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]


def _apply_increments(self, traj, feasible_stage, inc_smooth=None, inc_coll=None, inc_pose=None, alpha=None,
                      learning_rate=None, lr_schedule_mode=1, stage_idx=None):
    """calculateFeasibleIncrements | calculateEndPoseIncrements, then addIncrementsToTrajectory and
    handleJointLimits (torm/src/torm_optimizer.cpp:1555-1639, 1189-1268), in place on traj [N,T,7].
    alpha / learning_rate: python floats or [N] tensors (per-trajectory alpha_ / learning_rate_).
    stage_idx: optional int32 [N] device tensor choosing the stage per trajectory (0 feasible, 1 end-pose, < 0 skip);
    `feasible_stage` is then ignored (rlitg_torm_update_staged)."""
    pr = self.parameters
    N = traj.shape[0]
    ukey = (bool(feasible_stage), int(lr_schedule_mode))
    u = self._upd_cache.get(ukey) if hasattr(self, "_upd_cache") else None
    if u is None:  # the struct only depends on (stage, mode) and the parameters: built once (hot loop)
        u = RlitgTormUpdate()
        u.stage, u.lr_schedule_mode = (0 if feasible_stage else 1), int(lr_schedule_mode)
        u.smoothness_update_weight = pr.smoothness_update_weight_ / self.T  # the mains pass 30 / T
        u.obstacle_update_weight = pr.obstacle_update_weight_
        # TormOptimizer's constructor overwrites joint_update_limit_ when lr_schedule_mode_ == 0 (:104-106)
        u.joint_update_limit = 0.1 if lr_schedule_mode == 0 else pr.joint_update_limit_
        u.joint_min = (C.c_double * 7)(*fetch.LL)
        u.joint_max = (C.c_double * 7)(*fetch.UL)
        if not hasattr(self, "_upd_cache"):
            self._upd_cache = {}
        self._upd_cache[ukey] = u

    def vec(x, default):
        if x is None:
            x = default
        if torch.is_tensor(x):
            if x.device == self.device and x.dtype == traj.dtype and x.is_contiguous():
                return x
            return x.to(self.device, traj.dtype).contiguous()
        return torch.full((N,), float(x), device=self.device, dtype=traj.dtype)

    a = vec(alpha, 0.5)
    lr = vec(learning_rate, pr.learning_rate_)
    p = lambda t: None if t is None else t.data_ptr()
    code = _lib.F32 if traj.dtype == torch.float32 else _lib.F64
    assert traj.is_contiguous()
    with _lib.on_device(self.device):
        if stage_idx is None:
            check(lib.rlitg_torm_update(traj.data_ptr(), N, code, self.joint_costs_.handle, C.byref(self._cp), C.byref(u),
                                        p(inc_smooth), p(inc_coll), p(inc_pose), a.data_ptr(), lr.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
        else:
            assert stage_idx.dtype == torch.int32 and stage_idx.is_cuda and stage_idx.is_contiguous()
            check(lib.rlitg_torm_update_staged(traj.data_ptr(), N, code, self.joint_costs_.handle, C.byref(self._cp),
                                               C.byref(u), stage_idx.data_ptr(), p(inc_smooth), p(inc_coll), p(inc_pose),
                                               a.data_ptr(), lr.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return traj


TormBatch.applyIncrements = _apply_increments


def _local_optimize_tsgd(self, traj, maxiter, learning_rate=None, alpha=0.5, check_collision=True, lr_schedule_mode=3,
                         return_state=False, use_cuda_graph=None, stochastic=None, seed=0, rand=None, state=None, overlap=True,
                         graph_iterations=8, parts=None):
    """Lock-step two-stage gradient descent over the N trajectories of `traj` (in place): the loop body of
    TormOptimizer::localOptimizeTSGD (torm/src/torm_optimizer.cpp:397-598) entirely on the device -- cost +
    increments (kernels 3a / 3b), feasibility checks, the alpha_ / learning_rate_ schedule, the best-trajectory
    bookkeeping of :506-560 (a trajectory is recorded when its evaluator cost improves and it is collision free,
    within the velocity limits and, with use_singularity_check_, away from singularities) and the update step.
    Every trajectory carries its own optimiser state as [N] tensors:
      lr_schedule_mode 0  fixed weights (smoothness_update_weight_ / obstacle_update_weight_), joint_update_limit 0.1
                       1  adaptive: alpha_ +-0.1 steps on the feasibility pattern (:447-466), learning_rate_ halved /
                          doubled inside [lr/100, lr*1000] (:471-492), and the stage of each trajectory chosen from its
                          own history (:400-412) -- the batch then mixes feasible and end-pose updates
                       2  alpha_ as in 1, cosine learning rate with period maxiter / 2 (:493-496)
                       3  alpha_ = 0.5, fixed learning rate (:467-468, :497-498)
    stochastic (default: parameters.use_stochastic_descent_, true in every reference main): the collision increment of a
    feasible half-iteration takes ONE waypoint per trajectory (:1502-1515) -- collision_point_, the first colliding row of
    this half-iteration's isCurrentTrajectoryCollisionFree (:695-717), else a uniformly drawn free row.  The draws come
    from `rand` ([maxiter,2,N] float32 uniforms on the device; default: torch.rand seeded with `seed`; the reference
    calls random()); rlitg_torm_collision_point turns them into indices.  An index of -1 (the start row collides: the
    reference then writes row -1 of the increment matrix) takes no increment.
    state: dict(alpha=[N], learning_rate=[N]) carried over from a previous call -- alpha_ / learning_rate_ are members of
    the reference's optimiser and persist across localOptimizeTSGD calls (infea_col / infea_vel are locals, :383).
    There is no host synchronisation inside an iteration, so the loop can also be captured in a CUDA graph and replayed
    (use_cuda_graph=True; identical results): one graph holds `graph_iterations` iterations -- their draws and scheduled
    rates sit in a block buffer refilled by one device copy per replay -- so that the graph-to-graph boundary and the
    per-iteration refill are paid once per block; the iterations that do not fill a block run eagerly.
    overlap=False keeps every launch on one stream (measurement).
    parts: the batch is cut into `parts` contiguous slices that run the same loop on streams of their own (trajectories
    are independent, so the results are bit-identical): the latency-bound launches of one slice (update step,
    bookkeeping, collision point: at most one wave of CTAs each) then run under the throughput-bound cost kernels of
    another.  Only pays without the per-launch Python cost, i.e. inside a captured block -- default 1.
    Returns (best_traj [N,T,7], best_cost [N]); best_cost stays +inf for trajectories never found feasible.
    return_state=True appends a dict with the final alpha / lr / feasibility tensors and the draws."""
    pr = self.parameters
    N = traj.shape[0]
    dev, dt = self.device, traj.dtype
    assert traj.is_contiguous()
    mode = int(lr_schedule_mode)
    lr0 = float(pr.learning_rate_ if learning_rate is None else learning_rate)
    lr_max, lr_min = lr0 * 1000.0, lr0 / 100.0                      # constructor :99-101
    has_scene = self.scene is not None
    if stochastic is None:
        stochastic = bool(pr.use_stochastic_descent_)
    stochastic = bool(stochastic and has_scene)
    # ---- per-trajectory optimiser state, updated in place (the captured graph keeps pointing at these buffers)
    best_cost = torch.full((N,), float("inf"), device=dev, dtype=dt)
    best_traj = traj.clone()
    if state is not None and "learning_rate" in state:
        lr = state["learning_rate"].to(dev, dt).contiguous()
        al = state["alpha"].to(dev, dt).contiguous()
    else:
        lr = torch.full((N,), lr0, device=dev, dtype=dt)
        al = torch.full((N,), float(alpha), device=dev, dtype=dt)
    infea_col = torch.zeros(N, dtype=torch.uint8, device=dev)        # :383
    infea_vel = torch.zeros(N, dtype=torch.uint8, device=dev)
    stage_idx = torch.zeros(N, dtype=torch.int32, device=dev)
    KB = max(int(graph_iterations), 1) if use_cuda_graph else 1      # iterations per block (= per graph replay)
    lr_blk = torch.full((KB,), lr0, device=dev, dtype=dt)            # mode 2: the block's scheduled rates
    if mode == 2:                                                    # cosine with period maxiter / 2 (:493-496)
        period = max(maxiter // 2, 1)
        ph = (np.arange(max(maxiter, 1)) % period) / period
        lr_sched = torch.tensor(lr_min + 0.5 * (lr_max - lr_min) * (1.0 + np.cos(ph * np.pi)), device=dev, dtype=dt)
    zeros = None if has_scene else torch.zeros(N, self.T - 1, 7, device=dev, dtype=dt)

    if use_cuda_graph is None:
        use_cuda_graph = False
    sch = [RlitgTormSchedule(), RlitgTormSchedule()]
    for k in (0, 1):
        sch[k].lr_schedule_mode, sch[k].stage = mode, k
        sch[k].use_collision_check = int(has_scene and check_collision and pr.use_collision_check_)
        sch[k].use_velocity_check = int(pr.use_velocity_check_)
        sch[k].learning_rate, sch[k].learning_rate_min, sch[k].learning_rate_max = lr0, lr_min, lr_max
        sch[k].hits_per_trajectory = 1
    code = _lib.F32 if dt == torch.float32 else _lib.F64
    singular = torch.zeros(N, dtype=torch.uint8, device=dev) if pr.use_singularity_check_ else None
    need_point = bool(sch[0].use_collision_check or stochastic)
    any_hit = torch.zeros(N, dtype=torch.uint8, device=dev) if need_point else None
    stoch_idx = torch.zeros(N, dtype=torch.int32, device=dev) if stochastic else None
    if stochastic:
        if rand is None:
            g = torch.Generator(device=dev)
            g.manual_seed(int(seed))
            rand = torch.rand((max(maxiter, 1), 2, N), generator=g, device=dev, dtype=torch.float32)
        assert rand.shape == (max(maxiter, 1), 2, N) and rand.dtype == torch.float32 and rand.is_cuda
        u_blk = torch.empty(KB, 2, N, dtype=torch.float32, device=dev)   # the block's draws (fixed address: graph replay)
    base_off = (C.c_double * 3)(*pr.base_offset)

    # Every half-iteration has two independent parts that only read the trajectory: the end-pose / evaluator kernel, and
    # the collision-point test followed by the collision / smoothness kernel (which needs the test's waypoint choice).
    # They run on two streams and share one costs buffer (RLITG_STAGE_KEEP_COSTS).
    KEEP = STAGE_KEEP_COSTS
    sl_ = lambda t, s: None if t is None else t[s]

    def make_part(tb, s):
        """The loop body on trajectories traj[s]: `tb` is a view of this batch with output buffers of its own."""
        traj_p, best_traj_p, best_cost_p = traj[s], best_traj[s], best_cost[s]
        lr_p, al_p, infea_col_p, infea_vel_p, stage_idx_p = lr[s], al[s], infea_col[s], infea_vel[s], stage_idx[s]
        singular_p, any_hit_p, stoch_idx_p, zeros_p = sl_(singular, s), sl_(any_hit, s), sl_(stoch_idx, s), sl_(zeros, s)
        Np = traj_p.shape[0]
        side = torch.cuda.Stream(device=dev)
        ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

        def sub_step(stage, k):
            cur_stream = torch.cuda.current_stream(dev)
            if mode == 1:
                # every optimiser picks its own stage (:400-412) from the state before this half-iteration; the cost
                # kernels then only do, per trajectory, what that stage needs
                with _lib.on_device(dev):
                    check(lib.rlitg_torm_stage_decide(Np, code, lr_p.data_ptr(), infea_col_p.data_ptr(),
                                                      infea_vel_p.data_ptr(), C.byref(sch[stage]), stage_idx_p.data_ptr(),
                                                      cur_stream.cuda_stream))
                ts = stage_idx_p
                pose_mask, coll_mask = STAGE_POSE_GRAD | STAGE_EVAL, STAGE_SMOOTH | (STAGE_COLL if has_scene else 0)
            elif stage == 0:
                ts = None
                pose_mask, coll_mask = STAGE_EVAL, STAGE_SMOOTH | (STAGE_COLL if has_scene else 0)
            else:
                ts = None
                pose_mask, coll_mask = STAGE_POSE_GRAD | STAGE_EVAL, 0
            # the smoothness band rides in the end-pose / evaluator launch (it has the trajectory window in shared memory
            # anyway); what is left for the second call is the collision stage
            smooth_mask, coll_mask = coll_mask & STAGE_SMOOTH, coll_mask & ~STAGE_SMOOTH
            two = bool(overlap and (coll_mask or need_point))
            if two:
                ev_fork.record(cur_stream)
                side.wait_event(ev_fork)
                with torch.cuda.stream(side):
                    r = tb.cost_grad(traj_p, pose_mask | smooth_mask | (KEEP if coll_mask else 0), traj_stage=ts)
                    ev_join.record(side)
            else:
                r = tb.cost_grad(traj_p, pose_mask | smooth_mask | (KEEP if coll_mask else 0), traj_stage=ts)
            if need_point:
                with _lib.on_device(dev):
                    check(lib.rlitg_torm_collision_point(traj_p.data_ptr(), Np, self.T, code, self.scene.handle, base_off,
                                                         u_blk[k, stage][s].data_ptr() if stochastic else None, None,
                                                         any_hit_p.data_ptr(),
                                                         stoch_idx_p.data_ptr() if stochastic else None,
                                                         cur_stream.cuda_stream))
            rc = None
            if coll_mask:
                use_st = bool(stochastic and (coll_mask & STAGE_COLL))
                # one waypoint per trajectory takes the collision increment: only its seven stencil rows are walked
                rc = tb.cost_grad(traj_p, coll_mask | KEEP | (STAGE_COLL_ROW_ONLY if use_st else 0),
                                  stochastic_idx=stoch_idx_p if use_st else None, traj_stage=ts)
            if two:
                cur_stream.wait_event(ev_join)
            if singular_p is not None:
                from .pirl_env import manipulability
                yosh = manipulability(traj_p[:, 1:].reshape(-1, 7)).reshape(Np, self.T - 1)
                singular_p.copy_((yosh.abs() < pr.singularity_lower_bound_).any(dim=1))
            # stage choice (mode 1), feasibility flags, alpha_ / learning_rate_ and the best-trajectory record: one launch
            with _lib.on_device(dev):
                check(lib.rlitg_torm_bookkeep(traj_p.data_ptr(), best_traj_p.data_ptr(), Np, self.T, code,
                                              r["costs"].data_ptr(),
                                              any_hit_p.data_ptr() if sch[stage].use_collision_check else None,
                                              None if singular_p is None else singular_p.data_ptr(), al_p.data_ptr(),
                                              lr_p.data_ptr(), infea_col_p.data_ptr(), infea_vel_p.data_ptr(),
                                              best_cost_p.data_ptr(), None,
                                              lr_blk[k:k + 1].data_ptr() if mode == 2 else None, C.byref(sch[stage]),
                                              cur_stream.cuda_stream))
            # ---- increments -> trajectory (:562-576)
            if smooth_mask:
                inc_s = r["inc_smooth"]
                inc_c = rc["inc_coll"] if (rc is not None and rc["inc_coll"] is not None) else zeros_p
            else:
                inc_s = inc_c = None
            if mode == 1:
                tb.applyIncrements(traj_p, True, inc_s, inc_c, r["inc_pose"], al_p, lr_p, lr_schedule_mode=mode,
                                   stage_idx=stage_idx_p)
            elif stage == 0:
                tb.applyIncrements(traj_p, True, inc_s, inc_c, None, al_p, lr_p, lr_schedule_mode=mode)
            else:
                tb.applyIncrements(traj_p, False, inc_pose=r["inc_pose"], alpha=al_p, learning_rate=lr_p,
                                   lr_schedule_mode=mode)

        def iteration(k=0):
            sub_step(0, k)
            sub_step(1, k)

        return iteration

    # slices of the batch (boundaries at multiples of 16 trajectories: every slice keeps the alignment of the whole)
    n_parts = max(int(parts or 1), 1)
    per = -(-N // n_parts)
    per = -(-per // 16) * 16
    bounds = [(h, min(h + per, N)) for h in range(0, N, per)] if n_parts > 1 else [(0, N)]
    if len(bounds) == 1:
        part_iters = [make_part(self, slice(0, N))]
    else:
        import copy
        part_iters = []
        for (h0, h1) in bounds:
            tb = copy.copy(self)
            tb._buf, tb._call_cache = {}, {}
            part_iters.append(make_part(tb, slice(h0, h1)))
        part_streams = [torch.cuda.Stream(device=dev) for _ in bounds]
        ev_begin = torch.cuda.Event()
        ev_done = [torch.cuda.Event() for _ in bounds]

    def run_iterations(ks):
        """Iterations ks (indices into the block buffers) of every slice: the slices advance independently."""
        if len(part_iters) == 1:
            for k in ks:
                part_iters[0](k)
            return
        cur = torch.cuda.current_stream(dev)
        ev_begin.record(cur)
        for it, ps, ev in zip(part_iters, part_streams, ev_done):
            ps.wait_event(ev_begin)
            with torch.cuda.stream(ps):
                for k in ks:
                    it(k)
                ev.record(ps)
        for ev in ev_done:
            cur.wait_event(ev)

    def set_block(it0, n):
        # the draws / scheduled rates of iterations it0 .. it0 + n - 1 into the block buffers (device-to-device)
        if mode == 2:
            lr_blk[:n].copy_(lr_sched[it0:it0 + n])
        if stochastic:
            u_blk[:n].copy_(rand[it0:it0 + n])

    start = 0
    if use_cuda_graph and maxiter >= 1 + KB:
        # the first iteration runs eagerly: it is the warm-up a capture needs (kernel attributes, lazily built scene
        # volumes, output buffers of cost_grad) and it is real work, not thrown away
        set_block(0, 1)
        run_iterations([0])
        start = 1
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        cap_stream = torch.cuda.Stream(device=dev)
        cap_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap_stream):
            with torch.cuda.graph(graph, stream=cap_stream):
                run_iterations(list(range(KB)))
        torch.cuda.current_stream(dev).wait_stream(cap_stream)
        while start + KB <= maxiter:
            set_block(start, KB)
            graph.replay()
            start += KB
    for it in range(start, maxiter):
        set_block(it, 1)
        run_iterations([0])
    if return_state:
        return best_traj, best_cost, dict(alpha=al, learning_rate=lr, infea_col=infea_col.bool(), infea_vel=infea_vel.bool(),
                                          rand=rand if stochastic else None)
    return best_traj, best_cost


TormBatch.localOptimizeTSGD = _local_optimize_tsgd


def _iterative_exploration(self, init_trajs, new_trajectories=None, max_epochs=1, lr_schedule_mode=0,
                           return_threshold=-1.0, planning_time_limit=None, check_collision=True, seed=0):
    """TormOptimizer::iterativeExploration (torm/src/torm_optimizer.cpp:604-688) with the restarts as the batch axis:
    every row of `init_trajs` [N,T,7] is one optimiser's initial trajectory (Greedy / Linear initialiser, PIRL roll-out,
    ...).  Per epoch: localOptimizeTSGD(exploration_iter_) on all N (:634), stop when the wall clock passes
    `planning_time_limit` (:635) or any best cost is below `return_threshold` (:644-647), else replace the working
    trajectories by `new_trajectories(epoch) -> [N,T,7]` (the reference draws a new initial trajectory, :649-660) and go
    on; every optimiser keeps its best trajectory across epochs (best_trajectory_backup_cost_).  With
    new_trajectories=None this is `only_first`.  Returns dict(best_traj [N,T,7], best_cost [N], success [N] bool,
    argmin int, epochs int)."""
    import time as _time
    pr = self.parameters
    traj = init_trajs.clone()
    N = traj.shape[0]
    best_cost = torch.full((N,), float("inf"), device=self.device, dtype=traj.dtype)
    best_traj = traj.clone()
    t0 = _time.perf_counter()
    epochs = 0
    st = None  # alpha_ / learning_rate_ are optimiser members: they carry over from one localOptimizeTSGD call to the next
    for eph in range(max_epochs):
        bt, bc, st = self.localOptimizeTSGD(traj, pr.exploration_iter_, lr_schedule_mode=lr_schedule_mode,
                                            check_collision=check_collision, return_state=True, state=st, seed=seed + eph)
        better = bc < best_cost
        best_cost = torch.where(better, bc, best_cost)
        best_traj = torch.where(better[:, None, None], bt, best_traj)
        epochs = eph + 1
        if planning_time_limit is not None and _time.perf_counter() - t0 > planning_time_limit:
            break
        if return_threshold != -1 and bool((best_cost < return_threshold).any()):
            break
        if new_trajectories is None or eph + 1 >= max_epochs:
            break
        traj = new_trajectories(eph + 1).to(self.device, traj.dtype).contiguous().clone()
    return dict(best_traj=best_traj, best_cost=best_cost, success=torch.isfinite(best_cost),
                argmin=int(torch.argmin(best_cost).item()), epochs=epochs)


TormBatch.iterativeExploration = _iterative_exploration


lib.rlitg_torm_cost_grad_host.argtypes = [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.POINTER(RlitgTormParams), C.c_int, _vp,
                                          _vp, _vp, _vp, _vp]


def _cost_grad_host(self, traj, stage_mask, stochastic_idx=None, out=None):
    """rlitg_torm_cost_grad_host: `traj` is a HOST numpy array [N,T,7] (float32 / float64 matching the batch dtype);
    results come back as numpy arrays.  The end-to-end form for callers that keep trajectories on the host.
    out: optional dict of preallocated host arrays (inc_smooth / inc_coll / inc_pose [N,T-1,7], costs [N,10]) to fill --
    pinned ones (torch.empty(...).pin_memory().numpy()) give full PCIe bandwidth."""
    npdt = np.float32 if self.dtype == torch.float32 else np.float64
    traj = np.ascontiguousarray(traj, dtype=npdt)
    N, T, _ = traj.shape
    F = T - 1
    if getattr(self, "_targets_host", None) is None:
        self._targets_host = self.endPoses_desired_.cpu().numpy().astype(npdt)
    targets = self._targets_host
    out = dict(out) if out is not None else {}
    for k, bit in (("inc_smooth", STAGE_SMOOTH), ("inc_coll", STAGE_COLL), ("inc_pose", STAGE_POSE_GRAD)):
        if not (stage_mask & bit):
            out[k] = None
        elif out.get(k) is None:
            out[k] = np.empty((N, F, 7), dtype=npdt)
    if out.get("costs") is None:
        out["costs"] = np.empty((N, NCOST), dtype=npdt)
    st = None if stochastic_idx is None else np.ascontiguousarray(stochastic_idx, dtype=np.int32)
    p = lambda a: None if a is None else a.ctypes.data
    with _lib.on_device(self.device):
        check(lib.rlitg_torm_cost_grad_host(traj.ctypes.data, targets.ctypes.data, N,
                                            _lib.F32 if npdt == np.float32 else _lib.F64,
                                            None if self.scene is None else self.scene.handle,
                                            None if self.joint_costs_ is None else self.joint_costs_.handle,
                                            C.byref(self._cp), int(stage_mask), p(st), p(out["inc_smooth"]),
                                            p(out["inc_coll"]), p(out["inc_pose"]), out["costs"].ctypes.data))
    return out


TormBatch.cost_grad_host = _cost_grad_host
