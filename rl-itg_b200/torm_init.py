"""Greedy / Linear initial-trajectory generators of TORM on the GPU.

Host control flow follows TormOptimizer::getNewTrajectory (torm/src/torm_optimizer.cpp:939-1018, "Greedy") and
getJointInterpolatedTrajectory (:1021-1109, "Linear"); the per-candidate work -- fillInLinearInterpolation
(:784-807), the ranged getEndPoseCost (:1427-1446), performForwardKinematics + getCollisionCost (:1271-1300,
:1448-1463) -- runs for all candidates of a segment in one rlitg_torm_cost_grad launch, and the sequential
selection rule (with its pruning quirk) is replayed by rlitg_greedy_select.

IK candidates are an INPUT, as in the reference where they come from TRAC-IK random restarts (:963-970).
`dls_ik_candidates` is this build's own batched damped-least-squares generator on top of rlitg_pose_diff /
rlitg_pinv_step: new work, checked only by FK(candidate) ~ target (not a parity target).
After the segments are chosen the reference calls handleJointLimits (:1016, :1107): pass the TormBatch of the problem
as `batch` and the initialisers apply it (the update step with zero increments)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, fetch, pirl_env
from ._lib import check, lib
from .torm import NCOST, STAGE_COLL, STAGE_POSE, RlitgTormParams, TormParameters

lib.rlitg_greedy_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.c_void_p, C.c_void_p, C.c_void_p]


def simplified_points(T: int, gap: int = 10):
    """torm/src/torm_main.cpp:135-146: every gap-th target pose, always including the last one."""
    pts = list(range(gap, T, gap))
    if not pts or pts[-1] != T - 1:
        pts.append(T - 1)
    return pts


def fill_in_linear_interpolation(start_conf, goal_confs, num: int, continuous=fetch.CONTINUOUS):
    """fillInLinearInterpolation (:784-807) for C goals at once: rows 1..num of [C, num+1, 7] (row 0 = start).
    Continuous joints take the short way round."""
    st = start_conf.reshape(-1, 7)  # one start for all goals, or one per goal
    diff = goal_confs - st
    cj = torch.as_tensor(np.asarray(continuous, dtype=bool), device=goal_confs.device).reshape(1, 7)
    two_pi = 2 * np.pi
    wrapped = torch.where(diff < 0, two_pi - diff.abs(), -two_pi + diff.abs())
    diff = torch.where(cj & (diff.abs() > np.pi), wrapped, diff)
    theta = diff / float(num)
    t = torch.arange(0, num + 1, device=goal_confs.device, dtype=goal_confs.dtype).reshape(1, num + 1, 1)
    return st.reshape(-1, 1, 7) + t * theta.reshape(-1, 1, 7)


class SegmentScorer:
    """Scores C interpolated segments against a slice of the target path (pose + collision cost)."""

    def __init__(self, parameters: TormParameters, scene, device, dtype):
        self.parameters, self.scene = parameters, scene
        self.device, self.dtype = torch.device(device), dtype

    def score(self, segs, targets):
        """segs [C,L,7] (row 0 = segment start), targets [L,7] -> costs [C,10]."""
        segs = segs.contiguous()
        Cn, L, _ = segs.shape
        cp = self.parameters.to_c(L)
        costs = torch.empty(Cn, NCOST, device=self.device, dtype=self.dtype)
        mask = STAGE_POSE | (STAGE_COLL if self.scene is not None else 0)
        code = _lib.F32 if self.dtype == torch.float32 else _lib.F64
        tg = targets.contiguous()
        with _lib.on_device(self.device):
            check(lib.rlitg_torm_cost_grad(segs.data_ptr(), tg.data_ptr(), Cn, code,
                                           None if self.scene is None else self.scene.handle, None, C.byref(cp), mask,
                                           None, None, None, None, costs.data_ptr(), None,
                                           torch.cuda.current_stream().cuda_stream))
        return costs

    def select(self, costs, L):
        pr = self.parameters
        Cn = costs.shape[0]
        idx = torch.empty(1, dtype=torch.int32, device=self.device)
        best = torch.empty(1, dtype=self.dtype, device=self.device)
        code = _lib.F32 if self.dtype == torch.float32 else _lib.F64
        with _lib.on_device(self.device):
            check(lib.rlitg_greedy_select(costs.data_ptr(), Cn, L, code, pr.rotation_scale_factor_,
                                          pr.pose_cost_weight_, pr.collision_cost_weight_, idx.data_ptr(),
                                          best.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return idx, best


class TrajectoryInitializer:
    """getNewTrajectory (Greedy) and getJointInterpolatedTrajectory (Linear) for one problem."""

    def __init__(self, end_poses, parameters: TormParameters | None = None, scene=None, device="cuda:0",
                 dtype=torch.float32, gap=10):
        self.device, self.dtype = torch.device(device), dtype
        self.parameters = parameters or TormParameters()
        self.endPoses_desired_ = torch.as_tensor(end_poses).to(self.device, dtype).contiguous()
        self.T = int(self.endPoses_desired_.shape[0])
        self.simplified_points_ = simplified_points(self.T, gap)
        self.scorer = SegmentScorer(self.parameters, scene, self.device, dtype)

    def _handle_joint_limits(self, traj, batch):
        """handleJointLimits at the end of both initialisers (torm/src/torm_optimizer.cpp:1016, :1107): the update step with
        zero increments, on the TormBatch of the same problem."""
        if batch is not None:
            t = traj.unsqueeze(0).contiguous()
            batch.applyIncrements(t, False, inc_pose=torch.zeros(1, self.T - 1, 7, device=self.device, dtype=self.dtype))
            traj = t[0]
        return traj

    def getNewTrajectory(self, start_conf, candidates, batch=None):
        """Greedy.  start_conf [7]; candidates: list (one entry per simplified point) of [C_i,7] tensors of IK
        solutions for endPoses_desired_[simplified_points_[i]].  batch: the TormBatch of this problem -- handleJointLimits
        (:1016) is then applied to the result.  Returns (traj [T,7], chosen indices, best costs)."""
        traj = torch.empty(self.T, 7, device=self.device, dtype=self.dtype)
        traj[0] = torch.as_tensor(start_conf).to(self.device, self.dtype)
        start_point = 0
        chosen, bests = [], []
        for i, sp in enumerate(self.simplified_points_):
            cand = torch.as_tensor(candidates[i]).to(self.device, self.dtype)
            num = sp - start_point
            segs = fill_in_linear_interpolation(traj[start_point], cand, num, self.parameters.continuous)
            costs = self.scorer.score(segs, self.endPoses_desired_[start_point:sp + 1])
            idx, best = self.scorer.select(costs, num + 1)
            traj[start_point + 1:sp + 1] = segs[idx.long()[0], 1:]
            chosen.append(idx); bests.append(best)
            start_point = sp
        return self._handle_joint_limits(traj, batch), torch.cat(chosen), torch.cat(bests)

    def getJointInterpolatedTrajectory(self, start_conf, goal_candidates, batch=None):
        """Linear.  goal_candidates [C,7]: IK solutions of the last target pose (C = 1 with
        use_shortest_goal_conf, 30 otherwise).  batch: as in getNewTrajectory (:1107).
        Returns (traj [T,7], chosen index, best cost)."""
        st = torch.as_tensor(start_conf).to(self.device, self.dtype)
        cand = torch.as_tensor(goal_candidates).to(self.device, self.dtype).reshape(-1, 7)
        segs = fill_in_linear_interpolation(st, cand, self.T - 1, self.parameters.continuous)
        costs = self.scorer.score(segs, self.endPoses_desired_)
        idx, best = self.scorer.select(costs, self.T)
        return self._handle_joint_limits(segs[idx.long()[0]].clone(), batch), idx, best


def set_pirl_trajectory(end_poses, start_confs, policy, obsType="Conf+LinkPoses", t=1, gap=10, batch=None,
                        device="cuda:0", dtype=torch.float32):
    """TormOptimizer::setPIRLTrajectory (torm/src/torm_optimizer.cpp:898-936) for B start configurations of one problem
    at once (interpolation_mode_ 2, the RL-ITG initial trajectory): PirlInterpolator::interpolate
    (torm/src/pirl_interpolator.cpp:102-140, multi_action = 1) rolls the policy out over the simplified points -- one
    fused env step per point for all B -- the generated configurations become the rows at the simplified points
    (:922), fillInLinearInterpolation (:784-807) fills the rows between them, and handleJointLimits (:934) is applied
    when a TormBatch for the same problem is passed as `batch`.  `policy(obs [B,O] float32) -> action [B,7]`.
    Returns trajectories [B,T,7] (row 0 = start configuration)."""
    device = torch.device(device)
    poses = torch.as_tensor(end_poses).to(device, dtype).contiguous()
    T = int(poses.shape[0])
    sp = simplified_points(T, gap)
    st = torch.as_tensor(start_confs).to(device, dtype).reshape(-1, 7).contiguous()
    B = st.shape[0]
    # the env indexes its path from the start pose: row 0 = first target pose, rows 1.. = the simplified points
    path = torch.cat([poses[:1], poses[torch.as_tensor(sp, device=device)]], dim=0)
    env = pirl_env.PirlVecEnv(B, rewardType="Task", obsType=obsType, t=t, zDim=0, device=device, dtype=dtype)
    env.set_problems(path.unsqueeze(0).expand(B, -1, -1).contiguous(), None, np.full(B, len(sp), dtype=np.int32))
    gen = pirl_env.rollout_policy(env, policy, start_conf=st, n_steps=len(sp))   # [B, S, 7]
    traj = torch.empty(B, T, 7, device=device, dtype=dtype)
    traj[:, 0] = st
    start_point = 0
    for i, s_i in enumerate(sp):
        seg = fill_in_linear_interpolation(traj[:, start_point], gen[:, i], s_i - start_point)
        traj[:, start_point + 1:s_i + 1] = seg[:, 1:]   # the fill rewrites the goal row too: start + num * theta (:803-805)
        start_point = s_i
    if batch is not None:
        batch.applyIncrements(traj, False, inc_pose=torch.zeros(B, T - 1, 7, device=device, dtype=dtype))
    return traj


def dls_ik_candidates(target_pose, n, seed=0, iters=60, step_limit=0.3, tol_pos=1e-4, tol_rot=1e-3, device="cuda:0",
                      dtype=torch.float32, seeds=None):
    """Batched IK by truncated-SVD / QR pseudo-inverse steps from n random seeds (a GPU stand-in for the
    reference's TRAC-IK random restarts): q <- clip(q + clamp(pinv(J) twist)).  Returns (q [n,7], ok [n] bool)."""
    dev = torch.device(device)
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    ll = torch.tensor(fetch.LL, dtype=dtype, device=dev); ul = torch.tensor(fetch.UL, dtype=dtype, device=dev)
    if seeds is None:
        q = (torch.rand(n, 7, generator=g, dtype=torch.float64).to(dev, dtype)) * (ul - ll) + ll
    else:
        q = torch.as_tensor(seeds).to(dev, dtype).clone()
    tg = torch.as_tensor(target_pose).to(dev, dtype).reshape(1, 7).expand(q.shape[0], 7).contiguous()
    for _ in range(iters):
        tw = pirl_env.pose_diff(q, tg)
        dq = pirl_env.pinv_step(q, tw)
        m = dq.abs().amax(dim=1, keepdim=True).clamp_min(1e-12)
        dq = dq * torch.clamp(step_limit / m, max=1.0)
        q = torch.minimum(torch.maximum(q + dq, ll), ul)
    tw = pirl_env.pose_diff(q, tg)
    ok = (tw[:, :3].norm(dim=1) < tol_pos) & (tw[:, 3:].norm(dim=1) < tol_rot)
    return q, ok
