"""torch.library registration of the stateless kernels: `torch.ops.rlitg.*` (SURVEY §8b "C-ABI + torch op").

    torch.ops.rlitg.fk(q)              -> (pos [B,3], quat [B,4])     fk_srv, pirl/src/kinematic_solver.cpp:237-261
    torch.ops.rlitg.fk_jac(q)          -> (pos, rot [B,3,3], J [B,6,D]) TormIKSolver::fkSolver + getJacobian,
                                                                      torm/src/torm_ik_solver.cpp:156-158,173-175
    torch.ops.rlitg.allfk(q)           -> [B,D+1,7]                   allfk_srv, pirl/src/kinematic_solver.cpp:206-234
    torch.ops.rlitg.jacobian(q)        -> [B,6,D]                     getJacobian, torm/src/torm_ik_solver.cpp:173
    torch.ops.rlitg.nullspace_norm(err, q) -> [B]                     jaco_srv, pirl/src/kinematic_solver.cpp:332-351
    torch.ops.rlitg.pinv_step(q, twist, eps) -> [B,D]                 vikSolver, torm/src/torm_ik_solver.cpp:169
    torch.ops.rlitg.pose_diff(q, target) -> [B,6]                     KDL::diff, torm/src/torm_optimizer.cpp:1413
    torch.ops.rlitg.manipulability(q)  -> [B]                         checkSingularity, torm/src/torm_optimizer.cpp:748-765

Each op only has a CUDA implementation (it calls the C ABI of librlitg.so on the current stream) and a shape
function for meta / fake tensors; a CPU tensor raises NotImplementedError -- there is no CPU path.  The direct
functions in `kinematics` / `pirl_env` skip the dispatcher (a few µs per call) and are what the hot loops use; the
stateful entry points (scene handle, env state, smoothness band) stay methods of Scene / PirlVecEnv / TormBatch.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import kinematics as K
from . import pirl_env as P

_cuda = dict(mutates_args=(), device_types="cuda")


def _bd(q):
    if q.dim() != 2 or q.shape[1] not in (7, 8):
        raise ValueError("q must be [B,7] or [B,8]")
    return q.shape[0], q.shape[1]


@torch.library.custom_op("rlitg::fk", **_cuda)
def fk(q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return K.fk(q)


@fk.register_fake
def _(q):
    B, _D = _bd(q)
    return q.new_empty(B, 3), q.new_empty(B, 4)


@torch.library.custom_op("rlitg::fk_jac", **_cuda)
def fk_jac(q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    r = K.fk_jac(q)
    return r["pos"], r["rot"], r["jac"]


@fk_jac.register_fake
def _(q):
    B, D = _bd(q)
    return q.new_empty(B, 3), q.new_empty(B, 3, 3), q.new_empty(B, 6, D)


@torch.library.custom_op("rlitg::allfk", **_cuda)
def allfk(q: torch.Tensor) -> torch.Tensor:
    return K.allfk(q)


@allfk.register_fake
def _(q):
    B, D = _bd(q)
    return q.new_empty(B, D + 1, 7)


@torch.library.custom_op("rlitg::jacobian", **_cuda)
def jacobian(q: torch.Tensor) -> torch.Tensor:
    return K.jacobian(q)


@jacobian.register_fake
def _(q):
    B, D = _bd(q)
    return q.new_empty(B, 6, D)


@torch.library.custom_op("rlitg::nullspace_norm", **_cuda)
def nullspace_norm(err: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    return P.nullspace_norm(err, q)


@nullspace_norm.register_fake
def _(err, q):
    return q.new_empty(q.shape[0])


@torch.library.custom_op("rlitg::pinv_step", **_cuda)
def pinv_step(q: torch.Tensor, twist: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return P.pinv_step(q, twist, eps)


@pinv_step.register_fake
def _(q, twist, eps=1e-5):
    B, D = _bd(q)
    return q.new_empty(B, D)


@torch.library.custom_op("rlitg::pose_diff", **_cuda)
def pose_diff(q: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return P.pose_diff(q, target)


@pose_diff.register_fake
def _(q, target):
    return q.new_empty(q.shape[0], 6)


@torch.library.custom_op("rlitg::manipulability", **_cuda)
def manipulability(q: torch.Tensor) -> torch.Tensor:
    return P.manipulability(q)


@manipulability.register_fake
def _(q):
    return q.new_empty(q.shape[0])


OPS = ("fk", "fk_jac", "allfk", "jacobian", "nullspace_norm", "pinv_step", "pose_diff", "manipulability")
