"""Batched Fetch kinematics on the GPU + the reference's service-call signatures.

Tensor API (torch CUDA tensors, float32 or float64, leading batch axis):
    fk_jac(q) -> pos [B,3], rot [B,3,3], J [B,6,D]      (TormIKSolver::fkSolver + getJacobian,
                                                         torm/src/torm_ik_solver.cpp:156-158,173-175)
    fk(q) -> pos [B,3], quat [B,4] xyzw                  (fk_srv, pirl/src/kinematic_solver.cpp:237-261)
    allfk(q) -> [B, D+1, 7]                              (allfk_srv, pirl/src/kinematic_solver.cpp:206-234)
    jacobian(q) -> [B,6,D]                               (getJacobian)

Drop-in shims with the reference's call shape (python lists in, response object out):
    KinematicsServices().fk_srv(q).fk_result, .allfk_srv(q).allfk_result
                                                         (pirl/script/envs/PirlEnv.py:47-49,216,378,556)
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    raise TypeError("q must be float32 or float64, got %s" % t.dtype)


def _prep(q: torch.Tensor, soa: bool):
    if not q.is_cuda:
        raise ValueError("q must be a CUDA tensor (there is no CPU path); use the *_host functions for host data")
    if q.dim() != 2:
        raise ValueError("q must be 2-D: [B,D] (or [D,B] with q_soa=True)")
    q = q.contiguous()
    B, D = (q.shape[1], q.shape[0]) if soa else (q.shape[0], q.shape[1])
    if D not in (7, 8):
        raise ValueError("D must be 7 (arm) or 8 (torso + arm), got %d" % D)
    return q, B, D


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def fk_jac(q, want_pos=True, want_rot=True, want_quat=False, want_jac=True, q_soa=False, out_soa=False, out=None):
    """Fused FK + Jacobian.  Returns dict(pos, rot, quat, jac) with None for outputs not requested.
    ``out`` may hold preallocated tensors (keys as above) to avoid allocation in hot loops."""
    q, B, D = _prep(q, q_soa)
    out = out or {}
    kw = dict(device=q.device, dtype=q.dtype)

    def alloc(name, want, aos_shape, soa_shape):
        if not want:
            return None
        t = out.get(name)
        if t is None:
            t = torch.empty(soa_shape if out_soa else aos_shape, **kw)
        return t

    pos = alloc("pos", want_pos, (B, 3), (3, B))
    rot = alloc("rot", want_rot, (B, 3, 3), (9, B))
    quat = alloc("quat", want_quat, (B, 4), (4, B))
    jac = alloc("jac", want_jac, (B, 6, D), (6 * D, B))
    flags = (_lib.Q_SOA if q_soa else 0) | (_lib.OUT_SOA if out_soa else 0)
    with _lib.on_device(q.device):
        check(lib.rlitg_fk_jac(q.data_ptr(), B, D, _dtype_code(q), flags, _ptr(pos), _ptr(rot), _ptr(quat),
                               _ptr(jac), _stream()))
    return dict(pos=pos, rot=rot, quat=quat, jac=jac)


def fk(q, q_soa=False):
    r = fk_jac(q, want_pos=True, want_rot=False, want_quat=True, want_jac=False, q_soa=q_soa)
    return r["pos"], r["quat"]


def jacobian(q, q_soa=False):
    return fk_jac(q, want_pos=False, want_rot=False, want_quat=False, want_jac=True, q_soa=q_soa)["jac"]


def allfk(q, q_soa=False, out=None):
    q, B, D = _prep(q, q_soa)
    if out is None:
        out = torch.empty((B, D + 1, 7), device=q.device, dtype=q.dtype)
    with _lib.on_device(q.device):
        check(lib.rlitg_allfk(q.data_ptr(), B, D, _dtype_code(q), _lib.Q_SOA if q_soa else 0, out.data_ptr(),
                              _stream()))
    return out


def fk_jac_host(q, pos=None, rot=None, quat=None, jac=None):
    """End-to-end host path (rlitg_fk_jac_host): q and the outputs are HOST tensors (pinned for full PCIe
    bandwidth).  Outputs passed as None are not computed."""
    if q.is_cuda:
        raise ValueError("fk_jac_host takes host tensors")
    q = q.contiguous()
    B, D = q.shape
    check(lib.rlitg_fk_jac_host(q.data_ptr(), B, D, _dtype_code(q), 0, _ptr(pos), _ptr(rot), _ptr(quat), _ptr(jac)))
    return dict(pos=pos, rot=rot, quat=quat, jac=jac)


# ---------------------------------------------------------------------------------------------- service shims
class _Resp:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class KinematicsServices:
    """Stand-ins for the rospy.ServiceProxy objects of PirlEnv (pirl/script/envs/PirlEnv.py:42-49): same call
    shape (a python list of 7 floats in, a response object out), float64 like the .srv payloads, batch of one."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)

    def _q(self, query_config):
        return torch.tensor([list(query_config)], dtype=torch.float64, device=self.device)

    def fk_srv(self, query_config):
        pos, quat = fk(self._q(query_config))
        return _Resp(fk_result=tuple(torch.cat([pos[0], quat[0]]).tolist()), result=True)

    def allfk_srv(self, query_config):
        return _Resp(allfk_result=tuple(allfk(self._q(query_config))[0].reshape(-1).tolist()), result=True)
