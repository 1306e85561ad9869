"""GPU parity: kernel 1 (FK + Jacobian, all-link FK) through the C ABI vs the CPU oracle.
Tolerances (BASELINE.json north_star): FP32 <= 1e-5 relative with a 1e-6 absolute floor; FP64 build <= 1e-12."""
import numpy as np
import pytest

from conftest import random_configs

pytestmark = pytest.mark.gpu


def close(a, b, rel, floor):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    return float((err / (floor + rel * np.abs(b))).max())


def quat_close(a, b, tol):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    sign = np.sign(np.sum(a * b, axis=-1, keepdims=True)); sign[sign == 0] = 1
    return np.abs(a * sign - b).max() < tol


@pytest.mark.parametrize("D", [7, 8])
@pytest.mark.parametrize("dtype,rel,floor", [("float32", 1e-5, 1e-6), ("float64", 1e-12, 1e-13)])
@pytest.mark.parametrize("q_soa,out_soa", [(False, False), (True, True), (False, True), (True, False)])
@pytest.mark.parametrize("B", [1, 127, 128, 1000, 4099])
def test_fk_jac_parity(cuda_device, D, dtype, rel, floor, q_soa, out_soa, B):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import kinematics as K
    q = random_configs(B, D, seed=B + D)
    pos, R, J = O.fk_jac(q)
    quat = O.fk_quat(q)[:, 3:]
    tq = torch.tensor(q, dtype=getattr(torch, dtype), device=cuda_device)
    if q_soa:
        tq = tq.t().contiguous()
    r = K.fk_jac(tq, want_quat=True, q_soa=q_soa, out_soa=out_soa)
    torch.cuda.synchronize()
    gp, gr, gq, gj = [r[k].double().cpu().numpy() for k in ("pos", "rot", "quat", "jac")]
    if out_soa:
        gp, gr, gq, gj = gp.T, gr.T.reshape(B, 3, 3), gq.T, gj.T.reshape(B, 6, D)
    assert close(gp, pos, rel, floor) <= 1.0
    assert close(gr, R, rel, floor) <= 1.0
    assert close(gj, J, rel, floor) <= 1.0
    assert quat_close(gq, quat, 1e-5 if dtype == "float32" else 1e-12)


@pytest.mark.parametrize("D", [7, 8])
@pytest.mark.parametrize("dtype,tol", [("float32", 2e-6), ("float64", 1e-12)])
@pytest.mark.parametrize("B", [1, 129, 2050, 131072 + 77])  # the last one takes the persistent kernel + a tail
def test_allfk_parity(cuda_device, D, dtype, tol, B):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import kinematics as K
    q = random_configs(B, D, seed=17)
    want = O.allfk(q)
    for soa in (False, True):
        tq = torch.tensor(q, dtype=getattr(torch, dtype), device=cuda_device)
        got = K.allfk(tq.t().contiguous() if soa else tq, q_soa=soa).double().cpu().numpy()
        assert np.abs(got[..., :3] - want[..., :3]).max() < tol
        assert quat_close(got[..., 3:], want[..., 3:], 5 * tol)


def test_fk_fixture_pairs_on_gpu(cuda_device, fk_pairs):
    """The 62,610 + 526 + 5,000 reference (q -> pose) pairs, through the FP32 kernel."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import kinematics as K
    from test_oracle_golden import rot_err_angle
    for key, ptol, rtol in [("named", 2e-5, 1e-3), ("rand", 2e-5, 5e-5), ("obs", 2e-5, 2.5e-3)]:
        q, pose = fk_pairs[key + "_q"], fk_pairs[key + "_pose"]
        r = K.fk_jac(torch.tensor(q, dtype=torch.float32, device=cuda_device), want_jac=False)
        pos = r["pos"].double().cpu().numpy(); R = r["rot"].double().cpu().numpy()
        assert np.linalg.norm(pos - pose[:, :3], axis=1).max() < ptol + 2e-6
        assert rot_err_angle(R, O.quat_to_rot(pose[:, 3:])).max() < rtol + 2e-6
    # hello.yaml known answer in the FP64 build
    q = torch.tensor(fk_pairs["hello_init_q"][None], dtype=torch.float64, device=cuda_device)
    pos, quat = K.fk(q)
    assert np.abs(pos.cpu().numpy()[0] - [0.8, 0.45, 0.25]).max() < 1e-9
    assert np.abs(quat.cpu().numpy()[0] - [0, 0, 0, 1]).max() < 1e-9


def test_fk_jac_full_size_properties(cuda_device):
    """BASELINE config 2 size (1M configs): size-independent properties instead of an oracle sweep:
    R orthonormal, J angular columns unit, J consistent with central-difference FK on a sample,
    and identical results for AoS/SoA layouts and for the host-buffer entry point."""
    import torch
    from rlitg_b200 import kinematics as K
    B = 1 << 20
    g = torch.Generator(device="cpu"); g.manual_seed(0)
    from conftest import LL7, UL7
    q = (torch.rand(B, 7, generator=g, dtype=torch.float64) * torch.tensor(UL7 - LL7) + torch.tensor(LL7)).float()
    qd = q.to(cuda_device)
    r = K.fk_jac(qd)
    R = r["rot"]; J = r["jac"]
    eye = torch.eye(3, device=cuda_device)
    assert (R @ R.transpose(1, 2) - eye).abs().max().item() < 5e-6
    assert (J[:, 3:, :].norm(dim=1) - 1).abs().max().item() < 5e-6
    r2 = K.fk_jac(qd.t().contiguous(), q_soa=True, out_soa=True)
    assert torch.equal(r2["pos"].t().contiguous(), r["pos"])
    assert torch.equal(r2["jac"].t().reshape(B, 6, 7), J)
    # host path
    n = 300_000
    hp = torch.empty(n, 3).pin_memory(); hj = torch.empty(n, 6, 7).pin_memory(); hr = torch.empty(n, 3, 3).pin_memory()
    K.fk_jac_host(q[:n].contiguous().pin_memory(), pos=hp, rot=hr, jac=hj)
    assert torch.equal(hp, r["pos"][:n].cpu()) and torch.equal(hj, J[:n].cpu()) and torch.equal(hr, R[:n].cpu())
    # finite differences on a sample (FP64 build)
    idx = torch.arange(0, B, 4099)
    qs = q[idx].double().to(cuda_device)
    Js = K.jacobian(qs)
    h = 1e-6
    for j in range(7):
        d = torch.zeros(7, dtype=torch.float64, device=cuda_device); d[j] = h
        pp = K.fk_jac(qs + d, want_rot=False, want_jac=False)["pos"]
        pm = K.fk_jac(qs - d, want_rot=False, want_jac=False)["pos"]
        assert ((pp - pm) / (2 * h) - Js[:, :3, j]).abs().max().item() < 1e-7
    assert (Js.float() - J[idx.to(cuda_device)]).abs().max().item() < 2e-6


@pytest.mark.parametrize("D", [7, 8])
@pytest.mark.parametrize("dtype,rel,floor", [("float32", 1e-5, 1e-6), ("float64", 1e-12, 1e-13)])
def test_fk_jac_benchmark_size_against_oracle(cuda_device, D, dtype, rel, floor):
    """BASELINE config 2 as benchmarked: 1,048,576 configurations, array-of-structs in and out -> the PERSISTENT kernel
    (fk_jac_persist_kernel: >= 1,024 tiles) -- every pose, rotation and Jacobian entry against the CPU oracle, FP32 and
    the FP64 verification build, D = 7 and the D = 8 torso chain."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import kinematics as K, _lib
    import os
    O.set_num_threads(len(os.sched_getaffinity(0)))
    B = 1 << 20
    q = random_configs(B, D, seed=40 + D)
    if dtype == "float32":
        q = q.astype(np.float32).astype(np.float64)   # both sides see the same joint values
    pos, R, J = O.fk_jac(q)
    n0 = _lib.launch_count()
    r = K.fk_jac(torch.tensor(q, dtype=getattr(torch, dtype), device=cuda_device))
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 <= 2   # the persistent kernel (+ at most a tail launch), not a tile-per-CTA sweep
    gp, gr, gj = [r[k].double().cpu().numpy() for k in ("pos", "rot", "jac")]
    assert close(gp, pos, rel, floor) <= 1.0
    assert close(gr, R, rel, floor) <= 1.0
    assert close(gj, J, rel, floor) <= 1.0


def test_service_shims(cuda_device, fk_pairs):
    from oracle import oracle as O
    from rlitg_b200.kinematics import KinematicsServices
    srv = KinematicsServices(cuda_device)
    q = fk_pairs["hello_init_q"].tolist()
    res = srv.fk_srv(q)
    assert res.result is True and len(res.fk_result) == 7
    assert np.abs(np.array(res.fk_result) - O.fk_quat(np.array([q]))[0]).max() < 1e-12
    allr = srv.allfk_srv(q)
    assert len(allr.allfk_result) == 56
    assert np.abs(np.array(allr.allfk_result) - O.allfk(np.array([q]))[0].reshape(-1)).max() < 1e-12


def test_empty_batch(cuda_device):
    import torch
    from rlitg_b200 import kinematics as K
    r = K.fk_jac(torch.empty(0, 7, device=cuda_device))
    assert r["pos"].shape == (0, 3) and r["jac"].shape == (0, 6, 7)


def test_second_device_in_one_process():
    """Kernel attributes and host workspaces are per device: the same process can drive two GPUs (skipped on 1 GPU)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle as O
    from rlitg_b200 import kinematics as K
    q = random_configs(300000, 7, seed=2)
    pos, R, J = O.fk_jac(q[:1000])
    for dev in ("cuda:1", "cuda:0"):
        r = K.fk_jac(torch.tensor(q, dtype=torch.float32, device=dev))
        assert np.abs(r["jac"][:1000].cpu().numpy() - J).max() < 4e-6
        a = K.allfk(torch.tensor(q[:5000], dtype=torch.float64, device=dev))
        assert np.abs(a.cpu().numpy()[..., :3] - O.allfk(q[:5000])[..., :3]).max() < 1e-12
