"""torch.ops.rlitg.* (rl-itg_b200/ops.py): registration and shape functions on CPU, values on the GPU."""
import numpy as np
import pytest
import torch

import rlitg_b200.ops as ops


def test_ops_registered_with_shape_functions():
    q = torch.empty(5, 7, device="meta")
    q8 = torch.empty(3, 8, device="meta", dtype=torch.float64)
    for name in ops.OPS:
        assert hasattr(torch.ops.rlitg, name)
    pos, quat = torch.ops.rlitg.fk(q)
    assert pos.shape == (5, 3) and quat.shape == (5, 4)
    pos, rot, jac = torch.ops.rlitg.fk_jac(q8)
    assert rot.shape == (3, 3, 3) and jac.shape == (3, 6, 8) and jac.dtype == torch.float64
    assert torch.ops.rlitg.allfk(q8).shape == (3, 9, 7)
    assert torch.ops.rlitg.jacobian(q).shape == (5, 6, 7)
    assert torch.ops.rlitg.nullspace_norm(q, q).shape == (5,)
    assert torch.ops.rlitg.pinv_step(q, torch.empty(5, 6, device="meta")).shape == (5, 7)
    assert torch.ops.rlitg.pose_diff(q, torch.empty(5, 7, device="meta")).shape == (5, 6)
    assert torch.ops.rlitg.manipulability(q).shape == (5,)


def test_ops_have_no_cpu_path():
    with pytest.raises(NotImplementedError):
        torch.ops.rlitg.fk(torch.zeros(2, 7))
    with pytest.raises(NotImplementedError):
        torch.ops.rlitg.pinv_step(torch.zeros(2, 7), torch.zeros(2, 6))


@pytest.mark.gpu
def test_ops_match_oracle(cuda_device):
    from oracle import oracle as O
    from rlitg_b200 import synth
    q = synth.random_configs(257, seed=3)
    tq = torch.tensor(q, dtype=torch.float64, device=cuda_device)
    pos, rot, jac = torch.ops.rlitg.fk_jac(tq)
    rp, rr, rj = O.fk_jac(q)
    np.testing.assert_allclose(pos.cpu().numpy(), rp, atol=1e-12)
    np.testing.assert_allclose(rot.cpu().numpy().reshape(-1, 9), rr.reshape(-1, 9), atol=1e-12)
    np.testing.assert_allclose(jac.cpu().numpy(), rj, atol=1e-12)
    p2, quat = torch.ops.rlitg.fk(tq)
    assert torch.equal(p2, pos)
    np.testing.assert_allclose(torch.ops.rlitg.allfk(tq).cpu().numpy(), O.allfk(q), atol=1e-12)
    assert torch.equal(torch.ops.rlitg.jacobian(tq), jac)
    rng = np.random.default_rng(0)
    err = rng.normal(size=(257, 7))
    np.testing.assert_allclose(torch.ops.rlitg.nullspace_norm(torch.tensor(err, device=cuda_device), tq).cpu().numpy(),
                               O.nullspace_norm(err, q), rtol=1e-7, atol=1e-9)
