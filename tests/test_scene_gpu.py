"""GPU parity: scene staging (bit-packing, exact signed EDT -> distance records) and sphere-grid collision."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_scene(cuda_device, shape, origin, res, seed, max_prop=1.0):
    from oracle import oracle as O
    from rlitg_b200 import fetch, scene as S, synth
    occ, _ = synth.random_box_scene(shape, origin, res, seed=seed, n_clutter=3)
    sc = S.Scene(cuda_device)
    sc.set_occupancy(occ, origin, res, max_prop)
    link, center, radius = fetch.default_spheres()
    osc = O.Scene(O.build_field(occ, res, max_prop), origin, res, link, center, radius)
    return occ, sc, osc


@pytest.mark.parametrize("shape,res,max_prop", [((24, 40, 33), 0.05, 1.0), ((64, 64, 64), 0.04, 0.5),
                                                ((50, 70, 50), 0.02, 1.0)])
def test_field_bit_exact(cuda_device, shape, res, max_prop):
    origin = (-0.2, -1.0, 0.0)
    occ, sc, osc = make_scene(cuda_device, shape, origin, res, seed=3, max_prop=max_prop)
    got = sc.get_field()
    assert got.shape == osc.field.shape
    assert np.array_equal(got.view(np.uint32), osc.field.view(np.uint32))  # integer EDT + IEEE sqrt: bit-exact
    # bit-packed occupancy
    bits = sc.get_bits()
    nz, ny, nx = shape
    unpacked = ((bits[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(nz, ny, -1)[..., :nx]
    assert np.array_equal(unpacked.astype(np.uint8), occ)
    sc.close()


def test_field_full_size_properties(cuda_device):
    """128^3 (BASELINE config 4 grid): sign convention, gradient bounds, bit-exact vs the oracle."""
    occ, sc, osc = make_scene(cuda_device, (128, 128, 128), (-0.64, -1.28, -0.20), 0.02, seed=0)
    f = sc.get_field()
    assert np.array_equal(f.view(np.uint32), osc.field.view(np.uint32))
    active = f[..., 0] < 1e29
    inside = occ.astype(bool)
    assert (f[..., 0][active & inside] < 0).all() and (f[..., 0][active & ~inside] > 0).all()
    # central differences of the SIGNED field: across a solid face a component is exactly 1 (a box corner reaches
    # sqrt(3)); across a one-cell-thin wall it mixes the outside and the inside transforms and can reach 1.5
    comp = np.abs(f[..., 1:][active])
    assert comp.max() <= 1.5 + 1e-5
    assert (np.linalg.norm(f[..., 1:][active], axis=1) <= np.sqrt(3) + 1e-5).mean() > 0.999
    sc.close()


def test_empty_and_full_grids(cuda_device):
    from oracle import oracle as O
    from rlitg_b200 import scene as S
    for fill in (0, 1):
        occ = np.full((16, 16, 16), fill, dtype=np.uint8)
        sc = S.Scene(cuda_device)
        sc.set_occupancy(occ, (0, 0, 0), 0.1, 1.0)
        assert np.array_equal(sc.get_field().view(np.uint32), O.build_field(occ, 0.1, 1.0).view(np.uint32))
        sc.close()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_sphere_collide_parity(cuda_device, dtype):
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, synth
    occ, sc, osc = make_scene(cuda_device, (128, 128, 128), (-0.64, -1.28, -0.20), 0.02, seed=1)
    q = synth.random_configs(20000, seed=4)
    want = O.sphere_collide(q, osc, fetch.BASE_OFFSET)
    got = sc.collide(torch.tensor(q, dtype=getattr(torch, dtype), device=cuda_device)).cpu().numpy()
    mism = int((got != want).sum())
    assert 0.02 < want.mean() < 0.98  # the scene separates configurations
    # FP32 may pick a neighbouring cell for a sphere centre on a voxel boundary: count, do not hide
    assert mism <= (0 if dtype == "float64" else 20), mism
    sc.close()


def test_scene_errors(cuda_device):
    from rlitg_b200 import scene as S
    from rlitg_b200._lib import RlitgError
    sc = S.Scene(cuda_device, spheres=None)
    with pytest.raises(RlitgError):
        sc.set_spheres([5, 3], np.zeros((2, 3)), np.ones(2))  # not sorted by link
    with pytest.raises(RlitgError):
        sc.set_spheres([9], np.zeros((1, 3)), np.ones(1))  # link out of range
    import torch
    with pytest.raises(RlitgError):
        sc.collide(torch.zeros(4, 7, device=cuda_device))  # no grid / spheres yet
    sc.close()


def test_dilated_bit_volumes_and_smem_collide(cuda_device):
    """North-star item: the occupancy is staged as bit-packed volumes (one per sphere radius) and the binary test
    reads one bit per sphere -- from shared memory when the volumes fit (pirl_gazebo grid, 50 x 70 x 50 @ 0.02 m)."""
    import torch
    from oracle import oracle as O
    from rlitg_b200 import fetch, scene as S, synth
    g = S.PIRL_GAZEBO_GRID
    rng = np.random.default_rng(3)
    grid_xyz = np.zeros((50, 70, 50))
    grid_xyz[20:40, 10:60, 0:18] = 1.0          # a table slab, [ix][iy][iz] like occ_N.npy
    grid_xyz[25:30, 30:36, 18:30] = 1.0         # a box on it
    sc = S.Scene(cuda_device)
    sc.set_pirl_gazebo_occ(grid_xyz)
    field = sc.get_field()
    radii, bits = sc.get_dilated()
    link, center, radius = fetch.default_spheres()
    assert radii is not None and sorted(radii.tolist()) == sorted(set(np.float32(radius).tolist()))
    nz, ny, nx = sc.shape
    for k, r in enumerate(radii):
        unpacked = ((bits[k][..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(nz, ny, -1)[..., :nx].astype(bool)
        want = (field[..., 0] < 1e29) & (field[..., 0] < r)
        assert np.array_equal(unpacked, want), k
    occ = np.transpose(grid_xyz, (2, 1, 0)) != 0
    osc = O.Scene(O.build_field(occ, g["res"], 1.0), g["origin"], g["res"], link, center, radius)
    q = synth.random_configs(30000, seed=6)
    want = O.sphere_collide(q, osc, fetch.BASE_OFFSET)
    assert 0.02 < want.mean() < 0.98
    for dtype, tol in ((torch.float64, 0), (torch.float32, 20)):
        tq = torch.tensor(q, dtype=dtype, device=cuda_device)
        big = sc.collide(tq).cpu().numpy()          # >= 4096 configurations: shared-memory resident volumes
        small = torch.cat([sc.collide(tq[i:i + 1000]) for i in range(0, 30000, 1000)]).cpu().numpy()  # global path
        assert np.array_equal(big, small)
        assert int((big != want).sum()) <= tol
    sc.close()
