"""CPU tests: readers of the reference's file formats against raw fixture files copied from the reference's data
directory (tests/golden/raw) and against the golden arrays harvested by make_golden.py."""
import os

import numpy as np

from conftest import GOLDEN

RAW = os.path.join(GOLDEN, "raw")


def test_path_and_config_readers(init_traj, fk_pairs):
    from oracle import oracle as O
    from rlitg_b200 import io
    path = io.load_path(os.path.join(RAW, "path_square_head.txt"))
    assert path.shape == (12, 7)
    assert np.array_equal(path, init_traj["square_path"][:12])  # x,y,z;w,x,y,z -> x y z qx qy qz qw
    q = io.load_configs(os.path.join(RAW, "square_config.csv"))
    assert q.shape == (10, 7)
    # every start configuration is an IK solution of the first path pose
    pos, _ = O.fk(q)
    assert np.linalg.norm(pos - path[0, :3], axis=1).max() < 2e-5


def test_box_scene_voxelisation():
    from rlitg_b200 import io
    boxes = io.load_box_scene(os.path.join(RAW, "scene_0.txt"))
    assert boxes.shape == (1, 9) and np.allclose(boxes[0, :3], [0.8, 0.0, 0.316316218028])
    occ = io.voxelize_boxes(boxes, (64, 128, 64), (0.0, -1.27, 0.0), 0.02)
    vol = occ.sum() * 0.02 ** 3
    assert abs(vol - np.prod(boxes[0, 6:9])) / np.prod(boxes[0, 6:9]) < 0.08
    # a rotated box keeps its volume and changes its footprint
    b1 = boxes.copy(); b1[0, 6:9] = [0.3, 0.8, 0.4]  # small enough to stay inside the grid when rotated
    b2 = b1.copy(); b2[0, 5] = np.pi / 4
    occ1 = io.voxelize_boxes(b1, (64, 128, 64), (0.0, -1.27, 0.0), 0.02)
    occ2 = io.voxelize_boxes(b2, (64, 128, 64), (0.0, -1.27, 0.0), 0.02)
    assert abs(int(occ2.sum()) - int(occ1.sum())) < 0.05 * occ1.sum() and not np.array_equal(occ1, occ2)
    R = io.rpy_matrix(0.3, -0.2, 1.1)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12


def test_occ_npy_reader(tmp_path):
    from rlitg_b200 import io
    g = np.zeros((50, 70, 50)); g[3, 10, 7] = 1.0  # [ix][iy][iz] as save_pointcloud.py writes it
    p = tmp_path / "occ_0.npy"; np.save(p, g)
    occ, origin, res = io.load_occ_npy(str(p))
    assert occ.shape == (50, 70, 50) and occ[7, 10, 3] == 1 and occ.sum() == 1 and res == 0.02
