"""Readers for the reference's on-disk formats, so real problems -- not only synthetic ones -- can
drive the kernels.  Host side, numpy.

  path file        torm/launch/config/path/<name>: one pose per line `x,y,z;w,x,y,z`
                   (TormProblem::setTargetPoses, torm/src/torm_problem.cpp:154-179)
  config csv       start_conf/<name>_config.csv, initial_traj/*.csv, <k>_config.csv: comma separated joint rows
                   (torm/src/torm_generate_start_conf.cpp:94-106, torm/src/pirl_datagen.cpp:29-66)
  targetpose csv   <k>_targetpose.csv: `x,y,z,qx,qy,qz,qw` rows (torm/src/pirl_datagen.cpp:29-66)
  box scene        scene/scene_<n>.txt: `x y z roll pitch yaw sx sy sz` per box in base_link
                   (TormProblem::setCollisionObjects / makeCollisionObject, torm/src/torm_problem.cpp:77-107,121-150)
  occupancy        occ_<n>.npy float [50,70,50] indexed [ix][iy][iz] (pirl_gazebo/script/save_pointcloud.py:22-32)
All poses are returned as [x, y, z, qx, qy, qz, qw], the layout of include/rlitg.h."""
from __future__ import annotations

import numpy as np


def load_path(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            p, q = line.split(";")
            x, y, z = [float(v) for v in p.split(",")]
            w, qx, qy, qz = [float(v) for v in q.split(",")]
            rows.append((x, y, z, qx, qy, qz, w))
    return np.asarray(rows, dtype=np.float64).reshape(-1, 7)


def load_csv_rows(path, ncol=None):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                rows.append([float(v) for v in line.replace(";", ",").split(",")])
    a = np.asarray(rows, dtype=np.float64)
    if ncol is not None and a.size and a.shape[1] != ncol:
        raise ValueError("%s: expected %d columns, found %d" % (path, ncol, a.shape[1]))
    return a


def load_configs(path):
    return load_csv_rows(path, 7)


def load_target_poses(path):
    return load_csv_rows(path, 7)


def load_box_scene(path):
    """-> float64 [n_boxes, 9] = x y z roll pitch yaw sx sy sz"""
    rows = []
    with open(path) as f:
        for line in f:
            v = line.split()
            if v:
                rows.append([float(t) for t in v])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 9)


def rpy_matrix(roll, pitch, yaw):
    """tf::Quaternion::setRPY convention: R = Rz(yaw) Ry(pitch) Rx(roll)"""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def voxelize_boxes(boxes, shape_zyx, origin, res):
    """Occupancy uint8 [nz,ny,nx]: a cell is occupied when its centre lies inside an (oriented) box."""
    nz, ny, nx = shape_zyx
    xs = origin[0] + res * np.arange(nx)
    ys = origin[1] + res * np.arange(ny)
    zs = origin[2] + res * np.arange(nz)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    P = np.stack([X, Y, Z], axis=-1)
    occ = np.zeros(shape_zyx, dtype=np.uint8)
    for b in np.asarray(boxes, dtype=np.float64).reshape(-1, 9):
        R = rpy_matrix(b[3], b[4], b[5])
        local = (P - b[:3]) @ R  # R^T (p - c)
        occ |= np.all(np.abs(local) <= 0.5 * b[6:9], axis=-1).astype(np.uint8)
    return occ


def load_occ_npy(path):
    """pirl_gazebo occ_<n>.npy [ix][iy][iz] -> (occ uint8 [nz,ny,nx], origin of cell (0,0,0), res)"""
    g = np.load(path)
    return (np.transpose(g, (2, 1, 0)) != 0).astype(np.uint8), (0.21, -0.69, 0.21), 0.02
