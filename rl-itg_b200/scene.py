"""Collision scene handle: sphere table + occupancy grid / distance volume, resident on one GPU.

Replaces, as data, what the reference obtains from MoveIt (CollisionWorldHybrid / CollisionRobotHybrid,
torm/src/torm_optimizer.cpp:110-135) and, as a format, the pirl_gazebo occupancy grid
(pirl_gazebo/script/save_pointcloud.py:22-32: 50 x 70 x 50 cells @ 0.02 m from (0.2, -0.7, 0.2))."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, fetch
from ._lib import check, lib

lib.rlitg_scene_create.argtypes = [C.POINTER(C.c_void_p)]
lib.rlitg_scene_destroy.argtypes = [C.c_void_p]
lib.rlitg_scene_set_spheres.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
lib.rlitg_scene_set_occupancy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double,
                                          C.c_double]
lib.rlitg_scene_set_field.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double]
lib.rlitg_scene_get_field.argtypes = [C.c_void_p, C.c_void_p]
lib.rlitg_scene_get_bits.argtypes = [C.c_void_p, C.c_void_p]
lib.rlitg_scene_get_dilated.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.rlitg_sphere_grid_collide.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

# the pirl_gazebo grid geometry (cell centres): pirl_gazebo/script/save_pointcloud.py:22-27
PIRL_GAZEBO_GRID = dict(shape_zyx=(50, 70, 50), origin=(0.21, -0.69, 0.21), res=0.02)


class Scene:
    def __init__(self, device="cuda:0", spheres="default", base_offset=fetch.BASE_OFFSET):
        self.device = torch.device(device)
        self.base_offset = np.ascontiguousarray(base_offset, dtype=np.float64)
        self._h = C.c_void_p()
        with _lib.on_device(self.device):
            check(lib.rlitg_scene_create(C.byref(self._h)))
        self.shape = None
        self.origin = None
        self.res = None
        self.spheres = None
        if spheres == "default":
            self.set_spheres(*fetch.default_spheres())
        elif spheres is not None:
            self.set_spheres(*spheres)

    @property
    def handle(self):
        return self._h

    def set_spheres(self, link, center, radius):
        link = np.ascontiguousarray(link, dtype=np.int32)
        center = np.ascontiguousarray(center, dtype=np.float64)
        radius = np.ascontiguousarray(radius, dtype=np.float64)
        with _lib.on_device(self.device):
            check(lib.rlitg_scene_set_spheres(self._h, int(link.shape[0]), link.ctypes.data, center.ctypes.data,
                                              radius.ctypes.data))
        self.spheres = (link, center, radius)

    def set_occupancy(self, occ, origin, res, max_prop=1.0):
        """occ: uint8/bool [nz, ny, nx]; origin: centre of cell (0,0,0) in the planning frame."""
        occ = np.ascontiguousarray(np.asarray(occ) != 0, dtype=np.uint8)
        nz, ny, nx = occ.shape
        org = np.ascontiguousarray(origin, dtype=np.float64)
        with _lib.on_device(self.device):
            check(lib.rlitg_scene_set_occupancy(self._h, occ.ctypes.data, nx, ny, nz, org.ctypes.data, float(res),
                                                float(max_prop)))
        self.shape, self.origin, self.res = (nz, ny, nx), org, float(res)

    def set_pirl_gazebo_occ(self, grid_xyz, max_prop=1.0):
        """occ_N.npy as pirl_gazebo writes it: float [50,70,50] indexed [ix][iy][iz], cell k covering
        ws_low + [k, k+1) * 0.02 (pirl_gazebo/script/save_pointcloud.py:22-32)."""
        g = np.asarray(grid_xyz)
        self.set_occupancy(np.transpose(g, (2, 1, 0)), PIRL_GAZEBO_GRID["origin"], PIRL_GAZEBO_GRID["res"], max_prop)

    def set_field(self, field, origin, res):
        field = np.ascontiguousarray(field, dtype=np.float32)
        nz, ny, nx, four = field.shape
        assert four == 4
        org = np.ascontiguousarray(origin, dtype=np.float64)
        with _lib.on_device(self.device):
            check(lib.rlitg_scene_set_field(self._h, field.ctypes.data, nx, ny, nz, org.ctypes.data, float(res)))
        self.shape, self.origin, self.res = (nz, ny, nx), org, float(res)

    def get_field(self):
        nz, ny, nx = self.shape
        out = np.empty((nz, ny, nx, 4), dtype=np.float32)
        with _lib.on_device(self.device):
            check(lib.rlitg_scene_get_field(self._h, out.ctypes.data))
        return out

    def get_bits(self):
        nz, ny, nx = self.shape
        out = np.empty((nz, ny, (nx + 31) // 32), dtype=np.uint32)
        with _lib.on_device(self.device):
            check(lib.rlitg_scene_get_bits(self._h, out.ctypes.data))
        return out

    def get_dilated(self):
        """-> (radii float32 [ncls], bits uint32 [ncls, nz, ny, ceil(nx/32)]) or (None, None) when not built."""
        n = C.c_int32(0)
        with _lib.on_device(self.device):
            check(lib.rlitg_scene_get_dilated(self._h, C.byref(n), None, None))
            if n.value == 0:
                return None, None
            nz, ny, nx = self.shape
            radii = np.empty(n.value, dtype=np.float32)
            bits = np.empty((n.value, nz, ny, (nx + 31) // 32), dtype=np.uint32)
            check(lib.rlitg_scene_get_dilated(self._h, C.byref(n), radii.ctypes.data, bits.ctypes.data))
        return radii, bits

    def collide(self, q, out=None):
        """cc_srv for a batch: uint8 [B], 1 = some sphere penetrates an occupied region (`out`: optional result buffer)."""
        q = q.contiguous()
        B = q.shape[0]
        if out is None:
            out = torch.empty(B, dtype=torch.uint8, device=q.device)
        dt = _lib.F32 if q.dtype == torch.float32 else _lib.F64
        with _lib.on_device(q.device):
            check(lib.rlitg_sphere_grid_collide(q.data_ptr(), B, dt, self._h, self.base_offset.ctypes.data,
                                                out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out

    def close(self):
        if self._h:
            lib.rlitg_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
