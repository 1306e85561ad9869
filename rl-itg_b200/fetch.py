"""Fetch robot constants on the host side: joint limits as the reference's configuration files give them, and
the build-defined default collision-sphere table.

Limits: pirl/launch/pirl.yaml:9-14 (continuous joints [2,4,6] limited to +-3.141592, actions +-0.26),
fetch_moveit_assistant/config/joint_limits.yaml:5-39 (velocity limits), fetch.urdf:107 (torso 0..0.38615).
base offset: torm/launch/config/robot/fetch.yaml:18 fk_base_position (chain base torso_lift_link in base_link).

Sphere table: MoveIt's own decomposition (CollisionRobotDistanceField, torm/src/torm_optimizer.cpp:131-135) is
third-party, unpinned and unverifiable here, so the table is an INPUT of the scene.  The
default below is derived from the collision-mesh bounding boxes of the links the reference keeps
(start_collision_ = 3: elbow_flex_link .. finger links, torm/src/torm_optimizer.cpp:201-206; boxes in SURVEY
appendix A.5): spheres of radius = half the larger cross-section extent, evenly spaced along the longest box
axis.  24 spheres; finger links ride on gripper_link (finger joints fixed at 0.1,
torm/launch/config/problem/hello.yaml).
"""
from __future__ import annotations

import numpy as np

LL = np.array([-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592])
UL = np.array([1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592])
CONTINUOUS = np.array([0, 0, 1, 0, 1, 0, 1], dtype=np.int32)
VEL_LIMIT = np.array([1.256, 1.454, 1.571, 1.521, 1.571, 2.268, 2.268])
ACT_LL = np.full(7, -0.26)
ACT_UL = np.full(7, 0.26)
TORSO_RANGE = (0.0, 0.38615)
BASE_OFFSET = np.array([-0.086875, 0.0, 0.37743])

# (link index, box min, box max, number of spheres); link 7 = gripper_link
_BOXES = [
    (3, (-0.052, -0.052, -0.052), (0.166, 0.0705, 0.052), 5),     # elbow_flex_link
    (4, (-0.032, -0.087, -0.052), (0.1765, 0.052, 0.052), 5),     # forearm_roll_link
    (5, (-0.052, -0.054, -0.052), (0.1165, 0.067, 0.052), 4),     # wrist_flex_link
    (6, (-0.023, -0.0519, -0.052), (0.0195, 0.0519, 0.052), 2),   # wrist_roll_link
    (7, (-0.154, -0.063, -0.0399), (-0.032, 0.063, 0.0396), 4),   # gripper_link
    (7, (-0.029, -0.0018, -0.0132), (0.031, 0.0126, 0.0132), 2),  # r_gripper_finger_link in the gripper_link frame
    (7, (-0.029, -0.0126, -0.0132), (0.031, 0.0018, 0.0132), 2),  # l_gripper_finger_link in the gripper_link frame
]


def default_spheres():
    """-> (link int32 [S], center float64 [S,3], radius float64 [S]), sorted by link."""
    link, center, radius = [], [], []
    for lk, lo, hi, n in _BOXES:
        lo, hi = np.array(lo), np.array(hi)
        ext = hi - lo
        ax = int(np.argmax(ext))
        others = [k for k in range(3) if k != ax]
        r = 0.5 * max(ext[others[0]], ext[others[1]])
        inset = min(r, 0.5 * ext[ax])
        mid = 0.5 * (lo + hi)
        for t in np.linspace(lo[ax] + inset, hi[ax] - inset, n):
            c = mid.copy()
            c[ax] = t
            link.append(lk); center.append(c); radius.append(r)
    # the device table is stored as float32 (cx, cy, cz, r): return exactly representable values so that a
    # double-precision consumer (the oracle) sees the same table
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    return np.array(link, dtype=np.int32), f32(center), f32(radius)
