/*
 * rlitg.h -- C ABI of librlitg.so: B200 (sm_100a) kernels for RL-ITG's kinematics / cost hot path.
 *
 * The reference (MinsungYoon/RL-ITG) has no FFI for this path: the seams are ROS service proxies on the Python
 * side and TormIKSolver / TormOptimizer method calls on the C++ side.  Each entry point below
 * names the reference interface it replaces (paths relative to the reference root).  INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Functions WITHOUT a `_host` suffix take DEVICE pointers and a CUDA stream
 *     (`void* stream` = cudaStream_t; NULL = the legacy default stream); they are stream-ordered, never
 *     allocate and never synchronise.  `_host` functions take HOST pointers, copy in and out through an
 *     internal pinned/device workspace and return when the results are in the caller's buffers.
 *   - dtype: RLITG_F32 (production) or RLITG_F64 (verification build of the same kernels).
 *   - Return value: RLITG_OK (0) or a negative error code; rlitg_last_error() gives the message (thread local).
 *     The reference's convention is a `bool result` field the caller retries on (pirl_msgs/srv/fk.srv:4).
 *   - D = 7: arm chain torso_lift_link -> gripper_link (torm/launch/config/robot/fetch.yaml:2-16).
 *     D = 8: prismatic torso_lift_joint prepended (chain base_link -> gripper_link).
 *   - Poses are [x, y, z, qx, qy, qz, qw] (the fk.srv / allfk.srv layout, pirl/src/kinematic_solver.cpp:246-256).
 *   - There is no CPU fallback: every function fails with RLITG_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef RLITG_H_
#define RLITG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RLITG_API __attribute__((visibility("default")))
#else
#define RLITG_API
#endif

#define RLITG_VERSION 100

/* dtypes */
#define RLITG_F32 0
#define RLITG_F64 1

/* layout flags */
#define RLITG_Q_SOA 1   /* q is [D,B] instead of [B,D] */
#define RLITG_OUT_SOA 2 /* outputs are [3,B] / [9,B] / [4,B] / [6*D,B] instead of [B,3] / [B,9] / [B,4] / [B,6,D] */

/* error codes */
#define RLITG_OK 0
#define RLITG_ERR_NULL (-1)
#define RLITG_ERR_SHAPE (-2)
#define RLITG_ERR_DTYPE (-3)
#define RLITG_ERR_ALIGN (-4)
#define RLITG_ERR_CUDA (-5)
#define RLITG_ERR_ARG (-6)

RLITG_API int rlitg_version(void);
RLITG_API const char* rlitg_last_error(void);
/* number of kernels this library has launched in this process (all threads); bench.py reports the delta */
RLITG_API int64_t rlitg_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Kinematics
 * ------------------------------------------------------------------------------------------------------- */

/* Fused FK + geometric Jacobian over B configurations.  Any of pos / rot / quat / J may be NULL.
 *   pos [B,3]; rot [B,9] row-major; quat [B,4] xyzw (KDL GetQuaternion, sign not canonicalised);
 *   J [B,6,D] row-major, rows 0-2 linear, 3-5 angular, reference point = end-effector origin, chain-base frame.
 * Replaces TormIKSolver::fkSolver + getJacobian (torm/src/torm_ik_solver.cpp:156-158,173-175) and
 * KinematicsSolverSrv::fkSolver / getJacobian (pirl/src/kinematic_solver.cpp:110-112,202-204), i.e.
 * KDL::ChainFkSolverPos_recursive::JntToCart and KDL::ChainJntToJacSolver::JntToJac. */
RLITG_API int rlitg_fk_jac(const void* q, int64_t B, int D, int dtype, int flags, void* pos, void* rot, void* quat, void* J,
                 void* stream);

/* fk_srv: `/fk_solver` service, pirl_msgs/srv/fk.srv, pirl/src/kinematic_solver.cpp:237-261.
 * pose [B,7] = x y z qx qy qz qw. */
RLITG_API int rlitg_fk(const void* q, int64_t B, int D, int dtype, int flags, void* pos, void* quat, void* stream);

/* allfk_srv: `/allLinkfk_solver` service, pirl_msgs/srv/allfk.srv, pirl/src/kinematic_solver.cpp:206-234;
 * TormIKSolver::allfkSolver torm/src/torm_ik_solver.cpp:159-163.  out [B, D+1, 7]. */
RLITG_API int rlitg_allfk(const void* q, int64_t B, int D, int dtype, int flags, void* out, void* stream);

/* TormIKSolver::getJacobian torm/src/torm_ik_solver.cpp:173-175.  J [B,6,D]. */
RLITG_API int rlitg_jacobian(const void* q, int64_t B, int D, int dtype, int flags, void* J, void* stream);

/* Host-buffer form of rlitg_fk_jac (the end-to-end path: H2D of q, kernel, D2H of the outputs, chunked and
 * overlapped on internal streams).  Pinned host buffers give full PCIe bandwidth; pageable ones work. */
RLITG_API int rlitg_fk_jac_host(const void* q, int64_t B, int D, int dtype, int flags, void* pos, void* rot, void* quat,
                      void* J);

#ifdef __cplusplus
}
#endif
#endif /* RLITG_H_ */
