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

/* ---------------------------------------------------------------------------------------------------------
 * Scene: link-sphere table + occupancy grid / distance volume (device resident, owned by the handle)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct RlitgScene RlitgScene;

/* Creates an empty scene on the current CUDA device / frees it and its device memory. */
RLITG_API int rlitg_scene_create(RlitgScene** out);
RLITG_API int rlitg_scene_destroy(RlitgScene* scene);

/* Collision spheres (HOST pointers): link[S] = segment-tip index 0..7 (7 = gripper_link; the fingers ride on it),
 * sorted ascending; center [S,3] in the link frame; radius [S].  S <= 64.  Stands in for MoveIt's
 * CollisionRobotDistanceField sphere decomposition (torm/src/torm_optimizer.cpp:131-135,195-206), which is
 * third-party and unpinned: the table is an input. */
RLITG_API int rlitg_scene_set_spheres(RlitgScene* scene, int S, const int32_t* link, const double* center,
                                      const double* radius);

/* Occupancy volume (HOST uint8 [nz][ny][nx], non-zero = occupied; the pirl_gazebo occ_N.npy format,
 * pirl_gazebo/script/save_pointcloud.py:22-32) -> on device: bit-packed occupancy + exact signed Euclidean
 * distance records (d, dd/dx, dd/dy, dd/dz) per cell with MoveIt PropagationDistanceField semantics
 * (torm/src/torm_optimizer.cpp:110-120).  origin = centre of cell (0,0,0) in the planning frame (base_link). */
RLITG_API int rlitg_scene_set_occupancy(RlitgScene* scene, const uint8_t* occ, int nx, int ny, int nz,
                                        const double* origin, double res, double max_prop);

/* Same with the staging kernels (two exact EDT passes, distance records, bit-packing, dilated volumes) on `stream`; the
 * call synchronises that stream only (its host input and device temporaries must outlive the kernels), never the
 * device.  Launch errors of every staging kernel are reported. */
RLITG_API int rlitg_scene_set_occupancy_stream(RlitgScene* scene, const uint8_t* occ, int nx, int ny, int nz,
                                               const double* origin, double res, double max_prop, void* stream);

/* Distance volume given directly (HOST float [nz][ny][nx][4] records; d >= 1e30 marks an inactive cell). */
RLITG_API int rlitg_scene_set_field(RlitgScene* scene, const float* field, int nx, int ny, int nz,
                                    const double* origin, double res);
/* Read-back for parity tests: field_out float [nz][ny][nx][4]; bits_out uint32 [nz][ny][ceil(nx/32)]. */
RLITG_API int rlitg_scene_get_field(const RlitgScene* scene, float* field_out);
RLITG_API int rlitg_scene_get_bits(const RlitgScene* scene, uint32_t* bits_out);
/* The occupancy dilated by each distinct sphere radius (bit = signed distance < radius), bit-packed like the
 * occupancy: the volumes the binary collision test reads (one bit per sphere; staged in shared memory when they
 * fit, e.g. for the pirl_gazebo grid).  ncls_out int32[1] (0 = not built: more than 8 distinct radii);
 * radii_out float[ncls] and bits_out uint32[ncls][nz][ny][ceil(nx/32)] may be NULL. */
RLITG_API int rlitg_scene_get_dilated(const RlitgScene* scene, int32_t* ncls_out, float* radii_out,
                                      uint32_t* bits_out);

/* cc_srv: `/collision_check` service (pirl_msgs/srv/collision.srv, pirl/src/collision_checker.cpp:39-61) and
 * TormIKSolver::collisionChecking (torm/src/torm_ik_solver.cpp:59-66).  out[b] = 1 when any sphere of
 * configuration q[b] (D = 7, [B,7]) has signed distance - radius < 0.  An approximation of the reference's FCL
 * mesh test, not a parity target.  base_off = chain base -> planning frame (HOST double[3]). */
RLITG_API int rlitg_sphere_grid_collide(const void* q, int64_t B, int dtype, const RlitgScene* scene,
                                        const double* base_off, uint8_t* out, void* stream);

/* TormOptimizer::isCurrentTrajectoryCollisionFree for N trajectories (torm/src/torm_optimizer.cpp:695-717) and the
 * waypoint choice of use_stochastic_descent_ (:1502-1515) in one launch.  traj [N,T,7]; outputs (each nullable):
 *   first_hit [N] int32: first colliding row (the reference's collision_point_, row 0 = start) or -1;
 *   any_hit [N] uint8: 1 = the trajectory collides (feed rlitg_torm_bookkeep with hits_per_trajectory = 1);
 *   stochastic_idx [N] int32: the free-waypoint index for rlitg_torm_cost_grad -- first_hit - 1 when the trajectory
 *   collides (-1 when the start row collides: the reference then indexes row -1, here no increment is taken), else
 *   (int)(u[n] * (T - 2)) for uniform u [N] float32 in [0,1) (NULL: index 0) -- the reference draws random() there. */
RLITG_API int rlitg_torm_collision_point(const void* traj, int64_t N, int T, int dtype, const RlitgScene* scene,
                                         const double* base_off, const float* u, int32_t* first_hit, uint8_t* any_hit,
                                         int32_t* stochastic_idx, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Null-space imitation objective and pseudo-inverse velocity step
 * ------------------------------------------------------------------------------------------------------- */

/* jaco_srv: `/jaco_reward` service (pirl_msgs/srv/jaco.srv, pirl/src/kinematic_solver.cpp:332-351):
 * out[b] = || (I - J^T (J J^T)^-1 J) err[b] ||, J = Jacobian at q[b].  err, q [B,7]; out [B]. */
RLITG_API int rlitg_nullspace_norm(const void* err, const void* q, int64_t B, int dtype, void* out, void* stream);

/* TormIKSolver::vikSolver (torm/src/torm_ik_solver.cpp:169-171) = KDL::ChainIkSolverVel_pinv::CartToJnt:
 * dq[b] = pinv(J(q[b])) twist[b], singular values below eps (KDL default 1e-5) zeroed.  q [B,D], twist [B,6]
 * (linear, angular), dq [B,D]. */
RLITG_API int rlitg_pinv_step(const void* q, const void* twist, int64_t B, int D, int dtype, double eps, void* dq,
                              void* stream);

/* TormOptimizer::checkSingularity (torm/src/torm_optimizer.cpp:748-765): out[b] = |sqrt(det(J J^T))| at q[b]
 * (compare with singularity_lower_bound_).  q [B,D], out [B]. */
RLITG_API int rlitg_manipulability(const void* q, int64_t B, int D, int dtype, void* out, void* stream);

/* KDL::diff(FK(q[b]), target[b]) as used at torm/src/torm_optimizer.cpp:1413,1440 and
 * torm/src/utils/traj_evaluator.cpp:72,79: twist [B,6] = (p_t - p, R * rotvec(R^T R_t)).  target [B,7]. */
RLITG_API int rlitg_pose_diff(const void* q, const void* target, int64_t B, int D, int dtype, void* twist,
                              void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * PIRL path-wise IK environment step (PirlEnv_Fetch, pirl/script/envs/PirlEnv.py)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct RlitgPirlCfg {
    double ll[7], ul[7];    /* joint limits as the env sees them (pirl/launch/pirl.yaml:11-12) */
    int32_t cjoint[7];      /* 1 = continuous joint (pirl.yaml:9) */
    int32_t n_future;       /* t: future targets in the observation (PirlEnv.py:77,311-315) */
    int32_t z_dim;          /* scene latent length appended to the observation (:89,318) */
    int32_t obs_type;       /* 0 onlyConf, 1 Conf+LinkPoses, 2 Conf+EE, 3 onlyLinkPoses (:80-87) */
    int32_t train_mode;     /* mode == 'train' (:193) */
    int32_t refine_flag;    /* refineFlag (:221,233,238,253,258) */
    int32_t task_reward_flag, imitate_reward_flag, null_reward_flag; /* (:130-161) */
    double task_suc_r, dist_suc_r, rot_suc_r, imit_suc_r;
    double thr_early_stop, thr_dist, thr_rot, thr_imitate, zero_task_dist; /* (:123-128) */
    double dist_p[4], rot_p[4], imit_p[4], null_p[4];                      /* (:165-168) */
    double w_col, w_jlimit, w_early_stop;                                  /* (:170-172) */
    int32_t row_floats;     /* values per entry of the targets / demos tables: 0 or 7 = packed [B,Tmax,7]; 8 = rows padded to
                               [B,Tmax,8] (32-byte rows, 32-byte aligned tables); 32 = STEP RECORDS: targets is [B,Tmax,32] with
                               entry (b,t) = [target(t) 7 + pad | demo(t) 7 + pad | target(min(t+1, Tmax-1)) 7 + pad | pad 8]
                               (128 bytes in FP32, 128-byte aligned), the demos argument is ignored.  A per-thread gather costs a
                               whole 128-byte line on the B200, so everything one env step reads of the tables sits in one line */
    int32_t record_has_demo; /* row_floats = 32 only: the records' demo slots are valid (demonstrations were given) */
} RlitgPirlCfg;

/* Observation length for a configuration (PirlEnv.py:79-89). */
RLITG_API int rlitg_pirl_obs_dim(const RlitgPirlCfg* cfg);

/* One env.step(action) for B independent environments (PirlEnv.py:209-289), fused in one kernel:
 * _set_action (:184-207: add, single +-2pi wrap of continuous joints, clip, np.allclose joint-limit test, collision
 * test on the candidate) -> all-link FK (:216) -> _get_obs (:292-320) -> task / imitation / null-space rewards
 * (:495-542) -> done and info flags (:226-282).
 *   q [B,7] in/out; action [B,7] float32 (the policy's dtype, pirl/script/algos/sac/core.py:98-101);
 *   timestep [B] int32 in/out (incremented); max_timestep [B] int32;
 *   targets [B,Tmax,7], demos [B,Tmax,7] (may be NULL) -- [B,Tmax,8] with cfg->row_floats = 8 --, latent [B,z_dim]
 *   (NULL when z_dim = 0);
 *   scene may be NULL (no collision test); base_off HOST double[3];
 *   obs [B,O]; reward [B]; done [B] uint8; flags [B,8] uint8 = (is_col, is_jlimit, is_early_stop, is_timeout,
 *   is_dist_suc, is_rot_suc, is_imitate_suc, is_task_suc); counters [B,4] int32 in/out = (task, dist, rot, imitate)
 *   success counts (may be NULL); errs [B,4] = (dist_err, rot_err, imitate_err, nullspace objective) (may be NULL).
 * Floating tensors share dtype. */
RLITG_API int rlitg_pirl_env_step(const RlitgPirlCfg* cfg, int64_t B, int dtype, void* q, const float* action,
                                  int32_t* timestep, const int32_t* max_timestep, const void* targets,
                                  const void* demos, int Tmax, const void* latent, const RlitgScene* scene,
                                  const double* base_off, void* obs, void* reward, uint8_t* done, uint8_t* flags,
                                  int32_t* counters, void* errs, void* stream);

/* The same step with its argument list bound once (a vectorised PirlEnv_Fetch steps the same buffers for a whole
 * roll-out, PirlEnv.py:209-289 called from pirl/script/algos/sac/sac.py:285-334): rlitg_pirl_bind copies the
 * configuration and the pointers (the caller keeps the buffers alive and re-binds when one of them is replaced),
 * rlitg_pirl_step_bound is rlitg_pirl_env_step on them.  For bindings whose per-argument marshalling costs as much as
 * the 5 us kernel at 16,384 envs (ctypes). */
typedef struct RlitgPirlBound RlitgPirlBound;
RLITG_API int rlitg_pirl_bind(const RlitgPirlCfg* cfg, int64_t B, int dtype, void* q, int32_t* timestep,
                              const int32_t* max_timestep, const void* targets, const void* demos, int Tmax,
                              const void* latent, const RlitgScene* scene, const double* base_off, void* obs, void* reward,
                              uint8_t* done, uint8_t* flags, int32_t* counters, void* errs, RlitgPirlBound** out);
RLITG_API int rlitg_pirl_step_bound(const RlitgPirlBound* bound, const float* action, void* stream);
RLITG_API void rlitg_pirl_bound_destroy(RlitgPirlBound* bound);

/* Observation only (env.reset, PirlEnv.py:378-392; PirlInterpolator::getObs, torm/src/pirl_interpolator.cpp:176-248). */
RLITG_API int rlitg_pirl_obs(const RlitgPirlCfg* cfg, int64_t B, int dtype, const void* q, const int32_t* timestep,
                             const int32_t* max_timestep, const void* targets, int Tmax, const void* latent,
                             void* obs, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * TORM path-following objective and gradient (TormOptimizer / TormCost / traj_evaluator)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct RlitgTormParams {
    int32_t T;              /* target poses = trajectory rows; row 0 = start configuration (fixed) */
    int32_t D;              /* 7 */
    double dt;              /* time_duration_ (torm/include/torm/torm_parameters.h:83) */
    double rot_scale;       /* rotation_scale_factor_ 0.17 */
    double jac_weight;      /* jacobian_update_weight_ */
    double clearance;       /* min_clearence_ */
    double pinv_ridge;      /* pseudo_inverse_ridge_factor_ 1e-4 */
    double vik_eps;         /* KDL::ChainIkSolverVel_pinv eps 1e-5 */
    int32_t use_pinv;       /* use_pseudo_inverse_ */
    int32_t affine_axis_quirk; /* reproduce `axis = joint_transform * axis` (torm/src/torm_optimizer.cpp:1157) */
    double base_off[3];     /* chain base -> planning frame (torm/launch/config/robot/fetch.yaml:18) */
    int32_t cjoint[8];      /* 1 = continuous joint */
    double vel_limit[8];    /* fetch_moveit_assistant/config/joint_limits.yaml:5-39 */
} RlitgTormParams;

/* stage_mask bits */
#define RLITG_STAGE_POSE 1       /* getEndPoseCost(false): end-pose cost sums (torm/src/torm_optimizer.cpp:1399-1425) */
#define RLITG_STAGE_POSE_GRAD 2  /* getEndPoseCost(true) + calculateEndPoseIncrements (:1475-1482) */
#define RLITG_STAGE_SMOOTH 4     /* calculateSmoothnessIncrements (:1466-1473) */
#define RLITG_STAGE_COLL 8       /* performForwardKinematics (:1302-1382) + calculateCollisionIncrements (:1484-1553)
                                    + getCollisionCost (:1448-1463) */
#define RLITG_STAGE_EVAL 16      /* traj_evaluator::getCost (torm/src/utils/traj_evaluator.cpp:104-132) */
#define RLITG_STAGE_KEEP_COSTS 32 /* leave the cost entries of the stages that are off untouched instead of zeroing them:
                                    lets two calls with disjoint stages share one costs buffer (the optimiser loop runs the
                                    end-pose / evaluator call next to the collision-point test + collision call) */
#define RLITG_STAGE_COLL_ROW_ONLY 64 /* with stochastic_idx: compute the collision stage for that one waypoint only (chain
                                    walks of its seven stencil rows, potentials of one row); costs[5] is then 0.  What the optimiser
                                    loop needs with use_stochastic_descent_: getCollisionCost is not part of localOptimizeTSGD */
#define RLITG_TORM_NCOST 10
/* costs [N,10]: 0 eval position, 1 eval rotation, 2 velocity, 3 acceleration, 4 jerk (traj_evaluator means),
 *   5 collision cost (mean over the T rows of the summed sphere potentials), 6 weighted = [0] + 0.17 [1]
 *   (getWeightedPoseCost, the optimiser's cur_Cost), 7 velocity-limit violations, 8 / 9 end-pose position /
 *   rotation error SUMS over the free waypoints (getEndPoseCost = [8] + rot_scale [9]). */

/* Smoothness operator of TormCost (torm/src/torm_cost.cpp:46-77) with the scaling of
 * torm/src/torm_optimizer.cpp:211-236, built once per (T, dt, weights) on the host in FP64 and kept on device:
 * the 13-banded quad_cost_full_ over the T+10 padded rows and the inverse of its free block. */
typedef struct RlitgSmooth RlitgSmooth;
RLITG_API int rlitg_smooth_create(int T, double dt, const double* w3 /*vel, acc, jerk*/, double ridge,
                                  RlitgSmooth** out);
RLITG_API int rlitg_smooth_destroy(RlitgSmooth* s);
/* HOST read-back: Aband [(T+10),13] (A(i, i+k-6)), scale = max coeff of the unscaled inverse (the reference's
 * `// 96.77624`), Ainv [(T-1),(T-1)] (nullable). */
RLITG_API int rlitg_smooth_get(const RlitgSmooth* s, double* Aband, double* scale, double* Ainv);

/* Cost terms and increments for N candidate trajectories at once (one CTA per trajectory).
 *   traj [N,T,D] (row 0 = start, rows 1..T-1 free); targets [T,7] shared by all N;
 *   stochastic_idx [N] int32 (nullable): use_stochastic_descent_ -- the single free-waypoint index (0..T-2) whose
 *   spheres contribute to the collision increment (:1502-1515); NULL = every free waypoint;
 *   inc_smooth / inc_coll / inc_pose [N,T-1,D] (nullable when the stage is off); costs [N,10];
 *   active [N] float32 (nullable): mean number of active spheres (potential >= 1e-4) per evaluated waypoint.
 * Replaces the per-waypoint loops of TormOptimizer (in-process calls, one trajectory at a time) and
 * traj_evaluator; see the stage bits for the exact members. */
RLITG_API int rlitg_torm_cost_grad(const void* traj, const void* targets, int64_t N, int dtype,
                                   const RlitgScene* scene, const RlitgSmooth* smooth, const RlitgTormParams* p,
                                   int stage_mask, const int32_t* stochastic_idx, void* inc_smooth, void* inc_coll,
                                   void* inc_pose, void* costs, float* active, void* stream);

/* Same with the optimiser stage given per trajectory (traj_stage [N] int32 on the device, 0 = feasible, 1 = end-pose; from
 * rlitg_torm_stage_decide): trajectories in the end-pose stage skip the collision / smoothness kernel (their costs[5] is
 * 0 and their inc_smooth / inc_coll rows are left untouched), trajectories in the feasible stage skip the end-pose
 * increments (inc_pose rows untouched).  The evaluator and end-pose cost sums are computed for all.  For
 * lr_schedule_mode_ 1, where every optimiser picks its stage itself (torm/src/torm_optimizer.cpp:400-412). */
RLITG_API int rlitg_torm_cost_grad_staged(const void* traj, const void* targets, int64_t N, int dtype,
                                          const RlitgScene* scene, const RlitgSmooth* smooth, const RlitgTormParams* p,
                                          int stage_mask, const int32_t* traj_stage, const int32_t* stochastic_idx,
                                          void* inc_smooth, void* inc_coll, void* inc_pose, void* costs, float* active,
                                          void* stream);

/* The target path as frames, computed once per path and shared by every later call on it: frames [2T-1,12] =
 * (p, R columns) of the T target poses (KDL Rotation::Quaternion, no normalisation) followed by the T-1 inter-target
 * poses of traj_evaluator::setInterTargetPoses (torm/src/utils/traj_evaluator.cpp:244-257).  The reference rebuilds
 * both for every evaluator it constructs (twice per GD iteration, torm/src/torm_optimizer.cpp:421,432). */
RLITG_API int rlitg_torm_target_frames(const void* targets, int T, int dtype, void* frames, void* stream);
/* rlitg_torm_cost_grad_staged reading the path from `target_frames` (from rlitg_torm_target_frames; `targets` may then
 * be NULL).  target_frames = NULL is the plain call. */
RLITG_API int rlitg_torm_cost_grad_frames(const void* traj, const void* targets, const void* target_frames, int64_t N,
                                          int dtype, const RlitgScene* scene, const RlitgSmooth* smooth,
                                          const RlitgTormParams* p, int stage_mask, const int32_t* traj_stage,
                                          const int32_t* stochastic_idx, void* inc_smooth, void* inc_coll, void* inc_pose,
                                          void* costs, float* active, void* stream);

/* Host-buffer form of rlitg_torm_cost_grad: traj, targets, stochastic_idx and every output are HOST pointers (the
 * reference's TormOptimizer holds Eigen matrices on the host); copies in / out through an internal device workspace
 * and returns when the results are in the caller's buffers.  scene / smooth are the same device-resident handles. */
RLITG_API int rlitg_torm_cost_grad_host(const void* traj, const void* targets, int64_t N, int dtype,
                                        const RlitgScene* scene, const RlitgSmooth* smooth, const RlitgTormParams* p,
                                        int stage_mask, const int32_t* stochastic_idx, void* inc_smooth,
                                        void* inc_coll, void* inc_pose, void* costs);

/* Update step of the two-stage gradient descent, for N trajectories in place:
 * calculateFeasibleIncrements (torm/src/torm_optimizer.cpp:1555-1609) or calculateEndPoseIncrements (:1475-1482),
 * then addIncrementsToTrajectory (:1611-1639) and handleJointLimits (:1189-1268). */
typedef struct RlitgTormUpdate {
    int32_t stage;             /* 0 = feasible stage (smoothness + collision), 1 = end-pose stage */
    int32_t lr_schedule_mode;  /* lr_schedule_mode_: 0 = fixed weights, otherwise max-abs normalisation + alpha blend */
    double smoothness_update_weight; /* mode 0 only */
    double obstacle_update_weight;   /* mode 0 only */
    double joint_update_limit;       /* joint_update_limit_ */
    double joint_min[7], joint_max[7]; /* position bounds; entries of continuous joints are ignored */
} RlitgTormUpdate;
/*   traj [N,T,7] in/out (row 0 untouched); inc_* [N,T-1,7] from rlitg_torm_cost_grad (inc_pose already carries
 *   jacobian_update_weight_); alpha [N] (may be NULL = 0.5) and lr [N] = per-trajectory alpha_ / learning_rate_
 *   (for mode 0 pass parameters_->learning_rate_ in every entry).  Floating tensors share dtype. */
RLITG_API int rlitg_torm_update(void* traj, int64_t N, int dtype, const RlitgSmooth* smooth,
                                const RlitgTormParams* p, const RlitgTormUpdate* u, const void* inc_smooth,
                                const void* inc_coll, const void* inc_pose, const void* alpha, const void* lr,
                                void* stream);

/* Same, with the stage chosen per trajectory: stage_idx [N] int32 on the device, 0 = feasible, 1 = end-pose, negative =
 * leave the trajectory untouched (u->stage is ignored).  lr_schedule_mode_ 1 lets each optimiser switch stages on its
 * own feasibility history (torm/src/torm_optimizer.cpp:400-412), so a lock-step batch mixes both kinds of update. */
RLITG_API int rlitg_torm_update_staged(void* traj, int64_t N, int dtype, const RlitgSmooth* smooth,
                                       const RlitgTormParams* p, const RlitgTormUpdate* u, const int32_t* stage_idx,
                                       const void* inc_smooth, const void* inc_coll, const void* inc_pose,
                                       const void* alpha, const void* lr, void* stream);

/* Optimiser state of TormOptimizer::localOptimizeTSGD between a cost evaluation and the update step, for N trajectories
 * in lock step (one launch instead of the reference's scalar control flow, torm/src/torm_optimizer.cpp:400-560):
 *   stage_idx [N] int32 out (nullable): lr_schedule_mode_ 1 stage choice (:400-412) from the state BEFORE this call,
 *     0 = feasible, 1 = end-pose -- feed it to rlitg_torm_update_staged;
 *   infea_col / infea_vel [N] uint8 in/out: isCurrentTrajectoryCollisionFree (:436-440, any of hits [N*T], the
 *     per-waypoint flags of rlitg_sphere_grid_collide; NULL or use_collision_check = 0 -> feasible) and
 *     checkJointVelocityLimit (:441-445, costs[:,7] != 0);
 *   alpha, lr [N] in/out: alpha_ (:447-469) and learning_rate_ (:471-499); lr_iter (device scalar, nullable) is the
 *     cosine-schedule value of this iteration for mode 2 (a device value so that a captured CUDA graph can be replayed);
 *   best_cost [N], best_traj [N,T,7] in/out: best_trajectory_backup_cost_ / best_trajectory_ (:506-560), recorded when
 *     costs[:,6] improves and the trajectory is feasible and not flagged in singular [N] uint8 (nullable). */
typedef struct RlitgTormSchedule {
    int32_t lr_schedule_mode;   /* 0..3 */
    int32_t stage;              /* which half of the iteration this is: 0 (after :400-405) or 1 (:406-411) */
    int32_t use_collision_check, use_velocity_check;
    double learning_rate;       /* parameters_->learning_rate_ */
    double learning_rate_min, learning_rate_max; /* / 100 and * 1000 (:99-101) */
    int32_t hits_per_trajectory; /* 0: hits is [N*T] per-waypoint flags; 1: hits is [N] (any_hit of rlitg_torm_collision_point) */
    int32_t reserved_;
} RlitgTormSchedule;
/* The stage choice alone (:400-412), to be made BEFORE the cost evaluation of the half-iteration sch->stage so that
 * rlitg_torm_cost_grad_staged can skip the work the other stage would need; pass stage_idx = NULL to
 * rlitg_torm_bookkeep then. */
RLITG_API int rlitg_torm_stage_decide(int64_t N, int dtype, const void* lr, const uint8_t* infea_col,
                                      const uint8_t* infea_vel, const RlitgTormSchedule* sch, int32_t* stage_idx,
                                      void* stream);
RLITG_API int rlitg_torm_bookkeep(const void* traj, void* best_traj, int64_t N, int T, int dtype, const void* costs,
                                  const uint8_t* hits, const uint8_t* singular, void* alpha, void* lr,
                                  uint8_t* infea_col, uint8_t* infea_vel, void* best_cost, int32_t* stage_idx,
                                  const void* lr_iter, const RlitgTormSchedule* sch, void* stream);

/* Candidate selection of the Greedy / Linear initialisers: replays the sequential rule of
 * TormOptimizer::getNewTrajectory (torm/src/torm_optimizer.cpp:973-1003; same at :1066-1096) over C scored
 * candidates: eCost = ([8] + rot_scale [9]) / (Tseg - 1) (ranged getEndPoseCost :1427-1446), candidates with
 * eCost above the running best TOTAL cost are skipped (the reference's pruning quirk), total =
 * pose_w eCost + col_w [5].  costs [C,10] from rlitg_torm_cost_grad run on the C interpolated segments
 * ([C,Tseg,7], row 0 = segment start).  best_idx int32[1], best_cost [1] (nullable) on the device. */
RLITG_API int rlitg_greedy_select(const void* costs, int C, int Tseg, int dtype, double rot_scale, double pose_w,
                                  double col_w, int32_t* best_idx, void* best_cost, void* stream);

/* Measurement aid: launches `ctas` x 256 threads each running 16 independent FFMA chains of `iters` steps
 * (2 * ctas * 256 * iters * 16 flops); bench.py times it with CUDA events to get the FP32 FMA-pipe peak that
 * the TORM kernel's roofline fraction is quoted against. */
RLITG_API int rlitg_fma_probe(int ctas, int iters, float* out, void* stream);
/* Same for the FP64 FMA pipe (8 independent DFMA chains per thread: 2 * ctas * 256 * iters * 8 flops): the FP32 build of
 * the TORM collision kernel walks the chain in FP64, which is only affordable where this pipe is not vestigial. */
RLITG_API int rlitg_dfma_probe(int ctas, int iters, double* out, void* stream);

/* Measurement aid: streams `ntiles` tiles of `tile_bytes` from shared memory to `dst` with TMA bulk stores from
 * `ctas` persistent CTAs (the drain path of rlitg_fk_jac without arithmetic), optionally reading `read_floats`
 * floats per tile from `src` first.  Gives the HBM ceiling for fk_jac's 1 : 7.7 read : write mix. */
RLITG_API int rlitg_write_probe(void* dst, const void* src, int64_t ntiles, int tile_bytes, int read_floats,
                                int ctas, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Device memory helpers for hosts built WITHOUT the CUDA toolkit.  The reference's C++ packages are plain catkin
 * targets (torm/CMakeLists.txt:20-29, pirl/CMakeLists.txt:6-22: no CUDA language, no cudart); with these a
 * maintainer links librlitg.so alone.  include/rlitg.hpp (the C++ mirror of TormIKSolver / TormOptimizer) uses
 * nothing else.  Callers that already own device memory (torch, their own cudaMalloc) never need them: device
 * pointers and streams of the primary context are interchangeable.
 * ------------------------------------------------------------------------------------------------------- */
RLITG_API int rlitg_device_count(int* count);
RLITG_API int rlitg_set_device(int device);
RLITG_API int rlitg_dev_alloc(void** out, uint64_t bytes);
RLITG_API int rlitg_dev_free(void* ptr);
RLITG_API int rlitg_dev_zero(void* ptr, uint64_t bytes, void* stream);
/* stream-ordered copies; they return once the copy is queued unless the host buffer is pageable */
RLITG_API int rlitg_copy_h2d(void* dst_dev, const void* src_host, uint64_t bytes, void* stream);
RLITG_API int rlitg_copy_d2h(void* dst_host, const void* src_dev, uint64_t bytes, void* stream);
RLITG_API int rlitg_copy_d2d(void* dst_dev, const void* src_dev, uint64_t bytes, void* stream);
RLITG_API int rlitg_stream_sync(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLITG_H_ */
