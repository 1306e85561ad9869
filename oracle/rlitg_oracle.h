/*
 * rlitg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C FP64 restatement of the arithmetic on RL-ITG's kinematics / cost hot
 * path.  The reference (MinsungYoon/RL-ITG) cannot be compiled here: every file on
 * the path includes ROS / orocos-KDL / MoveIt / Eigen headers, none of which exist
 * in this image.  So each function below restates the
 * published algorithm of the third-party routine the reference calls, and cites
 * the reference call site it stands in for.
 *
 * Parity status:
 *   - FK (orc_fk / orc_allfk / orc_fk_quat): PINNED by the reference's own data
 *     fixtures (hello.yaml known answer, 82,610 (q -> pose) pairs; tests/golden).
 *   - TormCost smoothness matrix (orc_smooth_setup): PINNED by the 96.77624
 *     known answer at torm/src/torm_optimizer.cpp:235.
 *   - traj_evaluator costs: regression anchors only.
 *   - Jacobian, pinv step, null-space norm, PIRL env step, collision / end-pose
 *     increments: "parity unpinned" -- no reference artefact stores their values;
 *     pinned by self-consistency tests only (finite differences, J J+ J = J).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#ifndef RLITG_ORACLE_H_
#define RLITG_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- kinematics (KDL restatement; D = 7 arm chain torso_lift_link->gripper_link,
 *      D = 8 prepends the prismatic torso_lift_joint, chain base_link->gripper_link) */
int orc_fk(const double* q, int64_t B, int D, double* pos /*[B,3]*/, double* rot /*[B,9] row-major*/);
int orc_fk_quat(const double* q, int64_t B, int D, double* out /*[B,7] x y z qx qy qz qw*/);
int orc_allfk(const double* q, int64_t B, int D, double* out /*[B,D+1,7]*/);
int orc_jacobian(const double* q, int64_t B, int D, double* J /*[B,6,D] row-major*/);
int orc_fk_jac(const double* q, int64_t B, int D, double* pos, double* rot, double* J);
int orc_pose_diff(const double* q, const double* target /*[B,7]*/, int64_t B, int D, double* twist /*[B,6]*/);
int orc_pinv_step(const double* q, const double* twist /*[B,6]*/, int64_t B, int D, double eps,
                  double* dq /*[B,D]*/);
int orc_nullspace_norm(const double* err /*[B,7]*/, const double* q /*[B,7]*/, int64_t B, double* out /*[B]*/);
int orc_manipulability(const double* q, int64_t B, int D, double* out /*[B] sqrt(det(J J^T))*/);
void orc_quat_to_rot_kdl(const double* quat_xyzw, double* R9);
void orc_rot_to_quat_kdl(const double* R9, double* quat_xyzw);
double orc_rot_angle_kdl(const double* R9, double* axis3);

/* ---- scene: occupancy -> signed distance records, sphere lookup */
typedef struct {
    int S;                 /* collision spheres per configuration */
    const int32_t* link;   /* [S] segment-tip index 0..7 (7 = gripper_link) */
    const double* center;  /* [S,3] in the link frame */
    const double* radius;  /* [S] */
    int nx, ny, nz;
    double origin[3];      /* centre of cell (0,0,0), planning frame */
    double res;
    const float* field;    /* [nz][ny][nx][4] = (d, gx, gy, gz); d >= 1e30 = inactive */
} OrcScene;

int orc_build_field(const uint8_t* occ /*[nz][ny][nx]*/, int nx, int ny, int nz, double res,
                    double max_prop, float* field /*[nz][ny][nx][4]*/);
int orc_sphere_collide(const double* q /*[B,7]*/, int64_t B, const OrcScene* sc, const double* base_off,
                       uint8_t* out /*[B]*/);

/* ---- TORM */
typedef struct {
    int T;                  /* target poses = full trajectory rows (row 0 = start) */
    int D;                  /* 7 */
    double dt;              /* time_duration_ */
    double rot_scale;       /* rotation_scale_factor_ 0.17 */
    double jac_weight;      /* jacobian_update_weight_ */
    double clearance;       /* min_clearence_ */
    double pinv_ridge;      /* pseudo_inverse_ridge_factor_ 1e-4 */
    double vik_eps;         /* KDL ChainIkSolverVel_pinv eps 1e-5 */
    int use_pinv;           /* use_pseudo_inverse_ */
    int affine_axis_quirk;  /* SURVEY appendix B item 1 */
    double base_off[3];     /* chain base -> planning frame */
    int cjoint[8];          /* 1 = continuous joint */
    double vel_limit[8];
} OrcTormParams;

#define ORC_STAGE_POSE 1   /* getEndPoseCost(false): cost sums only */
#define ORC_STAGE_POSE_GRAD 2 /* getEndPoseCost(true) + calculateEndPoseIncrements */
#define ORC_STAGE_SMOOTH 4 /* calculateSmoothnessIncrements */
#define ORC_STAGE_COLL 8   /* performForwardKinematics + calculateCollisionIncrements + getCollisionCost */
#define ORC_STAGE_EVAL 16  /* traj_evaluator::getCost */
#define ORC_NCOST 10
/* costs[N,10]: 0 eval pose, 1 eval rot, 2 vel, 3 acc, 4 jerk, 5 collision cost, 6 weighted (0 + 0.17*1),
 *              7 n_vel_violation, 8 end-pose position sum, 9 end-pose rotation sum */

int orc_smooth_setup(int T, double dt, const double* w3, double ridge, double* Aband /*[(T+10),13]*/,
                     double* scale_out, double* Ainv /*nullable [(T-1),(T-1)]*/);
int orc_torm_cost_grad(const double* traj /*[N,T,D]*/, const double* targets /*[T,7]*/, const OrcScene* sc,
                       const OrcTormParams* p, const double* Aband, int stage_mask,
                       const int32_t* stochastic_idx /*nullable [N]*/, int64_t N, double* inc_smooth,
                       double* inc_coll, double* inc_pose /*[N,T-1,D] each, nullable*/,
                       double* costs /*[N,10]*/, double* active_spheres /*nullable [N]*/);
int orc_traj_eval(const double* traj, const double* targets, const OrcTormParams* p, int64_t N,
                  double* out /*[N,6] pose rot vel acc jerk nviol*/);

/* TORM update step (feasible or end-pose stage): traj [N,T,D] in/out; Ainv scaled [(T-1),(T-1)]; alpha, lr [N] */
int orc_torm_update(double* traj, int64_t N, int T, int D, const double* Ainv, int stage, int mode,
                    const double* inc_smooth, const double* inc_coll, const double* inc_pose, const double* alpha,
                    const double* lr, double sw, double ow, double limit, const int* cjoint, const double* jmin,
                    const double* jmax);

/* ---- PIRL path-wise IK env step */
typedef struct {
    double ll[7], ul[7];
    int cjoint[7];
    int n_future;          /* t */
    int z_dim;
    int obs_type;          /* 0 onlyConf, 1 Conf+LinkPoses, 2 Conf+EE, 3 onlyLinkPoses */
    int train_mode;
    int refine_flag;
    int task_reward_flag, imitate_reward_flag, null_reward_flag;
    double task_suc_r, dist_suc_r, rot_suc_r, imit_suc_r;
    double thr_early_stop, thr_dist, thr_rot, thr_imitate, zero_task_dist;
    double dist_p[4], rot_p[4], imit_p[4], null_p[4];
    double w_col, w_jlimit, w_early_stop;
} OrcPirlCfg;

int orc_pirl_obs_dim(const OrcPirlCfg* c);
int orc_pirl_env_step(const OrcPirlCfg* c, int64_t B, double* q /*[B,7] in/out*/, const float* action /*[B,7]*/,
                      int32_t* timestep /*[B] in/out*/, const int32_t* max_timestep /*[B]*/,
                      const double* targets /*[B,Tmax,7]*/, const double* demos /*[B,Tmax,7]*/, int Tmax,
                      const double* latent /*[B,z]*/, const OrcScene* sc /*nullable*/, const double* base_off,
                      double* obs /*[B,O]*/, double* reward /*[B]*/, uint8_t* done /*[B]*/,
                      uint8_t* flags /*[B,8]*/, int32_t* counters /*[B,4] in/out*/, double* errs /*[B,4]*/);
int orc_pirl_obs(const OrcPirlCfg* c, int64_t B, const double* q, const int32_t* timestep,
                 const int32_t* max_timestep, const double* targets, int Tmax, const double* latent,
                 double* obs);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
