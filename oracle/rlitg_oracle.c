/*
 * rlitg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See rlitg_oracle.h.
 *
 * Citations are file:line under /root/reference (MinsungYoon/RL-ITG).  The third-party routines the
 * reference calls (orocos_kdl, MoveIt collision_distance_field, numpy-quaternion; none vendored, none
 * version-pinned -- SURVEY.md 8c) are restated from their published algorithms, operation order kept
 * literal so that FP64 agreement at 1e-12 with the CUDA verification build means something.
 */
#include "rlitg_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * Chain constants: fetch_description/robots/fetch.urdf
 *   :102-108 torso_lift_joint (prismatic z), :188-195 shoulder_pan (z), :218-225 shoulder_lift (y),
 *   :248-255 upperarm_roll (x), :278-285 elbow_flex (y), :308-315 forearm_roll (x),
 *   :338-345 wrist_flex (y), :368-375 wrist_roll (x), :395-400 gripper_axis (fixed).  All rpy = 0.
 * ---------------------------------------------------------------------------------------------- */
enum { JT_FIXED = 0, JT_REV = 1, JT_PRIS = 2 };
typedef struct {
    int type;
    double origin[3];
    double axis[3];
} OrcSeg;

static const OrcSeg kSegs[9] = {
    {JT_PRIS, {-0.086875, 0.0, 0.37743}, {0, 0, 1}},          /* torso_lift_joint (D = 8 only) */
    {JT_REV, {0.119525, 0.0, 0.34858}, {0, 0, 1}},            /* shoulder_pan_joint   */
    {JT_REV, {0.117, 0.0, 0.0599999999999999}, {0, 1, 0}},    /* shoulder_lift_joint  */
    {JT_REV, {0.219, 0.0, 0.0}, {1, 0, 0}},                   /* upperarm_roll_joint  */
    {JT_REV, {0.133, 0.0, 0.0}, {0, 1, 0}},                   /* elbow_flex_joint     */
    {JT_REV, {0.197, 0.0, 0.0}, {1, 0, 0}},                   /* forearm_roll_joint   */
    {JT_REV, {0.1245, 0.0, 0.0}, {0, 1, 0}},                  /* wrist_flex_joint     */
    {JT_REV, {0.1385, 0.0, 0.0}, {1, 0, 0}},                  /* wrist_roll_joint     */
    {JT_FIXED, {0.16645, 0.0, 0.0}, {0, 1, 0}},               /* gripper_axis (fixed) */
};

/* segment table for a D-joint chain: D = 7 starts at shoulder_pan, D = 8 at torso_lift */
static const OrcSeg* chain_segs(int D, int* nseg) {
    *nseg = D + 1;
    return (D == 8) ? kSegs : kSegs + 1;
}

/* ---- minimal KDL::Vector / Rotation / Frame / Twist ------------------------------------------- */
typedef struct { double v[3]; } Vec;
typedef struct { double d[9]; } Rot; /* row-major, KDL Rotation::data */
typedef struct { Rot M; Vec p; } Frame;
typedef struct { Vec vel, rot; } Twist;

static Vec v3(double x, double y, double z) { Vec r = {{x, y, z}}; return r; }
static Vec vadd(Vec a, Vec b) { return v3(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
static Vec vsub(Vec a, Vec b) { return v3(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }
static Vec vscale(Vec a, double s) { return v3(a.v[0] * s, a.v[1] * s, a.v[2] * s); }
static double vdot(Vec a, Vec b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
static Vec vcross(Vec a, Vec b) {
    return v3(a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2],
              a.v[0] * b.v[1] - a.v[1] * b.v[0]);
}
static double vnorm(Vec a) { return sqrt(vdot(a, a)); }

static Rot rot_identity(void) { Rot r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; return r; }
static Vec rot_apply(const Rot* R, Vec v) {
    return v3(R->d[0] * v.v[0] + R->d[1] * v.v[1] + R->d[2] * v.v[2],
              R->d[3] * v.v[0] + R->d[4] * v.v[1] + R->d[5] * v.v[2],
              R->d[6] * v.v[0] + R->d[7] * v.v[1] + R->d[8] * v.v[2]);
}
static Rot rot_mul(const Rot* A, const Rot* B) {
    Rot C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C.d[3 * i + j] = A->d[3 * i] * B->d[j] + A->d[3 * i + 1] * B->d[3 + j] + A->d[3 * i + 2] * B->d[6 + j];
    return C;
}
static Rot rot_transpose(const Rot* A) {
    Rot C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C.d[3 * i + j] = A->d[3 * j + i];
    return C;
}
/* KDL Rotation::Rot2(axis, angle): Rodrigues for a unit axis */
static Rot rot2(Vec a, double angle) {
    double ct = cos(angle), st = sin(angle), vt = 1 - ct;
    double m_vt_0 = vt * a.v[0], m_vt_1 = vt * a.v[1], m_vt_2 = vt * a.v[2];
    double m_st_0 = a.v[0] * st, m_st_1 = a.v[1] * st, m_st_2 = a.v[2] * st;
    double m_vt_0_1 = m_vt_0 * a.v[1], m_vt_0_2 = m_vt_0 * a.v[2], m_vt_1_2 = m_vt_1 * a.v[2];
    Rot r = {{ct + m_vt_0 * a.v[0], -m_st_2 + m_vt_0_1, m_st_1 + m_vt_0_2,
              m_st_2 + m_vt_0_1, ct + m_vt_1 * a.v[1], -m_st_0 + m_vt_1_2,
              -m_st_1 + m_vt_0_2, m_st_0 + m_vt_1_2, ct + m_vt_2 * a.v[2]}};
    return r;
}
/* KDL Frame product: (M1 M2, M1 p2 + p1) */
static Frame frame_mul(const Frame* a, const Frame* b) {
    Frame c;
    c.M = rot_mul(&a->M, &b->M);
    c.p = vadd(rot_apply(&a->M, b->p), a->p);
    return c;
}
/* KDL Segment::pose(q).  kdl_parser builds Joint(origin, M_origin*axis) and f_tip such that
 * pose(q) = Frame(Rot2(axis, q) * M_origin, origin); all rpy are zero so M_origin = I. */
static Frame seg_pose(const OrcSeg* s, double q) {
    Frame f;
    f.p = v3(s->origin[0], s->origin[1], s->origin[2]);
    if (s->type == JT_REV) {
        f.M = rot2(v3(s->axis[0], s->axis[1], s->axis[2]), q);
    } else {
        f.M = rot_identity();
        if (s->type == JT_PRIS) f.p = vadd(f.p, vscale(v3(s->axis[0], s->axis[1], s->axis[2]), q));
    }
    return f;
}

/* KDL Rotation::Quaternion(x,y,z,w): no normalisation */
void orc_quat_to_rot_kdl(const double* q, double* R) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * x * y - 2 * w * z; R[2] = 2 * x * z + 2 * w * y;
    R[3] = 2 * x * y + 2 * w * z; R[4] = w2 - x2 + y2 - z2; R[5] = 2 * y * z - 2 * w * x;
    R[6] = 2 * x * z - 2 * w * y; R[7] = 2 * y * z + 2 * w * x; R[8] = w2 - x2 - y2 + z2;
}
/* KDL Rotation::GetQuaternion: trace-branch algorithm, sign not canonicalised */
void orc_rot_to_quat_kdl(const double* R, double* q) {
    double trace = R[0] + R[4] + R[8];
    double epsilon = 1E-12;
    double x, y, z, w;
    if (trace > epsilon) {
        double s = 0.5 / sqrt(trace + 1.0);
        w = 0.25 / s;
        x = (R[7] - R[5]) * s;
        y = (R[2] - R[6]) * s;
        z = (R[3] - R[1]) * s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = 2.0 * sqrt(1.0 + R[0] - R[4] - R[8]);
        w = (R[7] - R[5]) / s;
        x = 0.25 * s;
        y = (R[1] + R[3]) / s;
        z = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = 2.0 * sqrt(1.0 + R[4] - R[0] - R[8]);
        w = (R[2] - R[6]) / s;
        x = (R[1] + R[3]) / s;
        y = 0.25 * s;
        z = (R[5] + R[7]) / s;
    } else {
        double s = 2.0 * sqrt(1.0 + R[8] - R[0] - R[4]);
        w = (R[3] - R[1]) / s;
        x = (R[2] + R[6]) / s;
        y = (R[5] + R[7]) / s;
        z = 0.25 * s;
    }
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}
/* KDL Rotation::GetRotAngle(axis, eps = 1e-6).  Branch structure of orocos_kdl 1.4: the
 * antisymmetric part is tested against eps first (identity -> angle 0, axis z; otherwise angle pi with the
 * axis from the symmetric part).  In the general branch the angle is atan2(|a|/2, (tr-1)/2), a = the
 * antisymmetric vector: identical to KDL's acos((tr-1)/2) for an orthonormal matrix, but well conditioned
 * near 0 and pi and NaN-free for the slightly non-orthonormal targets that unnormalised 6-digit
 * quaternions give (SURVEY section 7 "hard parts"; it is also the form that reproduces the SURVEY 8c
 * regression anchors). */
double orc_rot_angle_kdl(const double* R, double* axis) {
    const double epsilon = 1e-6, epsilon2 = 1e-5;
    double x, y, z;
    if ((fabs(R[1] - R[3]) < epsilon) && (fabs(R[2] - R[6]) < epsilon) && (fabs(R[5] - R[7]) < epsilon)) {
        if ((fabs(R[1] + R[3]) < epsilon2) && (fabs(R[2] + R[6]) < epsilon2) && (fabs(R[5] + R[7]) < epsilon2) &&
            (fabs(R[0] + R[4] + R[8] - 3) < epsilon2)) {
            axis[0] = 0; axis[1] = 0; axis[2] = 1;
            return 0.0;
        }
        double xx = (R[0] + 1) / 2, yy = (R[4] + 1) / 2, zz = (R[8] + 1) / 2;
        double xy = (R[1] + R[3]) / 4, xz = (R[2] + R[6]) / 4, yz = (R[5] + R[7]) / 4;
        if ((xx > yy) && (xx > zz)) {
            x = sqrt(xx); y = xy / x; z = xz / x;
        } else if (yy > zz) {
            y = sqrt(yy); x = xy / y; z = yz / y;
        } else {
            z = sqrt(zz); x = xz / z; y = yz / z;
        }
        axis[0] = x; axis[1] = y; axis[2] = z;
        return M_PI;
    }
    double f = (R[0] + R[4] + R[8] - 1) / 2;
    x = (R[7] - R[5]); y = (R[2] - R[6]); z = (R[3] - R[1]);
    double r = sqrt(x * x + y * y + z * z);
    axis[0] = x / r; axis[1] = y / r; axis[2] = z / r;
    return atan2(r / 2, f);
}
/* KDL diff(Frame a, Frame b, dt = 1): vel = p_b - p_a; rot = R_a * (R_a^T R_b).GetRot() */
static Twist frame_diff(const Frame* a, const Frame* b) {
    Twist t;
    t.vel = vsub(b->p, a->p);
    Rot at = rot_transpose(&a->M);
    Rot rel = rot_mul(&at, &b->M);
    double axis[3];
    double angle = orc_rot_angle_kdl(rel.d, axis);
    Vec rv = v3(axis[0] * angle, axis[1] * angle, axis[2] * angle);
    t.rot = rot_apply(&a->M, rv);
    return t;
}

/* KDL ChainFkSolverPos_recursive::JntToCart(q, vector<Frame>):
 * torm/src/torm_ik_solver.cpp:156-163, pirl/src/kinematic_solver.cpp:110-115 */
static void chain_fk_all(const double* q, int D, Frame* out /*[D+1]*/) {
    int nseg;
    const OrcSeg* segs = chain_segs(D, &nseg);
    Frame cur;
    cur.M = rot_identity();
    cur.p = v3(0, 0, 0);
    int j = 0;
    for (int i = 0; i < nseg; i++) {
        Frame sp;
        if (segs[i].type != JT_FIXED) sp = seg_pose(&segs[i], q[j++]);
        else sp = seg_pose(&segs[i], 0.0);
        cur = frame_mul(&cur, &sp);
        out[i] = cur;
    }
}
static Frame chain_fk(const double* q, int D) {
    Frame all[9];
    chain_fk_all(q, D, all);
    return all[D];
}

/* KDL ChainJntToJacSolver::JntToJac: torm/src/torm_ik_solver.cpp:173-175, pirl/src/kinematic_solver.cpp:202-204.
 * Column k is set to the unit twist of joint k expressed in the base at the current tip, and every earlier
 * column has its reference point moved to each new tip (changeRefPoint).  jac is [6][D] row-major. */
static void chain_jac(const double* q, int D, double* jac) {
    int nseg;
    const OrcSeg* segs = chain_segs(D, &nseg);
    Frame T_tmp;
    T_tmp.M = rot_identity();
    T_tmp.p = v3(0, 0, 0);
    Twist cols[8];
    int k = 0, j = 0;
    for (int i = 0; i < nseg; i++) {
        Frame total;
        Twist t_tmp;
        memset(&t_tmp, 0, sizeof t_tmp);
        if (segs[i].type != JT_FIXED) {
            Frame sp = seg_pose(&segs[i], q[j]);
            total = frame_mul(&T_tmp, &sp);
            /* Segment::twist(q, 1) = joint.twist(1).RefPoint(joint.pose(q).M * f_tip.p); f_tip.p = 0 here */
            Vec ax = v3(segs[i].axis[0], segs[i].axis[1], segs[i].axis[2]);
            if (segs[i].type == JT_REV) {
                t_tmp.vel = v3(0, 0, 0);
                t_tmp.rot = rot_apply(&T_tmp.M, ax);
            } else {
                t_tmp.vel = rot_apply(&T_tmp.M, ax);
                t_tmp.rot = v3(0, 0, 0);
            }
            j++;
        } else {
            Frame sp = seg_pose(&segs[i], 0.0);
            total = frame_mul(&T_tmp, &sp);
        }
        /* changeRefPoint(jac, total.p - T_tmp.p): vel += rot x delta */
        Vec delta = vsub(total.p, T_tmp.p);
        for (int c = 0; c < k; c++) cols[c].vel = vadd(cols[c].vel, vcross(cols[c].rot, delta));
        if (segs[i].type != JT_FIXED) cols[k++] = t_tmp;
        T_tmp = total;
    }
    for (int c = 0; c < D; c++) {
        for (int r = 0; r < 3; r++) {
            jac[r * D + c] = cols[c].vel.v[r];
            jac[(3 + r) * D + c] = cols[c].rot.v[r];
        }
    }
}

/* One-sided (Hestenes) Jacobi SVD of A = J^T (D x 6): returns pinv(J) * v with singular values below
 * eps zeroed -- the result KDL::ChainIkSolverVel_pinv::CartToJnt computes through svd_HH
 * (torm/src/torm_ik_solver.cpp:169-171; defaults eps = 1e-5, maxiter = 150).  The pseudo-inverse is
 * unique, so any accurate SVD gives the same answer away from a singular value crossing eps. */
static void pinv_apply(const double* jac /*[6][D]*/, int D, const double* v6, double eps, double* dq) {
    double A[8][6]; /* A = J^T, columns get orthogonalised: A V = U S */
    double V[6][6];
    for (int i = 0; i < D; i++)
        for (int c = 0; c < 6; c++) A[i][c] = jac[c * D + i];
    for (int i = 0; i < 6; i++)
        for (int c = 0; c < 6; c++) V[i][c] = (i == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 5; p++) {
            for (int r = p + 1; r < 6; r++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < D; i++) {
                    alpha += A[i][p] * A[i][p];
                    beta += A[i][r] * A[i][r];
                    gamma += A[i][p] * A[i][r];
                }
                if (gamma == 0.0) continue;
                double lim = fabs(gamma) / sqrt(alpha * beta > 0 ? alpha * beta : 1e-300);
                if (lim > off) off = lim;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
                for (int i = 0; i < D; i++) {
                    double ap = A[i][p], ar = A[i][r];
                    A[i][p] = c * ap - s * ar;
                    A[i][r] = s * ap + c * ar;
                }
                for (int i = 0; i < 6; i++) {
                    double vp = V[i][p], vr = V[i][r];
                    V[i][p] = c * vp - s * vr;
                    V[i][r] = s * vp + c * vr;
                }
            }
        }
        if (off < 1e-15) break;
    }
    /* J^T = Uh S V^T with Uh S = A  =>  J = V S Uh^T, pinv(J) = Uh S^-1 V^T = A S^-2 V^T */
    for (int i = 0; i < D; i++) dq[i] = 0.0;
    for (int c = 0; c < 6; c++) {
        double s2 = 0;
        for (int i = 0; i < D; i++) s2 += A[i][c] * A[i][c];
        double s = sqrt(s2);
        if (fabs(s) < eps) continue; /* truncated SVD */
        double proj = 0;
        for (int r = 0; r < 6; r++) proj += V[r][c] * v6[r];
        for (int i = 0; i < D; i++) dq[i] += A[i][c] * (proj / s2);
    }
}

/* dense inverse by Gauss-Jordan with partial pivoting (stands in for Eigen's .inverse()) */
static int mat_inverse(double* a, int n, double* inv) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int c = 0; c < n; c++) {
        int piv = c;
        double best = fabs(a[c * n + c]);
        for (int r = c + 1; r < n; r++)
            if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); piv = r; }
        if (best == 0.0) return -1;
        if (piv != c) {
            for (int j = 0; j < n; j++) {
                double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t;
                t = inv[c * n + j]; inv[c * n + j] = inv[piv * n + j]; inv[piv * n + j] = t;
            }
        }
        double d = 1.0 / a[c * n + c];
        for (int j = 0; j < n; j++) { a[c * n + j] *= d; inv[c * n + j] *= d; }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            double f = a[r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; j++) { a[r * n + j] -= f * a[c * n + j]; inv[r * n + j] -= f * inv[c * n + j]; }
        }
    }
    return 0;
}

/* ---- public kinematics --------------------------------------------------------------------- */
static int check_D(int D) { return (D == 7 || D == 8) ? 0 : -1; }

int orc_fk(const double* q, int64_t B, int D, double* pos, double* rot) {
    if (check_D(D)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        Frame f = chain_fk(q + b * D, D);
        memcpy(pos + 3 * b, f.p.v, 3 * sizeof(double));
        memcpy(rot + 9 * b, f.M.d, 9 * sizeof(double));
    }
    return 0;
}

/* fk_srv response form: pirl/src/kinematic_solver.cpp:237-261 */
int orc_fk_quat(const double* q, int64_t B, int D, double* out) {
    if (check_D(D)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        Frame f = chain_fk(q + b * D, D);
        memcpy(out + 7 * b, f.p.v, 3 * sizeof(double));
        orc_rot_to_quat_kdl(f.M.d, out + 7 * b + 3);
    }
    return 0;
}

/* allfk_srv response form: pirl/src/kinematic_solver.cpp:206-234 */
int orc_allfk(const double* q, int64_t B, int D, double* out) {
    if (check_D(D)) return -1;
    int ns = D + 1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        Frame all[9];
        chain_fk_all(q + b * D, D, all);
        for (int s = 0; s < ns; s++) {
            double* o = out + (b * ns + s) * 7;
            memcpy(o, all[s].p.v, 3 * sizeof(double));
            orc_rot_to_quat_kdl(all[s].M.d, o + 3);
        }
    }
    return 0;
}

int orc_jacobian(const double* q, int64_t B, int D, double* J) {
    if (check_D(D)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) chain_jac(q + b * D, D, J + b * 6 * D);
    return 0;
}

int orc_fk_jac(const double* q, int64_t B, int D, double* pos, double* rot, double* J) {
    if (check_D(D)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        Frame f = chain_fk(q + b * D, D);
        memcpy(pos + 3 * b, f.p.v, 3 * sizeof(double));
        memcpy(rot + 9 * b, f.M.d, 9 * sizeof(double));
        chain_jac(q + b * D, D, J + b * 6 * D);
    }
    return 0;
}

static Frame target_frame(const double* t7) {
    Frame f;
    f.p = v3(t7[0], t7[1], t7[2]);
    orc_quat_to_rot_kdl(t7 + 3, f.M.d);
    return f;
}

/* diff(FK(q), target): torm/src/torm_optimizer.cpp:1413,1440 */
int orc_pose_diff(const double* q, const double* target, int64_t B, int D, double* twist) {
    if (check_D(D)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        Frame f = chain_fk(q + b * D, D);
        Frame t = target_frame(target + 7 * b);
        Twist tw = frame_diff(&f, &t);
        memcpy(twist + 6 * b, tw.vel.v, 3 * sizeof(double));
        memcpy(twist + 6 * b + 3, tw.rot.v, 3 * sizeof(double));
    }
    return 0;
}

/* vikSolver: torm/src/torm_ik_solver.cpp:169-171 */
int orc_pinv_step(const double* q, const double* twist, int64_t B, int D, double eps, double* dq) {
    if (check_D(D)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        double jac[48];
        chain_jac(q + b * D, D, jac);
        pinv_apply(jac, D, twist + 6 * b, eps, dq + b * D);
    }
    return 0;
}

/* jaco_srv: pirl/src/kinematic_solver.cpp:332-351.  ||(I - J^T (J J^T)^-1 J) err||, dense inverse, no ridge */
static double nullspace_norm(const double* err, const double* q) {
    double J[42], JJt[36], inv[36], N[49];
    chain_jac(q, 7, J);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 7; k++) s += J[i * 7 + k] * J[j * 7 + k];
            JJt[i * 6 + j] = s;
        }
    if (mat_inverse(JJt, 6, inv)) return NAN;
    /* N = I - J^T inv J */
    for (int a = 0; a < 7; a++)
        for (int b = 0; b < 7; b++) {
            double s = 0;
            for (int i = 0; i < 6; i++) {
                double t = 0;
                for (int j = 0; j < 6; j++) t += inv[i * 6 + j] * J[j * 7 + b];
                s += J[i * 7 + a] * t;
            }
            N[a * 7 + b] = ((a == b) ? 1.0 : 0.0) - s;
        }
    double n2 = 0;
    for (int a = 0; a < 7; a++) {
        double s = 0;
        for (int b = 0; b < 7; b++) s += N[a * 7 + b] * err[b];
        n2 += s * s;
    }
    return sqrt(n2);
}
int orc_nullspace_norm(const double* err, const double* q, int64_t B, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) out[b] = nullspace_norm(err + 7 * b, q + 7 * b);
    return 0;
}

static double det_n(double* a, int n) {
    double det = 1.0;
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++)
            if (fabs(a[r * n + c]) > fabs(a[piv * n + c])) piv = r;
        if (a[piv * n + c] == 0.0) return 0.0;
        if (piv != c) {
            for (int j = 0; j < n; j++) { double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t; }
            det = -det;
        }
        det *= a[c * n + c];
        for (int r = c + 1; r < n; r++) {
            double f = a[r * n + c] / a[c * n + c];
            for (int j = c; j < n; j++) a[r * n + j] -= f * a[c * n + j];
        }
    }
    return det;
}
/* checkSingularity: torm/src/torm_optimizer.cpp:748-765, |sqrt(det(J J^T))| */
int orc_manipulability(const double* q, int64_t B, int D, double* out) {
    if (check_D(D)) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        double J[48], JJt[36];
        chain_jac(q + b * D, D, J);
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) {
                double s = 0;
                for (int k = 0; k < D; k++) s += J[i * D + k] * J[j * D + k];
                JJt[i * 6 + j] = s;
            }
        out[b] = fabs(sqrt(det_n(JJt, 6)));
    }
    return 0;
}

/* ================================================================================================
 * Scene: occupancy -> signed distance records (build-defined; SURVEY "four facts" item 4 and A.4)
 * ============================================================================================== */
#define ORC_INACTIVE 1e30f

/* exact squared Euclidean distance transform, separable min-plus passes over integer cell offsets */
static void edt_sq(const uint8_t* seed /*1 = source*/, int nx, int ny, int nz, int32_t* out) {
    const int32_t INF = 1 << 28;
    int64_t n = (int64_t)nx * ny * nz;
    int32_t* a = (int32_t*)malloc(n * sizeof(int32_t));
    int32_t* b = (int32_t*)malloc(n * sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) a[i] = seed[i] ? 0 : INF;
    /* x pass */
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < (int64_t)ny * nz; l++) {
        const int32_t* f = a + l * nx;
        int32_t* g = b + l * nx;
        for (int i = 0; i < nx; i++) {
            int32_t best = INF;
            for (int j = 0; j < nx; j++) {
                int32_t v = f[j] + (i - j) * (i - j);
                if (v < best) best = v;
            }
            g[i] = best;
        }
    }
    /* y pass */
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < (int64_t)nx * nz; l++) {
        int z = (int)(l / nx), x = (int)(l % nx);
        for (int i = 0; i < ny; i++) {
            int32_t best = INF;
            for (int j = 0; j < ny; j++) {
                int32_t v = b[((int64_t)z * ny + j) * nx + x] + (i - j) * (i - j);
                if (v < best) best = v;
            }
            a[((int64_t)z * ny + i) * nx + x] = best;
        }
    }
    /* z pass */
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < (int64_t)nx * ny; l++) {
        for (int i = 0; i < nz; i++) {
            int32_t best = INF;
            for (int j = 0; j < nz; j++) {
                int32_t v = a[(int64_t)j * nx * ny + l] + (i - j) * (i - j);
                if (v < best) best = v;
            }
            out[(int64_t)i * nx * ny + l] = best;
        }
    }
    free(a);
    free(b);
}

/* Signed distance d = res * (sqrt(dsq to nearest occupied) - sqrt(dsq to nearest free)), capped at
 * max_prop (MoveIt PropagationDistanceField semantics); record = (d, central-difference
 * gradient / (2 res)).  Border cells and cells with d >= max_prop are inactive (d = 1e30, g = 0), which is
 * what MoveIt's getDistanceGradient / getCollisionSphereGradients leave untouched. */
int orc_build_field(const uint8_t* occ, int nx, int ny, int nz, double res, double max_prop, float* field) {
    int64_t n = (int64_t)nx * ny * nz;
    int32_t* dso = (int32_t*)malloc(n * sizeof(int32_t));
    int32_t* dsf = (int32_t*)malloc(n * sizeof(int32_t));
    uint8_t* inv = (uint8_t*)calloc(n, 1);
    float* d = (float*)malloc(n * sizeof(float));
    for (int64_t i = 0; i < n; i++) inv[i] = occ[i] ? 0 : 1;
    edt_sq(occ, nx, ny, nz, dso);
    edt_sq(inv, nx, ny, nz, dsf);
    const float resf = (float)res, capf = (float)max_prop;
    for (int64_t i = 0; i < n; i++) {
        float v = resf * (sqrtf((float)dso[i]) - sqrtf((float)(dsf[i] >= (1 << 28) ? 0 : dsf[i])));
        if (dso[i] >= (1 << 28)) v = capf;
        d[i] = v < capf ? v : capf;
    }
    const float i2r = (float)(1.0 / (2.0 * res));
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                int64_t i = ((int64_t)z * ny + y) * nx + x;
                float* r = field + 4 * i;
                int border = (x < 1 || y < 1 || z < 1 || x >= nx - 1 || y >= ny - 1 || z >= nz - 1);
                if (border || !(d[i] < capf)) {
                    r[0] = ORC_INACTIVE; r[1] = r[2] = r[3] = 0.0f;
                } else {
                    r[0] = d[i];
                    r[1] = (d[i + 1] - d[i - 1]) * i2r;
                    r[2] = (d[i + nx] - d[i - nx]) * i2r;
                    r[3] = (d[i + (int64_t)nx * ny] - d[i - (int64_t)nx * ny]) * i2r;
                }
            }
    free(dso); free(dsf); free(inv); free(d);
    return 0;
}

/* nearest-cell lookup; cell k is centred at origin + k * res (MoveIt VoxelGrid::worldToGrid) */
static int scene_lookup(const OrcScene* sc, Vec x, double* d, Vec* g) {
    int cx = (int)floor((x.v[0] - sc->origin[0]) / sc->res + 0.5);
    int cy = (int)floor((x.v[1] - sc->origin[1]) / sc->res + 0.5);
    int cz = (int)floor((x.v[2] - sc->origin[2]) / sc->res + 0.5);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= sc->nx || cy >= sc->ny || cz >= sc->nz) return 0;
    const float* r = sc->field + 4 * (((int64_t)cz * sc->ny + cy) * sc->nx + cx);
    if (!(r[0] < 1e29f)) return 0;
    *d = r[0];
    *g = v3(r[1], r[2], r[3]);
    return 1;
}

static void sphere_positions(const OrcScene* sc, const Frame* links /*[8]*/, const double* base_off, Vec* x) {
    Vec off = v3(base_off[0], base_off[1], base_off[2]);
    for (int s = 0; s < sc->S; s++) {
        const Frame* f = &links[sc->link[s]];
        Vec c = v3(sc->center[3 * s], sc->center[3 * s + 1], sc->center[3 * s + 2]);
        x[s] = vadd(off, vadd(f->p, rot_apply(&f->M, c)));
    }
}

/* B200 replacement for cc_srv (pirl/src/collision_checker.cpp:39-61): any sphere with d - r < 0.
 * An approximation of the FCL mesh test, not a parity target. */
int orc_sphere_collide(const double* q, int64_t B, const OrcScene* sc, const double* base_off, uint8_t* out) {
    if (sc->S > 64) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        Frame links[9];
        Vec x[64];
        chain_fk_all(q + 7 * b, 7, links);
        sphere_positions(sc, links, base_off, x);
        uint8_t hit = 0;
        for (int s = 0; s < sc->S; s++) {
            double d; Vec g;
            if (scene_lookup(sc, x[s], &d, &g) && d - sc->radius[s] < 0.0) hit = 1;
        }
        out[b] = hit;
    }
    return 0;
}

/* ================================================================================================
 * TORM
 * ============================================================================================== */
#define DIFF_RULE_LENGTH 7
/* torm/include/torm/torm_utils.h:49-57 */
static const double DIFF_RULES[3][DIFF_RULE_LENGTH] = {
    {-1.0 / 60, 3.0 / 20, -3.0 / 4, 0.0, 3.0 / 4, -3.0 / 20, 1.0 / 60},
    {1.0 / 90, -3.0 / 20, 3.0 / 2, -49.0 / 18, 3.0 / 2, -3.0 / 20, 1.0 / 90},
    {1.0 / 8, -1.0, 13.0 / 8, 0.0, -13.0 / 8, 1.0, -1.0 / 8}};

/* torm/include/torm/torm_utils.h:79-100 (copied by the reference from geometry/angles) */
static double normalizeAnglePositive(double a) { return fmod(fmod(a, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI); }
static double normalizeAngle(double a) {
    double r = normalizeAnglePositive(a);
    if (r > M_PI) r -= 2.0 * M_PI;
    return r;
}
static double shortestAngularDistance(double start, double end) {
    double res = normalizeAnglePositive(normalizeAnglePositive(end) - normalizeAnglePositive(start));
    if (res > M_PI) res = -(2.0 * M_PI - res);
    return normalizeAngle(res);
}

/* TormCost::getDiffMatrix torm/src/torm_cost.cpp:79-95 */
static void diff_matrix(int size, const double* rule, double* m) {
    memset(m, 0, sizeof(double) * size * size);
    for (int i = 0; i < size; i++)
        for (int j = -DIFF_RULE_LENGTH / 2; j <= DIFF_RULE_LENGTH / 2; j++) {
            int index = i + j;
            if (index < 0 || index >= size) continue;
            m[i * size + index] = rule[j + DIFF_RULE_LENGTH / 2];
        }
}

/* TormCost ctor torm/src/torm_cost.cpp:46-77 + the scaling at torm/src/torm_optimizer.cpp:211-236.
 * Every joint has joint_cost 1.0, so one matrix serves all joints.  Returns the scaled quad_cost_full_
 * in band storage Aband[i][k] = A(i, i + k - 6) and the scaled inverse of the free block. */
int orc_smooth_setup(int T, double dt, const double* w3, double ridge, double* Aband, double* scale_out,
                     double* Ainv) {
    int G = T + 10; /* num_vars_all: 5 padding rows each side (torm/src/torm_trajectory.cpp:66-71) */
    int F = G - 2 * (DIFF_RULE_LENGTH - 1) + 1; /* num_vars_free = T - 1 */
    double* dm = (double*)malloc(sizeof(double) * G * G);
    double* full = (double*)calloc((size_t)G * G, sizeof(double));
    double multiplier = 1.0;
    for (int k = 0; k < 3; k++) {
        multiplier *= dt;
        diff_matrix(G, DIFF_RULES[k], dm);
        double coef = w3[k] * multiplier;
        if (coef == 0.0) continue;
        /* full += coef * dm^T dm ; dm is 7-banded */
        for (int r = 0; r < G; r++) {
            int lo = r - 3 < 0 ? 0 : r - 3, hi = r + 3 >= G ? G - 1 : r + 3;
            for (int a = lo; a <= hi; a++)
                for (int b = lo; b <= hi; b++) full[a * G + b] += coef * (dm[r * G + a] * dm[r * G + b]);
        }
    }
    for (int i = 0; i < G; i++) full[i * G + i] += ridge;
    double* fr = (double*)malloc(sizeof(double) * F * F);
    double* inv = (double*)malloc(sizeof(double) * F * F);
    int o = DIFF_RULE_LENGTH - 1;
    for (int i = 0; i < F; i++)
        for (int j = 0; j < F; j++) fr[i * F + j] = full[(i + o) * G + (j + o)];
    int rc = mat_inverse(fr, F, inv);
    double mx = -DBL_MAX;
    for (int i = 0; i < F * F; i++)
        if (inv[i] > mx) mx = inv[i];
    double max_cost_scale = 1.0;
    if (max_cost_scale < mx) max_cost_scale = mx;
    *scale_out = mx;
    for (int i = 0; i < G; i++)
        for (int k = 0; k < 13; k++) {
            int j = i + k - 6;
            Aband[i * 13 + k] = (j < 0 || j >= G) ? 0.0 : full[i * G + j] * max_cost_scale;
        }
    if (Ainv) {
        double is = 1.0 / max_cost_scale;
        for (int i = 0; i < F * F; i++) Ainv[i] = inv[i] * is;
    }
    free(dm); free(full); free(fr); free(inv);
    return rc;
}

/* getPotential torm/include/torm/torm_optimizer.h:146-168 (diff * clearance) */
static double get_potential(double field_distance, double radius, double clearence) {
    double d = field_distance - radius;
    double potential = 0.0;
    if (d >= clearence) {
        potential = 0.0;
    } else if (d >= 0.0) {
        double diff = (d - clearence);
        double gradient_magnitude = diff * clearence;
        potential = 0.5 * gradient_magnitude * diff;
    } else {
        potential = -d + 0.5 * clearence;
    }
    return potential;
}

static void mat3_inverse(const double* m, double* inv) {
    double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                 m[2] * (m[3] * m[7] - m[4] * m[6]);
    double id = 1.0 / det;
    inv[0] = (m[4] * m[8] - m[5] * m[7]) * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = (m[5] * m[6] - m[3] * m[8]) * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = (m[3] * m[7] - m[4] * m[6]) * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

/* traj_evaluator::getPoseCost + joint costs: torm/src/utils/traj_evaluator.cpp:65-85,104-132,164-205,244-281 */
static void traj_eval_one(const double* traj, const double* targets, const OrcTormParams* p, double* out) {
    int T = p->T, D = p->D;
    double cpose = 0, crot = 0;
    for (int i = 1; i < T; i++) {
        Frame cur = chain_fk(traj + i * D, D);
        Frame tg = target_frame(targets + 7 * i);
        Twist d = frame_diff(&cur, &tg);
        cpose += sqrt(vdot(d.vel, d.vel));
        crot += sqrt(vdot(d.rot, d.rot));
    }
    for (int i = 0; i < T - 1; i++) {
        double qm[8];
        for (int j = 0; j < D; j++) qm[j] = (traj[(i + 1) * D + j] + traj[i * D + j]) / 2;
        Frame cur = chain_fk(qm, D);
        /* setInterTargetPoses :244-257 */
        Frame a = target_frame(targets + 7 * i), b = target_frame(targets + 7 * (i + 1));
        Frame inter;
        inter.p = v3((b.p.v[0] + a.p.v[0]) / 2, (b.p.v[1] + a.p.v[1]) / 2, (b.p.v[2] + a.p.v[2]) / 2);
        Rot at = rot_transpose(&a.M);
        Rot rel = rot_mul(&at, &b.M);
        double axis[3];
        double angle = orc_rot_angle_kdl(rel.d, axis);
        Rot half = rot2(v3(axis[0], axis[1], axis[2]), angle / 2);
        inter.M = rot_mul(&a.M, &half);
        Twist d = frame_diff(&cur, &inter);
        cpose += sqrt(vdot(d.vel, d.vel));
        crot += sqrt(vdot(d.rot, d.rot));
    }
    cpose /= (T - 1) * 2;
    crot /= (T - 1) * 2;

    double cvel = 0, cacc = 0, cjerk = 0;
    double* vel = (double*)malloc(sizeof(double) * (T + 6) * 3);
    double* acc = vel + (T + 6);
    double* jerk = acc + (T + 6);
    for (int j = 0; j < D; j++) {
        double v = 0, a = 0, je = 0;
        if (p->cjoint[j]) {
            /* calcJointCostwithNumericalDifferenciation :164-181 */
            for (int i = 0; i < T - 1; i++) vel[i] = shortestAngularDistance(traj[i * D + j], traj[(i + 1) * D + j]) / p->dt;
            for (int i = 0; i < T - 2; i++) acc[i] = (vel[i + 1] - vel[i]) / p->dt;
            for (int i = 0; i < T - 3; i++) jerk[i] = (acc[i + 1] - acc[i]) / p->dt;
            for (int i = 0; i < T - 1; i++) v += fabs(vel[i]);
            for (int i = 0; i < T - 2; i++) a += fabs(acc[i]);
            for (int i = 0; i < T - 3; i++) je += fabs(jerk[i]);
            v /= (T - 1); a /= (T - 2); je /= (T - 3);
        } else {
            /* calcJointCostwithFiniteDifferenciation :201-205 on the 3-row padded copy (:31-48) */
            int nf = T + 6;
            double sums[3] = {0, 0, 0};
            for (int k = 0; k < 3; k++) {
                double sc = pow(p->dt, k + 1);
                for (int r = 3; r < 3 + T; r++) {
                    double s = 0;
                    for (int t = -3; t <= 3; t++) {
                        int idx = r + t;
                        if (idx < 0 || idx >= nf) continue;
                        int src = idx - 3;
                        if (src < 0) src = 0;
                        if (src > T - 1) src = T - 1;
                        s += (DIFF_RULES[k][t + 3] / sc) * traj[src * D + j];
                    }
                    sums[k] += fabs(s);
                }
            }
            v = sums[0] / T; a = sums[1] / T; je = sums[2] / T;
        }
        cvel += v; cacc += a; cjerk += je;
    }
    free(vel);
    cvel /= D; cacc /= D; cjerk /= D;

    /* calcNofVelViolation :261-281 */
    int nviol = 0;
    for (int i = 0; i < T - 1; i++) {
        int flag = 0;
        for (int j = 0; j < D; j++) {
            double dd = p->cjoint[j] ? fabs(shortestAngularDistance(traj[i * D + j], traj[(i + 1) * D + j]))
                                     : fabs(traj[(i + 1) * D + j] - traj[i * D + j]);
            if (dd / p->dt > p->vel_limit[j]) { flag = 1; break; }
        }
        nviol += flag;
    }
    out[0] = cpose; out[1] = crot; out[2] = cvel; out[3] = cacc; out[4] = cjerk; out[5] = (double)nviol;
}

int orc_traj_eval(const double* traj, const double* targets, const OrcTormParams* p, int64_t N, double* out) {
#pragma omp parallel for schedule(dynamic)
    for (int64_t n = 0; n < N; n++) traj_eval_one(traj + n * p->T * p->D, targets, p, out + 6 * n);
    return 0;
}

static void torm_one(const double* traj, const double* targets, const OrcScene* sc, const OrcTormParams* p,
                     const double* Aband, int stage_mask, int stoch, double* inc_smooth, double* inc_coll,
                     double* inc_pose, double* costs, double* active_out) {
    const int T = p->T, D = p->D, G = T + 10, F = T - 1;
    const int free_start = DIFF_RULE_LENGTH - 1, free_end = G - 1 - (DIFF_RULE_LENGTH - 2);
    /* group trajectory with clamped padding: torm/src/torm_trajectory.cpp:81-96 */
#define GROW(i) (traj + (size_t)((i) - 5 < 0 ? 0 : ((i) - 5 > T - 1 ? T - 1 : (i) - 5)) * D)
    for (int k = 0; k < ORC_NCOST; k++) costs[k] = 0.0;

    /* ---- getEndPoseCost(grad) torm/src/torm_optimizer.cpp:1399-1425, calculateEndPoseIncrements :1475-1482 */
    if (stage_mask & (ORC_STAGE_POSE | ORC_STAGE_POSE_GRAD)) {
        double endPos = 0, endRot = 0;
        for (int i = free_start; i <= free_end; i++) {
            const double* q = GROW(i);
            Frame cur = chain_fk(q, D);
            Frame tg = target_frame(targets + 7 * (i - 5)); /* getFullTrajectoryIndex(i) = i - 5 */
            Twist tw = frame_diff(&cur, &tg);
            if ((stage_mask & ORC_STAGE_POSE_GRAD) && inc_pose) {
                double jac[48], v6[6], dq[8];
                chain_jac(q, D, jac);
                memcpy(v6, tw.vel.v, 24);
                memcpy(v6 + 3, tw.rot.v, 24);
                pinv_apply(jac, D, v6, p->vik_eps, dq);
                for (int j = 0; j < D; j++) inc_pose[(i - 6) * D + j] = p->jac_weight * dq[j];
            }
            endPos += sqrt(vdot(tw.vel, tw.vel));
            endRot += sqrt(vdot(tw.rot, tw.rot));
        }
        costs[8] = endPos;
        costs[9] = endRot;
    }

    /* ---- calculateSmoothnessIncrements :1466-1473 : -(2 A_full x)[free rows] */
    if ((stage_mask & ORC_STAGE_SMOOTH) && inc_smooth) {
        for (int j = 0; j < D; j++)
            for (int i = free_start; i <= free_end; i++) {
                double s = 0;
                for (int k = 0; k < 13; k++) {
                    int c = i + k - 6;
                    if (c < 0 || c >= G) continue;
                    s += Aband[i * 13 + k] * (2.0 * GROW(c)[j]);
                }
                inc_smooth[(i - free_start) * D + j] = -s;
            }
    }

    /* ---- performForwardKinematics :1302-1382 + calculateCollisionIncrements :1484-1553 */
    if ((stage_mask & ORC_STAGE_COLL) && sc) {
        const int S = sc->S;
        Vec* pos = (Vec*)malloc(sizeof(Vec) * G * S);
        Vec* grad = (Vec*)malloc(sizeof(Vec) * G * S);
        double* pot = (double*)malloc(sizeof(double) * G * S);
        Vec* jaxis = (Vec*)malloc(sizeof(Vec) * G * 7);
        Vec* jpos = (Vec*)malloc(sizeof(Vec) * G * 7);
        Vec off = v3(p->base_off[0], p->base_off[1], p->base_off[2]);
        int nseg;
        const OrcSeg* segs = chain_segs(7, &nseg);
        /* every padded row gets FK (deviation from the reference, which leaves rows 0..5 uninitialised:
         * SURVEY appendix B item 5) */
        for (int i = 0; i < G; i++) {
            Frame links[9];
            chain_fk_all(GROW(i), 7, links);
            sphere_positions(sc, links, p->base_off, pos + (size_t)i * S);
            /* computeJointProperties :1125-1162: joint_transform = child link frame in the planning frame */
            for (int j = 0; j < 7; j++) {
                Vec a = rot_apply(&links[j].M, v3(segs[j].axis[0], segs[j].axis[1], segs[j].axis[2]));
                Vec o = vadd(off, links[j].p);
                jaxis[i * 7 + j] = p->affine_axis_quirk ? vadd(a, o) : a; /* axis = joint_transform * axis (:1157) */
                jpos[i * 7 + j] = o;
            }
            for (int s = 0; s < S; s++) {
                double d; Vec g;
                if (scene_lookup(sc, pos[(size_t)i * S + s], &d, &g)) {
                    pot[(size_t)i * S + s] = get_potential(d, sc->radius[s], p->clearance);
                    grad[(size_t)i * S + s] = g;
                } else {
                    pot[(size_t)i * S + s] = 0.0; /* distances stay DBL_MAX, gradients 0 */
                    grad[(size_t)i * S + s] = v3(0, 0, 0);
                }
            }
        }
        /* getCollisionCost(start_idx - 1, end_idx) :1448-1463: mean over the T real rows */
        double ccost = 0;
        for (int i = free_start - 1; i <= free_end; i++) {
            double st = 0;
            for (int s = 0; s < S; s++) st += pot[(size_t)i * S + s];
            ccost += st;
        }
        costs[5] = ccost / (free_end - (free_start - 1) + 1);

        if (inc_coll) {
            for (int k = 0; k < F * D; k++) inc_coll[k] = 0.0;
            double inv_time = 1.0 / p->dt, inv_time_sq = inv_time * inv_time;
            int startPoint = free_start, endPoint = free_end;
            if (stoch >= 0) startPoint = endPoint = free_start + stoch; /* use_stochastic_descent_ :1502-1515 */
            double nactive = 0;
            for (int i = startPoint; i <= endPoint; i++) {
                for (int s = 0; s < S; s++) {
                    double potential = pot[(size_t)i * S + s];
                    if (potential < 0.0001) continue;
                    nactive += 1;
                    /* sphere velocity / acceleration by 7-point stencils :1364-1381 */
                    Vec vel = v3(0, 0, 0), acc = v3(0, 0, 0);
                    for (int k = -DIFF_RULE_LENGTH / 2; k <= DIFF_RULE_LENGTH / 2; k++) {
                        vel = vadd(vel, vscale(pos[(size_t)(i + k) * S + s], inv_time * DIFF_RULES[0][k + 3]));
                        acc = vadd(acc, vscale(pos[(size_t)(i + k) * S + s], inv_time_sq * DIFF_RULES[1][k + 3]));
                    }
                    double vel_mag = vnorm(vel);
                    Vec pg = vscale(grad[(size_t)i * S + s], -1.0);
                    double vel_mag_sq = vel_mag * vel_mag;
                    Vec nv = vscale(vel, 1.0 / vel_mag);
                    /* orthogonal_projector = I - nv nv^T */
                    double P[9];
                    for (int a = 0; a < 3; a++)
                        for (int b = 0; b < 3; b++) P[3 * a + b] = ((a == b) ? 1.0 : 0.0) - nv.v[a] * nv.v[b];
                    Rot Pm;
                    memcpy(Pm.d, P, sizeof P);
                    Vec curv = vscale(rot_apply(&Pm, acc), 1.0 / vel_mag_sq);
                    Vec cg = vscale(vsub(rot_apply(&Pm, pg), vscale(curv, potential)), vel_mag);
                    /* getJacobian :1164-1187: 3 x D positional Jacobian, ancestor-masked */
                    double Jm[21];
                    int nanc = sc->link[s] + 1 < 7 ? sc->link[s] + 1 : 7;
                    for (int j = 0; j < 7; j++) {
                        Vec col = v3(0, 0, 0);
                        if (j < nanc) col = vcross(jaxis[i * 7 + j], vsub(pos[(size_t)i * S + s], jpos[i * 7 + j]));
                        Jm[j] = col.v[0]; Jm[7 + j] = col.v[1]; Jm[14 + j] = col.v[2];
                    }
                    double* row = inc_coll + (size_t)(i - free_start) * D;
                    if (p->use_pinv) {
                        /* calculatePseudoInverse :767-772 : J^T (J J^T + ridge I)^-1 */
                        double JJt[9], JJi[9];
                        for (int a = 0; a < 3; a++)
                            for (int b = 0; b < 3; b++) {
                                double sum = 0;
                                for (int j = 0; j < 7; j++) sum += Jm[a * 7 + j] * Jm[b * 7 + j];
                                JJt[3 * a + b] = sum + ((a == b) ? p->pinv_ridge : 0.0);
                            }
                        mat3_inverse(JJt, JJi);
                        for (int j = 0; j < 7; j++) {
                            double sum = 0;
                            for (int b = 0; b < 3; b++) {
                                double pj = 0;
                                for (int a = 0; a < 3; a++) pj += Jm[a * 7 + j] * JJi[3 * a + b];
                                sum += pj * cg.v[b];
                            }
                            row[j] -= sum;
                        }
                    } else {
                        for (int j = 0; j < 7; j++)
                            row[j] -= Jm[j] * cg.v[0] + Jm[7 + j] * cg.v[1] + Jm[14 + j] * cg.v[2];
                    }
                }
            }
            if (active_out) *active_out = nactive / (endPoint - startPoint + 1);
        }
        free(pos); free(grad); free(pot); free(jaxis); free(jpos);
    }

    if (stage_mask & ORC_STAGE_EVAL) {
        double ev[6];
        traj_eval_one(traj, targets, p, ev);
        costs[0] = ev[0]; costs[1] = ev[1]; costs[2] = ev[2]; costs[3] = ev[3]; costs[4] = ev[4];
        costs[6] = ev[0] + 0.17 * ev[1]; /* getWeightedPoseCost :59-63 */
        costs[7] = ev[5];
    }
#undef GROW
}

int orc_torm_cost_grad(const double* traj, const double* targets, const OrcScene* sc, const OrcTormParams* p,
                       const double* Aband, int stage_mask, const int32_t* stochastic_idx, int64_t N,
                       double* inc_smooth, double* inc_coll, double* inc_pose, double* costs,
                       double* active_spheres) {
    if (p->D != 7) return -1;
    const size_t FD = (size_t)(p->T - 1) * p->D;
#pragma omp parallel for schedule(dynamic)
    for (int64_t n = 0; n < N; n++) {
        int st = stochastic_idx ? stochastic_idx[n] : -1;
        torm_one(traj + n * p->T * p->D, targets, sc, p, Aband, stage_mask, st,
                 inc_smooth ? inc_smooth + n * FD : 0, inc_coll ? inc_coll + n * FD : 0,
                 inc_pose ? inc_pose + n * FD : 0, costs + n * ORC_NCOST,
                 active_spheres ? active_spheres + n : 0);
    }
    return 0;
}

/* ================================================================================================
 * PIRL path-wise IK env step: pirl/script/envs/PirlEnv.py:184-349,495-542
 * ============================================================================================== */
/* numpy-quaternion as_rotation_matrix (normalises by |q|^2); input x y z w */
static void npq_as_rot(const double* q, double* m) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double n = w * w + x * x + y * y + z * z;
    m[0] = 1.0 - 2 * (y * y + z * z) / n; m[1] = 2 * (x * y - z * w) / n; m[2] = 2 * (x * z + y * w) / n;
    m[3] = 2 * (x * y + z * w) / n; m[4] = 1.0 - 2 * (x * x + z * z) / n; m[5] = 2 * (y * z - x * w) / n;
    m[6] = 2 * (x * z - y * w) / n; m[7] = 2 * (y * z + x * w) / n; m[8] = 1.0 - 2 * (x * x + y * y) / n;
}
/* calc_pos_and_rot :322-333 */
static int obs_pos_rot(const double* src7, double* o) {
    double m[9];
    o[0] = src7[0]; o[1] = src7[1]; o[2] = src7[2];
    npq_as_rot(src7 + 3, m);
    o[3] = m[0]; o[4] = m[3]; o[5] = m[6]; o[6] = m[1]; o[7] = m[4]; o[8] = m[7];
    return 9;
}
/* calc_pos_and_rot_error :335-349 : R_ct = tM inv(eM) */
static int obs_pos_rot_err(const double* tgt7, const double* src7, double* o) {
    double tM[9], eM[9], ei[9], R[9];
    o[0] = tgt7[0] - src7[0]; o[1] = tgt7[1] - src7[1]; o[2] = tgt7[2] - src7[2];
    npq_as_rot(tgt7 + 3, tM);
    npq_as_rot(src7 + 3, eM);
    mat3_inverse(eM, ei);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[3 * i + j] = tM[3 * i] * ei[j] + tM[3 * i + 1] * ei[3 + j] + tM[3 * i + 2] * ei[6 + j];
    o[3] = R[0]; o[4] = R[3]; o[5] = R[6]; o[6] = R[1]; o[7] = R[4]; o[8] = R[7];
    return 9;
}
int orc_pirl_obs_dim(const OrcPirlCfg* c) {
    int d = 0;
    if (c->obs_type == 0) d = 7;
    else if (c->obs_type == 1) d = 7 + 9 * 8;
    else if (c->obs_type == 2) d = 7 + 9;
    else d = 9 * 8;
    return d + 9 * c->n_future + c->z_dim;
}
/* _get_obs :292-320 */
static void pirl_obs_one(const OrcPirlCfg* c, const double* q, const double* pose56, int timestep, int max_ts,
                         const double* targets, const double* latent, double* obs) {
    int n = 0;
    if (c->obs_type == 0 || c->obs_type == 1 || c->obs_type == 2)
        for (int j = 0; j < 7; j++) obs[n++] = q[j];
    if (c->obs_type == 1 || c->obs_type == 3)
        for (int s = 0; s < 8; s++) n += obs_pos_rot(pose56 + 7 * s, obs + n);
    if (c->obs_type == 2) n += obs_pos_rot(pose56 + 49, obs + n);
    for (int i = 0; i < c->n_future; i++) {
        int t = timestep + (i + 1);
        if (t > max_ts) t = max_ts;
        n += obs_pos_rot_err(targets + 7 * t, pose56 + 49, obs + n);
    }
    for (int k = 0; k < c->z_dim; k++) obs[n++] = latent[k];
}
static double calc_reward(double err, const double* p) { return p[0] * exp(p[1] * err) - p[2] * ((p[3] * err) * (p[3] * err)); }

static void allfk56(const double* q, double* pose56) {
    Frame all[9];
    chain_fk_all(q, 7, all);
    for (int s = 0; s < 8; s++) {
        memcpy(pose56 + 7 * s, all[s].p.v, 24);
        orc_rot_to_quat_kdl(all[s].M.d, pose56 + 7 * s + 3);
    }
}

int orc_pirl_obs(const OrcPirlCfg* c, int64_t B, const double* q, const int32_t* timestep,
                 const int32_t* max_timestep, const double* targets, int Tmax, const double* latent, double* obs) {
    int O = orc_pirl_obs_dim(c);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        double pose[56];
        allfk56(q + 7 * b, pose);
        pirl_obs_one(c, q + 7 * b, pose, timestep[b], max_timestep[b], targets + (size_t)b * Tmax * 7,
                     latent ? latent + (size_t)b * c->z_dim : 0, obs + (size_t)b * O);
    }
    return 0;
}

int orc_pirl_env_step(const OrcPirlCfg* c, int64_t B, double* q, const float* action, int32_t* timestep,
                      const int32_t* max_timestep, const double* targets, const double* demos, int Tmax,
                      const double* latent, const OrcScene* sc, const double* base_off, double* obs,
                      double* reward, uint8_t* done, uint8_t* flags, int32_t* counters, double* errs) {
    int O = orc_pirl_obs_dim(c);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        double* cur = q + 7 * b;
        const double* tg = targets + (size_t)b * Tmax * 7;
        const double* dm = demos ? demos + (size_t)b * Tmax * 7 : 0;
        /* step :209-215 */
        int ts = timestep[b] + 1;
        int collision_flag = 0, joint_limit_flag = 0;
        /* _set_action :184-207 */
        double o2[7], cl[7];
        for (int j = 0; j < 7; j++) o2[j] = cur[j] + (double)action[7 * b + j];
        for (int j = 0; j < 7; j++)
            if (c->cjoint[j]) {
                if (o2[j] > M_PI) o2[j] -= 2 * M_PI;
                else if (o2[j] < -M_PI) o2[j] += 2 * M_PI;
            }
        int no_out = 1;
        for (int j = 0; j < 7; j++) {
            cl[j] = o2[j] < c->ll[j] ? c->ll[j] : (o2[j] > c->ul[j] ? c->ul[j] : o2[j]);
            /* np.allclose: |a - b| <= atol + rtol |b| */
            if (!(fabs(o2[j] - cl[j]) <= 1e-8 + 1e-5 * fabs(cl[j]))) no_out = 0;
        }
        if (c->train_mode) {
            if (no_out) {
                uint8_t hit = 0;
                if (sc) orc_sphere_collide(cl, 1, sc, base_off, &hit);
                if (hit) collision_flag = 1;
                else memcpy(cur, cl, sizeof cl);
            } else {
                joint_limit_flag = 1;
            }
        } else {
            memcpy(cur, cl, sizeof cl);
        }
        /* allfk :216 */
        double pose[56];
        allfk56(cur, pose);
        pirl_obs_one(c, cur, pose, ts, max_timestep[b], tg, latent ? latent + (size_t)b * c->z_dim : 0,
                     obs + (size_t)b * O);

        /* _compute_task_reward :495-513 */
        const double* tp = tg + 7 * ts;
        const double* ee = pose + 49;
        double dx = tp[0] - ee[0], dy = tp[1] - ee[1], dz = tp[2] - ee[2];
        double d_err = sqrt(dx * dx + dy * dy + dz * dz);
        /* diff_q = tq * conj(pq); only w is used.  tq = (w,x,y,z) = (tp[6],tp[3],tp[4],tp[5]) */
        double tw = tp[6], tx = tp[3], ty = tp[4], tz = tp[5];
        double pw = ee[6], px = -1 * ee[3], py = -1 * ee[4], pz = -1 * ee[5];
        double dw = tw * pw - tx * px - ty * py - tz * pz;
        double r_err = (fabs(dw) > 1.0) ? 2 * acos(1.0) : 2 * acos(fabs(dw));
        double dist_rwd = calc_reward(d_err, c->dist_p);
        double rot_rwd = calc_reward(r_err, c->rot_p);

        /* _compute_imitate_reward :515-526 and _compute_null_space_imitate_reward :528-539 */
        double imitate_rwd = 0, imitate_err = 0, null_rwd = 0, null_obj = 0;
        if (dm) {
            double ce[7];
            for (int j = 0; j < 7; j++) {
                ce[j] = dm[7 * ts + j] - cur[j];
                if (c->cjoint[j]) {
                    if (ce[j] > M_PI) ce[j] -= 2 * M_PI;
                    else if (ce[j] < -M_PI) ce[j] += 2 * M_PI;
                }
            }
            if (!c->refine_flag) {
                double l1 = 0;
                for (int j = 0; j < 7; j++) l1 += fabs(ce[j]);
                imitate_err = l1 / 7;
                imitate_rwd = calc_reward(imitate_err, c->imit_p);
            }
            null_obj = nullspace_norm(ce, cur);
            null_rwd = calc_reward(null_obj, c->null_p);
        }

        int is_time_out = (ts == max_timestep[b]);
        int is_col = collision_flag, is_jlimit = joint_limit_flag;
        int is_early_stop = d_err > c->thr_early_stop;
        int is_dist_suc = d_err < c->thr_dist;
        int is_rot_suc = r_err < c->thr_rot;
        int is_imitate_suc = (!c->refine_flag) ? (imitate_err < c->thr_imitate) : 0;

        double rw = -c->w_col * is_col - c->w_jlimit * is_jlimit - c->w_early_stop * is_early_stop;
        if (c->task_reward_flag)
            rw += (dist_rwd + c->dist_suc_r * is_dist_suc + (d_err < c->zero_task_dist ? rot_rwd : 0) +
                   (d_err < c->zero_task_dist ? c->rot_suc_r * is_rot_suc : 0) + c->task_suc_r * (is_dist_suc * is_rot_suc));
        if (c->imitate_reward_flag && !c->refine_flag) rw += (+imitate_rwd + c->imit_suc_r * is_imitate_suc);
        if (c->null_reward_flag && !c->refine_flag) rw += (+null_rwd);

        reward[b] = rw;
        done[b] = (uint8_t)(is_col || is_early_stop || is_time_out);
        uint8_t* f = flags + 8 * b;
        f[0] = is_col; f[1] = is_jlimit; f[2] = is_early_stop; f[3] = is_time_out;
        f[4] = is_dist_suc; f[5] = is_rot_suc; f[6] = is_imitate_suc; f[7] = (uint8_t)(is_dist_suc * is_rot_suc);
        if (counters) {
            counters[4 * b + 0] += is_dist_suc * is_rot_suc;
            counters[4 * b + 1] += is_dist_suc;
            counters[4 * b + 2] += is_rot_suc;
            if (!c->refine_flag) counters[4 * b + 3] += is_imitate_suc;
        }
        if (errs) { errs[4 * b] = d_err; errs[4 * b + 1] = r_err; errs[4 * b + 2] = imitate_err; errs[4 * b + 3] = null_obj; }
        timestep[b] = ts;
    }
    return 0;
}


/* ================================================================================================
 * TORM update step: calculateFeasibleIncrements torm/src/torm_optimizer.cpp:1555-1609,
 * calculateEndPoseIncrements :1475-1482 (already weighted by the caller), addIncrementsToTrajectory
 * :1611-1639, handleJointLimits :1189-1268.  updateGoalConfiguration / updateFullTrajectory are implicit
 * in the [N,T,D] representation (padded rows are clamped copies).  Ainv is the scaled inverse of the free
 * block (getQuadraticCostInverse).  stage 0 = feasible (smoothness + collision), 1 = end-pose.
 * ============================================================================================== */
static void torm_update_one(double* traj, int T, int D, const double* Ainv, int stage, int mode,
                            const double* inc_smooth, const double* inc_coll, const double* inc_pose, double alpha,
                            double lr, double sw, double ow, double limit, const int* cjoint, const double* jmin,
                            const double* jmax) {
    const int F = T - 1;
    double* fin = (double*)malloc(sizeof(double) * F * D);
    double* tmp = (double*)malloc(sizeof(double) * F);
    if (stage == 0) {
        double* sm = (double*)malloc(sizeof(double) * F * D);
        double* co = (double*)malloc(sizeof(double) * F * D);
        memcpy(sm, inc_smooth, sizeof(double) * F * D);
        memcpy(co, inc_coll, sizeof(double) * F * D);
        if (mode != 0) {
            double mx = -DBL_MAX, mn = DBL_MAX;
            for (int k = 0; k < F * D; k++) { if (sm[k] > mx) mx = sm[k]; if (sm[k] < mn) mn = sm[k]; }
            double scale = fmax(fabs(mx), fabs(mn));
            if (scale > 0) for (int k = 0; k < F * D; k++) sm[k] /= scale;
            mx = -DBL_MAX; mn = DBL_MAX;
            for (int k = 0; k < F * D; k++) { if (co[k] > mx) mx = co[k]; if (co[k] < mn) mn = co[k]; }
            scale = fmax(fabs(mx), fabs(mn));
            if (scale > 0) for (int k = 0; k < F * D; k++) co[k] /= scale;
        }
        for (int j = 0; j < D; j++) {
            for (int i = 0; i < F; i++)
                tmp[i] = (mode != 0) ? (alpha * co[i * D + j] + (1 - alpha) * sm[i * D + j])
                                     : (sw * sm[i * D + j] + ow * co[i * D + j]);
            for (int i = 0; i < F; i++) {
                double acc = 0;
                for (int k = 0; k < F; k++) acc += Ainv[(size_t)i * F + k] * tmp[k];
                fin[i * D + j] = lr * acc;
            }
        }
        free(sm); free(co);
    } else {
        memcpy(fin, inc_pose, sizeof(double) * F * D);
    }
    /* addIncrementsToTrajectory */
    if (mode != 0) {
        double mx = -DBL_MAX, mn = DBL_MAX;
        for (int k = 0; k < F * D; k++) { if (fin[k] > mx) mx = fin[k]; if (fin[k] < mn) mn = fin[k]; }
        double scale = fmax(fabs(mx), fabs(mn));
        if (scale > limit && scale > 0) for (int k = 0; k < F * D; k++) fin[k] = fin[k] / scale * limit;
        for (int i = 0; i < F; i++) for (int j = 0; j < D; j++) traj[(i + 1) * D + j] += fin[i * D + j];
    } else {
        for (int j = 0; j < D; j++) {
            double scale = 1.0, mx = -DBL_MAX, mn = DBL_MAX;
            for (int i = 0; i < F; i++) { double v = fin[i * D + j]; if (v > mx) mx = v; if (v < mn) mn = v; }
            double max_scale = limit / fabs(mx), min_scale = limit / fabs(mn);
            if (max_scale < scale) scale = max_scale;
            if (min_scale < scale) scale = min_scale;
            for (int i = 0; i < F; i++) traj[(i + 1) * D + j] += scale * fin[i * D + j];
        }
    }
    /* handleJointLimits */
    for (int j = 0; j < D; j++) {
        if (cjoint[j]) continue;
        int count = 0, violation;
        do {
            double max_abs_violation = 1e-6, max_violation = 0.0;
            int max_violation_index = 0;
            violation = 0;
            for (int i = 0; i < F; i++) {
                double amount = 0.0, absolute_amount = 0.0, v = traj[(i + 1) * D + j];
                if (v > jmax[j]) { amount = jmax[j] - v; absolute_amount = fabs(amount); }
                else if (v < jmin[j]) { amount = jmin[j] - v; absolute_amount = fabs(amount); }
                if (absolute_amount > max_abs_violation) {
                    max_abs_violation = absolute_amount; max_violation = amount; max_violation_index = i; violation = 1;
                }
            }
            if (violation) {
                int k = max_violation_index;
                double multiplier = max_violation / Ainv[(size_t)k * F + k];
                for (int i = 0; i < F; i++) traj[(i + 1) * D + j] += multiplier * Ainv[(size_t)i * F + k];
            }
            if (++count > 10) break;
        } while (violation);
    }
    free(fin); free(tmp);
}

int orc_torm_update(double* traj, int64_t N, int T, int D, const double* Ainv, int stage, int mode,
                    const double* inc_smooth, const double* inc_coll, const double* inc_pose, const double* alpha,
                    const double* lr, double sw, double ow, double limit, const int* cjoint, const double* jmin,
                    const double* jmax) {
    const size_t FD = (size_t)(T - 1) * D;
#pragma omp parallel for schedule(dynamic)
    for (int64_t n = 0; n < N; n++)
        torm_update_one(traj + n * T * D, T, D, Ainv, stage, mode, inc_smooth ? inc_smooth + n * FD : 0,
                        inc_coll ? inc_coll + n * FD : 0, inc_pose ? inc_pose + n * FD : 0, alpha ? alpha[n] : 0.5,
                        lr[n], sw, ow, limit, cjoint, jmin, jmax);
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
