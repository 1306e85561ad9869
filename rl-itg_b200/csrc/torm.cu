// Kernel 3: TORM path-following objective and gradient over N candidate trajectories x T waypoints
// (SURVEY.md 8a rows A4-A8).  One CTA per trajectory, one thread per waypoint, in two kernels with separate
// register budgets: 3a end-pose error + pinv step + trajectory evaluator, 3b sphere potentials + CHOMP increments
// + smoothness increments (both are launched by one rlitg_torm_cost_grad call, on the caller's stream).
//
// What one launch replaces in the reference (all serial, one trajectory at a time, FP64, KDL + MoveIt):
//   getEndPoseCost(grad)              torm/src/torm_optimizer.cpp:1399-1425   FK + KDL::diff + pinv velocity IK
//   calculateEndPoseIncrements        :1475-1482
//   performForwardKinematics          :1302-1382   sphere positions, distance-field queries, 7-point stencils
//   calculateCollisionIncrements      :1484-1553   CHOMP gradient per sphere through J^T (J J^T + ridge I)^-1
//   getCollisionCost                  :1448-1463
//   calculateSmoothnessIncrements     :1466-1473   -2 A_full x (TormCost, torm/src/torm_cost.cpp:46-77)
//   traj_evaluator::getCost           torm/src/utils/traj_evaluator.cpp:65-85,104-132,164-205,261-281
//
// Data movement per trajectory: the [T,7] rows are read once into shared memory (window with 6 halo rows for
// the 13-tap smoothness band and the 7-point stencils); sphere positions [T,S,3] live only in shared memory;
// the distance volume is gathered as one 16-byte record per sphere from L2; three [T-1,7] increment tiles and
// ten cost scalars are written.  Long paths (T > 256) are walked in 256-row chunks with recomputed halo rows.
// FP32-FMA bound on paper, L2-gather latency bound in practice.
#include <algorithm>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "fetch_chain.cuh"
#include "linalg.cuh"
#include "rlitg_internal.h"
#include "scene.cuh"

namespace rlitg {

constexpr int kTormMaxThreads = 256;
enum { R_WP_POS = 0, R_WP_ROT, R_MID_POS, R_MID_ROT, R_VEL_C, R_ACC_C, R_JERK_C, R_VEL_N, R_ACC_N, R_JERK_N, R_NVIOL_P };

// torm/include/torm/torm_utils.h:49-57
template <typename T> RLITG_DEV T diff_rule(int d, int k) {
    constexpr double R[3][7] = {{-1.0 / 60, 3.0 / 20, -3.0 / 4, 0.0, 3.0 / 4, -3.0 / 20, 1.0 / 60},
                                {1.0 / 90, -3.0 / 20, 3.0 / 2, -49.0 / 18, 3.0 / 2, -3.0 / 20, 1.0 / 90},
                                {1.0 / 8, -1.0, 13.0 / 8, 0.0, -13.0 / 8, 1.0, -1.0 / 8}};
    return T(R[d][k]);
}

template <typename T>
struct TormP {
    int Tn;
    T dt, inv_dt, inv_dt2, rot_scale, jac_w, clearance;
    double ridge, vik_eps;
    int use_pinv, quirk;
    T off[3];
    int cj[7];
    T vlim[7];
    T rule_v[7], rule_a[7], rule_j[7];  // DIFF_RULES / dt^(k+1) for the evaluator's finite differences
    T rv[7], ra[7];                     // the same velocity / acceleration rules for the sphere stencils (kernel 3b)
    // fixed-point sphere centres of the FP32 build (torm_coll.cuh): integer units of res / 2^fix_shift from the grid origin;
    // rvs / ras = rv / ra x the unit; a centre in metres = fix_org + units * fix_unit
    T rvs[7], ras[7], fix_unit, fix_org[3];
    double fix_scale;                   // units per metre
    int fix_shift;
    T sqrt_ridge, ridge_t;
    double off64[3];
};

RLITG_DEV float fmod_t(float a, float b) { return fmodf(a, b); }
RLITG_DEV double fmod_t(double a, double b) { return fmod(a, b); }

// torm/include/torm/torm_utils.h:79-100
template <typename T> RLITG_DEV T normalize_angle_positive(T a) {
    const T two_pi = T(2.0 * 3.14159265358979323846);
    return fmod_t(fmod_t(a, two_pi) + two_pi, two_pi);
}
template <typename T> RLITG_DEV T shortest_angular_distance(T start, T end) {
    const T pi = T(3.14159265358979323846), two_pi = T(2.0 * 3.14159265358979323846);
    if constexpr (sizeof(T) == 4) {
        // same value as the fmod chain below (the difference wrapped into [-pi, pi]) without five fmodf calls;
        // differs only in which of +-pi is returned for an exact half turn
        const float d = end - start;
        return fmaf(-rintf(d * 0.15915494309189535f), two_pi, d);
    }
    T res = normalize_angle_positive(normalize_angle_positive(end) - normalize_angle_positive(start));
    if (res > pi) res = -(two_pi - res);
    T a = normalize_angle_positive(res);
    if (a > pi) a -= two_pi;
    return a;
}

// getPotential torm/include/torm/torm_optimizer.h:146-168 (diff * clearance)
// The (CH + 12) x 7 trajectory window of a chunk (rows c0 - 6 .. c0 + CH + 5, clamped into the trajectory) into shared
// memory.  All of a thread's loads are issued before its first store: the plain copy loop went load / store / load in
// two or more serialised global round trips (12 % of the stall samples of the end-pose kernel, round 2).
template <typename T>
RLITG_DEV void stage_traj_window(T* __restrict__ s_traj, const T* __restrict__ tr, int c0, int Tn, int CH, int tid) {
    const int total = (CH + 12) * 7;
    for (int base = tid; base < total; base += 8 * CH) {
        T v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int idx = base + k * CH;
            if (idx < total) {
                const int r = idx / 7, j = idx - r * 7;
                v[k] = tr[min(max(c0 - 6 + r, 0), Tn - 1) * 7 + j];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int idx = base + k * CH;
            if (idx < total) s_traj[idx] = v[k];
        }
    }
}

template <typename T> RLITG_DEV T get_potential(T field_distance, T radius, T clearance) {
    const T d = field_distance - radius;
    if (d >= clearance) return T(0);
    if (d >= T(0)) {
        const T diff = d - clearance;
        const T gm = diff * clearance;
        return T(0.5) * gm * diff;
    }
    return -d + T(0.5) * clearance;
}

// rare path of kernel 3a's FP32 pseudo-inverse step (a singular value near KDL's truncation threshold): everything is
// rebuilt from the joint row so that the common path keeps neither the Jacobian nor the chain walk live across its QR
template <typename T>
__device__ __noinline__ void pinv_svd_from_row(const T* qrow, const T* tw, double eps, T* dq) {
    T q[7], J[42];
#pragma unroll
    for (int j = 0; j < 7; j++) q[j] = qrow[j];
    FkJac<T, 7> fj;
    walk_chain<T, 7>(q, fj);
    fill_jac<T, 7>(fj, J);
    pinv_svd_apply<T, 7>(J, tw, eps, dq);
}

template <typename T>
struct EeOnly {  // midpoint FK needs the end-effector frame only
    FrameC<T> ee;
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        if constexpr (K == 7) ee = f;
    }
};

template <typename T>
RLITG_DEV void target_frame(const T* __restrict__ t7, V3<T>& p, V3<T>& c0, V3<T>& c1, V3<T>& c2) {
    p = {t7[0], t7[1], t7[2]};
    quat_to_rot_kdl(t7[3], t7[4], t7[5], t7[6], c0, c1, c2);
}

// setInterTargetPoses torm/src/utils/traj_evaluator.cpp:244-257: mean position, R_a * Rot2(axis, angle / 2)
template <typename T>
RLITG_DEV void inter_target(const T* __restrict__ a7, const T* __restrict__ b7, V3<T>& p, V3<T>& c0, V3<T>& c1, V3<T>& c2) {
    V3<T> pa, a0, a1, a2, pb, b0, b1, b2;
    target_frame(a7, pa, a0, a1, a2);
    target_frame(b7, pb, b0, b1, b2);
    p = {(pb.x + pa.x) / T(2), (pb.y + pa.y) / T(2), (pb.z + pa.z) / T(2)};
    // rel = A^T B ; axis-angle by the same GetRotAngle branches as rot_diff_kdl
    const T m00 = dot(a0, b0), m01 = dot(a0, b1), m02 = dot(a0, b2);
    const T m10 = dot(a1, b0), m11 = dot(a1, b1), m12 = dot(a1, b2);
    const T m20 = dot(a2, b0), m21 = dot(a2, b1), m22 = dot(a2, b2);
    T x = m21 - m12, y = m02 - m20, z = m10 - m01, angle;
    const T eps = T(1e-6), eps2 = T(1e-5);
    if (abs_t(x) < eps && abs_t(y) < eps && abs_t(z) < eps) {
        if (abs_t(m01 + m10) < eps2 && abs_t(m02 + m20) < eps2 && abs_t(m12 + m21) < eps2 &&
            abs_t(m00 + m11 + m22 - T(3)) < eps2) {
            x = T(0); y = T(0); z = T(1); angle = T(0);
        } else {
            const T xx = (m00 + T(1)) / T(2), yy = (m11 + T(1)) / T(2), zz = (m22 + T(1)) / T(2);
            const T xy = (m01 + m10) / T(4), xz = (m02 + m20) / T(4), yz = (m12 + m21) / T(4);
            if (xx > yy && xx > zz) { x = sqrt_t(xx); y = xy / x; z = xz / x; }
            else if (yy > zz) { y = sqrt_t(yy); x = xy / y; z = yz / y; }
            else { z = sqrt_t(zz); x = xz / z; y = yz / z; }
            angle = T(3.14159265358979323846);
        }
    } else {
        const T f = (m00 + m11 + m22 - T(1)) / T(2);
        const T r = sqrt_t(x * x + y * y + z * z);
        angle = atan2_t(r / T(2), f);
        x /= r; y /= r; z /= r;
    }
    // KDL Rotation::Rot2(axis, angle / 2)
    T st, ct;
    sincos_joint(angle / T(2), st, ct);
    const T vt = T(1) - ct;
    const T h00 = ct + vt * x * x, h01 = -z * st + vt * x * y, h02 = y * st + vt * x * z;
    const T h10 = z * st + vt * x * y, h11 = ct + vt * y * y, h12 = -x * st + vt * y * z;
    const T h20 = -y * st + vt * x * z, h21 = x * st + vt * y * z, h22 = ct + vt * z * z;
    // C = A * H : column j of C = a0 h0j + a1 h1j + a2 h2j
    c0 = {a0.x * h00 + a1.x * h10 + a2.x * h20, a0.y * h00 + a1.y * h10 + a2.y * h20, a0.z * h00 + a1.z * h10 + a2.z * h20};
    c1 = {a0.x * h01 + a1.x * h11 + a2.x * h21, a0.y * h01 + a1.y * h11 + a2.y * h21, a0.z * h01 + a1.z * h11 + a2.z * h21};
    c2 = {a0.x * h02 + a1.x * h12 + a2.x * h22, a0.y * h02 + a1.y * h12 + a2.y * h22, a0.z * h02 + a1.z * h12 + a2.z * h22};
}

// Target frames precomputed once per path (rlitg_torm_target_frames): frames [2 Tn - 1][12] = (p, c0, c1, c2); rows
// 0 .. Tn-1 the target poses, rows Tn .. 2 Tn - 2 the evaluator's inter-target poses i + 1/2.  They depend on the path
// only, so the N trajectories of a batch share them instead of every waypoint thread rebuilding two quaternion
// frames, an axis-angle and a Rot2 per call (~300 instructions per waypoint).
template <typename T, bool kFrames>
RLITG_DEV void waypoint_target(const T* __restrict__ tg, int i, V3<T>& p, V3<T>& c0, V3<T>& c1, V3<T>& c2) {
    if constexpr (kFrames) {
        const T* f = tg + (size_t)i * 12;
        p = {f[0], f[1], f[2]}; c0 = {f[3], f[4], f[5]}; c1 = {f[6], f[7], f[8]}; c2 = {f[9], f[10], f[11]};
    } else {
        target_frame(tg + (size_t)i * 7, p, c0, c1, c2);
    }
}
template <typename T, bool kFrames>
RLITG_DEV void midpoint_target(const T* __restrict__ tg, int Tn, int i, V3<T>& p, V3<T>& c0, V3<T>& c1, V3<T>& c2) {
    if constexpr (kFrames) waypoint_target<T, true>(tg, Tn + i, p, c0, c1, c2);
    else inter_target(tg + (size_t)i * 7, tg + (size_t)(i + 1) * 7, p, c0, c1, c2);
}
template <typename T>
__global__ void torm_target_frames_kernel(const T* __restrict__ targets, int Tn, T* __restrict__ frames) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * Tn - 1) return;
    V3<T> p, c0, c1, c2;
    if (i < Tn) target_frame(targets + (size_t)i * 7, p, c0, c1, c2);
    else inter_target(targets + (size_t)(i - Tn) * 7, targets + (size_t)(i - Tn + 1) * 7, p, c0, c1, c2);
    T* f = frames + (size_t)i * 12;
    f[0] = p.x; f[1] = p.y; f[2] = p.z; f[3] = c0.x; f[4] = c0.y; f[5] = c0.z;
    f[6] = c1.x; f[7] = c1.y; f[8] = c1.z; f[9] = c2.x; f[10] = c2.y; f[11] = c2.z;
}

// block-wide sum of NR per-thread scalars -> tot[] valid on thread 0 (warp shuffles, then one smem round)
template <typename T, int NR>
RLITG_DEV void block_reduce(T (&red)[NR], T* s_red, T (&tot)[NR]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        T x = red[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        red[k] = x;
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NR; k++) s_red[warp * NR + k] = red[k];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NR; k++) {
            T x = T(0);
            for (int w = 0; w < nwarps; w++) x += s_red[w * NR + k];
            tot[k] = x;
        }
    }
}

#define TROW(row) (s_traj + (min(max((row), 0), Tn - 1) - c0 + 6) * 7)

// ---- kernel 3a: end-pose error / pinv step (A5) + trajectory evaluator (A8) -------------------------------
// Writes costs[0..4], [6..9]; costs[5] (collision) is zeroed when the collision kernel does not run.
constexpr int kNRedPose = 11;
// kGrad = false: no end-pose increments in this launch (the feasible half of a TSGD iteration asks for the evaluator only)
// -> the waypoint walk keeps the end-effector frame alone, no joint axes / origins, no Jacobian, no QR: a lighter kernel
// (fewer registers, more CTAs per SM) instead of a run-time branch around the heavy part.
template <typename T, int MAXREG, bool kFrames, bool kGrad>
__global__ void __launch_bounds__(kTormMaxThreads) __maxnreg__(MAXREG)
torm_pose_eval_kernel(TormP<T> P, const T* __restrict__ traj, const T* __restrict__ targets, int stage_mask,
                      int zero_coll_cost, T* __restrict__ inc_pose, T* __restrict__ costs,
                      const int32_t* __restrict__ traj_stage, const T* __restrict__ Aband, T* __restrict__ inc_smooth) {
    const int CH = blockDim.x, tid = threadIdx.x, Tn = P.Tn;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_traj = reinterpret_cast<T*>(smem_raw);  // [(CH + 12)][7]  rows c0-6 .. c0+CH+5, clamped
    T* s_v = s_traj + (CH + 12) * 7;             // [(CH + 3)][7]   first differences of continuous joints
    T* s_red = s_v + (CH + 3) * 7;               // [8][kNRedPose]
    const int64_t n = blockIdx.x;
    // traj_stage (nullable): per-trajectory stage of the optimiser; the end-pose increments are only computed for
    // trajectories in the end-pose stage (1), the collision / smoothness kernel only runs for those in stage 0
    const bool do_pgrad = (stage_mask & RLITG_STAGE_POSE_GRAD) != 0 && (!traj_stage || traj_stage[n] == 1);
    const bool do_eval = (stage_mask & RLITG_STAGE_EVAL) != 0;
    const T* tr = traj + n * (int64_t)Tn * 7;
    T* o_pose = inc_pose ? inc_pose + n * (int64_t)(Tn - 1) * 7 : nullptr;

    T red[kNRedPose];
#pragma unroll
    for (int k = 0; k < kNRedPose; k++) red[k] = T(0);

    for (int c0 = 0; c0 < Tn; c0 += CH) {
        __syncthreads();
        stage_traj_window<T>(s_traj, tr, c0, Tn, CH, tid);
        __syncthreads();
        const int i = c0 + tid;
        const bool valid = i < Tn;
        // ---- smoothness increments -(2 A_full x) on the free rows (torm_optimizer.cpp:1466-1473), when this launch is
        //      given the band: the trajectory window is already here, and the collision kernel (or, with one stochastic
        //      waypoint per trajectory, a whole extra 15 us launch of it) is spared the 13-tap band
        if (Aband != nullptr && inc_smooth != nullptr && valid && i >= 1 && (!traj_stage || traj_stage[n] == 0)) {
            const T* ab = Aband + (size_t)(i + 5) * 13;
            T* o_smooth = inc_smooth + n * (int64_t)(Tn - 1) * 7;
            T coef[13];
#pragma unroll
            for (int k = 0; k < 13; k++) coef[k] = ab[k];
#pragma unroll
            for (int j = 0; j < 7; j++) {
                T sacc = T(0);
#pragma unroll
                for (int k = 0; k < 13; k++) sacc += coef[k] * (T(2) * TROW(i + k - 6)[j]);
                o_smooth[(i - 1) * 7 + j] = -sacc;
            }
        }
        T q[7];
        if (valid) {
#pragma unroll
            for (int j = 0; j < 7; j++) q[j] = TROW(i)[j];
        }
        // ---- end-pose error at the waypoint: getEndPoseCost :1399-1425
        if (valid && i >= 1) {
            typename std::conditional<kGrad, FkJac<T, 7>, EeOnly<T>>::type fj;
            walk_chain<T, 7>(q, fj);
            V3<T> tp, t0, t1, t2;
            waypoint_target<T, kFrames>(targets, i, tp, t0, t1, t2);
            T angle;
            const V3<T> rv = rot_diff_kdl(fj.ee.c0, fj.ee.c1, fj.ee.c2, t0, t1, t2, angle);
            const V3<T> dp = tp - fj.ee.p;
            red[R_WP_POS] += sqrt_t(dot(dp, dp));
            red[R_WP_ROT] += sqrt_t(dot(rv, rv));
            if constexpr (kGrad) if (do_pgrad && o_pose) {
                T J[42], dq[7];
                fill_jac<T, 7>(fj, J);
                const T tw[6] = {dp.x, dp.y, dp.z, rv.x, rv.y, rv.z};
                if constexpr (sizeof(T) == 8) {
                    pinv_apply<T, 7, true>(J, tw, P.vik_eps, dq);  // verification build: the literal truncated SVD
                } else {
                    T v[6] = {tw[0], tw[1], tw[2], tw[3], tw[4], tw[5]};
                    if (pinv_qr_stream<T, 7>(J, v, dq) < T(1e-4)) pinv_svd_from_row<T>(TROW(i), tw, P.vik_eps, dq);
                }
#pragma unroll
                for (int j = 0; j < 7; j++) o_pose[(i - 1) * 7 + j] = P.jac_w * dq[j];
            }
        }
        if (do_eval) {
            // first differences of continuous joints (:164-181), shared with the neighbouring rows
            for (int e = tid; e < CH + 3; e += CH) {
                const int r = c0 + e;
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    T val = T(0);
                    if (P.cj[j] && r < Tn - 1) val = shortest_angular_distance(TROW(r)[j], TROW(r + 1)[j]) / P.dt;
                    s_v[e * 7 + j] = val;
                }
            }
            __syncthreads();
            if (valid) {
                if (i < Tn - 1) {  // midpoint pose error :76-82
                    T qm[7];
#pragma unroll
                    for (int j = 0; j < 7; j++) qm[j] = (TROW(i + 1)[j] + q[j]) / T(2);
                    EeOnly<T> ev;
                    walk_chain<T, 7>(qm, ev);
                    V3<T> ip, i0, i1, i2;
                    midpoint_target<T, kFrames>(targets, Tn, i, ip, i0, i1, i2);
                    T angle;
                    const V3<T> rv = rot_diff_kdl(ev.ee.c0, ev.ee.c1, ev.ee.c2, i0, i1, i2, angle);
                    const V3<T> dp = ip - ev.ee.p;
                    red[R_MID_POS] += sqrt_t(dot(dp, dp));
                    red[R_MID_ROT] += sqrt_t(dot(rv, rv));
                }
                bool viol = false;
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    if (P.cj[j]) {
                        const T* sv = s_v + tid * 7 + j;
                        const T v0 = sv[0], v1 = sv[7], v2 = sv[14];
                        const T a0 = (v1 - v0) / P.dt, a1 = (v2 - v1) / P.dt;
                        if (i < Tn - 1) red[R_VEL_C] += abs_t(v0);
                        if (i < Tn - 2) red[R_ACC_C] += abs_t(a0);
                        if (i < Tn - 3) red[R_JERK_C] += abs_t((a1 - a0) / P.dt);
                        if (i < Tn - 1 && abs_t(v0) > P.vlim[j]) viol = true;
                    } else {
                        T sv = T(0), sa = T(0), sj = T(0);
#pragma unroll
                        for (int k = -3; k <= 3; k++) {
                            const T xk = TROW(i + k)[j];
                            sv += P.rule_v[k + 3] * xk; sa += P.rule_a[k + 3] * xk; sj += P.rule_j[k + 3] * xk;
                        }
                        red[R_VEL_N] += abs_t(sv); red[R_ACC_N] += abs_t(sa); red[R_JERK_N] += abs_t(sj);
                        if (i < Tn - 1 && abs_t(TROW(i + 1)[j] - q[j]) / P.dt > P.vlim[j]) viol = true;
                    }
                }
                if (viol) red[R_NVIOL_P] += T(1);
            }
        }
    }
    T tot[kNRedPose];
    block_reduce<T, kNRedPose>(red, s_red, tot);
    if (tid == 0) {
        T* c = costs + n * RLITG_TORM_NCOST;
        const T tm1 = T(Tn - 1);
        T ev_pos = T(0), ev_rot = T(0), cvel = T(0), cacc = T(0), cjerk = T(0), nviol = T(0);
        if (do_eval) {
            ev_pos = (tot[R_WP_POS] + tot[R_MID_POS]) / (tm1 * T(2));
            ev_rot = (tot[R_WP_ROT] + tot[R_MID_ROT]) / (tm1 * T(2));
            cvel = (tot[R_VEL_C] / tm1 + tot[R_VEL_N] / T(Tn)) / T(7);
            cacc = (tot[R_ACC_C] / T(Tn - 2) + tot[R_ACC_N] / T(Tn)) / T(7);
            cjerk = (tot[R_JERK_C] / T(Tn - 3) + tot[R_JERK_N] / T(Tn)) / T(7);
            nviol = tot[R_NVIOL_P];
        }
        c[0] = ev_pos; c[1] = ev_rot; c[2] = cvel; c[3] = cacc; c[4] = cjerk;
        if (zero_coll_cost) c[5] = T(0);
        c[6] = do_eval ? ev_pos + T(0.17) * ev_rot : T(0);  // getWeightedPoseCost :59-63
        c[7] = nviol;
        c[8] = tot[R_WP_POS];
        c[9] = tot[R_WP_ROT];
    }
}

#undef TROW

}  // namespace rlitg

#include "torm_coll.cuh"

namespace rlitg {

// ---- smoothness operator (host, FP64): TormCost ctor + scaling -------------------------------------------
// (struct Smooth is declared in rlitg_internal.h: the update step in torm_update.cu shares it)

static const double kRules[3][7] = {{-1.0 / 60, 3.0 / 20, -3.0 / 4, 0.0, 3.0 / 4, -3.0 / 20, 1.0 / 60},
                                    {1.0 / 90, -3.0 / 20, 3.0 / 2, -49.0 / 18, 3.0 / 2, -3.0 / 20, 1.0 / 90},
                                    {1.0 / 8, -1.0, 13.0 / 8, 0.0, -13.0 / 8, 1.0, -1.0 / 8}};

static int smooth_build(Smooth& sm, int T, double dt, const double* w3, double ridge) {
    const int G = T + 10, F = T - 1, o = 6;
    // A_full = sum_d w_d dt^(d+1) D_d^T D_d + ridge I, kept as a 13-band (half bandwidth 6)
    std::vector<double> band((size_t)G * 13, 0.0);
    auto B = [&](int i, int j) -> double& { return band[(size_t)i * 13 + (j - i + 6)]; };
    double mult = 1.0;
    for (int d = 0; d < 3; d++) {
        mult *= dt;
        const double coef = w3[d] * mult;
        if (coef == 0.0) continue;
        for (int r = 0; r < G; r++)
            for (int a = -3; a <= 3; a++) {
                const int ca = r + a;
                if (ca < 0 || ca >= G) continue;
                for (int b = -3; b <= 3; b++) {
                    const int cb = r + b;
                    if (cb < 0 || cb >= G) continue;
                    B(ca, cb) += coef * (kRules[d][a + 3] * kRules[d][b + 3]);
                }
            }
    }
    for (int i = 0; i < G; i++) B(i, i) += ridge;
    // banded Cholesky of the free block A_free = A_full[6 : 6+F, 6 : 6+F] (SPD, half bandwidth 6)
    const int hb = 6;
    std::vector<double> L((size_t)F * (hb + 1), 0.0);  // L(i, j) for i-hb <= j <= i at L[i*(hb+1) + (j - i + hb)]
    auto Lr = [&](int i, int j) -> double& { return L[(size_t)i * (hb + 1) + (j - i + hb)]; };
    for (int i = 0; i < F; i++) {
        for (int j = std::max(0, i - hb); j <= i; j++) {
            double s = B(i + o, j + o);
            for (int k = std::max(0, i - hb); k < j; k++)
                if (k >= j - hb) s -= Lr(i, k) * Lr(j, k);
            if (i == j) {
                if (!(s > 0.0)) return set_error(RLITG_ERR_ARG, "smoothness matrix is not positive definite (all weights zero?)");
                Lr(i, i) = std::sqrt(s);
            } else {
                Lr(i, j) = s / Lr(j, j);
            }
        }
    }
    sm.inv.assign((size_t)F * F, 0.0);
    std::vector<double> y(F);
    double mx = -1e300;
    for (int c = 0; c < F; c++) {
        for (int i = 0; i < F; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = std::max(0, i - hb); k < i; k++) s -= Lr(i, k) * y[k];
            y[i] = s / Lr(i, i);
        }
        for (int i = F - 1; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k <= std::min(F - 1, i + hb); k++) s -= Lr(k, i) * y[k];
            y[i] = s / Lr(i, i);
        }
        for (int i = 0; i < F; i++) {
            sm.inv[(size_t)i * F + c] = y[i];
            mx = std::max(mx, y[i]);
        }
    }
    sm.scale = mx;
    const double mcs = std::max(1.0, mx);  // max_cost_scale starts at 1.0 (torm_optimizer.cpp:211)
    for (auto& x : band) x *= mcs;
    for (auto& x : sm.inv) x /= mcs;
    sm.band.swap(band);
    // Device layout of the factor of the UNSCALED free block (A_scaled^-1 b = (L L^T)^-1 b / mcs) for the column-oriented
    // substitution of torm_update_kernel: as soon as unknown i is known it is pushed into the six rows that use it, so
    // the only dependent operation per step is one FMA.  Two tables of F rows of 8 doubles, coefficients pre-divided
    // by the diagonal of the row they update (entries outside the matrix are 0):
    //   forward  row i: L(i+k,i) / L(i+k,i+k), k = 1..6 | 1 / L(i+7,i+7) | 1 / L(i,i)
    //   backward row i: L(i,i-k) / L(i-k,i-k), k = 1..6 | 1 / L(i-7,i-7) | 1 / L(i,i)
    std::vector<double> Lt((size_t)2 * F * 8, 0.0);
    for (int i = 0; i < F; i++) {
        double* fw = &Lt[(size_t)i * 8];
        double* bw = &Lt[((size_t)F + i) * 8];
        for (int k = 1; k <= 6; k++) {
            if (i + k < F) fw[k - 1] = Lr(i + k, i) / Lr(i + k, i + k);
            if (i - k >= 0) bw[k - 1] = Lr(i, i - k) / Lr(i - k, i - k);
        }
        if (i + 7 < F) fw[6] = 1.0 / Lr(i + 7, i + 7);
        if (i - 7 >= 0) bw[6] = 1.0 / Lr(i - 7, i - 7);
        fw[7] = bw[7] = 1.0 / Lr(i, i);
    }
    sm.chol.swap(Lt);
    // Partitioned substitution: the recurrence over F rows is the critical path of the update step, so the rows are cut
    // into P blocks that are solved at once with zero history; what a block's rows owe to the six rows in front of it
    // (forward) / behind it (backward) is linear in those six values and independent of the right-hand side: W below.
    // At least 12 rows per block (the last one then keeps >= 6: the boundary of a block lies inside its neighbour).
    // RLITG_UPD_PARTS = number of blocks (<= 8); the default is 1, the plain recurrence: measured at 1,024 x 200 the
    // partitioned form is no faster (feasible update 41.0 / 41.0 / 47.1 us with 8 / 4 / 2 blocks against 36.9 us; the
    // sequential recurrence is ~14 us of that, and the boundary walk, the spike rows from L2 and six more CTA barriers
    // cost what the shorter chains save) -- kept behind the knob, covered by tests/test_torm_gpu.py.
    {
        static const int pmax = [] { const char* e = getenv("RLITG_UPD_PARTS"); return e ? std::max(1, std::min(8, atoi(e))) : 1; }();
        int P = std::min(pmax, std::max(1, F / 12));
        int R = (F + P - 1) / P;
        while (P > 1 && F - (P - 1) * R < 6) { P--; R = (F + P - 1) / P; }
        sm.parts = P;
        sm.rows = R;
        sm.spike.assign((size_t)2 * F * 6, 0.0);
        std::vector<double> v(F);
        for (int pb = 0; pb < P; pb++) {
            const int s0 = pb * R, e0 = std::min(s0 + R, F);
            for (int m = 0; m < 6; m++) {
                if (pb >= 1) {  // L y = 0 on the block with y(s0 - 1 - m) = 1 in front of it
                    const int src = s0 - 1 - m;
                    for (int i = s0; i < e0; i++) {
                        double acc = 0.0;
                        for (int k = std::max(src, i - hb); k < i; k++) {
                            const double yk = k < s0 ? (k == src ? 1.0 : 0.0) : v[k];
                            acc -= Lr(i, k) * yk;
                        }
                        v[i] = acc / Lr(i, i);
                        sm.spike[((size_t)i) * 6 + m] = v[i];
                    }
                }
                if (pb <= P - 2) {  // L^T x = 0 on the block with x(e0 + m) = 1 behind it
                    const int src = e0 + m;
                    for (int i = e0 - 1; i >= s0; i--) {
                        double acc = 0.0;
                        for (int k = i + 1; k <= std::min(std::min(F - 1, i + hb), src); k++) {
                            const double xk = k >= e0 ? (k == src ? 1.0 : 0.0) : v[k];
                            acc -= Lr(k, i) * xk;
                        }
                        v[i] = acc / Lr(i, i);
                        sm.spike[((size_t)F + i) * 6 + m] = v[i];
                    }
                }
            }
        }
    }
    sm.mcs = mcs;
    sm.T = T;
    return RLITG_OK;
}

const double* smooth_inv_host(const Smooth* s) { return s->inv.data(); }
int smooth_T(const Smooth* s) { return s->T; }

template <typename T>
static TormP<T> convert_params(const RlitgTormParams& p) {
    TormP<T> o;
    o.Tn = p.T;
    o.dt = (T)p.dt; o.inv_dt = (T)(1.0 / p.dt); o.inv_dt2 = (T)((1.0 / p.dt) * (1.0 / p.dt));
    o.rot_scale = (T)p.rot_scale; o.jac_w = (T)p.jac_weight; o.clearance = (T)p.clearance;
    o.ridge = p.pinv_ridge; o.vik_eps = p.vik_eps;
    o.use_pinv = p.use_pinv; o.quirk = p.affine_axis_quirk;
    for (int k = 0; k < 3; k++) o.off[k] = (T)p.base_off[k];
    for (int j = 0; j < 7; j++) { o.cj[j] = p.cjoint[j]; o.vlim[j] = (T)p.vel_limit[j]; }
    for (int k = 0; k < 7; k++) {
        o.rule_v[k] = (T)(kRules[0][k] / std::pow(p.dt, 1));
        o.rule_a[k] = (T)(kRules[1][k] / std::pow(p.dt, 2));
        o.rule_j[k] = (T)(kRules[2][k] / std::pow(p.dt, 3));
        o.rv[k] = (T)((1.0 / p.dt) * kRules[0][k]);
        o.ra[k] = (T)((1.0 / p.dt) * (1.0 / p.dt) * kRules[1][k]);
        o.rvs[k] = o.ras[k] = T(0);
    }
    o.sqrt_ridge = (T)std::sqrt(p.pinv_ridge); o.ridge_t = (T)p.pinv_ridge;
    for (int k = 0; k < 3; k++) o.off64[k] = p.base_off[k];
    o.fix_unit = T(0); o.fix_org[0] = o.fix_org[1] = o.fix_org[2] = T(0); o.fix_scale = 0.0; o.fix_shift = 0;
    return o;
}

// The fixed-point grid of the sphere centres for one scene: 2^shift units per cell, the shift chosen so that +-10 m around
// the grid origin fit 31 bits (res 0.02 m: shift 22, unit 4.8e-9 m; res 0.01 m: shift 21).  The cell of a centre is then (units + half) >> shift.
template <typename T>
static void set_fixed_point(TormP<T>& P, const SceneDev& sd, double dt) {
    int shift = 8;
    while (shift < 26 && std::ldexp(5.12 / sd.res, shift + 1) < 2147483648.0) shift++;
    P.fix_shift = shift;
    P.fix_scale = std::ldexp(1.0, shift) / sd.res;
    const double unit = sd.res / std::ldexp(1.0, shift);
    P.fix_unit = (T)unit;
    for (int k = 0; k < 3; k++) P.fix_org[k] = (T)sd.origin[k];
    for (int k = 0; k < 7; k++) {
        P.rvs[k] = (T)((1.0 / dt) * kRules[0][k] * unit);
        P.ras[k] = (T)((1.0 / dt) * (1.0 / dt) * kRules[1][k] * unit);
    }
}

template <typename T>
static int launch_torm(const T* traj, const T* targets, int64_t N, const Scene* scene, const Smooth* smooth,
                       const RlitgTormParams* p, int stage_mask, const int32_t* stoch, T* inc_smooth, T* inc_coll,
                       T* inc_pose, T* costs, float* active, const int32_t* traj_stage, const T* frames, cudaStream_t st) {
    TormP<T> P = convert_params<T>(*p);
    const int CH = std::min(kTormMaxThreads, ((p->T + 31) / 32) * 32);
    const bool run_pose = (stage_mask & (RLITG_STAGE_POSE | RLITG_STAGE_POSE_GRAD | RLITG_STAGE_EVAL)) != 0;
    const bool run_coll = (stage_mask & (RLITG_STAGE_COLL | RLITG_STAGE_SMOOTH)) != 0;
    if (!run_pose && !run_coll) return set_error(RLITG_ERR_ARG, "stage_mask selects no stage");
    // use_stochastic_descent_ inside the optimiser loop (RLITG_STAGE_COLL_ROW_ONLY): one waypoint per trajectory takes the
    // collision increment -> kernel 3c (a warp per trajectory, spheres on lanes); the smoothness stage, if asked for, keeps
    // the thread-per-waypoint kernel without its collision part
    const bool row_only = (stage_mask & RLITG_STAGE_COLL) && (stage_mask & RLITG_STAGE_COLL_ROW_ONLY) && stoch && inc_coll;
    // the smoothness band rides in the end-pose / evaluator kernel whenever that kernel runs (RLITG_TORM_SMOOTH_IN_POSE=0:
    // always in the collision kernel)
    static const int sip = [] { const char* e = getenv("RLITG_TORM_SMOOTH_IN_POSE"); return e ? atoi(e) : 1; }();
    const bool smooth_in_pose = sip && run_pose && (stage_mask & RLITG_STAGE_SMOOTH) && inc_smooth != nullptr && smooth != nullptr;
    const T* pose_band = smooth_in_pose
        ? reinterpret_cast<const T*>(sizeof(T) == 4 ? (const void*)smooth->d_band32 : (const void*)smooth->d_band64) : nullptr;
    int coll_kernel_mask = smooth_in_pose ? (stage_mask & ~RLITG_STAGE_SMOOTH) : stage_mask;
    // does any kernel of the collision side run (and own costs[5])?
    const bool coll_side_runs = row_only || (run_coll && (coll_kernel_mask & (RLITG_STAGE_COLL | RLITG_STAGE_SMOOTH)) != 0);
    // When both kernels run they are forked onto two streams (joined again before returning to the caller's
    // stream order): the pose kernel's CTAs fill the SMs that the collision kernel's last, partial wave leaves idle.
    // The side stream and its two events are shared by every caller on a device, so the fork .. join sequence is
    // serialised on the host (launches are asynchronous: the lock is held for microseconds).
    struct Fork { std::mutex mu; cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
    static Fork forks[64];
    std::unique_lock<std::mutex> fork_lock;
    static const int concurrent = [] { const char* e = getenv("RLITG_TORM_CONCURRENT"); return e ? atoi(e) : 1; }();
    cudaStream_t st_pose = st;
    Fork* fk = nullptr;
    if (run_pose && coll_side_runs && concurrent) {
        int dev = 0;
        cudaGetDevice(&dev);
        fk = &forks[dev & 63];
        fork_lock = std::unique_lock<std::mutex>(fk->mu);
        if (!fk->side) {
            if (cudaStreamCreateWithFlags(&fk->side, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&fk->fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&fk->join, cudaEventDisableTiming) != cudaSuccess)
                return set_cuda_error("torm fork stream / events");
        }
        st_pose = fk->side;
        if (cudaEventRecord(fk->fork, st) != cudaSuccess || cudaStreamWaitEvent(st_pose, fk->fork, 0) != cudaSuccess)
            return set_cuda_error("torm fork");
    }
    if (row_only) {
        static const int fk64r = [] { const char* e = getenv("RLITG_TORM_FK64"); return e ? atoi(e) : 1; }();
        SceneDev sd = scene_dev(scene);
        set_fixed_point(P, sd, p->dt);
        const int S = sd.S;
        const size_t smem = sizeof(T) * kRowWarps * (size_t)(7 * ((4 * S) | 1) + 42);
        using RowK = void (*)(TormP<T>, const T*, SceneDev, int, int, int64_t, const int32_t*, T*, T*, float*, const int32_t*);
        RowK rk;
        if constexpr (sizeof(T) == 4) rk = fk64r ? torm_coll_row_kernel<T, true> : torm_coll_row_kernel<T, false>;
        else rk = torm_coll_row_kernel<T, false>;
        int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(rk), smem, "torm_coll_row_kernel");
        if (rc_attr) return rc_attr;
        coll_kernel_mask &= ~RLITG_STAGE_COLL;
        const bool other_coll_kernel = (coll_kernel_mask & RLITG_STAGE_SMOOTH) != 0;
        // who zeroes what: 1 = costs[5] only, 2 = the pose entries too; nothing when the smoothness launch below does it
        const int zero = other_coll_kernel ? 0 : ((run_pose || (stage_mask & RLITG_STAGE_KEEP_COSTS)) ? 1 : 2);
        rk<<<(unsigned)((N + kRowWarps - 1) / kRowWarps), kRowWarps * 32, smem, st>>>(P, traj, sd, S, zero, N, stoch, inc_coll,
                                                                                   costs, active, traj_stage);
        int rc = check_launch("torm_coll_row_kernel");
        if (rc) return rc;
    }
    auto launch_pose = [&]() -> int {
        const size_t smem = sizeof(T) * ((size_t)(CH + 12) * 7 + (size_t)(CH + 3) * 7 + 8 * kNRedPose);
        // register budget of the pose kernel: 2 CTAs/SM (128 regs) or 3 (96 regs at 224 threads); RLITG_TORM_POSE_OCC
        static const int occ = [] { const char* e = getenv("RLITG_TORM_POSE_OCC"); return e ? atoi(e) : 2; }();
        using PoseK = void (*)(TormP<T>, const T*, const T*, int, int, T*, T*, const int32_t*, const T*, T*);
        constexpr int kR2 = sizeof(T) == 4 ? 128 : 255, kR3 = sizeof(T) == 4 ? 96 : 255, kRL = sizeof(T) == 4 ? 72 : 255;
        const bool grad = (stage_mask & RLITG_STAGE_POSE_GRAD) != 0;
        PoseK kern;
        if (!grad) kern = frames ? torm_pose_eval_kernel<T, kRL, true, false> : torm_pose_eval_kernel<T, kRL, false, false>;
        else if (frames) kern = (sizeof(T) == 4 && occ == 3) ? torm_pose_eval_kernel<T, kR3, true, true> : torm_pose_eval_kernel<T, kR2, true, true>;
        else kern = (sizeof(T) == 4 && occ == 3) ? torm_pose_eval_kernel<T, kR3, false, true> : torm_pose_eval_kernel<T, kR2, false, true>;
        {
            int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "torm_pose_eval_kernel");
            if (rc_attr) return rc_attr;
        }
        kern<<<(unsigned)N, CH, smem, st_pose>>>(P, traj, frames ? frames : targets, stage_mask,
                                                 (coll_side_runs || (stage_mask & RLITG_STAGE_KEEP_COSTS)) ? 0 : 1, inc_pose, costs,
                                                 traj_stage, smooth_in_pose ? pose_band : nullptr,
                                                 smooth_in_pose ? inc_smooth : nullptr);
        return check_launch("torm_pose_eval_kernel");
    };
    // Launch order of the two concurrent kernels.  The heavy end-pose kernel (pinv step: 48 us alone, two CTAs per SM like
    // the collision kernel) goes FIRST: the collision kernel's CTAs then move in as its CTAs retire and the two kernels'
    // partial last waves overlap (full call 156.5 -> 146.6 us; a high-priority side stream gives 148.5).  The light
    // evaluator-only kernel (23 us, four CTAs per SM) goes SECOND, into the gaps of the collision kernel: first, it would
    // simply run ahead of it (feasible stage 130 vs 122 us).  RLITG_TORM_ORDER = 0 / 1 forces collision-first / pose-first.
    static const int pose_first_env = [] { const char* e = getenv("RLITG_TORM_ORDER"); return e ? atoi(e) : -1; }();
    const bool pose_first = pose_first_env < 0 ? (stage_mask & RLITG_STAGE_POSE_GRAD) != 0 : pose_first_env != 0;
    if (run_pose && pose_first) {
        int rc = launch_pose();
        if (rc) return rc;
    }
    if (run_coll && (coll_kernel_mask & (RLITG_STAGE_COLL | RLITG_STAGE_SMOOTH))) {
        const int stage_mask_pose = stage_mask;
        stage_mask = coll_kernel_mask;
        SceneDev sd{};
        int S = 0;
        if (stage_mask & RLITG_STAGE_COLL) {
            sd = scene_dev(scene);
            S = sd.S;
            set_fixed_point(P, sd, p->dt);
        }
        static const int cocc = [] { const char* e = getenv("RLITG_TORM_COLL_OCC"); return e ? atoi(e) : 2; }();
        const int PW = (sizeof(T) == 4 && cocc == 3) ? 3 : 4;  // three-value sphere rows: what lets a third CTA fit in shared memory
        const size_t smem = sizeof(T) * ((size_t)(CH + 12) * 7 + 8 * 2 + (size_t)(std::min(CH, (int)p->T) + 6) * ((PW * S) | 1));
        if (smem > 220 * 1024) return set_error(RLITG_ERR_SHAPE, "too many spheres for the shared-memory tile");
        // FP32 build: chain walk of phase 1 in FP64 (accuracy of the 1 / dt^2 stencils, torm_coll.cuh); RLITG_TORM_FK64=0
        // keeps it in FP32 (measurement knob).  RLITG_TORM_COLL_OCC: 2 CTAs/SM (128 registers, four-value sphere rows with the
        // cell index) or 3 (96 registers + three-value rows: 21 warps/SM, measured 125 vs 108 us -- more resident warps do
        // not pay on this kernel)
        static const int fk64 = [] { const char* e = getenv("RLITG_TORM_FK64"); return e ? atoi(e) : 1; }();
        using KernT = void (*)(TormP<T>, const T*, SceneDev, int, const T*, int, int, const int32_t*, T*, T*, T*, float*,
                               const int32_t*);
        KernT kern;
        if constexpr (sizeof(T) == 4) {
            if (cocc == 3) kern = fk64 ? torm_coll_kernel<T, true, 96, 3> : torm_coll_kernel<T, false, 96, 3>;
            else kern = fk64 ? torm_coll_kernel<T, true, 128, 4> : torm_coll_kernel<T, false, 128, 4>;
        } else {
            kern = torm_coll_kernel<T, false, 255, 4>;
        }
        {
            int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "torm_coll_kernel");
            if (rc_attr) return rc_attr;
        }
        const T* band = nullptr;
        if (stage_mask & RLITG_STAGE_SMOOTH)
            band = reinterpret_cast<const T*>(sizeof(T) == 4 ? (const void*)smooth->d_band32 : (const void*)smooth->d_band64);
        const int zero_pose = (run_pose || (stage_mask & RLITG_STAGE_KEEP_COSTS)) ? 0 : 1;
        kern<<<(unsigned)N, CH, smem, st>>>(P, traj, sd, S, band, stage_mask, zero_pose, stoch, inc_smooth,
                                            row_only ? nullptr : inc_coll, costs, row_only ? nullptr : active, traj_stage);
        int rc = check_launch("torm_coll_kernel");
        if (rc) return rc;
        stage_mask = stage_mask_pose;
    }
    if (run_pose && !pose_first) {
        int rc = launch_pose();
        if (rc) return rc;
    }
    if (fk) {
        if (cudaEventRecord(fk->join, st_pose) != cudaSuccess || cudaStreamWaitEvent(st, fk->join, 0) != cudaSuccess)
            return set_cuda_error("torm join");
    }
    return RLITG_OK;
}

}  // namespace rlitg

using namespace rlitg;

extern "C" {

RLITG_API int rlitg_smooth_create(int T, double dt, const double* w3, double ridge, RlitgSmooth** out) {
    if (!out || !w3) return set_error(RLITG_ERR_NULL, "out / w3 is NULL");
    if (T < 5 || T > 8192) return set_error(RLITG_ERR_SHAPE, "T must be in 5..8192");
    if (!(dt > 0)) return set_error(RLITG_ERR_ARG, "dt must be > 0");
    Smooth* s = new Smooth();
    int rc = smooth_build(*s, T, dt, w3, ridge);
    if (rc) { delete s; return rc; }
    const size_t nb = s->band.size();
    std::vector<float> b32(nb);
    for (size_t i = 0; i < nb; i++) b32[i] = (float)s->band[i];
    if (cudaMalloc(&s->d_band64, nb * 8) != cudaSuccess || cudaMalloc(&s->d_band32, nb * 4) != cudaSuccess ||
        cudaMemcpy(s->d_band64, s->band.data(), nb * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(s->d_band32, b32.data(), nb * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        rc = set_cuda_error("rlitg_smooth_create");
        rlitg_smooth_destroy(reinterpret_cast<RlitgSmooth*>(s));
        return rc;
    }
    // for the update step: banded Cholesky factor (preconditioning solve) and the dense scaled inverse (rows are
    // the columns handleJointLimits adds, torm/src/torm_optimizer.cpp:1256-1263)
    const size_t nl = s->chol.size(), ni = s->inv.size(), ns = s->spike.size();
    if (cudaMalloc(&s->d_chol, nl * 8) != cudaSuccess || cudaMalloc(&s->d_inv64, ni * 8) != cudaSuccess ||
        cudaMalloc(&s->d_spike, ns * 8) != cudaSuccess ||
        cudaMemcpy(s->d_spike, s->spike.data(), ns * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(s->d_chol, s->chol.data(), nl * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(s->d_inv64, s->inv.data(), ni * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        rc = set_cuda_error("rlitg_smooth_create");
        rlitg_smooth_destroy(reinterpret_cast<RlitgSmooth*>(s));
        return rc;
    }
    *out = reinterpret_cast<RlitgSmooth*>(s);
    return RLITG_OK;
}

RLITG_API int rlitg_smooth_destroy(RlitgSmooth* h) {
    Smooth* s = reinterpret_cast<Smooth*>(h);
    if (!s) return RLITG_OK;
    if (s->d_band64) cudaFree(s->d_band64);
    if (s->d_band32) cudaFree(s->d_band32);
    if (s->d_inv64) cudaFree(s->d_inv64);
    if (s->d_chol) cudaFree(s->d_chol);
    if (s->d_spike) cudaFree(s->d_spike);
    delete s;
    return RLITG_OK;
}

RLITG_API int rlitg_smooth_get(const RlitgSmooth* h, double* Aband, double* scale, double* Ainv) {
    const Smooth* s = reinterpret_cast<const Smooth*>(h);
    if (!s) return set_error(RLITG_ERR_NULL, "smooth handle is NULL");
    if (Aband) std::copy(s->band.begin(), s->band.end(), Aband);
    if (scale) *scale = s->scale;
    if (Ainv) std::copy(s->inv.begin(), s->inv.end(), Ainv);
    return RLITG_OK;
}

RLITG_API int rlitg_torm_cost_grad(const void* traj, const void* targets, int64_t N, int dtype,
                                   const RlitgScene* scene, const RlitgSmooth* smooth, const RlitgTormParams* p,
                                   int stage_mask, const int32_t* stochastic_idx, void* inc_smooth, void* inc_coll,
                                   void* inc_pose, void* costs, float* active, void* stream) {
    return rlitg_torm_cost_grad_staged(traj, targets, N, dtype, scene, smooth, p, stage_mask, nullptr, stochastic_idx,
                                       inc_smooth, inc_coll, inc_pose, costs, active, stream);
}

RLITG_API int rlitg_torm_cost_grad_staged(const void* traj, const void* targets, int64_t N, int dtype,
                                          const RlitgScene* scene, const RlitgSmooth* smooth, const RlitgTormParams* p,
                                          int stage_mask, const int32_t* traj_stage, const int32_t* stochastic_idx,
                                          void* inc_smooth, void* inc_coll, void* inc_pose, void* costs, float* active,
                                          void* stream) {
    return rlitg_torm_cost_grad_frames(traj, targets, nullptr, N, dtype, scene, smooth, p, stage_mask, traj_stage,
                                       stochastic_idx, inc_smooth, inc_coll, inc_pose, costs, active, stream);
}

RLITG_API int rlitg_torm_target_frames(const void* targets, int T, int dtype, void* frames, void* stream) {
    if (!targets || !frames) return set_error(RLITG_ERR_NULL, "targets / frames is NULL");
    if (T < 2 || T > 8192) return set_error(RLITG_ERR_SHAPE, "T must be in 2..8192");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((2 * T - 1 + 127) / 128);
    if (dtype == RLITG_F32) torm_target_frames_kernel<float><<<grid, 128, 0, st>>>((const float*)targets, T, (float*)frames);
    else if (dtype == RLITG_F64) torm_target_frames_kernel<double><<<grid, 128, 0, st>>>((const double*)targets, T, (double*)frames);
    else return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    return check_launch("torm_target_frames_kernel");
}

RLITG_API int rlitg_torm_cost_grad_frames(const void* traj, const void* targets, const void* target_frames, int64_t N,
                                          int dtype, const RlitgScene* scene, const RlitgSmooth* smooth,
                                          const RlitgTormParams* p, int stage_mask, const int32_t* traj_stage,
                                          const int32_t* stochastic_idx, void* inc_smooth, void* inc_coll, void* inc_pose,
                                          void* costs, float* active, void* stream) {
    if (!p) return set_error(RLITG_ERR_NULL, "params is NULL");
    if (p->D != 7) return set_error(RLITG_ERR_SHAPE, "TORM runs the 7-joint arm chain (D = 7)");
    // the evaluator's T - 2 / T - 3 divisors and the smoothness band need five rows; the end-pose and collision terms are
    // per waypoint and also serve the initialisers' short segments (getNewTrajectory scores segments of 2..10 rows)
    if (p->T < 2 || p->T > 8192) return set_error(RLITG_ERR_SHAPE, "T must be in 2..8192");
    if (p->T < 5 && (stage_mask & (RLITG_STAGE_EVAL | RLITG_STAGE_SMOOTH)))
        return set_error(RLITG_ERR_SHAPE, "the evaluator and smoothness stages need T >= 5");
    if (N < 0 || N > 0x7fffffff) return set_error(RLITG_ERR_SHAPE, "N out of range");
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    if (N == 0) return RLITG_OK;
    if (!traj || (!targets && !target_frames) || !costs) return set_error(RLITG_ERR_NULL, "traj / targets / costs is NULL");
    const Scene* sc = reinterpret_cast<const Scene*>(scene);
    const Smooth* sm = reinterpret_cast<const Smooth*>(smooth);
    if ((stage_mask & RLITG_STAGE_COLL) && !scene_ready(sc, true, true))
        return set_error(RLITG_ERR_ARG, "the collision stage needs a scene with a grid and a sphere table");
    if (stage_mask & RLITG_STAGE_SMOOTH) {
        if (!sm || !inc_smooth) return set_error(RLITG_ERR_NULL, "the smoothness stage needs a smooth handle and inc_smooth");
        if (smooth_T(sm) != p->T) return set_error(RLITG_ERR_SHAPE, "smooth handle was built for a different T");
    }
    if ((stage_mask & RLITG_STAGE_POSE_GRAD) && !inc_pose) return set_error(RLITG_ERR_NULL, "inc_pose is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32)
        return launch_torm<float>((const float*)traj, (const float*)targets, N, sc, sm, p, stage_mask, stochastic_idx,
                                  (float*)inc_smooth, (float*)inc_coll, (float*)inc_pose, (float*)costs, active, traj_stage,
                                  (const float*)target_frames, st);
    return launch_torm<double>((const double*)traj, (const double*)targets, N, sc, sm, p, stage_mask, stochastic_idx,
                               (double*)inc_smooth, (double*)inc_coll, (double*)inc_pose, (double*)costs, active, traj_stage,
                               (const double*)target_frames, st);
}

}  // extern "C"
