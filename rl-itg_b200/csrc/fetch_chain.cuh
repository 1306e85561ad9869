// Fetch arm kinematic chain, walked in registers.
//
// Constants: fetch_description/robots/fetch.urdf (torso_lift :102-108, shoulder_pan :188-195, shoulder_lift
// :218-225, upperarm_roll :248-255, elbow_flex :278-285, forearm_roll :308-315, wrist_flex :338-345, wrist_roll
// :368-375, gripper_axis :395-400).  Every joint rpy is zero and every axis is a coordinate axis, so
// KDL's Segment::pose(q) = T(origin) * Rot2(axis, q) (what the reference evaluates through
// KDL::ChainFkSolverPos_recursive, torm/src/torm_ik_solver.cpp:156-163) reduces to a translation along the
// current columns followed by a planar rotation of two columns.  The rotation is kept as three column vectors.
//
// D = 7: chain torso_lift_link -> gripper_link (the chain the reference runs).
// D = 8: the prismatic torso_lift_joint is prepended (chain base_link -> gripper_link).
#pragma once
#include <type_traits>

#include "rlitg_common.cuh"

namespace rlitg {

enum { AX_X = 0, AX_Y = 1, AX_Z = 2 };

template <int J> using jidx = std::integral_constant<int, J>;

// arm joint j (0..6): axis and origin (in the parent link frame)
__host__ __device__ constexpr int arm_axis(int j) {
    return j == 0 ? AX_Z : ((j == 1 || j == 3 || j == 5) ? AX_Y : AX_X);
}
template <typename T> __host__ __device__ constexpr T arm_ox(int j) {
    return j == 0 ? T(0.119525) : j == 1 ? T(0.117) : j == 2 ? T(0.219) : j == 3 ? T(0.133)
         : j == 4 ? T(0.197) : j == 5 ? T(0.1245) : T(0.1385);
}
template <typename T> __host__ __device__ constexpr T arm_oz(int j) {
    return j == 0 ? T(0.34858) : j == 1 ? T(0.0599999999999999) : T(0);
}
template <typename T> __host__ __device__ constexpr T tip_ox() { return T(0.16645); }
// torso_lift_joint origin in base_link (also fk_base_position of torm/launch/config/robot/fetch.yaml:18)
template <typename T> __host__ __device__ constexpr T torso_ox() { return T(-0.086875); }
template <typename T> __host__ __device__ constexpr T torso_oz() { return T(0.37743); }

template <typename T>
struct FrameC {  // rotation columns + origin
    V3<T> c0, c1, c2, p;
};

// Rotates the frame by angle (s, c) about its own AXIS.
template <int AXIS, typename T>
RLITG_DEV void rotate_axis(FrameC<T>& f, T s, T c) {
    if (AXIS == AX_Z) {
        const V3<T> a = f.c0, b = f.c1;
        f.c0 = {c * a.x + s * b.x, c * a.y + s * b.y, c * a.z + s * b.z};
        f.c1 = {c * b.x - s * a.x, c * b.y - s * a.y, c * b.z - s * a.z};
    } else if (AXIS == AX_Y) {
        const V3<T> a = f.c0, b = f.c2;
        f.c0 = {c * a.x - s * b.x, c * a.y - s * b.y, c * a.z - s * b.z};
        f.c2 = {s * a.x + c * b.x, s * a.y + c * b.y, s * a.z + c * b.z};
    } else {
        const V3<T> a = f.c1, b = f.c2;
        f.c1 = {c * a.x + s * b.x, c * a.y + s * b.y, c * a.z + s * b.z};
        f.c2 = {c * b.x - s * a.x, c * b.y - s * a.y, c * b.z - s * a.z};
    }
}

template <int AXIS, typename T>
RLITG_DEV V3<T> axis_of(const FrameC<T>& f) {
    return AXIS == AX_Z ? f.c2 : (AXIS == AX_Y ? f.c1 : f.c0);
}

template <typename T, int J, int D, bool kFastTrig64, typename QT, typename Visit>
RLITG_DEV void arm_step(FrameC<T>& f, const QT* q, Visit& visit) {
    constexpr int A = D - 7;  // index offset of arm joint 0 in q
    constexpr T ox = arm_ox<T>(J), oz = arm_oz<T>(J);
    // translate by the joint origin (y component is zero for every arm joint)
    f.p = axpy(f.p, ox, f.c0);
    if (oz != T(0)) f.p = axpy(f.p, oz, f.c2);
    T s, c;
    if constexpr (kFastTrig64 && sizeof(T) == 8) sincos_joint(T(q[A + J]), s, c, FastTrig64{});
    else sincos_joint(T(q[A + J]), s, c);
    rotate_axis<arm_axis(J)>(f, s, c);
    visit(jidx<A + J>{}, f);
    if constexpr (J < 6) arm_step<T, J + 1, D, kFastTrig64>(f, q, visit);
}

// Walks the chain.  visit(jidx<k>, frame) is called with the frame of the child link of joint k
// (k = 0..D-1, i.e. KDL segment k's tip frame) and finally with k = D for gripper_link.
// kFastTrig64: a double walk inside an FP32 production kernel uses the lean FP64 sincos (rlitg_common.cuh).
// The joint values may be stored in another precision than the walk (QT: e.g. FP32 rows in shared memory, FP64 walk).
template <typename T, int D, bool kFastTrig64 = false, typename QT, typename Visit>
RLITG_DEV void walk_chain(const QT* q, Visit& visit) {
    static_assert(D == 7 || D == 8, "Fetch chain has 7 (arm) or 8 (torso + arm) joints");
    FrameC<T> f;
    f.c0 = {T(1), T(0), T(0)};
    f.c1 = {T(0), T(1), T(0)};
    f.c2 = {T(0), T(0), T(1)};
    f.p = {T(0), T(0), T(0)};
    if constexpr (D == 8) {
        f.p = {torso_ox<T>(), T(0), torso_oz<T>() + T(q[0])};
        visit(jidx<0>{}, f);
    }
    arm_step<T, 0, D, kFastTrig64>(f, q, visit);
    f.p = axpy(f.p, tip_ox<T>(), f.c0);
    visit(jidx<D>{}, f);
}

// joint kind / world axis for chain joint k of a D-joint chain
template <int D> __host__ __device__ constexpr bool joint_is_prismatic(int k) { return D == 8 && k == 0; }
template <int D> __host__ __device__ constexpr int joint_axis(int k) { return (D == 8 && k == 0) ? AX_Z : arm_axis(k - (D - 7)); }

// FK + geometric Jacobian (KDL::ChainJntToJacSolver::JntToJac semantics, torm/src/torm_ik_solver.cpp:173-175):
// column k = [z_k x (p_ee - p_k); z_k] for a revolute joint, [z_k; 0] for the prismatic torso.
template <typename T, int D>
struct FkJac {
    V3<T> z[D];
    V3<T> o[D];
    FrameC<T> ee;
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        if constexpr (K < D) {
            z[K] = axis_of<joint_axis<D>(K)>(f);
            o[K] = f.p;
        } else {
            ee = f;
        }
    }
    // J(r, k) for r in 0..5
    template <int K>
    RLITG_DEV void column(V3<T>& lin, V3<T>& ang) const {
        if constexpr (joint_is_prismatic<D>(K)) {
            lin = z[K];
            ang = {T(0), T(0), T(0)};
        } else {
            lin = cross(z[K], ee.p - o[K]);
            ang = z[K];
        }
    }
};

// row-major [6][D] Jacobian from the walked chain
template <typename T, int D, int K = 0>
RLITG_DEV void fill_jac_cols(const FkJac<T, D>& fj, T* jv) {
    V3<T> lin, ang;
    fj.template column<K>(lin, ang);
    jv[0 * D + K] = lin.x; jv[1 * D + K] = lin.y; jv[2 * D + K] = lin.z;
    jv[3 * D + K] = ang.x; jv[4 * D + K] = ang.y; jv[5 * D + K] = ang.z;
    if constexpr (K + 1 < D) fill_jac_cols<T, D, K + 1>(fj, jv);
}
template <typename T, int D>
RLITG_DEV void fill_jac(const FkJac<T, D>& fj, T (&jv)[6 * D]) { fill_jac_cols<T, D, 0>(fj, jv); }

}  // namespace rlitg
