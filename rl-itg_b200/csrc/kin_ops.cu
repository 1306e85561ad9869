// Per-configuration kinematic operators that sit next to FK / Jacobian on the hot path:
//   pose_diff       KDL::diff(FK(q), target)              torm/src/torm_optimizer.cpp:1413,1440
//   pinv_step       KDL::ChainIkSolverVel_pinv::CartToJnt torm/src/torm_ik_solver.cpp:169-171
//   nullspace_norm  jaco_srv                              pirl/src/kinematic_solver.cpp:332-351
// One thread per configuration; these are small-batch helpers (the fused TORM / PIRL kernels inline the same
// device functions), so loads are plain array-of-structs.
#include "fetch_chain.cuh"
#include "linalg.cuh"
#include "rlitg_internal.h"

namespace rlitg {

template <typename T, int D>
__global__ void __launch_bounds__(128) pose_diff_kernel(const T* __restrict__ q, const T* __restrict__ target,
                                                        int64_t B, T* __restrict__ twist) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T qv[D];
#pragma unroll
    for (int j = 0; j < D; j++) qv[j] = q[b * D + j];
    FkJac<T, D> fj;
    walk_chain<T, D>(qv, fj);
    const T* t7 = target + b * 7;
    V3<T> t0, t1, t2;
    quat_to_rot_kdl(t7[3], t7[4], t7[5], t7[6], t0, t1, t2);
    T angle;
    const V3<T> rv = rot_diff_kdl(fj.ee.c0, fj.ee.c1, fj.ee.c2, t0, t1, t2, angle);
    T* o = twist + b * 6;
    o[0] = t7[0] - fj.ee.p.x; o[1] = t7[1] - fj.ee.p.y; o[2] = t7[2] - fj.ee.p.z;
    o[3] = rv.x; o[4] = rv.y; o[5] = rv.z;
}

template <typename T, int D>
__global__ void __launch_bounds__(128) pinv_step_kernel(const T* __restrict__ q, const T* __restrict__ twist,
                                                        int64_t B, double eps, T* __restrict__ dq) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T qv[D];
#pragma unroll
    for (int j = 0; j < D; j++) qv[j] = q[b * D + j];
    FkJac<T, D> fj;
    walk_chain<T, D>(qv, fj);
    T J[6 * D], v[6], out[D];
    fill_jac<T, D>(fj, J);
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = twist[b * 6 + i];
    pinv_apply<T, D, sizeof(T) == 8>(J, v, eps, out);
#pragma unroll
    for (int j = 0; j < D; j++) dq[b * D + j] = out[j];
}

template <typename T>
__global__ void __launch_bounds__(128) nullspace_kernel(const T* __restrict__ err, const T* __restrict__ q, int64_t B,
                                                        T* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T qv[7], e[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { qv[j] = q[b * 7 + j]; e[j] = err[b * 7 + j]; }
    FkJac<T, 7> fj;
    walk_chain<T, 7>(qv, fj);
    T J[42];
    fill_jac<T, 7>(fj, J);
    out[b] = nullspace_norm<T, 7>(J, e);
}

// sqrt(det(J J^T)) -- the "yosh" manipulability measure of TormOptimizer::checkSingularity
// (torm/src/torm_optimizer.cpp:748-765).  With J^T = Q R (modified Gram-Schmidt), det(J J^T) = prod(R_kk)^2.
template <typename T, int D>
__global__ void __launch_bounds__(128) manipulability_kernel(const T* __restrict__ q, int64_t B, T* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T qv[D];
#pragma unroll
    for (int j = 0; j < D; j++) qv[j] = q[b * D + j];
    FkJac<T, D> fj;
    walk_chain<T, D>(qv, fj);
    T J[6 * D], R[36];
    fill_jac<T, D>(fj, J);
    mgs_rows<T, D>(J, R);
    T prod = T(1);
#pragma unroll
    for (int k = 0; k < 6; k++) prod *= R[k * 6 + k];
    out[b] = abs_t(prod);
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 127) / 128); }

template <typename T>
static int run_pose_diff(const void* q, const void* target, int64_t B, int D, void* twist, cudaStream_t st) {
    if (D == 7) pose_diff_kernel<T, 7><<<nblk(B), 128, 0, st>>>((const T*)q, (const T*)target, B, (T*)twist);
    else pose_diff_kernel<T, 8><<<nblk(B), 128, 0, st>>>((const T*)q, (const T*)target, B, (T*)twist);
    return check_launch("pose_diff_kernel");
}
template <typename T>
static int run_pinv(const void* q, const void* twist, int64_t B, int D, double eps, void* dq, cudaStream_t st) {
    if (D == 7) pinv_step_kernel<T, 7><<<nblk(B), 128, 0, st>>>((const T*)q, (const T*)twist, B, eps, (T*)dq);
    else pinv_step_kernel<T, 8><<<nblk(B), 128, 0, st>>>((const T*)q, (const T*)twist, B, eps, (T*)dq);
    return check_launch("pinv_step_kernel");
}

}  // namespace rlitg

using namespace rlitg;

static int check_args(int64_t B, int D, int dtype) {
    if (B < 0) return set_error(RLITG_ERR_SHAPE, "B must be >= 0");
    if (D != 7 && D != 8) return set_error(RLITG_ERR_SHAPE, "D must be 7 (arm) or 8 (torso + arm)");
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    return RLITG_OK;
}

extern "C" {

RLITG_API int rlitg_nullspace_norm(const void* err, const void* q, int64_t B, int dtype, void* out, void* stream) {
    int rc = check_args(B, 7, dtype);
    if (rc) return rc;
    if (B == 0) return RLITG_OK;
    if (!err || !q || !out) return set_error(RLITG_ERR_NULL, "err / q / out is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32) nullspace_kernel<float><<<nblk(B), 128, 0, st>>>((const float*)err, (const float*)q, B, (float*)out);
    else nullspace_kernel<double><<<nblk(B), 128, 0, st>>>((const double*)err, (const double*)q, B, (double*)out);
    return check_launch("nullspace_kernel");
}

RLITG_API int rlitg_manipulability(const void* q, int64_t B, int D, int dtype, void* out, void* stream) {
    int rc = check_args(B, D, dtype);
    if (rc) return rc;
    if (B == 0) return RLITG_OK;
    if (!q || !out) return set_error(RLITG_ERR_NULL, "q / out is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32) {
        if (D == 7) manipulability_kernel<float, 7><<<nblk(B), 128, 0, st>>>((const float*)q, B, (float*)out);
        else manipulability_kernel<float, 8><<<nblk(B), 128, 0, st>>>((const float*)q, B, (float*)out);
    } else {
        if (D == 7) manipulability_kernel<double, 7><<<nblk(B), 128, 0, st>>>((const double*)q, B, (double*)out);
        else manipulability_kernel<double, 8><<<nblk(B), 128, 0, st>>>((const double*)q, B, (double*)out);
    }
    return check_launch("manipulability_kernel");
}

RLITG_API int rlitg_pinv_step(const void* q, const void* twist, int64_t B, int D, int dtype, double eps, void* dq,
                              void* stream) {
    int rc = check_args(B, D, dtype);
    if (rc) return rc;
    if (B == 0) return RLITG_OK;
    if (!q || !twist || !dq) return set_error(RLITG_ERR_NULL, "q / twist / dq is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == RLITG_F32 ? run_pinv<float>(q, twist, B, D, eps, dq, st) : run_pinv<double>(q, twist, B, D, eps, dq, st);
}

RLITG_API int rlitg_pose_diff(const void* q, const void* target, int64_t B, int D, int dtype, void* twist,
                              void* stream) {
    int rc = check_args(B, D, dtype);
    if (rc) return rc;
    if (B == 0) return RLITG_OK;
    if (!q || !target || !twist) return set_error(RLITG_ERR_NULL, "q / target / twist is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == RLITG_F32 ? run_pose_diff<float>(q, target, B, D, twist, st) : run_pose_diff<double>(q, target, B, D, twist, st);
}

}  // extern "C"
