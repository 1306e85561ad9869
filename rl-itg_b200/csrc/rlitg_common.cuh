// Common device helpers for the rlitg sm_100a kernels: small vector algebra, range-limited sincos,
// rotation <-> quaternion conversions in the KDL conventions the reference relies on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RLITG_DEV __device__ __forceinline__

namespace rlitg {

template <typename T>
struct V3 {
    T x, y, z;
};

template <typename T> RLITG_DEV V3<T> mk(T x, T y, T z) { return V3<T>{x, y, z}; }
template <typename T> RLITG_DEV V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> RLITG_DEV V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> RLITG_DEV V3<T> operator*(V3<T> a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T> RLITG_DEV V3<T> operator*(T s, V3<T> a) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T> RLITG_DEV T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> RLITG_DEV V3<T> cross(V3<T> a, V3<T> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// a + s * b
template <typename T> RLITG_DEV V3<T> axpy(V3<T> a, T s, V3<T> b) { return {a.x + s * b.x, a.y + s * b.y, a.z + s * b.z}; }

RLITG_DEV float rsqrt_t(float x) { return rsqrtf(x); }
RLITG_DEV double rsqrt_t(double x) { return 1.0 / sqrt(x); }
RLITG_DEV float sqrt_t(float x) { return sqrtf(x); }
RLITG_DEV double sqrt_t(double x) { return sqrt(x); }
RLITG_DEV float abs_t(float x) { return fabsf(x); }
RLITG_DEV double abs_t(double x) { return fabs(x); }
RLITG_DEV float atan2_t(float y, float x) { return atan2f(y, x); }
RLITG_DEV double atan2_t(double y, double x) { return atan2(y, x); }
RLITG_DEV float acos_t(float x) { return acosf(x); }
RLITG_DEV double acos_t(double x) { return acos(x); }
RLITG_DEV float exp_t(float x) { return __expf(x); }
RLITG_DEV double exp_t(double x) { return exp(x); }
RLITG_DEV float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
RLITG_DEV double fma_t(double a, double b, double c) { return fma(a, b, c); }
RLITG_DEV float floor_t(float x) { return floorf(x); }
RLITG_DEV double floor_t(double x) { return floor(x); }

// sin / cos for joint angles.  Joint values are bounded (|q| <= pi after the reference's wrapping, a few
// multiples at worst), so the float path is a 3-term Cody-Waite reduction by pi/2 plus minimax polynomials on
// [-pi/4, pi/4] (about 1 ulp), with none of libdevice's large-argument slow path or its stack frame.
RLITG_DEV void sincos_joint(float x, float& s, float& c) {
    const float k = rintf(x * 0.636619772367581f);
    float r = fmaf(k, -1.5707962512969971f, x);
    r = fmaf(k, -7.5497894158615964e-08f, r);
    r = fmaf(k, -5.3903029534742384e-15f, r);
    const float r2 = r * r;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    sp = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    cp = fmaf(cp, r2, -0.5f);
    cp = fmaf(cp, r2, 1.0f);
    const int q = (int)k;
    const float ss = (q & 1) ? cp : sp;
    const float cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}
RLITG_DEV void sincos_joint(double x, double& s, double& c) { sincos(x, &s, &c); }

// Lean FP64 sin / cos for the FP32 production kernels that walk the chain in double (TORM collision kernel: sphere
// centres accurate to ~1e-11 m so that the 1 / dt^2-scaled stencils keep their digits).  Two-constant Cody-Waite
// reduction by pi/2 and the fdlibm kernel polynomials on [-pi/4, pi/4] (|error| < 1e-15 for |x| < 1e4): ~20 FP64
// instructions, no large-argument path, no stack frame -- unlike libdevice's sincos, which the FP64 verification
// build keeps using.
struct FastTrig64 {};
RLITG_DEV void sincos_joint(double x, double& s, double& c, FastTrig64) {
    const double k = rint(x * 0.63661977236758138);
    double r = fma(k, -1.5707963267948966, x);
    r = fma(k, -6.123233995736766e-17, r);
    const double z = r * r;
    double sp = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    sp = fma(sp, z, 2.75573137070700676789e-06);
    sp = fma(sp, z, -1.98412698298579493134e-04);
    sp = fma(sp, z, 8.33333333332248946124e-03);
    sp = fma(sp, z, -1.66666666666666324348e-01);
    sp = fma(sp * z, r, r);
    double cp = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    cp = fma(cp, z, -2.75573143513906633035e-07);
    cp = fma(cp, z, 2.48015872894767294178e-05);
    cp = fma(cp, z, -1.38888888888741095749e-03);
    cp = fma(cp, z, 4.16666666666666019037e-02);
    cp = fma(cp * z, z, fma(z, -0.5, 1.0));
    const int q = (int)k;
    const double ss = (q & 1) ? cp : sp;
    const double cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// KDL Rotation::GetQuaternion (trace-branch, sign not canonicalised).  R given by columns c0 c1 c2.
template <typename T>
RLITG_DEV void rot_to_quat_kdl(V3<T> c0, V3<T> c1, V3<T> c2, T& x, T& y, T& z, T& w) {
    // R(i,j): row i of column j
    const T r00 = c0.x, r10 = c0.y, r20 = c0.z, r01 = c1.x, r11 = c1.y, r21 = c1.z, r02 = c2.x, r12 = c2.y, r22 = c2.z;
    const T trace = r00 + r11 + r22;
    if (trace > T(1e-12)) {
        const T s = T(0.5) * rsqrt_t(trace + T(1));
        w = T(0.25) / s;
        x = (r21 - r12) * s;
        y = (r02 - r20) * s;
        z = (r10 - r01) * s;
    } else if (r00 > r11 && r00 > r22) {
        const T s = T(2) * sqrt_t(T(1) + r00 - r11 - r22);
        w = (r21 - r12) / s; x = T(0.25) * s; y = (r01 + r10) / s; z = (r02 + r20) / s;
    } else if (r11 > r22) {
        const T s = T(2) * sqrt_t(T(1) + r11 - r00 - r22);
        w = (r02 - r20) / s; x = (r01 + r10) / s; y = T(0.25) * s; z = (r12 + r21) / s;
    } else {
        const T s = T(2) * sqrt_t(T(1) + r22 - r00 - r11);
        w = (r10 - r01) / s; x = (r02 + r20) / s; y = (r12 + r21) / s; z = T(0.25) * s;
    }
}

// KDL Rotation::Quaternion(x, y, z, w): no normalisation.  Output columns.
template <typename T>
RLITG_DEV void quat_to_rot_kdl(T x, T y, T z, T w, V3<T>& c0, V3<T>& c1, V3<T>& c2) {
    const T x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    c0 = {w2 + x2 - y2 - z2, T(2) * x * y + T(2) * w * z, T(2) * x * z - T(2) * w * y};
    c1 = {T(2) * x * y - T(2) * w * z, w2 - x2 + y2 - z2, T(2) * y * z + T(2) * w * x};
    c2 = {T(2) * x * z + T(2) * w * y, T(2) * y * z - T(2) * w * x, w2 - x2 - y2 + z2};
}

// KDL diff(R_a, R_b) = R_a * (R_a^T R_b).GetRot(), GetRot as restated in oracle/rlitg_oracle.c
// (orocos_kdl 1.4 branch structure, angle by atan2).  Returns the rotation vector in the base frame and its
// norm (= the angle).
template <typename T>
RLITG_DEV V3<T> rot_diff_kdl(V3<T> a0, V3<T> a1, V3<T> a2, V3<T> b0, V3<T> b1, V3<T> b2, T& angle) {
    // rel = A^T B, rel(i,j) = dot(a_i, b_j)
    const T m00 = dot(a0, b0), m01 = dot(a0, b1), m02 = dot(a0, b2);
    const T m10 = dot(a1, b0), m11 = dot(a1, b1), m12 = dot(a1, b2);
    const T m20 = dot(a2, b0), m21 = dot(a2, b1), m22 = dot(a2, b2);
    T x = m21 - m12, y = m02 - m20, z = m10 - m01;
    V3<T> rv;
    const T eps = T(1e-6), eps2 = T(1e-5);
    if (abs_t(x) < eps && abs_t(y) < eps && abs_t(z) < eps) {
        if (abs_t(m01 + m10) < eps2 && abs_t(m02 + m20) < eps2 && abs_t(m12 + m21) < eps2 &&
            abs_t(m00 + m11 + m22 - T(3)) < eps2) {
            angle = T(0);
            return {T(0), T(0), T(0)};
        }
        const T xx = (m00 + T(1)) / T(2), yy = (m11 + T(1)) / T(2), zz = (m22 + T(1)) / T(2);
        const T xy = (m01 + m10) / T(4), xz = (m02 + m20) / T(4), yz = (m12 + m21) / T(4);
        if (xx > yy && xx > zz) { x = sqrt_t(xx); y = xy / x; z = xz / x; }
        else if (yy > zz) { y = sqrt_t(yy); x = xy / y; z = yz / y; }
        else { z = sqrt_t(zz); x = xz / z; y = yz / z; }
        angle = T(3.14159265358979323846);
        rv = {x * angle, y * angle, z * angle};
    } else {
        const T f = (m00 + m11 + m22 - T(1)) / T(2);
        const T r = sqrt_t(x * x + y * y + z * z);
        angle = atan2_t(r / T(2), f);
        const T sc = angle / r;
        rv = {x * sc, y * sc, z * sc};
    }
    // express in the base frame: A * rv
    return {a0.x * rv.x + a1.x * rv.y + a2.x * rv.z, a0.y * rv.x + a1.y * rv.y + a2.y * rv.z,
            a0.z * rv.x + a1.z * rv.y + a2.z * rv.z};
}

// ---- streaming accesses: data touched once per launch should not age in L2 like ordinary lines (evict-first) -----
template <typename U>
RLITG_DEV void st_stream(U* p, U v) { __stcs(p, v); }

// ---- TMA (bulk async copy) helpers: shared::cta -> global, used to drain staged output tiles --------------
RLITG_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
RLITG_DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// The stores carry an L2 evict-first policy: the outputs are written once and never re-read by the kernel, and letting
// them age in the 126 MB L2 like ordinary lines cost 8 % of the fk_jac step (45.3 -> 41.7 us at 1M configs, measured;
// evict_last was neutral).
RLITG_DEV void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
                 : "memory");
}
RLITG_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
RLITG_DEV void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

}  // namespace rlitg
