// Small dense algebra for the pseudo-inverse steps of the hot path.
//   * FP32 build: modified Gram-Schmidt QR of J^T (D x 6) held in registers.  pinv(J) v = Q R^-T v and the
//     null-space projector I - Q Q^T come out with error ~ cond(J) * eps -- not cond^2 as the normal equations
//     J J^T would give in FP32 -- and without any FP64 instruction on the hot path.
//   * FP64 build (verification): the literal algorithms -- one-sided Jacobi SVD with KDL's singular-value
//     truncation (ChainIkSolverVel_pinv, eps 1e-5, torm/src/torm_ik_solver.cpp:169-171) for the velocity step,
//     a 6x6 Cholesky solve of J J^T for the null-space norm (pirl/src/kinematic_solver.cpp:347).
//   * The FP32 build falls back to the FP64 SVD when a diagonal entry of R says a singular value may be near
//     KDL's truncation threshold, so the truncated-SVD semantics hold at singular configurations.
#pragma once
#include "rlitg_common.cuh"

namespace rlitg {

// A = J J^T (lower triangle, packed row-major: A[i*(i+1)/2 + j], j <= i) for J [6][D]
template <typename T, int D>
RLITG_DEV void jjt6(const T (&J)[6 * D], double (&A)[21]) {
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s += (double)J[i * D + k] * (double)J[j * D + k];
            A[i * (i + 1) / 2 + j] = s;
        }
}

// In-place Cholesky A = L L^T (packed lower).  Returns the smallest pivot ratio min_i(d_i) / max_i(A_ii) where
// d_i is the i-th pivot before the square root: ~ (sigma_min / sigma_max)^2 scale, used as the rank test.
RLITG_DEV double chol6(double (&A)[21]) {
    double amax = 0.0, dmin = 1e300;
#pragma unroll
    for (int i = 0; i < 6; i++) amax = fmax(amax, A[i * (i + 1) / 2 + i]);
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= A[j * (j + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
        dmin = fmin(dmin, d);
        const double l = sqrt(fmax(d, 1e-300));
        A[j * (j + 1) / 2 + j] = l;
        const double il = 1.0 / l;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = A[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= A[i * (i + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
            A[i * (i + 1) / 2 + j] = s * il;
        }
    }
    return dmin / fmax(amax, 1e-300);
}

RLITG_DEV void chol6_solve(const double (&L)[21], double (&b)[6]) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[i * (i + 1) / 2 + k] * b[k];
        b[i] = s / L[i * (i + 1) / 2 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = b[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= L[k * (k + 1) / 2 + i] * b[k];
        b[i] = s / L[i * (i + 1) / 2 + i];
    }
}


// Modified Gram-Schmidt on the rows of Q (in: the 6 rows of J, each of length D; out: orthonormal rows).
// R is upper triangular, packed R[k][j] (j >= k) at R[k * 6 + j].  Returns min_k R_kk / max_k R_kk.
template <typename T, int D>
RLITG_DEV T mgs_rows(T (&Q)[6 * D], T (&R)[36]) {
    T rmin = T(1e30), rmax = T(0);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        T n2 = T(0);
#pragma unroll
        for (int i = 0; i < D; i++) n2 += Q[k * D + i] * Q[k * D + i];
        const T r = sqrt_t(n2);
        R[k * 6 + k] = r;
        rmin = r < rmin ? r : rmin;
        rmax = r > rmax ? r : rmax;
        const T inv = T(1) / (r > T(1e-30) ? r : T(1e-30));
#pragma unroll
        for (int i = 0; i < D; i++) Q[k * D + i] *= inv;
#pragma unroll
        for (int j = k + 1; j < 6; j++) {
            T d = T(0);
#pragma unroll
            for (int i = 0; i < D; i++) d += Q[k * D + i] * Q[j * D + i];
            R[k * 6 + j] = d;
#pragma unroll
            for (int i = 0; i < D; i++) Q[j * D + i] -= d * Q[k * D + i];
        }
    }
    return rmin / (rmax > T(1e-30) ? rmax : T(1e-30));
}

// dq = pinv(J) v with singular values below eps zeroed: one-sided (Hestenes) Jacobi SVD of A = J^T (D x 6).
// (A V = U S  =>  pinv(J) = A S^-2 V^T.)  Every loop has compile-time bounds, so A and V live in registers; the
// function is kept out of line because it is the rare path of the FP32 build (a warp that takes it must not
// stretch the kernel's tail: ~6 sweeps of 15 rotations in FP32).  Arithmetic in T.
template <typename T, int D>
__device__ __noinline__ void pinv_svd_apply(const T* Jin, const T* v6, double eps, T* dq) {
    T A[D][6], V[6][6];
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int c = 0; c < 6; c++) A[i][c] = Jin[c * D + i];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int c = 0; c < 6; c++) V[i][c] = (i == c) ? T(1) : T(0);
    const T tol = sizeof(T) == 4 ? T(1e-7) : T(1e-15);
    for (int sweep = 0; sweep < 40; sweep++) {
        T off = T(0);
#pragma unroll
        for (int p = 0; p < 5; p++) {
#pragma unroll
            for (int r = p + 1; r < 6; r++) {
                T alpha = T(0), beta = T(0), gamma = T(0);
#pragma unroll
                for (int i = 0; i < D; i++) {
                    alpha += A[i][p] * A[i][p];
                    beta += A[i][r] * A[i][r];
                    gamma += A[i][p] * A[i][r];
                }
                const T ab = alpha * beta;
                const T lim = abs_t(gamma) * rsqrt_t(ab > T(0) ? ab : T(1e-30));
                if (gamma != T(0)) {
                    off = lim > off ? lim : off;
                    const T zeta = (beta - alpha) / (T(2) * gamma);
                    const T tt = (zeta >= T(0) ? T(1) : T(-1)) / (abs_t(zeta) + sqrt_t(T(1) + zeta * zeta));
                    const T c = rsqrt_t(T(1) + tt * tt), sn = c * tt;
#pragma unroll
                    for (int i = 0; i < D; i++) {
                        const T ap = A[i][p], ar = A[i][r];
                        A[i][p] = c * ap - sn * ar;
                        A[i][r] = sn * ap + c * ar;
                    }
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        const T vp = V[i][p], vr = V[i][r];
                        V[i][p] = c * vp - sn * vr;
                        V[i][r] = sn * vp + c * vr;
                    }
                }
            }
        }
        if (off < tol) break;
    }
    T acc[D];
#pragma unroll
    for (int i = 0; i < D; i++) acc[i] = T(0);
#pragma unroll
    for (int c = 0; c < 6; c++) {
        T s2 = T(0);
#pragma unroll
        for (int i = 0; i < D; i++) s2 += A[i][c] * A[i][c];
        T proj = T(0);
#pragma unroll
        for (int r = 0; r < 6; r++) proj += V[r][c] * v6[r];
        const T w = (sqrt_t(s2) < T(eps)) ? T(0) : proj / s2;  // truncated SVD: singular values below eps dropped
#pragma unroll
        for (int i = 0; i < D; i++) acc[i] += A[i][c] * w;
    }
#pragma unroll
    for (int i = 0; i < D; i++) dq[i] = acc[i];
}

// dq = pinv(J) v.  FP32: min-norm solution through the QR of J^T (J = R^T Q, so dq = Q^T-rows * y with
// R^T y = v); equal to the Moore-Penrose result for full row rank.  When the smallest diagonal of R is below
// 1e-4 of the largest (a singular value may be near KDL's 1e-5 truncation) the truncated FP64 SVD runs instead.
// FP64 (kAlwaysSvd): the literal SVD path.
template <typename T, int D, bool kAlwaysSvd>
RLITG_DEV void pinv_apply(const T (&J)[6 * D], const T (&v6)[6], double eps, T (&dq)[D]) {
    if constexpr (kAlwaysSvd) {
        T Jc[6 * D], vc[6];
#pragma unroll
        for (int i = 0; i < 6 * D; i++) Jc[i] = J[i];
#pragma unroll
        for (int i = 0; i < 6; i++) vc[i] = v6[i];
        pinv_svd_apply<T, D>(Jc, vc, eps, dq);
    } else {
        T Q[6 * D], R[36];
#pragma unroll
        for (int i = 0; i < 6 * D; i++) Q[i] = J[i];
        const T ratio = mgs_rows<T, D>(Q, R);
        if (ratio < T(1e-4)) {  // rare: copies keep J / v6 out of local memory on the common path
            T Jc[6 * D], vc[6], out[D];
#pragma unroll
            for (int i = 0; i < 6 * D; i++) Jc[i] = J[i];
#pragma unroll
            for (int i = 0; i < 6; i++) vc[i] = v6[i];
            pinv_svd_apply<T, D>(Jc, vc, eps, out);
#pragma unroll
            for (int i = 0; i < D; i++) dq[i] = out[i];
            return;
        }
        T y[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            T sacc = v6[k];
#pragma unroll
            for (int i = 0; i < k; i++) sacc -= R[i * 6 + k] * y[i];
            y[k] = sacc / R[k * 6 + k];
        }
#pragma unroll
        for (int i = 0; i < D; i++) {
            T sacc = T(0);
#pragma unroll
            for (int k = 0; k < 6; k++) sacc += y[k] * Q[k * D + i];
            dq[i] = sacc;
        }
    }
}

// Streaming form of the same QR solve for register-starved callers (kernel 3a): modified Gram-Schmidt over the rows of
// J (in Q, destroyed) with the forward substitution R^T y = v and the accumulation dq = sum_k y_k q_k folded into the
// sweep, so neither R nor the finished rows stay live (peak: the 6 x D rows + 13 scalars instead of 2 x 6 x D + 36).
// Returns min_k R_kk / max_k R_kk; the caller takes the truncated-SVD path when it is below 1e-4 (dq is then discarded).
template <typename T, int D>
RLITG_DEV T pinv_qr_stream(T (&Q)[6 * D], T (&v)[6], T (&dq)[D]) {
    T rmin = T(1e30), rmax = T(0);
#pragma unroll
    for (int i = 0; i < D; i++) dq[i] = T(0);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        T n2 = T(0);
#pragma unroll
        for (int i = 0; i < D; i++) n2 += Q[k * D + i] * Q[k * D + i];
        const T r = sqrt_t(n2);
        rmin = r < rmin ? r : rmin;
        rmax = r > rmax ? r : rmax;
        const T inv = T(1) / (r > T(1e-30) ? r : T(1e-30));
        const T y = v[k] * inv;
#pragma unroll
        for (int i = 0; i < D; i++) {
            Q[k * D + i] *= inv;
            dq[i] += y * Q[k * D + i];
        }
#pragma unroll
        for (int j = k + 1; j < 6; j++) {
            T d = T(0);
#pragma unroll
            for (int i = 0; i < D; i++) d += Q[k * D + i] * Q[j * D + i];
#pragma unroll
            for (int i = 0; i < D; i++) Q[j * D + i] -= d * Q[k * D + i];
            v[j] -= d * y;
        }
    }
    return rmin / (rmax > T(1e-30) ? rmax : T(1e-30));
}

// || (I - J^T (J J^T)^-1 J) e ||   (jaco_srv, pirl/src/kinematic_solver.cpp:332-351)
template <typename T, int D>
RLITG_DEV T nullspace_norm(const T (&J)[6 * D], const T (&e)[D]) {
    if constexpr (sizeof(T) == 4) {
        // e - Q^T Q e with the orthonormal rows of the QR of J^T.  Streaming sweep: as soon as row k is orthonormal it is
        // projected out of the residual and of the rows behind it and is dead -- no R, no finished rows stay live
        T Q[6 * D], r[D];
#pragma unroll
        for (int i = 0; i < 6 * D; i++) Q[i] = J[i];
#pragma unroll
        for (int i = 0; i < D; i++) r[i] = e[i];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            T n2 = T(0);
#pragma unroll
            for (int i = 0; i < D; i++) n2 += Q[k * D + i] * Q[k * D + i];
            const T rn = sqrt_t(n2);
            const T inv = T(1) / (rn > T(1e-30) ? rn : T(1e-30));
            T c = T(0);
#pragma unroll
            for (int i = 0; i < D; i++) { Q[k * D + i] *= inv; c += Q[k * D + i] * r[i]; }
#pragma unroll
            for (int i = 0; i < D; i++) r[i] -= c * Q[k * D + i];
#pragma unroll
            for (int j = k + 1; j < 6; j++) {
                T d = T(0);
#pragma unroll
                for (int i = 0; i < D; i++) d += Q[k * D + i] * Q[j * D + i];
#pragma unroll
                for (int i = 0; i < D; i++) Q[j * D + i] -= d * Q[k * D + i];
            }
        }
        T n2 = T(0);
#pragma unroll
        for (int i = 0; i < D; i++) n2 += r[i] * r[i];
        return sqrt_t(n2);
    } else {
        double A[21];
        jjt6<T, D>(J, A);
        chol6(A);
        double y[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s += (double)J[i * D + k] * (double)e[k];
            y[i] = s;
        }
        chol6_solve(A, y);
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) {
            double s = (double)e[k];
#pragma unroll
            for (int i = 0; i < 6; i++) s -= (double)J[i * D + k] * y[i];
            n2 += s * s;
        }
        return (T)sqrt(n2);
    }
}

}  // namespace rlitg
