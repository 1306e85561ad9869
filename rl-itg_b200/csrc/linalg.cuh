// Small dense algebra for the pseudo-inverse steps of the hot path.
//   * 6x6 SPD solve of (J J^T) y = b by Cholesky -- always in FP64: the normal equations square cond(J), and
//     FP32 there would cost ~cond^2 * 6e-8; 6x6 in FP64 is ~300 DFMA per configuration.
//   * one-sided Jacobi SVD pseudo-inverse with KDL's singular-value truncation (ChainIkSolverVel_pinv, eps 1e-5,
//     torm/src/torm_ik_solver.cpp:169-171), used when the Cholesky pivots say J is near rank-deficient and by the
//     FP64 verification build.
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

// dq = pinv(J) v with singular values below eps zeroed: one-sided Jacobi on A = J^T (D x 6), FP64.
// (A V = U S  =>  pinv(J) = A S^-2 V^T.)  Not inlined: it is the rare path of the FP32 build.
template <typename T, int D>
__device__ __noinline__ void pinv_svd_apply(const T* J, const T* v6, double eps, T* dq) {
    double A[D][6], V[6][6];
    for (int i = 0; i < D; i++)
        for (int c = 0; c < 6; c++) A[i][c] = (double)J[c * D + i];
    for (int i = 0; i < 6; i++)
        for (int c = 0; c < 6; c++) V[i][c] = (i == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 5; p++)
            for (int r = p + 1; r < 6; r++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < D; i++) {
                    alpha += A[i][p] * A[i][p];
                    beta += A[i][r] * A[i][r];
                    gamma += A[i][p] * A[i][r];
                }
                if (gamma == 0.0) continue;
                const double ab = alpha * beta;
                const double lim = fabs(gamma) / sqrt(ab > 0 ? ab : 1e-300);
                off = fmax(off, lim);
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
                for (int i = 0; i < D; i++) {
                    const double ap = A[i][p], ar = A[i][r];
                    A[i][p] = c * ap - s * ar;
                    A[i][r] = s * ap + c * ar;
                }
                for (int i = 0; i < 6; i++) {
                    const double vp = V[i][p], vr = V[i][r];
                    V[i][p] = c * vp - s * vr;
                    V[i][r] = s * vp + c * vr;
                }
            }
        if (off < 1e-15) break;
    }
    double acc[D];
    for (int i = 0; i < D; i++) acc[i] = 0.0;
    for (int c = 0; c < 6; c++) {
        double s2 = 0;
        for (int i = 0; i < D; i++) s2 += A[i][c] * A[i][c];
        if (sqrt(s2) < eps) continue;
        double proj = 0;
        for (int r = 0; r < 6; r++) proj += V[r][c] * (double)v6[r];
        for (int i = 0; i < D; i++) acc[i] += A[i][c] * (proj / s2);
    }
    for (int i = 0; i < D; i++) dq[i] = (T)acc[i];
}

// dq = pinv(J) v.  Fast path J^T (J J^T)^-1 v (equal to the Moore-Penrose result for full row rank); falls back
// to the truncated SVD when the smallest Cholesky pivot ratio says a singular value may be near KDL's eps.
// kAlwaysSvd = literal path for the verification build.
template <typename T, int D, bool kAlwaysSvd>
RLITG_DEV void pinv_apply(const T (&J)[6 * D], const T (&v6)[6], double eps, T (&dq)[D]) {
    if constexpr (kAlwaysSvd) {
        pinv_svd_apply<T, D>(J, v6, eps, dq);
    } else {
        double A[21];
        jjt6<T, D>(J, A);
        const double ratio = chol6(A);
        // sigma_min^2 / sigma_max^2 >~ ratio; KDL truncates sigma < 1e-5 with sigma_max = O(1): be conservative
        if (ratio < 1e-8) {
            pinv_svd_apply<T, D>(J, v6, eps, dq);
            return;
        }
        double y[6];
#pragma unroll
        for (int i = 0; i < 6; i++) y[i] = (double)v6[i];
        chol6_solve(A, y);
#pragma unroll
        for (int k = 0; k < D; k++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 6; i++) s += (double)J[i * D + k] * y[i];
            dq[k] = (T)s;
        }
    }
}

// || (I - J^T (J J^T)^-1 J) e ||   (jaco_srv, pirl/src/kinematic_solver.cpp:332-351)
template <typename T, int D>
RLITG_DEV T nullspace_norm(const T (&J)[6 * D], const T (&e)[D]) {
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

}  // namespace rlitg
