// TORM update step on the device: what localOptimizeTSGD does between two cost evaluations,
//   calculateFeasibleIncrements   torm/src/torm_optimizer.cpp:1555-1609   max-abs normalisation, alpha blend,
//                                                                         A^-1 preconditioning, learning rate
//   calculateEndPoseIncrements    :1475-1482                              (weight already applied by kernel 3a)
//   addIncrementsToTrajectory     :1611-1639                              joint_update_limit_ scaling
//   handleJointLimits             :1189-1268                              <= 11 corrective A^-1 columns per joint
// for N trajectories at once, so the [N,T,7] tensor never leaves the GPU between iterations.
// updateGoalConfiguration / updateFullTrajectory (:1118-1123, :774-777) are implicit: padded rows are clamped reads.
//
// One CTA per trajectory.  The preconditioning A_free^-1 b is a banded Cholesky solve (A_free is SPD with half
// bandwidth 6) instead of the reference's dense (T-1)^2 product: O(13 T) per joint, in FP64 (7 serial chains per
// trajectory; the kernel is latency bound and tiny next to kernel 3).  The corrective columns of handleJointLimits
// are rows of the dense scaled inverse kept on the device (symmetric, so a column read is contiguous).
#include <mutex>

#include "rlitg_internal.h"

namespace rlitg {

struct UpdP {
    int Tn, stage, mode;
    double sw, ow, limit, mcs;
    int cj[7];
    double jmin[7], jmax[7];
};

struct SpikeP {   // partitioned substitution (smooth_build): blocks of `rows` rows, spike [2][F][6]
    int parts, rows;
    const double* spike;
};

constexpr int kUpdGroup = 64;  // threads that work on one trajectory

// barrier over the 64 threads of one trajectory's group (hardware barrier 1 + group; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kUpdGroup) : "memory");
}

__device__ __forceinline__ double group_max(double v, double* s_tmp, int g, int gt) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    group_sync(g);
    if ((gt & 31) == 0) s_tmp[gt >> 5] = v;
    group_sync(g);
    return fmax(s_tmp[0], s_tmp[1]);
}

// PER trajectories per CTA, 64 threads each.  Everything but the banded solve is done by a trajectory's own group
// with group barriers.  The solve is one thread per (trajectory, joint) and bound by the FP64 pipe's issue rate once
// several CTAs share an SM (a warp-wide FP64 instruction costs the same for 7 active lanes as for 32), so the
// 7 x PER solves of a CTA are packed into the lanes of warp 0: PER = 4 -> 28 lanes per FP64 instruction.
template <typename T, int PER>
__global__ void __launch_bounds__(kUpdGroup * PER) __maxnreg__(PER > 0 ? 96 : 0)
torm_update_kernel(UpdP P, T* __restrict__ traj, int64_t N, const T* __restrict__ inc_smooth,
                   const T* __restrict__ inc_coll, const T* __restrict__ inc_pose, const T* __restrict__ alpha,
                   const T* __restrict__ lr, const double* __restrict__ chol, const double* __restrict__ inv,
                   const int32_t* __restrict__ stage_idx, int tab_in_smem, SpikeP SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tn = P.Tn, F = Tn - 1, tid = threadIdx.x;
    const int g = tid / kUpdGroup, gt = tid - g * kUpdGroup;
    const size_t FD = (size_t)F * 7;
    const size_t FDp = FD + (FD & 1);  // even number of doubles: every region stays 16-byte aligned
    // shared memory: PER x { fin [F][7] right-hand side -> final increments, tr [F][7] free rows of the trajectory }
    // then ONE substitution table [F][8] at a time (tab_in_smem): every step of the solve reads one 64-byte row, and a
    // first-touch L2 round trip per row (a 128-byte line = two rows) was most of the solve's latency.
    // fin carries 8 zero rows in front and behind, so that the solve's look-ahead reads need no bounds test
    constexpr int kPad = 8 * 7;
    const size_t stride = FDp + 2 * kPad + FDp;  // doubles per trajectory: padded fin, then tr
    double* base = reinterpret_cast<double*>(smem_raw);
    double* fin = base + (size_t)g * stride + kPad;
    double* tr = fin + FDp + kPad;
    double4* s_tab = reinterpret_cast<double4*>(base + (size_t)PER * stride);
    __shared__ double s_tmp_all[PER][2];
    __shared__ double s_best_all[PER][kUpdGroup];
    __shared__ int s_idx_all[PER][kUpdGroup];
    __shared__ int s_stage[PER];
    __shared__ double s_wb[8][36];  // partitioned substitution: the 6 x 6 boundary rows of every inner block
    double* s_tmp = s_tmp_all[g];
    double* s_best = s_best_all[g];
    int* s_idx = s_idx_all[g];

    const int64_t n = (int64_t)blockIdx.x * PER + g;
    // per-trajectory stage (lr_schedule_mode_ 1 lets every optimiser pick its own, torm_optimizer.cpp:400-412);
    // a negative entry leaves the trajectory untouched
    const int stage = n < N ? (stage_idx ? stage_idx[n] : P.stage) : -1;
    if (gt == 0) s_stage[g] = stage;
    const bool active = stage >= 0;
    T* trow = traj + (active ? n : 0) * (int64_t)Tn * 7 + 7;  // first free row
    const double4* g_tab = reinterpret_cast<const double4*>(chol);  // [2][F] rows of 2 x double4
    if (tab_in_smem)
        for (int k = tid; k < 2 * F; k += kUpdGroup * PER) s_tab[k] = g_tab[k];  // forward table

    if (gt < kPad) { fin[gt - kPad] = 0.0; fin[FD + gt] = 0.0; }
    if (active) {
        for (int k = gt; k < (int)FD; k += kUpdGroup) tr[k] = (double)trow[k];
        if (stage == 0) {
            const T* s = inc_smooth + n * FD;
            const T* c = inc_coll + n * FD;
            double ss = 1.0, sc = 1.0;
            if (P.mode != 0) {
                double ms = 0.0, mc = 0.0;
                for (int k = gt; k < (int)FD; k += kUpdGroup) { ms = fmax(ms, fabs((double)s[k])); mc = fmax(mc, fabs((double)c[k])); }
                ss = group_max(ms, s_tmp, g, gt);
                sc = group_max(mc, s_tmp, g, gt);
            }
            const double a = alpha ? (double)alpha[n] : 0.5;
            // plain FP64 issues at 1/16 of the FP32 rate here and a division is ~30 such instructions: one reciprocal
            // per thread instead of two divisions per element
            const double ws = P.mode != 0 ? (1 - a) * (ss > 0 ? 1.0 / ss : 1.0) : P.sw;
            const double wc = P.mode != 0 ? a * (sc > 0 ? 1.0 / sc : 1.0) : P.ow;
            for (int k = gt; k < (int)FD; k += kUpdGroup) fin[k] = wc * (double)c[k] + ws * (double)s[k];
        } else {
            const T* p = inc_pose + n * FD;
            for (int k = gt; k < (int)FD; k += kUpdGroup) fin[k] = (double)p[k];
        }
    }
    __syncthreads();

    // ---- banded Cholesky solve per (trajectory, joint) column: L y = b, then L^T x = y.  The recurrence is the
    // critical path (an FP64 FMA has ~40 cycles of latency here), so the substitution is column oriented: six
    // running rows live in registers, already divided by their diagonal; once unknown i is final it is pushed into
    // all six with independent FMAs and the first of them IS unknown i+1 -- one dependent FMA per step, no division.
    // Coefficient rows (smooth_build) are fetched one step ahead.
    // Partitioned (SP.parts > 1): the F rows are cut into blocks of SP.rows rows; warp w runs the recurrence of block w
    // with zero history (all blocks at once: the chain is F / parts steps long instead of F), then warp 0 walks the
    // block boundaries -- the true last six values of block p are its zero-history values plus W (six values in front
    // of the block), W = spike table of smooth_build -- and finally every warp adds W (boundary) to the rest of its block.
    // The same, mirrored, for L^T x = y.  One block = the plain sequential substitution.
    const int lane = tid & 31, wp = tid >> 5;
    const int sg = lane / 7, sj = lane - sg * 7;   // solve lane -> (trajectory of this CTA, joint)
    const bool col_ok = lane < 7 * PER && s_stage[sg < PER ? sg : 0] == 0;
    const int NP = SP.parts, R = SP.rows;
    const bool solver = col_ok && wp < NP;
    double* sfin = base + (size_t)(sg < PER ? sg : 0) * stride + kPad + sj;  // column sj of that trajectory, row stride 7
    const int s0 = wp * R, e0 = min(s0 + R, F);    // this warp's block
    // (rows -8..-1 and F..F+7 of sfin are zero; table rows outside 0..F-1 are never needed: their products are with
    //  those zeros, so the row index is simply clamped)
    if (solver) {
        const double4* tab = tab_in_smem ? s_tab : g_tab;  // one row = 2 x double4
        // ---- forward: rows s0..s0+6 start as b_r / L(r,r); a0 is the row being finalised.  Rows behind the block
        // enter as zeros: they belong to the next warp, which overwrites them while this one runs
        double a[7];
#pragma unroll
        for (int r = 0; r < 7; r++) a[r] = (s0 + r < e0 ? sfin[(s0 + r) * 7] : 0.0) * tab[2 * min(s0 + r, F - 1) + 1].w;
        double4 ca = tab[2 * s0], cb = tab[2 * s0 + 1];
        double nb_rhs = s0 + 7 < e0 ? sfin[(s0 + 7) * 7] : 0.0;
        for (int i = s0; i < e0; i++) {
            const int in = min(i + 1, F - 1);
            const double4 na = tab[2 * in], nb = tab[2 * in + 1];
            const double nrhs = i + 8 < e0 ? sfin[(i + 8) * 7] : 0.0;
            const double y = a[0];
            sfin[i * 7] = y;
            a[0] = a[1] - ca.x * y; a[1] = a[2] - ca.y * y; a[2] = a[3] - ca.z * y; a[3] = a[4] - ca.w * y;
            a[4] = a[5] - cb.x * y; a[5] = a[6] - cb.y * y;
            a[6] = nb_rhs * cb.z;  // row i+7 enters the window: b / L(i+7,i+7)
            ca = na; cb = nb; nb_rhs = nrhs;
        }
    }
    __syncthreads();
    if (tab_in_smem)
        for (int k = tid; k < 2 * F; k += kUpdGroup * PER) s_tab[k] = g_tab[2 * F + k];  // backward table
    if (NP > 1) {
        const double* Wf = SP.spike;  // [F][6]
        // the 6 x 6 spike rows warp 0 needs for block wp's boundary, fetched by block wp's own warp (all blocks at once)
        if (wp >= 1 && wp < NP - 1) {
            for (int k = lane; k < 36; k += 32) s_wb[wp][k] = __ldg(Wf + (size_t)(e0 - 1 - k / 6) * 6 + (k % 6));
        }
        __syncthreads();
        if (wp == 0 && col_ok) {
            // boundaries, block by block: the last six rows of block p become final here (block 0 is final already)
            for (int p = 1; p < NP - 1; p++) {
                const int pe = min(p * R + R, F);
                double gb[6], t[6];
#pragma unroll
                for (int m = 0; m < 6; m++) gb[m] = sfin[(p * R - 1 - m) * 7];
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    const double* w = &s_wb[p][q * 6];
                    t[q] = sfin[(pe - 1 - q) * 7] + ((w[0] * gb[0] + w[1] * gb[1]) + (w[2] * gb[2] + w[3] * gb[3]) + (w[4] * gb[4] + w[5] * gb[5]));
                }
#pragma unroll
                for (int q = 0; q < 6; q++) sfin[(pe - 1 - q) * 7] = t[q];
            }
        }
        __syncthreads();
        if (solver && wp >= 1) {
            double gb[6];
#pragma unroll
            for (int m = 0; m < 6; m++) gb[m] = sfin[(s0 - 1 - m) * 7];
            const int ce = wp < NP - 1 ? e0 - 6 : e0;  // the last six rows of an inner block are final already
            const double2* w2 = reinterpret_cast<const double2*>(Wf + (size_t)s0 * 6);  // rows of 48 bytes
#pragma unroll 4
            for (int i = s0; i < ce; i++) {
                const double2 wa = __ldg(w2 + 3 * (i - s0)), wb2 = __ldg(w2 + 3 * (i - s0) + 1), wc2 = __ldg(w2 + 3 * (i - s0) + 2);
                sfin[i * 7] += (wa.x * gb[0] + wa.y * gb[1]) + (wb2.x * gb[2] + wb2.y * gb[3]) + (wc2.x * gb[4] + wc2.y * gb[5]);
            }
        }
    }
    __syncthreads();
    if (solver) {
        // ---- backward: rows e0-1 .. e0-7 start as y_r / L(r,r); rows in front of the block enter as zeros
        const double4* tbk = tab_in_smem ? s_tab : g_tab + 2 * (size_t)F;
        const int e = e0 - 1;
        double a[7];
#pragma unroll
        for (int r = 0; r < 7; r++) a[r] = (e - r >= s0 ? sfin[(e - r) * 7] : 0.0) * tbk[2 * max(e - r, 0) + 1].w;
        double4 ca = tbk[2 * e], cb = tbk[2 * e + 1];
        double nb_rhs = e - 7 >= s0 ? sfin[(e - 7) * 7] : 0.0;
        for (int i = e; i >= s0; i--) {
            const int in = max(i - 1, 0);
            const double4 na = tbk[2 * in], nb = tbk[2 * in + 1];
            const double nrhs = i - 8 >= s0 ? sfin[(i - 8) * 7] : 0.0;
            const double x = a[0];
            sfin[i * 7] = x;
            a[0] = a[1] - ca.x * x; a[1] = a[2] - ca.y * x; a[2] = a[3] - ca.z * x; a[3] = a[4] - ca.w * x;
            a[4] = a[5] - cb.x * x; a[5] = a[6] - cb.y * x;
            a[6] = nb_rhs * cb.z;  // row i-7 enters the window
            ca = na; cb = nb; nb_rhs = nrhs;
        }
    }
    __syncthreads();
    if (NP > 1) {
        const double* Wb = SP.spike + (size_t)F * 6;  // [F][6]
        if (wp >= 1 && wp < NP - 1) {
            for (int k = lane; k < 36; k += 32) s_wb[wp][k] = __ldg(Wb + (size_t)(s0 + k / 6) * 6 + (k % 6));
        }
        __syncthreads();
        if (wp == 0 && col_ok) {
            // boundaries from the back: the first six rows of block p become final (the last block is final already)
            for (int p = NP - 2; p >= 1; p--) {
                const int ps = p * R, pe = min(ps + R, F);
                double hb[6], t[6];
#pragma unroll
                for (int m = 0; m < 6; m++) hb[m] = sfin[(pe + m) * 7];
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    const double* w = &s_wb[p][q * 6];
                    t[q] = sfin[(ps + q) * 7] + ((w[0] * hb[0] + w[1] * hb[1]) + (w[2] * hb[2] + w[3] * hb[3]) + (w[4] * hb[4] + w[5] * hb[5]));
                }
#pragma unroll
                for (int q = 0; q < 6; q++) sfin[(ps + q) * 7] = t[q];
            }
        }
        __syncthreads();
        if (solver && wp < NP - 1) {
            double hb[6];
#pragma unroll
            for (int m = 0; m < 6; m++) hb[m] = sfin[(e0 + m) * 7];
            const int cs = wp >= 1 ? s0 + 6 : s0;  // the first six rows of an inner block are final already
            const double2* w2 = reinterpret_cast<const double2*>(Wb + (size_t)cs * 6);
#pragma unroll 4
            for (int i = cs; i < e0; i++) {
                const double2 wa = __ldg(w2 + 3 * (i - cs)), wb2 = __ldg(w2 + 3 * (i - cs) + 1), wc2 = __ldg(w2 + 3 * (i - cs) + 2);
                sfin[i * 7] += (wa.x * hb[0] + wa.y * hb[1]) + (wb2.x * hb[2] + wb2.y * hb[3]) + (wc2.x * hb[4] + wc2.y * hb[5]);
            }
        }
        __syncthreads();
    }
    if (!active) return;  // no CTA-wide barrier below this line

    if (stage == 0) {
        const double rate = (double)lr[n] / P.mcs;
        for (int k = gt; k < (int)FD; k += kUpdGroup) fin[k] *= rate;
        group_sync(g);
    }

    // ---- addIncrementsToTrajectory :1611-1639
    if (P.mode != 0) {
        double m = 0.0;
        for (int k = gt; k < (int)FD; k += kUpdGroup) m = fmax(m, fabs(fin[k]));
        const double scale = group_max(m, s_tmp, g, gt);
        const double f = (scale > P.limit && scale > 0) ? P.limit / scale : 1.0;
        for (int k = gt; k < (int)FD; k += kUpdGroup) tr[k] += fin[k] * f;
    } else {
        for (int j = 0; j < 7; j++) {
            double mx = -1e300, mn = 1e300;
            for (int i = gt; i < F; i += kUpdGroup) { mx = fmax(mx, fin[i * 7 + j]); mn = fmin(mn, fin[i * 7 + j]); }
            mx = group_max(mx, s_tmp, g, gt);
            mn = -group_max(-mn, s_tmp, g, gt);
            double scale = 1.0;
            const double max_scale = P.limit / fabs(mx), min_scale = P.limit / fabs(mn);
            if (max_scale < scale) scale = max_scale;
            if (min_scale < scale) scale = min_scale;
            for (int i = gt; i < F; i += kUpdGroup) tr[i * 7 + j] += scale * fin[i * 7 + j];
        }
    }
    group_sync(g);

    // ---- handleJointLimits :1189-1268 (continuous joints are skipped)
    for (int j = 0; j < 7; j++) {
        if (P.cj[j]) continue;
        for (int count = 0; count <= 10; count++) {
            double best = 1e-6;
            int idx = -1;
            for (int i = gt; i < F; i += kUpdGroup) {
                const double v = tr[i * 7 + j];
                double amount = 0.0;
                if (v > P.jmax[j]) amount = P.jmax[j] - v;
                else if (v < P.jmin[j]) amount = P.jmin[j] - v;
                if (fabs(amount) > best) { best = fabs(amount); idx = i; }  // first maximum wins
            }
            // winner over the group: largest violation, smallest row index on ties (warp shuffles, then the two warps)
            double wb = idx >= 0 ? best : 0.0;
            int wi = idx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, wb, o);
                const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
                if (oi >= 0 && (wi < 0 || ob > wb || (ob == wb && oi < wi))) { wb = ob; wi = oi; }
            }
            if ((gt & 31) == 0) { s_best[gt >> 5] = wb; s_idx[gt >> 5] = wi; }
            group_sync(g);
            {
                const double b0 = s_best[0], b1 = s_best[1];
                const int i0 = s_idx[0], i1 = s_idx[1];
                wb = b0; wi = i0;
                if (i1 >= 0 && (wi < 0 || b1 > wb || (b1 == wb && i1 < wi))) { wb = b1; wi = i1; }
            }
            group_sync(g);
            if (wi < 0) break;
            const double v = tr[wi * 7 + j];
            const double viol = v > P.jmax[j] ? P.jmax[j] - v : P.jmin[j] - v;
            const double mult = viol / inv[(size_t)wi * F + wi];
            group_sync(g);
            for (int i = gt; i < F; i += kUpdGroup) tr[i * 7 + j] += mult * inv[(size_t)wi * F + i];
            group_sync(g);
        }
    }
    group_sync(g);
    for (int k = gt; k < (int)FD; k += kUpdGroup) trow[k] = (T)tr[k];
}

template <typename T, int PER>
static int launch_update_per(const UpdP& P, const Smooth* sm, T* traj, int64_t N, const T* inc_smooth, const T* inc_coll,
                             const T* inc_pose, const T* alpha, const T* lr, const int32_t* stage_idx, size_t smem,
                             int tab_in_smem, cudaStream_t st) {
    auto kern = torm_update_kernel<T, PER>;
    int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "torm_update_kernel");
    if (rc_attr) return rc_attr;
    const unsigned grid = (unsigned)((N + PER - 1) / PER);
    // one block of the partitioned substitution per warp; RLITG_UPD_PARTS=1 forces the sequential recurrence (measurement)
    static const int parts_env = [] { const char* e = getenv("RLITG_UPD_PARTS"); return e ? atoi(e) : 0; }();
    SpikeP SP;
    SP.parts = (parts_env == 1 || sm->parts > kUpdGroup * PER / 32) ? 1 : sm->parts;
    SP.rows = SP.parts > 1 ? sm->rows : P.Tn - 1;
    SP.spike = sm->d_spike;
    kern<<<grid, kUpdGroup * PER, smem, st>>>(P, traj, N, inc_smooth, inc_coll, inc_pose, alpha, lr, sm->d_chol,
                                              sm->d_inv64, stage_idx, tab_in_smem, SP);
    return check_launch("torm_update_kernel");
}

template <typename T>
static int launch_update(const RlitgTormUpdate* u, const RlitgTormParams* p, const Smooth* sm, T* traj, int64_t N,
                         const T* inc_smooth, const T* inc_coll, const T* inc_pose, const T* alpha, const T* lr,
                         const int32_t* stage_idx, cudaStream_t st) {
    UpdP P;
    P.Tn = p->T; P.stage = u->stage; P.mode = u->lr_schedule_mode;
    P.sw = u->smoothness_update_weight; P.ow = u->obstacle_update_weight; P.limit = u->joint_update_limit;
    P.mcs = sm->mcs;
    for (int j = 0; j < 7; j++) { P.cj[j] = p->cjoint[j]; P.jmin[j] = u->joint_min[j]; P.jmax[j] = u->joint_max[j]; }
    const size_t F = (size_t)(p->T - 1);
    const size_t FDp = F * 7 + ((F * 7) & 1);
    auto bytes = [&](int per, bool tab) { return sizeof(double) * ((size_t)per * (2 * FDp + 2 * 8 * 7) + (tab ? F * 8 : 0)); };
    // four trajectories per CTA (lane-packed solve) when some trajectory may take the feasible stage and two such CTAs
    // fit an SM; otherwise one per CTA (measured faster for the solve-free end-pose update: 29 vs 35 us per 1,024).
    // The substitution tables go to shared memory whenever they fit, longer paths read them through L1 / L2.
    const bool may_solve = stage_idx != nullptr || u->stage == 0;
    if (may_solve && N >= 4 && bytes(4, true) <= 110 * 1024)
        return launch_update_per<T, 4>(P, sm, traj, N, inc_smooth, inc_coll, inc_pose, alpha, lr, stage_idx, bytes(4, true), 1, st);
    const bool tab = may_solve && bytes(1, true) <= 200 * 1024;  // no solve, no table
    if (bytes(1, tab) > 200 * 1024) return set_error(RLITG_ERR_SHAPE, "T too large for the update step's shared-memory tile");
    return launch_update_per<T, 1>(P, sm, traj, N, inc_smooth, inc_coll, inc_pose, alpha, lr, stage_idx, bytes(1, tab), tab ? 1 : 0, st);
}

}  // namespace rlitg

using namespace rlitg;

static int torm_update_impl(void* traj, int64_t N, int dtype, const RlitgSmooth* smooth, const RlitgTormParams* p,
                            const RlitgTormUpdate* u, const int32_t* stage_idx, const void* inc_smooth,
                            const void* inc_coll, const void* inc_pose, const void* alpha, const void* lr,
                            void* stream) {
    if (!p || !u || !smooth) return set_error(RLITG_ERR_NULL, "params / update / smooth is NULL");
    if (p->D != 7) return set_error(RLITG_ERR_SHAPE, "TORM runs the 7-joint arm chain (D = 7)");
    if (N < 0 || N > 0x7fffffff) return set_error(RLITG_ERR_SHAPE, "N out of range");
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    if (!stage_idx && u->stage != 0 && u->stage != 1) return set_error(RLITG_ERR_ARG, "stage must be 0 (feasible) or 1 (end-pose)");
    const Smooth* sm = reinterpret_cast<const Smooth*>(smooth);
    if (sm->T != p->T) return set_error(RLITG_ERR_SHAPE, "smooth handle was built for a different T");
    if (N == 0) return RLITG_OK;
    if (!traj) return set_error(RLITG_ERR_NULL, "traj is NULL");
    const bool feas = stage_idx || u->stage == 0, endp = stage_idx || u->stage == 1;
    if (feas && (!inc_smooth || !inc_coll || !lr)) return set_error(RLITG_ERR_NULL, "feasible stage needs inc_smooth, inc_coll and lr");
    if (endp && !inc_pose) return set_error(RLITG_ERR_NULL, "end-pose stage needs inc_pose");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32)
        return launch_update<float>(u, p, sm, (float*)traj, N, (const float*)inc_smooth, (const float*)inc_coll,
                                    (const float*)inc_pose, (const float*)alpha, (const float*)lr, stage_idx, st);
    return launch_update<double>(u, p, sm, (double*)traj, N, (const double*)inc_smooth, (const double*)inc_coll,
                                 (const double*)inc_pose, (const double*)alpha, (const double*)lr, stage_idx, st);
}

extern "C" RLITG_API int rlitg_torm_update(void* traj, int64_t N, int dtype, const RlitgSmooth* smooth,
                                           const RlitgTormParams* p, const RlitgTormUpdate* u, const void* inc_smooth,
                                           const void* inc_coll, const void* inc_pose, const void* alpha,
                                           const void* lr, void* stream) {
    return torm_update_impl(traj, N, dtype, smooth, p, u, nullptr, inc_smooth, inc_coll, inc_pose, alpha, lr, stream);
}

extern "C" RLITG_API int rlitg_torm_update_staged(void* traj, int64_t N, int dtype, const RlitgSmooth* smooth,
                                                  const RlitgTormParams* p, const RlitgTormUpdate* u,
                                                  const int32_t* stage_idx, const void* inc_smooth,
                                                  const void* inc_coll, const void* inc_pose, const void* alpha,
                                                  const void* lr, void* stream) {
    if (!stage_idx) return set_error(RLITG_ERR_NULL, "stage_idx is NULL");
    return torm_update_impl(traj, N, dtype, smooth, p, u, stage_idx, inc_smooth, inc_coll, inc_pose, alpha, lr, stream);
}

// ---- per-trajectory optimiser state of localOptimizeTSGD (torm/src/torm_optimizer.cpp:400-560) --------------------
// One CTA per trajectory replaces the scalar host control flow of the reference between the cost evaluation and the
// update step: stage choice (lr_schedule_mode_ 1, :400-412, from the state BEFORE this sub-step), feasibility flags
// (:436-446), alpha_ (:447-469), learning_rate_ (:471-499) and the best-trajectory record (:506-560).
namespace rlitg {
struct BookP {
    int T, mode, stage, use_col, use_vel, hits_per_traj;
    double lr0, lr_min, lr_max;
};

template <typename T>
__global__ void __launch_bounds__(128)
torm_bookkeep_kernel(BookP P, const T* __restrict__ traj, T* __restrict__ best_traj, const T* __restrict__ costs,
                     const uint8_t* __restrict__ hits, const uint8_t* __restrict__ singular, T* __restrict__ alpha,
                     T* __restrict__ lr, uint8_t* __restrict__ infea_col, uint8_t* __restrict__ infea_vel,
                     T* __restrict__ best_cost, int32_t* __restrict__ stage_idx, const T* __restrict__ lr_iter) {
    const int64_t n = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ int s_hit;
    __shared__ int s_better;
    if (tid == 0) s_hit = 0;
    __syncthreads();
    if (P.use_col && hits) {
        int h = 0;
        if (P.hits_per_traj) {
            if (tid == 0) h = hits[n];
        } else {
            for (int i = tid; i < P.T; i += blockDim.x) h |= hits[n * P.T + i];
        }
        if (h) s_hit = 1;  // benign race: every writer stores 1
    }
    __syncthreads();
    if (tid == 0) {
        const T lr_old = lr[n];
        const bool col_old = infea_col[n] != 0, vel_old = infea_vel[n] != 0;
        if (stage_idx) {
            bool feasible_stage;
            if (P.stage == 0) feasible_stage = !(P.mode == 1 && lr_old < T(1e-2) && !col_old && !vel_old);
            else feasible_stage = (P.mode == 1 && lr_old == T(P.lr_max) && col_old);
            stage_idx[n] = feasible_stage ? 0 : 1;
        }
        const bool col = s_hit != 0;
        const bool vel = P.use_vel && costs[n * RLITG_TORM_NCOST + 7] != T(0);
        infea_col[n] = col; infea_vel[n] = vel;
        T a = alpha[n];
        if (P.mode == 1 || P.mode == 2) {
            if (col && !vel) a = a < T(0.5) ? T(0.5) : (a + T(0.1) < T(1) ? a + T(0.1) : T(1));
            else if (!col && vel) a = a > T(0.5) ? T(0.5) : (a - T(0.1) > T(0) ? a - T(0.1) : T(0));
            else a = T(0.5);
        } else if (P.mode == 3) {
            a = T(0.5);
        }
        alpha[n] = a;
        T l = lr_old;
        if (P.mode == 1) {
            if (!col && !vel) l = l > T(P.lr0) ? T(P.lr0) : (l / T(2) > T(P.lr_min) ? l / T(2) : T(P.lr_min));
            else l = l * T(2) < T(P.lr_max) ? l * T(2) : T(P.lr_max);
        } else if (P.mode == 2) {
            l = lr_iter ? lr_iter[0] : T(P.lr0);
        } else if (P.mode == 3) {
            l = T(P.lr0);
        }
        lr[n] = l;
        const T cur = costs[n * RLITG_TORM_NCOST + 6];
        const bool ok = !col && !vel && !(singular && singular[n]);
        const bool better = ok && cur < best_cost[n];
        if (better) best_cost[n] = cur;
        s_better = better;
    }
    __syncthreads();
    if (s_better) {
        const int64_t off = n * (int64_t)P.T * 7;
        for (int k = tid; k < P.T * 7; k += blockDim.x) best_traj[off + k] = traj[off + k];
    }
}
}  // namespace rlitg

namespace rlitg {
template <typename T>
__global__ void torm_stage_decide_kernel(int64_t N, int mode, int stage, T lr_max, const T* __restrict__ lr,
                                         const uint8_t* __restrict__ infea_col, const uint8_t* __restrict__ infea_vel,
                                         int32_t* __restrict__ stage_idx) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const bool col = infea_col[n] != 0, vel = infea_vel[n] != 0;
    bool feasible_stage;
    if (stage == 0) feasible_stage = !(mode == 1 && lr[n] < T(1e-2) && !col && !vel);
    else feasible_stage = (mode == 1 && lr[n] == lr_max && col);
    stage_idx[n] = feasible_stage ? 0 : 1;
}
}  // namespace rlitg

extern "C" RLITG_API int rlitg_torm_stage_decide(int64_t N, int dtype, const void* lr, const uint8_t* infea_col,
                                                 const uint8_t* infea_vel, const RlitgTormSchedule* sch,
                                                 int32_t* stage_idx, void* stream) {
    if (!sch) return set_error(RLITG_ERR_NULL, "schedule is NULL");
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    if (N < 0) return set_error(RLITG_ERR_SHAPE, "N out of range");
    if (N == 0) return RLITG_OK;
    if (!lr || !infea_col || !infea_vel || !stage_idx) return set_error(RLITG_ERR_NULL, "lr / infea_* / stage_idx is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((N + 255) / 256);
    if (dtype == RLITG_F32)
        torm_stage_decide_kernel<float><<<grid, 256, 0, st>>>(N, sch->lr_schedule_mode, sch->stage, (float)sch->learning_rate_max,
                                                              (const float*)lr, infea_col, infea_vel, stage_idx);
    else
        torm_stage_decide_kernel<double><<<grid, 256, 0, st>>>(N, sch->lr_schedule_mode, sch->stage, sch->learning_rate_max,
                                                               (const double*)lr, infea_col, infea_vel, stage_idx);
    return check_launch("torm_stage_decide_kernel");
}

extern "C" RLITG_API int rlitg_torm_bookkeep(const void* traj, void* best_traj, int64_t N, int T, int dtype,
                                             const void* costs, const uint8_t* hits, const uint8_t* singular,
                                             void* alpha, void* lr, uint8_t* infea_col, uint8_t* infea_vel,
                                             void* best_cost, int32_t* stage_idx, const void* lr_iter,
                                             const RlitgTormSchedule* sch, void* stream) {
    if (!sch) return set_error(RLITG_ERR_NULL, "schedule is NULL");
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    if (N < 0 || N > 0x7fffffff || T < 2) return set_error(RLITG_ERR_SHAPE, "N / T out of range");
    if (sch->stage != 0 && sch->stage != 1) return set_error(RLITG_ERR_ARG, "stage must be 0 or 1");
    if (sch->lr_schedule_mode < 0 || sch->lr_schedule_mode > 3) return set_error(RLITG_ERR_ARG, "lr_schedule_mode must be 0..3");
    if (N == 0) return RLITG_OK;
    if (!traj || !best_traj || !costs || !alpha || !lr || !infea_col || !infea_vel || !best_cost)
        return set_error(RLITG_ERR_NULL, "traj / best_traj / costs / alpha / lr / infea_* / best_cost is NULL");
    BookP P;
    P.T = T; P.mode = sch->lr_schedule_mode; P.stage = sch->stage;
    P.use_col = sch->use_collision_check; P.use_vel = sch->use_velocity_check;
    P.hits_per_traj = sch->hits_per_trajectory;
    P.lr0 = sch->learning_rate; P.lr_min = sch->learning_rate_min; P.lr_max = sch->learning_rate_max;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32)
        torm_bookkeep_kernel<float><<<(unsigned)N, 128, 0, st>>>(P, (const float*)traj, (float*)best_traj, (const float*)costs,
                                                                 hits, singular, (float*)alpha, (float*)lr, infea_col, infea_vel,
                                                                 (float*)best_cost, stage_idx, (const float*)lr_iter);
    else
        torm_bookkeep_kernel<double><<<(unsigned)N, 128, 0, st>>>(P, (const double*)traj, (double*)best_traj, (const double*)costs,
                                                                  hits, singular, (double*)alpha, (double*)lr, infea_col, infea_vel,
                                                                  (double*)best_cost, stage_idx, (const double*)lr_iter);
    return check_launch("torm_bookkeep_kernel");
}

// ---- candidate selection of the Greedy / Linear initialisers --------------------------------
// TormOptimizer::getNewTrajectory torm/src/torm_optimizer.cpp:973-1003 (and :1066-1096) walks the candidates in
// order with a running best TOTAL cost that it compares against each candidate's POSE-ONLY cost (SURVEY appendix B
// item 9).  The rule is sequential, so one thread replays it over the scored batch; the scoring itself
// (interpolated segments through rlitg_torm_cost_grad) is the parallel part.
namespace rlitg {
template <typename T>
__global__ void greedy_select_kernel(const T* __restrict__ costs, int C, int Tseg, double rot_scale, double pose_w,
                                     double col_w, int32_t* __restrict__ best_idx, T* __restrict__ best_cost) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double best = INFINITY;
    int bi = -1;
    for (int j = 0; j < C; j++) {
        const T* c = costs + (size_t)j * RLITG_TORM_NCOST;
        const double eCost = ((double)c[8] + rot_scale * (double)c[9]) / (double)(Tseg - 1);  // ranged getEndPoseCost
        if (best < eCost) continue;
        const double cost = pose_w * eCost + col_w * (double)c[5];
        if (j == 0 || best > cost) { best = cost; bi = j; }
    }
    *best_idx = bi;
    if (best_cost) *best_cost = (T)best;
}
}  // namespace rlitg

extern "C" RLITG_API int rlitg_greedy_select(const void* costs, int C, int Tseg, int dtype, double rot_scale,
                                             double pose_w, double col_w, int32_t* best_idx, void* best_cost,
                                             void* stream) {
    if (!costs || !best_idx) return set_error(RLITG_ERR_NULL, "costs / best_idx is NULL");
    if (C < 1 || Tseg < 2) return set_error(RLITG_ERR_SHAPE, "C must be >= 1 and Tseg >= 2");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32) greedy_select_kernel<float><<<1, 32, 0, st>>>((const float*)costs, C, Tseg, rot_scale, pose_w, col_w, best_idx, (float*)best_cost);
    else if (dtype == RLITG_F64) greedy_select_kernel<double><<<1, 32, 0, st>>>((const double*)costs, C, Tseg, rot_scale, pose_w, col_w, best_idx, (double*)best_cost);
    else return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    return check_launch("greedy_select_kernel");
}

// ---- host-buffer form of rlitg_torm_cost_grad (the reference's callers hold Eigen matrices on the host) --------
namespace rlitg {
struct HostWs {
    std::mutex mu;
    cudaStream_t st = nullptr, st2 = nullptr;  // two streams: the chunks of a call alternate between them
    cudaEvent_t ev = nullptr;
    void* buf = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (!st && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return set_cuda_error("cudaStreamCreate");
        if (!st2 && cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking) != cudaSuccess) return set_cuda_error("cudaStreamCreate");
        if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return set_cuda_error("cudaEventCreate");
        if (cap < bytes) {
            if (buf) cudaFree(buf);
            buf = nullptr; cap = 0;
            if (cudaMalloc(&buf, bytes) != cudaSuccess) return set_cuda_error("cudaMalloc(torm host workspace)");
            cap = bytes;
        }
        return RLITG_OK;
    }
};
static HostWs g_torm_wss[64];  // one workspace per device
}  // namespace rlitg

extern "C" RLITG_API int rlitg_torm_cost_grad_host(const void* traj, const void* targets, int64_t N, int dtype,
                                                   const RlitgScene* scene, const RlitgSmooth* smooth,
                                                   const RlitgTormParams* p, int stage_mask,
                                                   const int32_t* stochastic_idx, void* inc_smooth, void* inc_coll,
                                                   void* inc_pose, void* costs) {
    if (!p) return set_error(RLITG_ERR_NULL, "params is NULL");
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    if (N < 0) return set_error(RLITG_ERR_SHAPE, "N out of range");
    if (N == 0) return RLITG_OK;
    if (!traj || !targets || !costs) return set_error(RLITG_ERR_NULL, "traj / targets / costs is NULL");
    const size_t es = dtype == RLITG_F32 ? 4 : 8;
    const size_t T = (size_t)p->T, F = T - 1;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_traj = al(es * N * T * 7), b_tg = al(es * T * 7), b_inc = al(es * N * F * 7), b_cost = al(es * N * RLITG_TORM_NCOST),
                 b_st = al(4 * (size_t)N);
    int ws_dev = 0;
    cudaGetDevice(&ws_dev);
    HostWs& g_torm_ws = g_torm_wss[ws_dev & 63];
    std::lock_guard<std::mutex> lk(g_torm_ws.mu);
    int rc = g_torm_ws.ensure(b_traj + b_tg + 3 * b_inc + b_cost + b_st);
    if (rc) return rc;
    char* base = static_cast<char*>(g_torm_ws.buf);
    char *d_traj = base, *d_tg = d_traj + b_traj, *d_is = d_tg + b_tg, *d_ic = d_is + b_inc, *d_ip = d_ic + b_inc,
         *d_cost = d_ip + b_inc, *d_st = d_cost + b_cost;
    // The trajectories go through in chunks that alternate between two streams: chunk k's increments travel back over
    // PCIe (17 MB per 1,024 x 200 call: the longest leg) while chunk k + 1 computes and chunk k + 2 is on its way in.
    const int64_t nchunk = N >= 512 ? 4 : 1;
    const int64_t per = (N + nchunk - 1) / nchunk;
    cudaStream_t sts[2] = {g_torm_ws.st, g_torm_ws.st2};
    if (cudaMemcpyAsync(d_tg, targets, es * T * 7, cudaMemcpyHostToDevice, sts[0]) != cudaSuccess ||
        cudaEventRecord(g_torm_ws.ev, sts[0]) != cudaSuccess || cudaStreamWaitEvent(sts[1], g_torm_ws.ev, 0) != cudaSuccess)
        return set_cuda_error("H2D targets");
    const char* h_traj = static_cast<const char*>(traj);
    for (int64_t c = 0; c < nchunk; c++) {
        const int64_t n0 = c * per, n = std::min(per, N - n0);
        if (n <= 0) break;
        cudaStream_t st = sts[c & 1];
        const size_t o_tr = es * n0 * T * 7, o_inc = es * n0 * F * 7, o_cost = es * n0 * RLITG_TORM_NCOST;
        if (cudaMemcpyAsync(d_traj + o_tr, h_traj + o_tr, es * n * T * 7, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            (stochastic_idx && cudaMemcpyAsync(d_st + 4 * n0, stochastic_idx + n0, 4 * (size_t)n, cudaMemcpyHostToDevice, st) != cudaSuccess))
            return set_cuda_error("H2D trajectories");
        rc = rlitg_torm_cost_grad(d_traj + o_tr, d_tg, n, dtype, scene, smooth, p, stage_mask,
                                  stochastic_idx ? reinterpret_cast<const int32_t*>(d_st) + n0 : nullptr,
                                  inc_smooth ? d_is + o_inc : nullptr, inc_coll ? d_ic + o_inc : nullptr,
                                  inc_pose ? d_ip + o_inc : nullptr, d_cost + o_cost, nullptr, st);
        if (rc) {
            cudaStreamSynchronize(sts[0]);
            cudaStreamSynchronize(sts[1]);
            return rc;
        }
        cudaError_t e = cudaMemcpyAsync(static_cast<char*>(costs) + o_cost, d_cost + o_cost, es * n * RLITG_TORM_NCOST, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && inc_smooth && (stage_mask & RLITG_STAGE_SMOOTH))
            e = cudaMemcpyAsync(static_cast<char*>(inc_smooth) + o_inc, d_is + o_inc, es * n * F * 7, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && inc_coll && (stage_mask & RLITG_STAGE_COLL))
            e = cudaMemcpyAsync(static_cast<char*>(inc_coll) + o_inc, d_ic + o_inc, es * n * F * 7, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && inc_pose && (stage_mask & RLITG_STAGE_POSE_GRAD))
            e = cudaMemcpyAsync(static_cast<char*>(inc_pose) + o_inc, d_ip + o_inc, es * n * F * 7, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) {
            cudaStreamSynchronize(sts[0]);
            cudaStreamSynchronize(sts[1]);
            return set_cuda_error("D2H increments");
        }
    }
    const cudaError_t e0 = cudaStreamSynchronize(sts[0]), e1 = cudaStreamSynchronize(sts[1]);
    if (e0 != cudaSuccess || e1 != cudaSuccess) return set_cuda_error("D2H increments");
    return RLITG_OK;
}
