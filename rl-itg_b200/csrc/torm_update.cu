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

constexpr int kUpdThreads = 64;

__device__ __forceinline__ double block_max(double v, double* s_tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = s_tmp[0];
    for (int w = 1; w < kUpdThreads / 32; w++) r = fmax(r, s_tmp[w]);
    return r;
}

template <typename T>
__global__ void __launch_bounds__(kUpdThreads)
torm_update_kernel(UpdP P, T* __restrict__ traj, const T* __restrict__ inc_smooth, const T* __restrict__ inc_coll,
                   const T* __restrict__ inc_pose, const T* __restrict__ alpha, const T* __restrict__ lr,
                   const double* __restrict__ chol, const double* __restrict__ inv,
                   const int32_t* __restrict__ stage_idx, int tab_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tn = P.Tn, F = Tn - 1, tid = threadIdx.x;
    double* fin = reinterpret_cast<double*>(smem_raw);  // [F][7] right-hand side -> final increments
    // second region, [F][8] doubles: during the solve it holds ONE substitution table at a time (tab_in_smem) -- every
    // step reads one 64-byte row, and a first-touch L2 round trip per row (a 128-byte line = two rows) was most of the
    // solve's latency -- and afterwards the free rows of the trajectory ([F][7]).  120 F bytes per CTA in all: 24 KB
    // at T = 200, so that a whole 1,024-trajectory batch is resident in one wave.
    double* tr = fin + (size_t)F * 7 + (F & 1);         // (F & 1): 7 F doubles is odd for odd F; keep 16-byte alignment
    double4* s_tab = reinterpret_cast<double4*>(tr);
    __shared__ double s_tmp[kUpdThreads / 32];
    __shared__ double s_best[kUpdThreads];
    __shared__ int s_idx[kUpdThreads];
    const int64_t n = blockIdx.x;
    // per-trajectory stage (lr_schedule_mode_ 1 lets every optimiser pick its own, torm_optimizer.cpp:400-412);
    // a negative entry leaves the trajectory untouched
    const int stage = stage_idx ? stage_idx[n] : P.stage;
    if (stage < 0) return;
    T* trow = traj + n * (int64_t)Tn * 7 + 7;  // first free row
    const size_t FD = (size_t)F * 7;
    const double4* g_tab = reinterpret_cast<const double4*>(chol);  // [2][F] rows of 2 x double4
    if (stage == 0 && tab_in_smem)
        for (int k = tid; k < 2 * F; k += kUpdThreads) s_tab[k] = g_tab[k];  // forward table

    if (stage == 0) {
        const T* s = inc_smooth + n * FD;
        const T* c = inc_coll + n * FD;
        double ss = 1.0, sc = 1.0;
        if (P.mode != 0) {
            double ms = 0.0, mc = 0.0;
            for (int k = tid; k < (int)FD; k += kUpdThreads) { ms = fmax(ms, fabs((double)s[k])); mc = fmax(mc, fabs((double)c[k])); }
            ss = block_max(ms, s_tmp);
            sc = block_max(mc, s_tmp);
        }
        const double a = alpha ? (double)alpha[n] : 0.5;
        for (int k = tid; k < (int)FD; k += kUpdThreads) {
            double sv = (double)s[k], cv = (double)c[k];
            if (P.mode != 0) {
                if (ss > 0) sv /= ss;
                if (sc > 0) cv /= sc;
                fin[k] = a * cv + (1 - a) * sv;
            } else {
                fin[k] = P.sw * sv + P.ow * cv;
            }
        }
        __syncthreads();
        // banded Cholesky solve per joint column: L y = b, then L^T x = y.  One thread per joint.  The recurrence is
        // the critical path (an FP64 FMA has ~40 cycles of latency here), so the substitution is column oriented: six
        // running rows live in registers, already divided by their diagonal; once unknown i is final it is pushed
        // into all six with independent FMAs and the first of them IS unknown i+1 -- one dependent FMA per step.
        // Coefficient rows (smooth_build) are fetched one step ahead.
        if (tid < 7) {
            const int j = tid;
            const double4* tab = tab_in_smem ? s_tab : g_tab;  // one row = 2 x double4
            auto rdiag = [&](const double4* t, int r) { return t[2 * r + 1].w; };
            auto rhs_at = [&](int r) { return (r >= 0 && r < F) ? fin[r * 7 + j] : 0.0; };
            // ---- forward: rows 0..5 start as b_r / L(r,r); a0 is the row being finalised
            double a0 = rhs_at(0) * rdiag(tab, 0);
            double a1 = F > 1 ? rhs_at(1) * rdiag(tab, 1) : 0.0, a2 = F > 2 ? rhs_at(2) * rdiag(tab, 2) : 0.0;
            double a3 = F > 3 ? rhs_at(3) * rdiag(tab, 3) : 0.0, a4 = F > 4 ? rhs_at(4) * rdiag(tab, 4) : 0.0;
            double a5 = F > 5 ? rhs_at(5) * rdiag(tab, 5) : 0.0, a6 = F > 6 ? rhs_at(6) * rdiag(tab, 6) : 0.0;
            double4 ca = tab[0], cb = tab[1];
            double nb_rhs = rhs_at(7);
            for (int i = 0; i < F; i++) {
                const int in = i + 1 < F ? i + 1 : i;
                const double4 na = tab[2 * in], nb = tab[2 * in + 1];
                const double nrhs = rhs_at(i + 8);
                const double y = a0;
                fin[i * 7 + j] = y;
                a0 = a1 - ca.x * y; a1 = a2 - ca.y * y; a2 = a3 - ca.z * y; a3 = a4 - ca.w * y; a4 = a5 - cb.x * y;
                a5 = a6 - cb.y * y;
                a6 = nb_rhs * cb.z;  // row i+7 enters the window: b / L(i+7,i+7)
                ca = na; cb = nb; nb_rhs = nrhs;
            }
        }
        __syncthreads();
        if (tab_in_smem)
            for (int k = tid; k < 2 * F; k += kUpdThreads) s_tab[k] = g_tab[2 * F + k];  // backward table
        __syncthreads();
        if (tid < 7) {
            const int j = tid;
            auto rdiag = [&](const double4* t, int r) { return t[2 * r + 1].w; };
            auto rhs_at = [&](int r) { return (r >= 0 && r < F) ? fin[r * 7 + j] : 0.0; };
            // ---- backward: rows F-1 .. F-6 start as y_r / L(r,r)
            const double4* tbk = tab_in_smem ? s_tab : g_tab + 2 * (size_t)F;
            double a0, a1, a2, a3, a4, a5, a6, nb_rhs;
            double4 ca, cb;
            const int e = F - 1;
            a0 = rhs_at(e) * rdiag(tbk, e);
            a1 = e - 1 >= 0 ? rhs_at(e - 1) * rdiag(tbk, e - 1) : 0.0; a2 = e - 2 >= 0 ? rhs_at(e - 2) * rdiag(tbk, e - 2) : 0.0;
            a3 = e - 3 >= 0 ? rhs_at(e - 3) * rdiag(tbk, e - 3) : 0.0; a4 = e - 4 >= 0 ? rhs_at(e - 4) * rdiag(tbk, e - 4) : 0.0;
            a5 = e - 5 >= 0 ? rhs_at(e - 5) * rdiag(tbk, e - 5) : 0.0; a6 = e - 6 >= 0 ? rhs_at(e - 6) * rdiag(tbk, e - 6) : 0.0;
            ca = tbk[2 * e]; cb = tbk[2 * e + 1];
            nb_rhs = rhs_at(e - 7);
            for (int i = e; i >= 0; i--) {
                const int in = i > 0 ? i - 1 : 0;
                const double4 na = tbk[2 * in], nb = tbk[2 * in + 1];
                const double nrhs = rhs_at(i - 8);
                const double x = a0;
                fin[i * 7 + j] = x;
                a0 = a1 - ca.x * x; a1 = a2 - ca.y * x; a2 = a3 - ca.z * x; a3 = a4 - ca.w * x; a4 = a5 - cb.x * x;
                a5 = a6 - cb.y * x;
                a6 = nb_rhs * cb.z;  // row i-7 enters the window
                ca = na; cb = nb; nb_rhs = nrhs;
            }
        }
        __syncthreads();  // the table region is free again
        const double rate = (double)lr[n] / P.mcs;
        for (int k = tid; k < (int)FD; k += kUpdThreads) fin[k] *= rate;
    } else {
        const T* p = inc_pose + n * FD;
        for (int k = tid; k < (int)FD; k += kUpdThreads) fin[k] = (double)p[k];
    }
    for (int k = tid; k < (int)FD; k += kUpdThreads) tr[k] = (double)trow[k];
    __syncthreads();

    // ---- addIncrementsToTrajectory :1611-1639
    if (P.mode != 0) {
        double m = 0.0;
        for (int k = tid; k < (int)FD; k += kUpdThreads) m = fmax(m, fabs(fin[k]));
        const double scale = block_max(m, s_tmp);
        const bool clip = scale > P.limit && scale > 0;
        for (int k = tid; k < (int)FD; k += kUpdThreads) tr[k] += clip ? fin[k] / scale * P.limit : fin[k];
    } else {
        for (int j = 0; j < 7; j++) {
            double mx = -1e300, mn = 1e300;
            for (int i = tid; i < F; i += kUpdThreads) { mx = fmax(mx, fin[i * 7 + j]); mn = fmin(mn, fin[i * 7 + j]); }
            mx = block_max(mx, s_tmp);
            mn = -block_max(-mn, s_tmp);
            double scale = 1.0;
            const double max_scale = P.limit / fabs(mx), min_scale = P.limit / fabs(mn);
            if (max_scale < scale) scale = max_scale;
            if (min_scale < scale) scale = min_scale;
            for (int i = tid; i < F; i += kUpdThreads) tr[i * 7 + j] += scale * fin[i * 7 + j];
        }
    }
    __syncthreads();

    // ---- handleJointLimits :1189-1268 (continuous joints are skipped)
    for (int j = 0; j < 7; j++) {
        if (P.cj[j]) continue;
        for (int count = 0; count <= 10; count++) {
            double best = 1e-6, amount_best = 0.0;
            int idx = -1;
            for (int i = tid; i < F; i += kUpdThreads) {
                const double v = tr[i * 7 + j];
                double amount = 0.0;
                if (v > P.jmax[j]) amount = P.jmax[j] - v;
                else if (v < P.jmin[j]) amount = P.jmin[j] - v;
                if (fabs(amount) > best) { best = fabs(amount); amount_best = amount; idx = i; }  // first maximum wins
            }
            s_best[tid] = idx >= 0 ? best : 0.0;
            s_idx[tid] = idx;
            __syncthreads();
            // every thread resolves the winner the same way: largest violation, smallest row index on ties
            double wb = 0.0;
            int wi = -1;
            for (int t = 0; t < kUpdThreads; t++) {
                const int ti = s_idx[t];
                if (ti >= 0 && (s_best[t] > wb || (s_best[t] == wb && ti < wi))) { wb = s_best[t]; wi = ti; }
            }
            __syncthreads();
            if (wi < 0) break;
            const double v = tr[wi * 7 + j];
            const double viol = v > P.jmax[j] ? P.jmax[j] - v : P.jmin[j] - v;
            (void)amount_best;
            const double mult = viol / inv[(size_t)wi * F + wi];
            __syncthreads();
            for (int i = tid; i < F; i += kUpdThreads) tr[i * 7 + j] += mult * inv[(size_t)wi * F + i];
            __syncthreads();
        }
    }
    __syncthreads();
    for (int k = tid; k < (int)FD; k += kUpdThreads) trow[k] = (T)tr[k];
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
    size_t smem = sizeof(double) * 2 * F * 7;
    if (smem > 200 * 1024) return set_error(RLITG_ERR_SHAPE, "T too large for the update step's shared-memory tile");
    const size_t smem_tab = sizeof(double) * (F * 7 + (F & 1) + F * 8);
    const int tab_in_smem = smem_tab <= 200 * 1024;  // longer paths read the tables through L1 / L2
    smem = tab_in_smem ? smem_tab : smem + 8;
    auto kern = torm_update_kernel<T>;
    {
        int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "torm_update_kernel");
        if (rc_attr) return rc_attr;
    }
    kern<<<(unsigned)N, kUpdThreads, smem, st>>>(P, traj, inc_smooth, inc_coll, inc_pose, alpha, lr, sm->d_chol, sm->d_inv64,
                                                 stage_idx, tab_in_smem);
    return check_launch("torm_update_kernel");
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
    cudaStream_t st = nullptr;
    void* buf = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (!st && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return set_cuda_error("cudaStreamCreate");
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
    cudaStream_t st = g_torm_ws.st;
    if (cudaMemcpyAsync(d_traj, traj, es * N * T * 7, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_tg, targets, es * T * 7, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        (stochastic_idx && cudaMemcpyAsync(d_st, stochastic_idx, 4 * (size_t)N, cudaMemcpyHostToDevice, st) != cudaSuccess))
        return set_cuda_error("H2D trajectories");
    rc = rlitg_torm_cost_grad(d_traj, d_tg, N, dtype, scene, smooth, p, stage_mask,
                              stochastic_idx ? reinterpret_cast<const int32_t*>(d_st) : nullptr,
                              inc_smooth ? d_is : nullptr, inc_coll ? d_ic : nullptr, inc_pose ? d_ip : nullptr, d_cost,
                              nullptr, st);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(costs, d_cost, es * N * RLITG_TORM_NCOST, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && inc_smooth && (stage_mask & RLITG_STAGE_SMOOTH)) e = cudaMemcpyAsync(inc_smooth, d_is, es * N * F * 7, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && inc_coll && (stage_mask & RLITG_STAGE_COLL)) e = cudaMemcpyAsync(inc_coll, d_ic, es * N * F * 7, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && inc_pose && (stage_mask & RLITG_STAGE_POSE_GRAD)) e = cudaMemcpyAsync(inc_pose, d_ip, es * N * F * 7, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return set_cuda_error("D2H increments");
    return RLITG_OK;
}
