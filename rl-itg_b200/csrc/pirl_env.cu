// Kernel 2: the PIRL path-wise IK environment step, fused.
//
// One thread per environment does what PirlEnv_Fetch.step does with three ROS service round-trips
// (pirl/script/envs/PirlEnv.py:209-289; cc_srv :196, allfk_srv :216, jaco_srv :536):
//   _set_action :184-207 -> all-link FK -> _get_obs :292-320 -> _compute_task_reward :495-513,
//   _compute_imitate_reward :515-526, _compute_null_space_imitate_reward :528-539 (+ jaco_srv,
//   pirl/src/kinematic_solver.cpp:332-351) -> reward / done / info flags :226-282.
// The same observation math exists in C++ at torm/src/pirl_interpolator.cpp:176-274 (float32 output).
//
// HBM-bound from ~256k envs on (536 B/env at the trained setting: 67 % of the measured HBM peak at 1M envs),
// latency-bound at 16k envs (one wave of 4-warp SMs).  What round 2's per-instruction stall samples changed
// (tools/ncu_sass_stalls.py, profiles/r02_pirl_*_stalls*.txt):
//   * load schedule: q / action / timestep / max_timestep are issued together and consumed together, then the three
//     table rows the timestep points at go out under the chain walk (ld_issue, load_row7);
//   * the table rows are 32-byte records read with one L2-only 256-bit load each;
//   * trained layout (kFast): the observation row is assembled in registers four values at a time (RowW) and each warp
//     drains its 32 rows with one bulk store issued before the reward arithmetic;
//   * other layouts: rows staged in shared memory at an odd stride (conflict-free), 128-bit cooperative copy-out.
#include <climits>
#include <cstdlib>
#include <type_traits>

#include "fetch_chain.cuh"
#include "linalg.cuh"
#include "rlitg_internal.h"
#include "scene.cuh"

namespace rlitg {

constexpr int kEnvTile = 128;

template <typename T>
struct PirlCfgT {
    T ll[7], ul[7];
    int cj[7];
    int n_future, z_dim, obs_type, train_mode, refine_flag, task_flag, imit_flag, null_flag;
    T task_suc_r, dist_suc_r, rot_suc_r, imit_suc_r;
    T thr_early, thr_dist, thr_rot, thr_imit, zero_task;
    T dist_p[4], rot_p[4], imit_p[4], null_p[4];
    T w_col, w_jl, w_es;
    int O;   // observation length
    int RS;  // floats per entry of the targets table: 7 (packed rows), 8 (32-byte rows) or 32 (128-byte step records)
    int rec_has_demo;  // RS = 32: the records carry demonstration rows
};

static int obs_dim_of(const RlitgPirlCfg& c) {
    int d = 0;
    if (c.obs_type == 0) d = 7;
    else if (c.obs_type == 1) d = 7 + 9 * 8;
    else if (c.obs_type == 2) d = 7 + 9;
    else d = 9 * 8;
    return d + 9 * c.n_future + c.z_dim;
}

template <typename T>
static PirlCfgT<T> convert_cfg(const RlitgPirlCfg& c) {
    PirlCfgT<T> o;
    for (int j = 0; j < 7; j++) { o.ll[j] = (T)c.ll[j]; o.ul[j] = (T)c.ul[j]; o.cj[j] = c.cjoint[j]; }
    o.n_future = c.n_future; o.z_dim = c.z_dim; o.obs_type = c.obs_type; o.train_mode = c.train_mode;
    o.refine_flag = c.refine_flag; o.task_flag = c.task_reward_flag; o.imit_flag = c.imitate_reward_flag;
    o.null_flag = c.null_reward_flag;
    o.task_suc_r = (T)c.task_suc_r; o.dist_suc_r = (T)c.dist_suc_r; o.rot_suc_r = (T)c.rot_suc_r; o.imit_suc_r = (T)c.imit_suc_r;
    o.thr_early = (T)c.thr_early_stop; o.thr_dist = (T)c.thr_dist; o.thr_rot = (T)c.thr_rot; o.thr_imit = (T)c.thr_imitate;
    o.zero_task = (T)c.zero_task_dist;
    for (int k = 0; k < 4; k++) { o.dist_p[k] = (T)c.dist_p[k]; o.rot_p[k] = (T)c.rot_p[k]; o.imit_p[k] = (T)c.imit_p[k]; o.null_p[k] = (T)c.null_p[k]; }
    o.w_col = (T)c.w_col; o.w_jl = (T)c.w_jlimit; o.w_es = (T)c.w_early_stop;
    o.O = obs_dim_of(c);
    o.RS = c.row_floats == 8 ? 8 : (c.row_floats == 32 ? 32 : 7);
    o.rec_has_demo = c.record_has_demo;
    return o;
}

// calc_reward: a exp(b e) - c (k e)^2   (PirlEnv.py:541-542)
template <typename T>
RLITG_DEV T calc_reward(T err, const T* p) {
    T e;
    if constexpr (sizeof(T) == 8) e = exp(p[1] * err); else e = expf(p[1] * err);
    const T ke = p[3] * err;
    return p[0] * e - p[2] * (ke * ke);
}

// Loads that must be ISSUED where they stand.  ptxas otherwise sinks the q / action loads behind the first use of the
// timestep (ncu source page, round 2: three serialised DRAM round trips at the head of every thread -- timestep, then
// the target rows, then q and the action -- 34 % of all stall samples at 1M envs).  volatile asm keeps program order.
RLITG_DEV float ld_issue(const float* p) { float v; asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
RLITG_DEV double ld_issue(const double* p) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
RLITG_DEV int ld_issue(const int32_t* p) { int v; asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

// one 7-value row of the target / demo tables.  Padded rows (kRow8) are 32-byte aligned: one 256-bit load that touches
// exactly one sector (FP32), instead of seven 4-byte loads of a row that straddles two (the 28-byte rows of a
// [B,Tmax,7] table start at multiples of 28 bytes: 1.41 x the algorithmic DRAM traffic at 1M envs, round 1).
// The row width is a template parameter: as a run-time branch both paths were if-converted into predicated loads of
// the same registers, and the second set waited for the first to land (a fourth round trip).
template <typename T, int kRW>
RLITG_DEV void load_row7(const T* __restrict__ row, T (&o)[7]) {
    if constexpr (kRW != 7) {
        if constexpr (sizeof(T) == 4) {
            // one 256-bit load (sm_100: LDG.E.256) = one request for the row's single 32-byte sector
            // L2-only (.cg): a 32-byte gather has no reuse in L1, and keeping the gathers out of it took the 1M-env step from
            // 160 to 132 us in the same run (DRAM bytes unchanged; .nc / L1::no_allocate / .cs: 160 / 150 / 162 us)
            float pad;
            asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(pad)
                         : "l"(row));
        } else {
            const double2* r2 = reinterpret_cast<const double2*>(row);
            const double2 a = __ldcg(r2), b = __ldcg(r2 + 1), c = __ldcg(r2 + 2), d = __ldcg(r2 + 3);
            o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 7; k++) o[k] = __ldg(row + k);
    }
}

// numpy-quaternion as_rotation_matrix (normalises by |q|^2), columns out
template <typename T>
RLITG_DEV void npq_rot(const T* t7, V3<T>& c0, V3<T>& c1, V3<T>& c2) {
    const T x = t7[3], y = t7[4], z = t7[5], w = t7[6];
    const T n = w * w + x * x + y * y + z * z;
    const T s = T(2) / n;
    c0 = {T(1) - s * (y * y + z * z), s * (x * y + z * w), s * (x * z - y * w)};
    c1 = {s * (x * y - z * w), T(1) - s * (x * x + z * z), s * (y * z + x * w)};
    c2 = {s * (x * z + y * w), s * (y * z - x * w), T(1) - s * (x * x + y * y)};
}

// the end-effector frame only: with link poses in the observation the Jacobian is rebuilt from the staged row
template <typename T>
struct EeKeep {
    FrameC<T> ee;
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        if constexpr (K == 7) ee = f;
    }
};

// row-major [6][7] Jacobian from the link poses staged for the observation: link k's entry holds (p, c0, c1) of the very
// frame joint k turns in, so z_k = c0 / c1 / c0 x c1 by the joint's axis and column k = [z_k x (p_ee - p_k); z_k]
template <typename T>
RLITG_DEV void jac_from_link_rows(const T* __restrict__ lrow, const V3<T>& eep, T (&J)[42]) {
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const T* f = lrow + 9 * k;
        const V3<T> p = {f[0], f[1], f[2]}, c0 = {f[3], f[4], f[5]}, c1 = {f[6], f[7], f[8]};
        const V3<T> z = arm_axis(k) == AX_X ? c0 : (arm_axis(k) == AX_Y ? c1 : cross(c0, c1));
        const V3<T> lin = cross(z, eep - p);
        J[0 * 7 + k] = lin.x; J[1 * 7 + k] = lin.y; J[2 * 7 + k] = lin.z;
        J[3 * 7 + k] = z.x; J[4 * 7 + k] = z.y; J[5 * 7 + k] = z.z;
    }
}

// Observation row assembled in aligned groups of four: values arrive in row order at compile-time positions and every
// completed group leaves as ONE 128-bit shared-memory store (the trained layout, kFast: 88 values = 22 stores instead
// of 88, rows densely packed so that a warp's 32 rows drain with a single bulk store)
template <typename T>
struct RowW {
    T* row;
    T w[4];
    template <int P>
    RLITG_DEV void put(T v) {
        w[P & 3] = v;
        if constexpr ((P & 3) == 3) {
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(row + (P - 3)) = make_float4(w[0], w[1], w[2], w[3]);
            } else {
                *reinterpret_cast<double2*>(row + (P - 3)) = make_double2(w[0], w[1]);
                *reinterpret_cast<double2*>(row + (P - 1)) = make_double2(w[2], w[3]);
            }
        }
    }
};

template <typename T>
RLITG_DEV void put_conf(RowW<T>& rw, const T (&c)[7]) {
    rw.template put<0>(c[0]); rw.template put<1>(c[1]); rw.template put<2>(c[2]); rw.template put<3>(c[3]);
    rw.template put<4>(c[4]); rw.template put<5>(c[5]); rw.template put<6>(c[6]);
}

template <typename T, bool kLinkObs, bool kFast>
struct PirlVisit {
    RowW<T> rw;  // kFast only
    T* row;
    int link_off, ee_off;
    bool do_coll;
    const SceneDev* sc;
    const float4* sph;
    const int* link_start;
    V3<T> off;
    bool hit;
    typename std::conditional<kLinkObs, EeKeep<T>, FkJac<T, 7>>::type fj;
    template <int K>
    RLITG_DEV void operator()(jidx<K> k, const FrameC<T>& f) {
        fj(k, f);
        if constexpr (kFast) {  // layout [conf 7 | 8 links x 9 | target 9]: link K at 7 + 9 K
            constexpr int P = 7 + 9 * K;
            rw.template put<P + 0>(f.p.x); rw.template put<P + 1>(f.p.y); rw.template put<P + 2>(f.p.z);
            rw.template put<P + 3>(f.c0.x); rw.template put<P + 4>(f.c0.y); rw.template put<P + 5>(f.c0.z);
            rw.template put<P + 6>(f.c1.x); rw.template put<P + 7>(f.c1.y); rw.template put<P + 8>(f.c1.z);
        } else if (link_off >= 0) {  // calc_pos_and_rot :322-333: p, R[:,0], R[:,1]
            T* o = row + link_off + 9 * K;
            o[0] = f.p.x; o[1] = f.p.y; o[2] = f.p.z;
            o[3] = f.c0.x; o[4] = f.c0.y; o[5] = f.c0.z;
            o[6] = f.c1.x; o[7] = f.c1.y; o[8] = f.c1.z;
        }
        if (!kFast && K == 7 && ee_off >= 0) {
            T* o = row + ee_off;
            o[0] = f.p.x; o[1] = f.p.y; o[2] = f.p.z;
            o[3] = f.c0.x; o[4] = f.c0.y; o[5] = f.c0.z;
            o[6] = f.c1.x; o[7] = f.c1.y; o[8] = f.c1.z;
        }
        if (do_coll) {
            const GridMap<T> gm(*sc);
            for (int s = link_start[K]; s < link_start[K + 1]; s++) {
                const float4 c = sph[s];
                const V3<T> x = {off.x + f.p.x + f.c0.x * T(c.x) + f.c1.x * T(c.y) + f.c2.x * T(c.z),
                                 off.y + f.p.y + f.c0.y * T(c.x) + f.c1.y * T(c.y) + f.c2.y * T(c.z),
                                 off.z + f.p.z + f.c0.z * T(c.x) + f.c1.z * T(c.y) + f.c2.z * T(c.z)};
                if (sphere_hits(*sc, gm, x, s, c.w, sc->dil)) hit = true;
            }
        }
    }
};

// kLinkObs: the observation carries the link poses (obs types 1 and 3; 1 is the trained setting) -- the 42 joint origin /
// axis values of the Jacobian then live in the staged observation row instead of registers (5 CTAs/SM instead of 4)
// kFast: the trained layout (obs type 1, one future target, no latent: 88 values) -- the row is assembled with 128-bit
// shared-memory stores into densely packed rows and every warp drains its 32 rows with one bulk store (TMA), without a
// CTA barrier and without the copy-out loop (410 of 2,170 instructions per thread, and the instruction footprint of the
// thread drops below the 32 KB instruction cache: 43 % of the stall samples at 16,384 envs were instruction fetches)
template <typename T, bool kStep, int TILE, int MINB, bool kLinkObs, int kRW, bool kFast>
__global__ void __launch_bounds__(TILE, MINB) pirl_env_kernel(
    PirlCfgT<T> cfg, int64_t B, T* __restrict__ q, const float* __restrict__ action, int32_t* __restrict__ timestep,
    const int32_t* __restrict__ max_timestep, const T* __restrict__ targets, const T* __restrict__ demos, int Tmax,
    const T* __restrict__ latent, SceneDev sc, int has_scene, T ox, T oy, T oz, T* __restrict__ obs,
    T* __restrict__ reward, uint8_t* __restrict__ done, uint8_t* __restrict__ flags, int32_t* __restrict__ counters,
    T* __restrict__ errs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int O = cfg.O, OS = kFast ? O : (O | 1);
    T* s_obs = reinterpret_cast<T*>(smem_raw);  // [TILE, OS]
    // sphere table behind the observation rows, allocated only with a scene: without it the trained layout fits FIVE
    // 128-thread CTAs per SM (5 x (45,056 + 1,024 reserved) bytes) instead of four
    float4* s_sph = reinterpret_cast<float4*>(s_obs + (size_t)TILE * OS);  // [kMaxSpheres]
    int* s_start = reinterpret_cast<int*>(s_sph + kMaxSpheres);              // [kNumLinks + 1]
    if (has_scene) {
        for (int i = threadIdx.x; i < sc.S; i += blockDim.x) s_sph[i] = sc.spheres[i];
        if (threadIdx.x <= kNumLinks) s_start[threadIdx.x] = sc.link_start[threadIdx.x];
        __syncthreads();
    }
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * TILE;
    const int n = (int)min((int64_t)TILE, B - base);
    const int64_t b = base + tid;
    const unsigned wmask = __ballot_sync(0xffffffffu, tid < n);  // lanes of this warp that own an environment
    if (tid < n) {
        T* row = s_obs + tid * OS;
        // first round trip: q, the action and the timestep together
        T cur[7];
#pragma unroll
        for (int j = 0; j < 7; j++) cur[j] = ld_issue(q + b * 7 + j);
        float act[7];
        if constexpr (kStep) {
#pragma unroll
            for (int j = 0; j < 7; j++) act[j] = ld_issue(action + b * 7 + j);
        }
        int ts = ld_issue(timestep + b);
        const int mts = ld_issue(max_timestep + b);
        // kRW = 32: 128-byte step records [target(t) | demo(t) | target(min(t + 1, Tmax - 1)) | pad], one per (env, t).
        // A per-thread gather costs a whole 128-byte line on this part whatever its size (tools/probes/gather_probe.cu:
        // 4 L2 sectors and 128 DRAM bytes per 32-byte -- or 16-byte -- row), so the three rows a step needs sit in ONE line
        constexpr int RS = kRW;
        const bool has_demo = kRW == 32 ? cfg.rec_has_demo != 0 : demos != nullptr;
        const T* tg = targets + (int64_t)b * Tmax * RS;

        const int ot = cfg.obs_type;
        PirlVisit<T, kLinkObs, kFast> v;
        v.row = row;
        v.rw.row = row;
        v.link_off = (ot == 1) ? 7 : (ot == 3 ? 0 : -1);
        v.ee_off = (ot == 2) ? 7 : -1;
        v.sc = &sc; v.sph = s_sph; v.link_start = s_start;
        v.off = {ox, oy, oz};
        v.hit = false;
        v.do_coll = false;

        bool collision_flag = false, joint_limit_flag = false;
        // loads whose addresses are known once the timestep is: issued before the chain walk so that their latency
        // overlaps the arithmetic instead of adding three dependent round trips at the end of the thread
        T pf_tp[7], pf_dm[7], pf_ft[7];
        int4 pf_cnt = make_int4(0, 0, 0, 0);
        if constexpr (kStep) {
            // _set_action :184-207
            T cl[7];
            bool no_out = true;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                T o2 = cur[j] + T(act[j]);
                // scheduling tie: the first consumer of q / action also consumes the timestep, so that all three
                // loads are in flight together (see ld_issue); INT_MIN is not a timestep
                if (j == 0 && (ts == INT_MIN || mts == INT_MIN)) o2 = T(0);
                if (cfg.cj[j]) {
                    const T pi = T(3.14159265358979323846);
                    if (o2 > pi) o2 -= T(2) * pi;
                    else if (o2 < -pi) o2 += T(2) * pi;
                }
                const T c = o2 < cfg.ll[j] ? cfg.ll[j] : (o2 > cfg.ul[j] ? cfg.ul[j] : o2);
                // np.allclose(o2, clip(o2)): |a - b| <= 1e-8 + 1e-5 |b|
                if (!(abs_t(o2 - c) <= T(1e-8) + T(1e-5) * abs_t(c))) no_out = false;
                cl[j] = c;
            }
            // second round trip, under the chain walk: the rows the timestep points at (placed behind the first use
            // of q and the action so that the scheduler does not hold those loads back behind the wait for the timestep)
            ts += 1;
            {
                const int tsp = min(max(ts, 0), Tmax - 1);
                load_row7<T, kRW>(tg + (size_t)tsp * RS, pf_tp);
                if (has_demo) {
                    if constexpr (kRW == 32) load_row7<T, kRW>(tg + (size_t)tsp * RS + 8, pf_dm);
                    else load_row7<T, kRW>(demos + ((int64_t)b * Tmax + tsp) * RS, pf_dm);
                }
                if (counters) pf_cnt = __ldcg(reinterpret_cast<const int4*>(counters + b * 4));
                if constexpr (kFast) {
                    const int tf = min(max(min(ts + 1, mts), 0), Tmax - 1);
                    // records: the future target of step t is slot 2 of the record just read, except at the end of a path
                    if (kRW == 32 && tf == tsp + 1) load_row7<T, kRW>(tg + (size_t)tsp * RS + 16, pf_ft);
                    else load_row7<T, kRW>(tg + (size_t)tf * RS, pf_ft);
                }
            }
            // one walk of the configuration the step will keep: the clipped candidate (test mode, or inside the
            // limits), else the unchanged one (joint-limit rejection); a second walk only after a collision
            const bool use_cl = !cfg.train_mode || no_out;
            joint_limit_flag = !use_cl;
            T qv[7];
#pragma unroll
            for (int j = 0; j < 7; j++) qv[j] = use_cl ? cl[j] : cur[j];
            v.do_coll = cfg.train_mode && no_out && has_scene != 0;
            if constexpr (kFast) put_conf<T>(v.rw, qv);
            walk_chain<T, 7>(qv, v);
            if (v.hit) {
                collision_flag = true;
                v.do_coll = false;
                if constexpr (kFast) put_conf<T>(v.rw, cur);
                walk_chain<T, 7>(cur, v);  // the move is rejected: pose of the unchanged configuration
            } else {
#pragma unroll
                for (int j = 0; j < 7; j++) cur[j] = qv[j];
            }
#pragma unroll
            for (int j = 0; j < 7; j++) q[b * 7 + j] = cur[j];
            timestep[b] = ts;
        } else {
            if constexpr (kFast) {
                const int tf = min(max(min(ts + 1, mts), 0), Tmax - 1);
                load_row7<T, kRW>(tg + (size_t)tf * RS, pf_ft);
                put_conf<T>(v.rw, cur);
            }
            walk_chain<T, 7>(cur, v);
        }
        const FrameC<T>& ee = v.fj.ee;

        // ---- _get_obs :292-320
        if constexpr (kFast) {
            // the single future target: index min(timestep + 1, max_timestep), loaded ahead with the other rows
            V3<T> m0, m1, m2;
            npq_rot(pf_ft, m0, m1, m2);
            const V3<T> r0 = {ee.c0.x, ee.c1.x, ee.c2.x}, r1 = {ee.c0.y, ee.c1.y, ee.c2.y};
            v.rw.template put<79>(pf_ft[0] - ee.p.x);
            v.rw.template put<80>(pf_ft[1] - ee.p.y);
            v.rw.template put<81>(pf_ft[2] - ee.p.z);
            v.rw.template put<82>(m0.x * r0.x + m1.x * r0.y + m2.x * r0.z);
            v.rw.template put<83>(m0.y * r0.x + m1.y * r0.y + m2.y * r0.z);
            v.rw.template put<84>(m0.z * r0.x + m1.z * r0.y + m2.z * r0.z);
            v.rw.template put<85>(m0.x * r1.x + m1.x * r1.y + m2.x * r1.z);
            v.rw.template put<86>(m0.y * r1.x + m1.y * r1.y + m2.y * r1.z);
            v.rw.template put<87>(m0.z * r1.x + m1.z * r1.y + m2.z * r1.z);
            // the rows of this warp are complete: start their bulk store now, under the reward arithmetic (the wait for
            // the engine's shared-memory reads at the end of the kernel was 13 % of the stall samples at 16,384 envs)
            fence_proxy_async_smem();  // every writer: its row becomes visible to the async proxy
            __syncwarp(wmask);
            if ((tid & 31) == 0) {
                const int warp = tid >> 5;
                bulk_store(obs + (base + 32 * warp) * O, s_obs + (size_t)32 * warp * O,
                           (uint32_t)(min(32, n - 32 * warp) * O * sizeof(T)));
                bulk_commit();
            }
        } else {
            if (ot != 3) {
    #pragma unroll
                for (int j = 0; j < 7; j++) row[j] = cur[j];
            }
            int fo = (ot == 0) ? 7 : (ot == 1 ? 79 : (ot == 2 ? 16 : 72));
            for (int i = 0; i < cfg.n_future; i++) {
                int t = ts + (i + 1);
                if (t > mts) t = mts;
                t = min(max(t, 0), Tmax - 1);  // memory-safety guard: a caller stepping a finished episode stays in its path
                T t7[7];
                const int tbase = min(max(ts, 0), Tmax - 1);  // kStep: the record of this step (ts is already incremented)
                if (kRW == 32 && kStep && i == 0 && t == tbase + 1) load_row7<T, kRW>(tg + (size_t)tbase * RS + 16, t7);
                else load_row7<T, kRW>(tg + (size_t)t * RS, t7);
                // calc_pos_and_rot_error :335-349: R_ct = tM inv(eM); columns 0 and 1
                V3<T> m0, m1, m2;
                npq_rot(t7, m0, m1, m2);
                T* o = row + fo + 9 * i;
                o[0] = t7[0] - ee.p.x; o[1] = t7[1] - ee.p.y; o[2] = t7[2] - ee.p.z;
                // inv(eM) = eM^T: column c of R_ct = tM * (row c of eM)
                const V3<T> r0 = {ee.c0.x, ee.c1.x, ee.c2.x}, r1 = {ee.c0.y, ee.c1.y, ee.c2.y};
                o[3] = m0.x * r0.x + m1.x * r0.y + m2.x * r0.z;
                o[4] = m0.y * r0.x + m1.y * r0.y + m2.y * r0.z;
                o[5] = m0.z * r0.x + m1.z * r0.y + m2.z * r0.z;
                o[6] = m0.x * r1.x + m1.x * r1.y + m2.x * r1.z;
                o[7] = m0.y * r1.x + m1.y * r1.y + m2.y * r1.z;
                o[8] = m0.z * r1.x + m1.z * r1.y + m2.z * r1.z;
            }
            fo += 9 * cfg.n_future;
            for (int k = 0; k < cfg.z_dim; k++) row[fo + k] = latent[b * cfg.z_dim + k];

        }
        if constexpr (kStep) {
            // ---- _compute_task_reward :495-513 (target index = timestep, loaded above into pf_tp with the index clamped
            // into the path as a memory-safety guard -- the reference raises IndexError when a finished episode is stepped)
            T tp[7];
#pragma unroll
            for (int k = 0; k < 7; k++) tp[k] = pf_tp[k];
            const T dx = tp[0] - ee.p.x, dy = tp[1] - ee.p.y, dz = tp[2] - ee.p.z;
            const T d_err = sqrt_t(dx * dx + dy * dy + dz * dz);
            T ex, ey, ez, ew;
            rot_to_quat_kdl(ee.c0, ee.c1, ee.c2, ex, ey, ez, ew);
            // (tq * conj(pq)).w = tw pw + tx px + ty py + tz pz
            const T dw = tp[6] * ew + tp[3] * ex + tp[4] * ey + tp[5] * ez;
            const T adw = abs_t(dw);
            const T r_err = (adw > T(1)) ? T(0) : T(2) * acos_t(adw);
            const T dist_rwd = calc_reward(d_err, cfg.dist_p);
            const T rot_rwd = calc_reward(r_err, cfg.rot_p);

            // ---- imitation :515-526 and null-space imitation :528-539
            T imitate_rwd = T(0), imitate_err = T(0), null_rwd = T(0), null_obj = T(0);
            if (has_demo) {
                const T* dm = pf_dm;
                T ce[7];
                T l1 = T(0);
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    T e = dm[j] - cur[j];
                    if (cfg.cj[j]) {
                        const T pi = T(3.14159265358979323846);
                        if (e > pi) e -= T(2) * pi;
                        else if (e < -pi) e += T(2) * pi;
                    }
                    ce[j] = e;
                    l1 += abs_t(e);
                }
                if (!cfg.refine_flag) {
                    imitate_err = l1 / T(7);
                    imitate_rwd = calc_reward(imitate_err, cfg.imit_p);
                }
                // the reference calls jaco_srv on every step (:224); it is only needed when it enters the
                // reward or the caller asks for it through errs
                if ((cfg.null_flag && !cfg.refine_flag) || errs != nullptr) {
                    T J[42];
                    if constexpr (kLinkObs) jac_from_link_rows<T>(row + v.link_off, ee.p, J);
                    else fill_jac<T, 7>(v.fj, J);
                    null_obj = nullspace_norm<T, 7>(J, ce);
                    null_rwd = calc_reward(null_obj, cfg.null_p);
                }
            }

            const int is_time_out = (ts == mts);
            const int is_col = collision_flag, is_jl = joint_limit_flag;
            const int is_es = d_err > cfg.thr_early;
            const int is_ds = d_err < cfg.thr_dist;
            const int is_rs = r_err < cfg.thr_rot;
            const int is_is = (!cfg.refine_flag) ? (imitate_err < cfg.thr_imit) : 0;
            T rw = -cfg.w_col * T(is_col) - cfg.w_jl * T(is_jl) - cfg.w_es * T(is_es);
            if (cfg.task_flag) {
                const bool near_t = d_err < cfg.zero_task;
                rw += (dist_rwd + cfg.dist_suc_r * T(is_ds) + (near_t ? rot_rwd : T(0)) +
                       (near_t ? cfg.rot_suc_r * T(is_rs) : T(0)) + cfg.task_suc_r * T(is_ds * is_rs));
            }
            if (cfg.imit_flag && !cfg.refine_flag) rw += (imitate_rwd + cfg.imit_suc_r * T(is_is));
            if (cfg.null_flag && !cfg.refine_flag) rw += null_rwd;

            reward[b] = rw;
            done[b] = (uint8_t)(is_col || is_es || is_time_out);
            const uint32_t lo = (uint32_t)is_col | ((uint32_t)is_jl << 8) | ((uint32_t)is_es << 16) | ((uint32_t)is_time_out << 24);
            const uint32_t hi = (uint32_t)is_ds | ((uint32_t)is_rs << 8) | ((uint32_t)is_is << 16) | ((uint32_t)(is_ds * is_rs) << 24);
            *reinterpret_cast<uint2*>(flags + b * 8) = make_uint2(lo, hi);
            if (counters) {
                int4 c = pf_cnt;
                c.x += is_ds * is_rs; c.y += is_ds; c.z += is_rs;
                if (!cfg.refine_flag) c.w += is_is;
                *reinterpret_cast<int4*>(counters + b * 4) = c;
            }
            if (errs) {
                errs[b * 4 + 0] = d_err; errs[b * 4 + 1] = r_err; errs[b * 4 + 2] = imitate_err; errs[b * 4 + 3] = null_obj;
            }
        }
    }
    if constexpr (kFast) {
        // every warp drained its own (densely packed) rows with one bulk store, no CTA barrier; the shared memory must
        // stay allocated until the engine has read it
        if ((tid & 31) == 0 && tid < n) bulk_wait_read_all();
        return;
    }
    __syncthreads();
    // coalesced copy-out of the staged rows
    T* dst = obs + base * O;
    if (sizeof(T) == 4 && (O & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        // 128-bit stores: a row holds O / 4 whole float4 groups, so a group never straddles two rows
        const int G = O >> 2;
        int r = tid / G, g = tid - r * G;
        const int dr = TILE / G, dg = TILE - dr * G;
        for (int i = tid; i < n * G; i += TILE) {
            const T* sp = s_obs + r * OS + 4 * g;
            st_stream(reinterpret_cast<float4*>(dst) + i, make_float4((float)sp[0], (float)sp[1], (float)sp[2], (float)sp[3]));
            r += dr; g += dg;
            if (g >= G) { g -= G; r++; }
        }
    } else {
        // (row, column) advance incrementally instead of dividing by the runtime O
        int r = tid / O, c = tid - r * O;
        const int dr = TILE / O, dc = TILE - dr * O;
        for (int i = tid; i < n * O; i += TILE) {
            st_stream(dst + i, s_obs[r * OS + c]);
            r += dr; c += dc;
            if (c >= O) { c -= O; r++; }
        }
    }
}

// ---- kernel 2b: the same step with EIGHT LANES PER ENVIRONMENT, for small batches ----------------------------------
// Below one wave the thread-per-env kernel's time is the serial instruction chain of one thread (~2,200 instructions
// per warp at ~7 cycles each with one warp per scheduler: profiles/r02_pirl_16k_thread_full.json).  Here the chain is
// cut (to ~1,370 instructions per warp) by data-parallel splits inside a group of 8 lanes (4 environments per warp,
// group shuffles only):
//   sin / cos        lane j = joint j
//   chain walk       lane r = coordinate r: every row of [c0 c1 c2 | p] evolves independently under right-multiplication
//                    by the joint transforms (3 lanes busy, ~45 dependent FMAs instead of ~230); link frames go to a
//                    per-environment shared-memory scratch that every lane reads afterwards
//   collision test   sphere s on lane s mod 8
//   observation      link k on lane k (9 values each), staged in shared memory, written back 8 floats per store
//   Jacobian / QR    column j on lane j; the dot products of the modified Gram-Schmidt sweep and of the null-space
//                    projection are 3-step group reductions
// The scalar parts (task reward, flags) are executed redundantly by all lanes: free in latency, but 5 x the total work of
// the thread-per-env kernel -- which is why this kernel only wins while the GPU is nearly empty (launch_pirl).
constexpr int kLaneEnvs = 16;  // environments per CTA: 128 threads

template <typename T>
RLITG_DEV T group_sum(T x, unsigned gmask) {
    x += __shfl_xor_sync(gmask, x, 1);
    x += __shfl_xor_sync(gmask, x, 2);
    x += __shfl_xor_sync(gmask, x, 4);
    return x;
}

// FK of one environment by its lane group: qv = joint value of lane l (< 7); frames fr[k][12] = (p, c0, c1, c2) of link k
template <typename T>
RLITG_DEV void group_fk(T qv, int l, unsigned gmask, T* __restrict__ sc, T* __restrict__ fr) {
    if (l < 7) {
        T sn, cs;
        sincos_joint(qv, sn, cs);
        sc[2 * l] = sn;
        sc[2 * l + 1] = cs;
    }
    __syncwarp(gmask);
    if (l < 3) {
        const int r = l;
        T e0 = r == 0 ? T(1) : T(0), e1 = r == 1 ? T(1) : T(0), e2 = r == 2 ? T(1) : T(0), pp = T(0);
#pragma unroll
        for (int j = 0; j < 7; j++) {
            pp = fma_t(arm_ox<T>(j), e0, pp);
            if (arm_oz<T>(j) != T(0)) pp = fma_t(arm_oz<T>(j), e2, pp);
            const T sn = sc[2 * j], cs = sc[2 * j + 1];
            if (arm_axis(j) == AX_Z) { const T a = e0, b = e1; e0 = cs * a + sn * b; e1 = cs * b - sn * a; }
            else if (arm_axis(j) == AX_Y) { const T a = e0, b = e2; e0 = cs * a - sn * b; e2 = sn * a + cs * b; }
            else { const T a = e1, b = e2; e1 = cs * a + sn * b; e2 = cs * b - sn * a; }
            T* f = fr + j * 12;
            f[r] = pp; f[3 + r] = e0; f[6 + r] = e1; f[9 + r] = e2;
        }
        pp = fma_t(tip_ox<T>(), e0, pp);
        T* f = fr + 7 * 12;
        f[r] = pp; f[3 + r] = e0; f[6 + r] = e1; f[9 + r] = e2;
    }
    __syncwarp(gmask);
}

template <typename T, bool kStep, int kRW>
__global__ void __launch_bounds__(kLaneEnvs * 8) pirl_env_lanes_kernel(
    PirlCfgT<T> cfg, int64_t B, T* __restrict__ q, const float* __restrict__ action, int32_t* __restrict__ timestep,
    const int32_t* __restrict__ max_timestep, const T* __restrict__ targets, const T* __restrict__ demos, int Tmax,
    const T* __restrict__ latent, SceneDev sc, int has_scene, T ox, T oy, T oz, T* __restrict__ obs,
    T* __restrict__ reward, uint8_t* __restrict__ done, uint8_t* __restrict__ flags, int32_t* __restrict__ counters,
    T* __restrict__ errs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int O = cfg.O;
    T* s_fr = reinterpret_cast<T*>(smem_raw);        // [kLaneEnvs][8][12] link frames
    T* s_sc = s_fr + kLaneEnvs * 96;                 // [kLaneEnvs][14] sin, cos
    T* s_obs = s_sc + kLaneEnvs * 14;                // [kLaneEnvs][O] observation rows
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_link[kMaxSpheres];
    if (has_scene) {
        for (int i = threadIdx.x; i < sc.S; i += blockDim.x) {
            s_sph[i] = sc.spheres[i];
            int link = 0;
            while (link < kNumLinks - 1 && i >= sc.link_start[link + 1]) link++;
            s_link[i] = link;
        }
        __syncthreads();
    }
    const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
    const unsigned gmask = 0xffu << (8 * ((threadIdx.x & 31) >> 3));
    const int64_t b = (int64_t)blockIdx.x * kLaneEnvs + g;
    if (b >= B) return;  // whole groups leave together; there is no CTA-wide barrier below
    T* fr = s_fr + g * 96;
    T* scv = s_sc + g * 14;
    T* row = s_obs + (size_t)g * O;
    constexpr int RS = kRW;
    const bool has_demo = kRW == 32 ? cfg.rec_has_demo != 0 : demos != nullptr;
    const T* tg = targets + (int64_t)b * Tmax * RS;
    const int ot = cfg.obs_type;

    T cur = l < 7 ? ld_issue(q + b * 7 + l) : T(0);
    float act = 0.f;
    if constexpr (kStep) {
        if (l < 7) act = ld_issue(action + b * 7 + l);
    }
    int ts = ld_issue(timestep + b);
    const int mts = ld_issue(max_timestep + b);
    bool collision_flag = false, joint_limit_flag = false;
    T tp[7], dm[7];
    int cnt = 0;
    if constexpr (kStep) {
        // _set_action :184-207, joint l on lane l
        T o2 = cur + T(act);
        if (ts == INT_MIN) o2 = T(0);  // scheduling tie, as in the thread-per-env kernel
        const int j = l < 7 ? l : 0;
        if (l < 7 && cfg.cj[j]) {
            const T pi = T(3.14159265358979323846);
            if (o2 > pi) o2 -= T(2) * pi;
            else if (o2 < -pi) o2 += T(2) * pi;
        }
        const T lo = cfg.ll[j], hi = cfg.ul[j];
        const T c = l < 7 ? (o2 < lo ? lo : (o2 > hi ? hi : o2)) : T(0);
        // np.allclose(o2, clip(o2)): |a - b| <= 1e-8 + 1e-5 |b|
        const bool ok = l >= 7 || (abs_t(o2 - c) <= T(1e-8) + T(1e-5) * abs_t(c));
        const bool no_out = (__ballot_sync(gmask, ok) & gmask) == gmask;
        ts += 1;
        const int tsp = min(max(ts, 0), Tmax - 1);
        load_row7<T, kRW>(tg + (size_t)tsp * RS, tp);  // every lane holds the whole row (one broadcast transaction per load)
        if (has_demo) {
            if constexpr (kRW == 32) load_row7<T, kRW>(tg + (size_t)tsp * RS + 8, dm);
            else load_row7<T, kRW>(demos + ((int64_t)b * Tmax + tsp) * RS, dm);
        }
        if (counters && l < 4) cnt = counters[b * 4 + l];
        if (cfg.train_mode) {
            if (no_out) {
                group_fk<T>(c, l, gmask, scv, fr);
                bool hit = false;
                if (has_scene) {
                    const GridMap<T> gm(sc);
                    for (int s = l; s < sc.S; s += 8) {
                        const float4 cc = s_sph[s];
                        const T* f = fr + s_link[s] * 12;
                        const V3<T> x = {ox + f[0] + f[3] * T(cc.x) + f[6] * T(cc.y) + f[9] * T(cc.z),
                                         oy + f[1] + f[4] * T(cc.x) + f[7] * T(cc.y) + f[10] * T(cc.z),
                                         oz + f[2] + f[5] * T(cc.x) + f[8] * T(cc.y) + f[11] * T(cc.z)};
                        if (sphere_hits(sc, gm, x, s, cc.w, sc.dil)) hit = true;
                    }
                }
                if (__ballot_sync(gmask, hit) & gmask) {
                    collision_flag = true;
                    __syncwarp(gmask);
                    group_fk<T>(cur, l, gmask, scv, fr);  // the move is rejected: pose of the unchanged configuration
                } else {
                    cur = c;
                }
            } else {
                joint_limit_flag = true;
                group_fk<T>(cur, l, gmask, scv, fr);
            }
        } else {
            cur = c;
            group_fk<T>(cur, l, gmask, scv, fr);
        }
        if (l < 7) q[b * 7 + l] = cur;
        if (l == 0) timestep[b] = ts;
    } else {
        group_fk<T>(cur, l, gmask, scv, fr);
    }
    const T* ee = fr + 7 * 12;  // p, c0, c1, c2 of gripper_link

    // ---- _get_obs :292-320, staged in shared memory
    if (ot != 3 && l < 7) row[l] = cur;
    {
        const int link_off = (ot == 1) ? 7 : (ot == 3 ? 0 : -1);
        if (link_off >= 0) {  // calc_pos_and_rot :322-333: p, R[:,0], R[:,1] of link l
#pragma unroll
            for (int k = 0; k < 9; k++) row[link_off + 9 * l + k] = fr[l * 12 + k];
        } else if (ot == 2 && l == 7) {
#pragma unroll
            for (int k = 0; k < 9; k++) row[7 + k] = ee[k];
        }
    }
    int fo = (ot == 0) ? 7 : (ot == 1 ? 79 : (ot == 2 ? 16 : 72));
    for (int i = l; i < cfg.n_future; i += 8) {  // future target i on lane i mod 8
        int t = ts + (i + 1);
        if (t > mts) t = mts;
        t = min(max(t, 0), Tmax - 1);  // memory-safety guard: a caller stepping a finished episode stays in its path
        T t7[7];
        const int tbase = min(max(ts, 0), Tmax - 1);
        if (kRW == 32 && kStep && i == 0 && t == tbase + 1) load_row7<T, kRW>(tg + (size_t)tbase * RS + 16, t7);
        else load_row7<T, kRW>(tg + (size_t)t * RS, t7);
        // calc_pos_and_rot_error :335-349: R_ct = tM inv(eM); columns 0 and 1
        V3<T> m0, m1, m2;
        npq_rot(t7, m0, m1, m2);
        T* o = row + fo + 9 * i;
        o[0] = t7[0] - ee[0]; o[1] = t7[1] - ee[1]; o[2] = t7[2] - ee[2];
        // inv(eM) = eM^T: column c of R_ct = tM * (row c of eM); row 0 of eM = (c0.x, c1.x, c2.x)
        const V3<T> r0 = {ee[3], ee[6], ee[9]}, r1 = {ee[4], ee[7], ee[10]};
        o[3] = m0.x * r0.x + m1.x * r0.y + m2.x * r0.z;
        o[4] = m0.y * r0.x + m1.y * r0.y + m2.y * r0.z;
        o[5] = m0.z * r0.x + m1.z * r0.y + m2.z * r0.z;
        o[6] = m0.x * r1.x + m1.x * r1.y + m2.x * r1.z;
        o[7] = m0.y * r1.x + m1.y * r1.y + m2.y * r1.z;
        o[8] = m0.z * r1.x + m1.z * r1.y + m2.z * r1.z;
    }
    fo += 9 * cfg.n_future;
    for (int k = l; k < cfg.z_dim; k += 8) row[fo + k] = latent[b * cfg.z_dim + k];
    __syncwarp(gmask);
    for (int k = l; k < O; k += 8) st_stream(obs + b * O + k, row[k]);

    if constexpr (kStep) {
        // ---- _compute_task_reward :495-513, redundantly on every lane
        const T dx = tp[0] - ee[0], dy = tp[1] - ee[1], dz = tp[2] - ee[2];
        const T d_err = sqrt_t(dx * dx + dy * dy + dz * dz);
        T ex, ey, ez, ew;
        rot_to_quat_kdl(V3<T>{ee[3], ee[4], ee[5]}, V3<T>{ee[6], ee[7], ee[8]}, V3<T>{ee[9], ee[10], ee[11]}, ex, ey, ez, ew);
        const T dw = tp[6] * ew + tp[3] * ex + tp[4] * ey + tp[5] * ez;
        const T adw = abs_t(dw);
        const T r_err = (adw > T(1)) ? T(0) : T(2) * acos_t(adw);
        const T dist_rwd = calc_reward(d_err, cfg.dist_p);
        const T rot_rwd = calc_reward(r_err, cfg.rot_p);

        // ---- imitation :515-526 and null-space imitation :528-539, joint l on lane l
        T imitate_rwd = T(0), imitate_err = T(0), null_rwd = T(0), null_obj = T(0);
        if (has_demo) {
            T e = T(0);
            if (l < 7) {
                T dv = dm[0];
#pragma unroll
                for (int k = 1; k < 7; k++) dv = (l == k) ? dm[k] : dv;
                e = dv - cur;
                if (cfg.cj[l]) {
                    const T pi = T(3.14159265358979323846);
                    if (e > pi) e -= T(2) * pi;
                    else if (e < -pi) e += T(2) * pi;
                }
            }
            const T l1 = group_sum(abs_t(e), gmask);
            if (!cfg.refine_flag) {
                imitate_err = l1 / T(7);
                imitate_rwd = calc_reward(imitate_err, cfg.imit_p);
            }
            if ((cfg.null_flag && !cfg.refine_flag) || errs != nullptr) {
                // column l of the Jacobian (KDL JntToJac): [z x (p_ee - o); z]; lane 7 carries a zero column
                T Q[6];
#pragma unroll
                for (int k = 0; k < 6; k++) Q[k] = T(0);
                if (l < 7) {
                    const int ci = (l == 0) ? 2 : ((l & 1) ? 1 : 0);  // axis z, y, x, y, x, y, x -> rotation column
                    const T* f = fr + l * 12;
                    const V3<T> z = {f[3 + 3 * ci], f[4 + 3 * ci], f[5 + 3 * ci]};
                    const V3<T> dlt = {ee[0] - f[0], ee[1] - f[1], ee[2] - f[2]};
                    const V3<T> lin = cross(z, dlt);
                    Q[0] = lin.x; Q[1] = lin.y; Q[2] = lin.z; Q[3] = z.x; Q[4] = z.y; Q[5] = z.z;
                }
                // modified Gram-Schmidt over the six rows of J (this lane holds one column), then e - Q^T Q e
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const T n2 = group_sum(Q[k] * Q[k], gmask);
                    const T rn = sqrt_t(n2);
                    Q[k] *= T(1) / (rn > T(1e-30) ? rn : T(1e-30));
#pragma unroll
                    for (int jj = k + 1; jj < 6; jj++) {
                        const T d = group_sum(Q[k] * Q[jj], gmask);
                        Q[jj] -= d * Q[k];
                    }
                }
                T r = e;
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const T c = group_sum(Q[k] * r, gmask);
                    r -= c * Q[k];
                }
                null_obj = sqrt_t(group_sum(r * r, gmask));
                null_rwd = calc_reward(null_obj, cfg.null_p);
            }
        }

        const int is_time_out = (ts == mts);
        const int is_col = collision_flag, is_jl = joint_limit_flag;
        const int is_es = d_err > cfg.thr_early;
        const int is_ds = d_err < cfg.thr_dist;
        const int is_rs = r_err < cfg.thr_rot;
        const int is_is = (!cfg.refine_flag) ? (imitate_err < cfg.thr_imit) : 0;
        T rw = -cfg.w_col * T(is_col) - cfg.w_jl * T(is_jl) - cfg.w_es * T(is_es);
        if (cfg.task_flag) {
            const bool near_t = d_err < cfg.zero_task;
            rw += (dist_rwd + cfg.dist_suc_r * T(is_ds) + (near_t ? rot_rwd : T(0)) +
                   (near_t ? cfg.rot_suc_r * T(is_rs) : T(0)) + cfg.task_suc_r * T(is_ds * is_rs));
        }
        if (cfg.imit_flag && !cfg.refine_flag) rw += (imitate_rwd + cfg.imit_suc_r * T(is_is));
        if (cfg.null_flag && !cfg.refine_flag) rw += null_rwd;
        if (l == 0) {
            reward[b] = rw;
            done[b] = (uint8_t)(is_col || is_es || is_time_out);
            const uint32_t lo32 = (uint32_t)is_col | ((uint32_t)is_jl << 8) | ((uint32_t)is_es << 16) | ((uint32_t)is_time_out << 24);
            const uint32_t hi32 = (uint32_t)is_ds | ((uint32_t)is_rs << 8) | ((uint32_t)is_is << 16) | ((uint32_t)(is_ds * is_rs) << 24);
            *reinterpret_cast<uint2*>(flags + b * 8) = make_uint2(lo32, hi32);
            if (errs) {
                errs[b * 4 + 0] = d_err; errs[b * 4 + 1] = r_err; errs[b * 4 + 2] = imitate_err; errs[b * 4 + 3] = null_obj;
            }
        }
        if (counters && l < 4) {
            const int add = l == 0 ? is_ds * is_rs : (l == 1 ? is_ds : (l == 2 ? is_rs : (cfg.refine_flag ? 0 : is_is)));
            counters[b * 4 + l] = cnt + add;
        }
    }
}

template <typename T, bool kStep>
static int launch_pirl(const RlitgPirlCfg* cfg, int64_t B, T* q, const float* action, int32_t* timestep,
                       const int32_t* max_timestep, const T* targets, const T* demos, int Tmax, const T* latent,
                       const Scene* scene, const double* base_off, T* obs, T* reward, uint8_t* done, uint8_t* flags,
                       int32_t* counters, T* errs, cudaStream_t st) {
    PirlCfgT<T> c = convert_cfg<T>(*cfg);
    // small batches (down to the reference's own mode of use, one environment): eight lanes per environment (kernel 2b).
    // Measured as 50 steps per CUDA graph (profiles/README.md): 5.9 vs 6.4 us at 1,024 envs, equal at 4,096, and
    // SLOWER from there on (10.6 vs 6.6 us at 16,384: 8 x the threads with redundant scalar work, 5.6 M warp
    // instructions against 1.1 M) -- so the switch point sits at 2,048; RLITG_PIRL_LANES_MAX moves it
    static const int64_t lanes_max = [] { const char* e = getenv("RLITG_PIRL_LANES_MAX"); return e ? atoll(e) : 2048ll; }();
    if (B <= lanes_max && c.O <= 1024) {
        auto kl = c.RS == 32 ? pirl_env_lanes_kernel<T, kStep, 32>
                             : (c.RS == 8 ? pirl_env_lanes_kernel<T, kStep, 8> : pirl_env_lanes_kernel<T, kStep, 7>);
        const size_t sm = sizeof(T) * kLaneEnvs * (size_t)(96 + 14 + c.O);
        int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kl), sm, "pirl_env_lanes_kernel");
        if (rc_attr) return rc_attr;
        SceneDev sdl{};
        int hs = 0;
        double offl[3] = {0, 0, 0};
        if (scene) { sdl = scene_dev(scene); hs = 1; offl[0] = base_off[0]; offl[1] = base_off[1]; offl[2] = base_off[2]; }
        const unsigned gl = (unsigned)((B + kLaneEnvs - 1) / kLaneEnvs);
        kl<<<gl, kLaneEnvs * 8, sm, st>>>(c, B, q, action, timestep, max_timestep, targets, demos, Tmax, latent, sdl, hs,
                                          (T)offl[0], (T)offl[1], (T)offl[2], obs, reward, done, flags, counters, errs);
        return check_launch("pirl_env_lanes_kernel");
    }
    // template axes: link poses in the observation (the Jacobian's joint frames then live in the staged row, not in
    // registers) and the row width of the target / demo tables.  The CTA shape stays 128 threads x 4 CTAs/SM: with the
    // staged observation rows (89 floats per env in the trained setting) shared memory allows 6 x 96 = 18 or 2 x 320 = 20
    // warps/SM at 96 registers, and both measured SLOWER at 1M envs (197 / 216 vs 168 us).
    const bool link_obs = c.obs_type == 1 || c.obs_type == 3;
    constexpr int tile = kEnvTile;
    constexpr int kMinB = sizeof(T) == 4 ? 4 : 1;
    using EnvK = void (*)(PirlCfgT<T>, int64_t, T*, const float*, int32_t*, const int32_t*, const T*, const T*, int, const T*,
                          SceneDev, int, T, T, T, T*, T*, uint8_t*, uint8_t*, int32_t*, T*);
    // the trained layout (obs type 1, one future target, no latent) takes the 128-bit staging + bulk-store path
    static const int pfast = [] { const char* e = getenv("RLITG_PIRL_FAST"); return e ? atoi(e) : 1; }();
    const bool fast = pfast && c.obs_type == 1 && c.n_future == 1 && c.z_dim == 0 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0;
    EnvK kern;
    constexpr int kMinBF = sizeof(T) == 4 ? 5 : 1;
#define RLITG_PIRL_PICK(MB, LO, FA) \
    (c.RS == 32 ? pirl_env_kernel<T, kStep, tile, MB, LO, 32, FA> \
                : (c.RS == 8 ? pirl_env_kernel<T, kStep, tile, MB, LO, 8, FA> : pirl_env_kernel<T, kStep, tile, MB, LO, 7, FA>))
    if (fast) kern = RLITG_PIRL_PICK(kMinBF, true, true);
    else if (link_obs) kern = RLITG_PIRL_PICK(kMinB, true, false);
    else kern = RLITG_PIRL_PICK(kMinB, false, false);
#undef RLITG_PIRL_PICK
    size_t smem = sizeof(T) * tile * (size_t)(fast ? c.O : (c.O | 1)) +
                        (scene ? kMaxSpheres * sizeof(float4) + 16 * sizeof(int) : 0);
    if (smem > 200 * 1024) return set_error(RLITG_ERR_SHAPE, "observation too long for the shared-memory tile");
    // resident CTAs per SM are set through the shared-memory request: FOUR (16 warps) -- five fit the trained layout, and
    // measured slower from 262k envs on (40.5 vs 37.2 us; 141.7 vs 132.0 us at 1M: less L1 for the strided q / action
    // rows); RLITG_PIRL_SMEM_MIN moves the floor
    static const size_t smem_min = [] { const char* e = getenv("RLITG_PIRL_SMEM_MIN"); return e ? (size_t)atoll(e) : (size_t)46800; }();
    if (sizeof(T) == 4 && smem < smem_min) smem = smem_min;
    {
        int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "pirl_env_kernel");
        if (rc_attr) return rc_attr;
    }
    SceneDev sd{};
    int has_scene = 0;
    double off[3] = {0, 0, 0};
    if (scene) {
        sd = scene_dev(scene);
        has_scene = 1;
        off[0] = base_off[0]; off[1] = base_off[1]; off[2] = base_off[2];
    }
    const unsigned grid = (unsigned)((B + tile - 1) / tile);
    kern<<<grid, tile, smem, st>>>(c, B, q, action, timestep, max_timestep, targets, demos, Tmax, latent, sd,
                                       has_scene, (T)off[0], (T)off[1], (T)off[2], obs, reward, done, flags, counters, errs);
    return check_launch("pirl_env_kernel");
}

static int check_cfg(const RlitgPirlCfg* cfg) {
    if (!cfg) return set_error(RLITG_ERR_NULL, "cfg is NULL");
    if (cfg->obs_type < 0 || cfg->obs_type > 3) return set_error(RLITG_ERR_ARG, "obs_type must be 0..3");
    if (cfg->n_future < 0 || cfg->n_future > 64 || cfg->z_dim < 0 || cfg->z_dim > 4096)
        return set_error(RLITG_ERR_ARG, "n_future / z_dim out of range");
    if (cfg->row_floats != 0 && cfg->row_floats != 7 && cfg->row_floats != 8 && cfg->row_floats != 32)
        return set_error(RLITG_ERR_ARG, "row_floats must be 0 / 7 (packed rows), 8 (32-byte rows) or 32 (128-byte step records)");
    return RLITG_OK;
}

}  // namespace rlitg

using namespace rlitg;

extern "C" {

RLITG_API int rlitg_pirl_obs_dim(const RlitgPirlCfg* cfg) {
    if (check_cfg(cfg)) return RLITG_ERR_ARG;
    return obs_dim_of(*cfg);
}

RLITG_API int rlitg_pirl_env_step(const RlitgPirlCfg* cfg, int64_t B, int dtype, void* q, const float* action,
                                  int32_t* timestep, const int32_t* max_timestep, const void* targets,
                                  const void* demos, int Tmax, const void* latent, const RlitgScene* scene,
                                  const double* base_off, void* obs, void* reward, uint8_t* done, uint8_t* flags,
                                  int32_t* counters, void* errs, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (B < 0 || Tmax < 1) return set_error(RLITG_ERR_SHAPE, "B must be >= 0 and Tmax >= 1");
    if (B == 0) return RLITG_OK;
    if (!q || !action || !timestep || !max_timestep || !targets || !obs || !reward || !done || !flags)
        return set_error(RLITG_ERR_NULL, "a required buffer is NULL");
    if (cfg->z_dim > 0 && !latent) return set_error(RLITG_ERR_NULL, "latent is NULL but z_dim > 0");
    const Scene* sc = reinterpret_cast<const Scene*>(scene);
    if (sc && !scene_ready(sc, true, true)) return set_error(RLITG_ERR_ARG, "scene needs a grid and a sphere table");
    if (sc && !base_off) return set_error(RLITG_ERR_NULL, "base_off is NULL");
    if ((reinterpret_cast<uintptr_t>(flags) & 7) || (counters && (reinterpret_cast<uintptr_t>(counters) & 15)))
        return set_error(RLITG_ERR_ALIGN, "flags must be 8-byte and counters 16-byte aligned");
    if (cfg->row_floats == 8 && ((reinterpret_cast<uintptr_t>(targets) & 31) || (demos && (reinterpret_cast<uintptr_t>(demos) & 31))))
        return set_error(RLITG_ERR_ALIGN, "padded target / demo rows (row_floats = 8) must be 32-byte aligned");
    if (cfg->row_floats == 32 && (reinterpret_cast<uintptr_t>(targets) & 127))
        return set_error(RLITG_ERR_ALIGN, "step records (row_floats = 32) must be 128-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32)
        return launch_pirl<float, true>(cfg, B, (float*)q, action, timestep, max_timestep, (const float*)targets,
                                        (const float*)demos, Tmax, (const float*)latent, sc, base_off, (float*)obs,
                                        (float*)reward, done, flags, counters, (float*)errs, st);
    if (dtype == RLITG_F64)
        return launch_pirl<double, true>(cfg, B, (double*)q, action, timestep, max_timestep, (const double*)targets,
                                         (const double*)demos, Tmax, (const double*)latent, sc, base_off, (double*)obs,
                                         (double*)reward, done, flags, counters, (double*)errs, st);
    return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
}

// ---- bound step: the argument list of rlitg_pirl_env_step held by a handle ------------------------------------------
// At 16,384 envs the kernel takes ~5 us and a binding that marshals twenty arguments per call (ctypes: ~4 us) makes the
// step launch-bound.  A vectorised env steps the same buffers thousands of times: bind them once, pass the action.
struct PirlBound {
    RlitgPirlCfg cfg;
    int64_t B;
    int dtype, Tmax;
    void *q, *obs, *reward, *errs;
    int32_t *timestep, *counters;
    const int32_t* max_timestep;
    const void *targets, *demos, *latent;
    const RlitgScene* scene;
    double base_off[3];
    bool has_off;
    uint8_t *done, *flags;
};

RLITG_API int rlitg_pirl_bind(const RlitgPirlCfg* cfg, int64_t B, int dtype, void* q, int32_t* timestep,
                              const int32_t* max_timestep, const void* targets, const void* demos, int Tmax,
                              const void* latent, const RlitgScene* scene, const double* base_off, void* obs, void* reward,
                              uint8_t* done, uint8_t* flags, int32_t* counters, void* errs, RlitgPirlBound** out) {
    if (!out) return set_error(RLITG_ERR_NULL, "out is NULL");
    int rc = check_cfg(cfg);
    if (rc) return rc;
    PirlBound* b = new PirlBound();
    b->cfg = *cfg; b->B = B; b->dtype = dtype; b->Tmax = Tmax;
    b->q = q; b->obs = obs; b->reward = reward; b->errs = errs;
    b->timestep = timestep; b->counters = counters; b->max_timestep = max_timestep;
    b->targets = targets; b->demos = demos; b->latent = latent; b->scene = scene;
    b->has_off = base_off != nullptr;
    for (int k = 0; k < 3; k++) b->base_off[k] = base_off ? base_off[k] : 0.0;
    b->done = done; b->flags = flags;
    *out = reinterpret_cast<RlitgPirlBound*>(b);
    return RLITG_OK;
}

RLITG_API int rlitg_pirl_step_bound(const RlitgPirlBound* h, const float* action, void* stream) {
    if (!h) return set_error(RLITG_ERR_NULL, "bound handle is NULL");
    const PirlBound* b = reinterpret_cast<const PirlBound*>(h);
    return rlitg_pirl_env_step(&b->cfg, b->B, b->dtype, b->q, action, b->timestep, b->max_timestep, b->targets, b->demos, b->Tmax,
                               b->latent, b->scene, b->has_off ? b->base_off : nullptr, b->obs, b->reward, b->done, b->flags,
                               b->counters, b->errs, stream);
}

RLITG_API void rlitg_pirl_bound_destroy(RlitgPirlBound* h) { delete reinterpret_cast<PirlBound*>(h); }

RLITG_API int rlitg_pirl_obs(const RlitgPirlCfg* cfg, int64_t B, int dtype, const void* q, const int32_t* timestep,
                             const int32_t* max_timestep, const void* targets, int Tmax, const void* latent,
                             void* obs, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (B < 0 || Tmax < 1) return set_error(RLITG_ERR_SHAPE, "B must be >= 0 and Tmax >= 1");
    if (B == 0) return RLITG_OK;
    if (!q || !timestep || !max_timestep || !targets || !obs) return set_error(RLITG_ERR_NULL, "a required buffer is NULL");
    if (cfg->z_dim > 0 && !latent) return set_error(RLITG_ERR_NULL, "latent is NULL but z_dim > 0");
    if (cfg->row_floats == 8 && (reinterpret_cast<uintptr_t>(targets) & 31))
        return set_error(RLITG_ERR_ALIGN, "padded target rows (row_floats = 8) must be 32-byte aligned");
    if (cfg->row_floats == 32 && (reinterpret_cast<uintptr_t>(targets) & 127))
        return set_error(RLITG_ERR_ALIGN, "step records (row_floats = 32) must be 128-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32)
        return launch_pirl<float, false>(cfg, B, (float*)q, nullptr, (int32_t*)timestep, max_timestep,
                                         (const float*)targets, nullptr, Tmax, (const float*)latent, nullptr, nullptr,
                                         (float*)obs, nullptr, nullptr, nullptr, nullptr, nullptr, st);
    if (dtype == RLITG_F64)
        return launch_pirl<double, false>(cfg, B, (double*)q, nullptr, (int32_t*)timestep, max_timestep,
                                          (const double*)targets, nullptr, Tmax, (const double*)latent, nullptr, nullptr,
                                          (double*)obs, nullptr, nullptr, nullptr, nullptr, nullptr, st);
    return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
}

}  // extern "C"
