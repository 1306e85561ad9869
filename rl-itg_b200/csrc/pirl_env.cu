// Kernel 2: the PIRL path-wise IK environment step, fused.
//
// One thread per environment does what PirlEnv_Fetch.step does with three ROS service round-trips
// (pirl/script/envs/PirlEnv.py:209-289; cc_srv :196, allfk_srv :216, jaco_srv :536):
//   _set_action :184-207 -> all-link FK -> _get_obs :292-320 -> _compute_task_reward :495-513,
//   _compute_imitate_reward :515-526, _compute_null_space_imitate_reward :528-539 (+ jaco_srv,
//   pirl/src/kinematic_solver.cpp:332-351) -> reward / done / info flags :226-282.
// The same observation math exists in C++ at torm/src/pirl_interpolator.cpp:176-274 (float32 output).
//
// HBM-bound in principle (536 B/env at the trained setting), launch/latency-bound at 16k envs.  Observation rows
// are staged in shared memory (odd row stride -> conflict-free) and written back coalesced.
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
    int RS;  // floats per target / demo row: 7, or 8 (rows padded to 32 bytes: every row is one or two whole DRAM sectors)
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
    o.RS = c.row_floats == 8 ? 8 : 7;
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

// one 7-value row of the target / demo tables.  Padded rows (RS = 8) are 32-byte aligned: two 128-bit loads that touch
// exactly one sector (FP32), instead of seven 4-byte loads of a row that straddles two (the 28-byte rows of a
// [B,Tmax,7] table start at multiples of 28 bytes: 1.41 x the algorithmic DRAM traffic at 1M envs, round 1)
template <typename T>
RLITG_DEV void load_row7(const T* __restrict__ row, int RS, T (&o)[7]) {
    if (RS == 8) {
        if constexpr (sizeof(T) == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(row)), b = __ldg(reinterpret_cast<const float4*>(row) + 1);
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z;
        } else {
            const double2* r2 = reinterpret_cast<const double2*>(row);
            const double2 a = __ldg(r2), b = __ldg(r2 + 1), c = __ldg(r2 + 2), d = __ldg(r2 + 3);
            o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 7; k++) o[k] = row[k];
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

template <typename T>
struct PirlVisit {
    T* row;
    int link_off, ee_off;
    bool do_coll;
    const SceneDev* sc;
    const float4* sph;
    const int* link_start;
    V3<T> off;
    bool hit;
    FkJac<T, 7> fj;
    template <int K>
    RLITG_DEV void operator()(jidx<K> k, const FrameC<T>& f) {
        fj(k, f);
        if (link_off >= 0) {  // calc_pos_and_rot :322-333: p, R[:,0], R[:,1]
            T* o = row + link_off + 9 * K;
            o[0] = f.p.x; o[1] = f.p.y; o[2] = f.p.z;
            o[3] = f.c0.x; o[4] = f.c0.y; o[5] = f.c0.z;
            o[6] = f.c1.x; o[7] = f.c1.y; o[8] = f.c1.z;
        }
        if (K == 7 && ee_off >= 0) {
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

template <typename T, bool kStep>
__global__ void __launch_bounds__(kEnvTile, sizeof(T) == 4 ? 4 : 1) pirl_env_kernel(
    PirlCfgT<T> cfg, int64_t B, T* __restrict__ q, const float* __restrict__ action, int32_t* __restrict__ timestep,
    const int32_t* __restrict__ max_timestep, const T* __restrict__ targets, const T* __restrict__ demos, int Tmax,
    const T* __restrict__ latent, SceneDev sc, int has_scene, T ox, T oy, T oz, T* __restrict__ obs,
    T* __restrict__ reward, uint8_t* __restrict__ done, uint8_t* __restrict__ flags, int32_t* __restrict__ counters,
    T* __restrict__ errs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int O = cfg.O, OS = O | 1;
    T* s_obs = reinterpret_cast<T*>(smem_raw);  // [kEnvTile, OS]
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_start[kNumLinks + 1];
    if (has_scene) {
        for (int i = threadIdx.x; i < sc.S; i += blockDim.x) s_sph[i] = sc.spheres[i];
        if (threadIdx.x <= kNumLinks) s_start[threadIdx.x] = sc.link_start[threadIdx.x];
        __syncthreads();
    }
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kEnvTile;
    const int n = (int)min((int64_t)kEnvTile, B - base);
    const int64_t b = base + tid;
    if (tid < n) {
        T* row = s_obs + tid * OS;
        T cur[7];
#pragma unroll
        for (int j = 0; j < 7; j++) cur[j] = q[b * 7 + j];
        int ts = timestep[b];
        const int mts = max_timestep[b];
        const int RS = cfg.RS;
        const T* tg = targets + (int64_t)b * Tmax * RS;

        const int ot = cfg.obs_type;
        PirlVisit<T> v;
        v.row = row;
        v.link_off = (ot == 1) ? 7 : (ot == 3 ? 0 : -1);
        v.ee_off = (ot == 2) ? 7 : -1;
        v.sc = &sc; v.sph = s_sph; v.link_start = s_start;
        v.off = {ox, oy, oz};
        v.hit = false;
        v.do_coll = false;

        bool collision_flag = false, joint_limit_flag = false;
        // loads whose addresses are known once the timestep is: issued before the chain walk so that their latency
        // overlaps the arithmetic instead of adding three dependent round trips at the end of the thread
        T pf_tp[7], pf_dm[7];
        int4 pf_cnt = make_int4(0, 0, 0, 0);
        if constexpr (kStep) {
            ts += 1;
            {
                const int tsp = min(max(ts, 0), Tmax - 1);
                load_row7(tg + (size_t)tsp * RS, RS, pf_tp);
                if (demos != nullptr) load_row7(demos + ((int64_t)b * Tmax + tsp) * RS, RS, pf_dm);
                if (counters) pf_cnt = *reinterpret_cast<const int4*>(counters + b * 4);
            }
            // _set_action :184-207
            T cl[7];
            bool no_out = true;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                T o2 = cur[j] + T(action[b * 7 + j]);
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
            if (cfg.train_mode) {
                if (no_out) {
                    v.do_coll = has_scene != 0;
                    walk_chain<T, 7>(cl, v);
                    if (v.hit) {
                        collision_flag = true;
                        v.do_coll = false;
                        walk_chain<T, 7>(cur, v);  // the move is rejected: pose of the unchanged configuration
                    } else {
#pragma unroll
                        for (int j = 0; j < 7; j++) cur[j] = cl[j];
                    }
                } else {
                    joint_limit_flag = true;
                    walk_chain<T, 7>(cur, v);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 7; j++) cur[j] = cl[j];
                walk_chain<T, 7>(cur, v);
            }
#pragma unroll
            for (int j = 0; j < 7; j++) q[b * 7 + j] = cur[j];
            timestep[b] = ts;
        } else {
            walk_chain<T, 7>(cur, v);
        }
        const FrameC<T>& ee = v.fj.ee;

        // ---- _get_obs :292-320
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
            load_row7(tg + (size_t)t * RS, RS, t7);
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
            if (demos != nullptr) {
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
                    fill_jac<T, 7>(v.fj, J);
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
    __syncthreads();
    // coalesced copy-out of the staged rows
    T* dst = obs + base * O;
    if (sizeof(T) == 4 && (O & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        // 128-bit stores: a row holds O / 4 whole float4 groups, so a group never straddles two rows
        const int G = O >> 2;
        int r = tid / G, g = tid - r * G;
        const int dr = kEnvTile / G, dg = kEnvTile - dr * G;
        for (int i = tid; i < n * G; i += kEnvTile) {
            const T* sp = s_obs + r * OS + 4 * g;
            st_stream(reinterpret_cast<float4*>(dst) + i, make_float4((float)sp[0], (float)sp[1], (float)sp[2], (float)sp[3]));
            r += dr; g += dg;
            if (g >= G) { g -= G; r++; }
        }
    } else {
        // (row, column) advance incrementally instead of dividing by the runtime O
        int r = tid / O, c = tid - r * O;
        const int dr = kEnvTile / O, dc = kEnvTile - dr * O;
        for (int i = tid; i < n * O; i += kEnvTile) {
            st_stream(dst + i, s_obs[r * OS + c]);
            r += dr; c += dc;
            if (c >= O) { c -= O; r++; }
        }
    }
}

template <typename T, bool kStep>
static int launch_pirl(const RlitgPirlCfg* cfg, int64_t B, T* q, const float* action, int32_t* timestep,
                       const int32_t* max_timestep, const T* targets, const T* demos, int Tmax, const T* latent,
                       const Scene* scene, const double* base_off, T* obs, T* reward, uint8_t* done, uint8_t* flags,
                       int32_t* counters, T* errs, cudaStream_t st) {
    PirlCfgT<T> c = convert_cfg<T>(*cfg);
    const size_t smem = sizeof(T) * kEnvTile * (size_t)(c.O | 1);
    if (smem > 200 * 1024) return set_error(RLITG_ERR_SHAPE, "observation too long for the shared-memory tile");
    auto kern = pirl_env_kernel<T, kStep>;
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
    const unsigned grid = (unsigned)((B + kEnvTile - 1) / kEnvTile);
    kern<<<grid, kEnvTile, smem, st>>>(c, B, q, action, timestep, max_timestep, targets, demos, Tmax, latent, sd,
                                       has_scene, (T)off[0], (T)off[1], (T)off[2], obs, reward, done, flags, counters, errs);
    return check_launch("pirl_env_kernel");
}

static int check_cfg(const RlitgPirlCfg* cfg) {
    if (!cfg) return set_error(RLITG_ERR_NULL, "cfg is NULL");
    if (cfg->obs_type < 0 || cfg->obs_type > 3) return set_error(RLITG_ERR_ARG, "obs_type must be 0..3");
    if (cfg->n_future < 0 || cfg->n_future > 64 || cfg->z_dim < 0 || cfg->z_dim > 4096)
        return set_error(RLITG_ERR_ARG, "n_future / z_dim out of range");
    if (cfg->row_floats != 0 && cfg->row_floats != 7 && cfg->row_floats != 8)
        return set_error(RLITG_ERR_ARG, "row_floats must be 0 / 7 (packed rows) or 8 (rows padded to 32 bytes)");
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
    if (cfg->row_floats == 8 && ((reinterpret_cast<uintptr_t>(targets) & 15) || (demos && (reinterpret_cast<uintptr_t>(demos) & 15))))
        return set_error(RLITG_ERR_ALIGN, "padded target / demo rows (row_floats = 8) must be 16-byte aligned");
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

RLITG_API int rlitg_pirl_obs(const RlitgPirlCfg* cfg, int64_t B, int dtype, const void* q, const int32_t* timestep,
                             const int32_t* max_timestep, const void* targets, int Tmax, const void* latent,
                             void* obs, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (B < 0 || Tmax < 1) return set_error(RLITG_ERR_SHAPE, "B must be >= 0 and Tmax >= 1");
    if (B == 0) return RLITG_OK;
    if (!q || !timestep || !max_timestep || !targets || !obs) return set_error(RLITG_ERR_NULL, "a required buffer is NULL");
    if (cfg->z_dim > 0 && !latent) return set_error(RLITG_ERR_NULL, "latent is NULL but z_dim > 0");
    if (cfg->row_floats == 8 && (reinterpret_cast<uintptr_t>(targets) & 15))
        return set_error(RLITG_ERR_ALIGN, "padded target rows (row_floats = 8) must be 16-byte aligned");
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
