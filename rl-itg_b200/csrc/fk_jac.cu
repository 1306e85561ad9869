// Kernel 1: batched Fetch forward kinematics + geometric Jacobian.
//
// Replaces, for a whole batch at once, the reference's per-configuration KDL calls
//   TormIKSolver::fkSolver / allfkSolver / getJacobian   torm/src/torm_ik_solver.cpp:156-175
//   KinematicsSolverSrv::fkSolver_srv / allLinkfkSolver_srv   pirl/src/kinematic_solver.cpp:206-261
//
// One thread per configuration walks the chain in registers (fetch_chain.cuh).  HBM-bound: 4*(D + 12 + 6D)
// bytes per configuration in FP32 (244 B at D = 7).  Array-of-structs outputs ([B,3], [B,9], [B,6,D], the
// layout a drop-in caller holds) are staged per 128-configuration tile in shared memory -- 64-bit stores at a
// 42-word thread stride are bank-conflict free -- and drained with TMA bulk stores (cp.async.bulk
// shared::cta -> global), one per output array, so the write-back is fully coalesced and asynchronous to the
// threads.  Struct-of-arrays outputs ([3,B], [9,B], [6D,B]) are stored directly (already coalesced).
#include <algorithm>
#include <cstdlib>

#include "fetch_chain.cuh"
#include "rlitg_internal.h"

namespace rlitg {

constexpr int kTile = 128;  // configurations per CTA

template <typename T>
RLITG_DEV T ld_stream(const T* p) { return __ldg(p); }  // (__ldcs, evict-first, measured 8 % slower on the whole step)

// vectorised shared-memory tile write: N values per thread at thread stride N (see header comment)
template <typename T, int N>
RLITG_DEV void stage_row(T* tile, int tid, const T (&v)[N]) {
    T* dst = tile + tid * N;
    if constexpr (sizeof(T) == 4 && (N % 4) == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if constexpr (sizeof(T) == 4 && (N % 2) == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 2) *reinterpret_cast<float2*>(dst + i) = make_float2(v[i], v[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) dst[i] = v[i];
    }
}

// drain a staged [n, N] tile to global memory
template <typename T>
RLITG_DEV void drain_tile(T* gdst, const T* tile, int n_elems, bool bulk, int tid, int nthreads) {
    if (bulk) {
        if (tid == 0) bulk_store(gdst, tile, (uint32_t)(n_elems * sizeof(T)));
    } else {
        for (int i = tid; i < n_elems; i += nthreads) st_stream(gdst + i, tile[i]);
    }
}

template <typename T, int D, bool QSOA, bool OSOA>
__global__ void __launch_bounds__(kTile) fk_jac_kernel(const T* __restrict__ q, int64_t B, T* __restrict__ pos,
                                                       T* __restrict__ rot, T* __restrict__ quat,
                                                       T* __restrict__ J, int allow_bulk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sJ = reinterpret_cast<T*>(smem_raw);   // [kTile, 6D]
    T* sRot = sJ + kTile * 6 * D;             // [kTile, 9]
    T* sQuat = sRot + kTile * 9;              // [kTile, 4]
    T* sPos = sQuat + kTile * 4;              // [kTile, 3]
    T* sQ = sPos + kTile * 3;                 // [kTile, D]  (AoS input staging)

    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kTile;
    const int n = (int)min((int64_t)kTile, B - base);
    const int64_t b = base + tid;
    const bool live = tid < n;

    T qv[D];
    if constexpr (QSOA) {
#pragma unroll
        for (int j = 0; j < D; j++) qv[j] = live ? ld_stream(q + (int64_t)j * B + b) : T(0);
    } else {
        // coalesced tile load, then a conflict-free strided read (D is odd or 8 -> at worst 8-way on D = 8)
        const T* src = q + base * D;
        for (int i = tid; i < n * D; i += kTile) sQ[i] = ld_stream(src + i);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < D; j++) qv[j] = live ? sQ[tid * D + j] : T(0);
    }

    FkJac<T, D> fj;
    walk_chain<T, D>(qv, fj);

    if constexpr (OSOA) {
        if (!live) return;
        if (pos) { st_stream(pos + b, fj.ee.p.x); st_stream(pos + B + b, fj.ee.p.y); st_stream(pos + 2 * B + b, fj.ee.p.z); }
        if (rot) {
            st_stream(rot + 0 * B + b, fj.ee.c0.x); st_stream(rot + 1 * B + b, fj.ee.c1.x); st_stream(rot + 2 * B + b, fj.ee.c2.x);
            st_stream(rot + 3 * B + b, fj.ee.c0.y); st_stream(rot + 4 * B + b, fj.ee.c1.y); st_stream(rot + 5 * B + b, fj.ee.c2.y);
            st_stream(rot + 6 * B + b, fj.ee.c0.z); st_stream(rot + 7 * B + b, fj.ee.c1.z); st_stream(rot + 8 * B + b, fj.ee.c2.z);
        }
        if (quat) {
            T x, y, z, w;
            rot_to_quat_kdl(fj.ee.c0, fj.ee.c1, fj.ee.c2, x, y, z, w);
            st_stream(quat + b, x); st_stream(quat + B + b, y); st_stream(quat + 2 * B + b, z); st_stream(quat + 3 * B + b, w);
        }
        if (J) {
            T jv[6 * D];
            fill_jac(fj, jv);
#pragma unroll
            for (int i = 0; i < 6 * D; i++) st_stream(J + (int64_t)i * B + b, jv[i]);
        }
    } else {
        if (J) {
            T jv[6 * D];
            fill_jac(fj, jv);
            stage_row<T, 6 * D>(sJ, tid, jv);
        }
        if (rot) {
            const T rv[9] = {fj.ee.c0.x, fj.ee.c1.x, fj.ee.c2.x, fj.ee.c0.y, fj.ee.c1.y,
                             fj.ee.c2.y, fj.ee.c0.z, fj.ee.c1.z, fj.ee.c2.z};
            stage_row<T, 9>(sRot, tid, rv);
        }
        if (quat) {
            T qq[4];
            rot_to_quat_kdl(fj.ee.c0, fj.ee.c1, fj.ee.c2, qq[0], qq[1], qq[2], qq[3]);
            stage_row<T, 4>(sQuat, tid, qq);
        }
        if (pos) {
            const T pv[3] = {fj.ee.p.x, fj.ee.p.y, fj.ee.p.z};
            stage_row<T, 3>(sPos, tid, pv);
        }
        // bulk (TMA) drain needs 16-byte sizes: a full tile always qualifies, a partial tile falls back
        const bool bulk = allow_bulk && n == kTile;
        if (bulk) fence_proxy_async_smem();
        __syncthreads();
        if (J) drain_tile(J + base * 6 * D, sJ, n * 6 * D, bulk, tid, kTile);
        if (rot) drain_tile(rot + base * 9, sRot, n * 9, bulk, tid, kTile);
        if (quat) drain_tile(quat + base * 4, sQuat, n * 4, bulk, tid, kTile);
        if (pos) drain_tile(pos + base * 3, sPos, n * 3, bulk, tid, kTile);
        if (bulk && tid == 0) {
            bulk_commit();
            bulk_wait_read_all();
        }
    }
}


// ---- persistent variant for array-of-structs outputs ------------------------------------------------------
// One CTA per SM slot loops over 128-configuration tiles: the next tile's joint vector is prefetched into
// registers before the current tile is walked, and the staged outputs drain through TMA bulk stores while the
// next tile computes (NSTAGE staging buffers; a buffer is reused once its bulk group has finished reading).
// This removes the load -> compute -> store serialisation and the CTA launch gaps of the one-tile-per-CTA kernel.
template <int N>
RLITG_DEV void bulk_wait_read_pending() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

template <typename T, int D, bool QSOA, int NSTAGE>
__global__ void __launch_bounds__(kTile) fk_jac_persist_kernel(const T* __restrict__ q, int64_t B, int64_t ntiles,
                                                               T* __restrict__ pos, T* __restrict__ rot,
                                                               T* __restrict__ quat, T* __restrict__ J, int dbg) {
    constexpr int kStageElems = kTile * (6 * D + 9 + 4 + 3);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* stage0 = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x;
    int64_t tile = blockIdx.x;
    // joint vectors are prefetched two tiles ahead: under the kernel's write stream a read takes longer than one
    // tile's worth of arithmetic (a one-tile prefetch still showed 20 % long-scoreboard stalls at its first use)
    T qn[D], qn2[D];
    auto load_q = [&](int64_t t, T (&dst)[D]) {
        const int64_t b = t * kTile + tid;
#pragma unroll
        for (int j = 0; j < D; j++) dst[j] = QSOA ? ld_stream(q + (int64_t)j * B + b) : ld_stream(q + b * D + j);
    };
    if (tile < ntiles) load_q(tile, qn);
    if (tile + gridDim.x < ntiles) load_q(tile + gridDim.x, qn2);
    for (int it = 0; tile < ntiles; tile += gridDim.x, it++) {
        T qv[D];
#pragma unroll
        for (int j = 0; j < D; j++) { qv[j] = qn[j]; qn[j] = qn2[j]; }
        if (tile + 2 * (int64_t)gridDim.x < ntiles) load_q(tile + 2 * (int64_t)gridDim.x, qn2);

        FkJac<T, D> fj;
        walk_chain<T, D>(qv, fj);

        T* sJ = stage0 + (it % NSTAGE) * kStageElems;
        T* sRot = sJ + kTile * 6 * D;
        T* sQuat = sRot + kTile * 9;
        T* sPos = sQuat + kTile * 4;
        if (it >= NSTAGE) {  // the bulk group that last read this buffer must have drained
            if (tid == 0) bulk_wait_read_pending<NSTAGE - 1>();
            __syncthreads();
        }
        if (J) {
            T jv[6 * D];
            fill_jac(fj, jv);
            stage_row<T, 6 * D>(sJ, tid, jv);
        }
        if (rot) {
            const T rv[9] = {fj.ee.c0.x, fj.ee.c1.x, fj.ee.c2.x, fj.ee.c0.y, fj.ee.c1.y,
                             fj.ee.c2.y, fj.ee.c0.z, fj.ee.c1.z, fj.ee.c2.z};
            stage_row<T, 9>(sRot, tid, rv);
        }
        if (quat) {
            T qq[4];
            rot_to_quat_kdl(fj.ee.c0, fj.ee.c1, fj.ee.c2, qq[0], qq[1], qq[2], qq[3]);
            stage_row<T, 4>(sQuat, tid, qq);
        }
        if (pos) {
            const T pv[3] = {fj.ee.p.x, fj.ee.p.y, fj.ee.p.z};
            stage_row<T, 3>(sPos, tid, pv);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0 && !(dbg & 1)) {
            const int64_t base = tile * kTile;
            if (J) bulk_store(J + base * 6 * D, sJ, (uint32_t)(kTile * 6 * D * sizeof(T)));
            if (rot) bulk_store(rot + base * 9, sRot, (uint32_t)(kTile * 9 * sizeof(T)));
            if (quat) bulk_store(quat + base * 4, sQuat, (uint32_t)(kTile * 4 * sizeof(T)));
            if (pos) bulk_store(pos + base * 3, sPos, (uint32_t)(kTile * 3 * sizeof(T)));
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_read_all();
}

// ---- all-link FK: [B, D+1, 7] = (pos, quat xyzw) per segment tip (allLinkfkSolver_srv) --------------------
template <typename T, int D>
struct AllFkVisit {
    T* row;  // this thread's padded shared-memory row
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        T x, y, z, w;
        rot_to_quat_kdl(f.c0, f.c1, f.c2, x, y, z, w);
        T* o = row + K * 7;
        o[0] = f.p.x; o[1] = f.p.y; o[2] = f.p.z; o[3] = x; o[4] = y; o[5] = z; o[6] = w;
    }
};

// persistent all-link FK (array-of-structs input or output, >= 1024 full tiles): same structure as
// fk_jac_persist_kernel -- registers -> conflict-light 128-bit staging stores -> TMA bulk drain, q prefetched.
template <typename T, int D>
struct AllFkRegs {
    T v[(D + 1) * 7];
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        T x, y, z, w;
        rot_to_quat_kdl(f.c0, f.c1, f.c2, x, y, z, w);
        v[K * 7 + 0] = f.p.x; v[K * 7 + 1] = f.p.y; v[K * 7 + 2] = f.p.z;
        v[K * 7 + 3] = x; v[K * 7 + 4] = y; v[K * 7 + 5] = z; v[K * 7 + 6] = w;
    }
};

template <typename T, int D, bool QSOA, int NSTAGE>
__global__ void __launch_bounds__(kTile) allfk_persist_kernel(const T* __restrict__ q, int64_t B, int64_t ntiles,
                                                              T* __restrict__ out) {
    constexpr int N = (D + 1) * 7;
    constexpr int kStageElems = kTile * N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* stage0 = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x;
    int64_t tile = blockIdx.x;
    T qn[D], qn2[D];
    auto load_q = [&](int64_t t, T (&dst)[D]) {
        const int64_t b = t * kTile + tid;
#pragma unroll
        for (int j = 0; j < D; j++) dst[j] = QSOA ? ld_stream(q + (int64_t)j * B + b) : ld_stream(q + b * D + j);
    };
    if (tile < ntiles) load_q(tile, qn);
    if (tile + gridDim.x < ntiles) load_q(tile + gridDim.x, qn2);
    for (int it = 0; tile < ntiles; tile += gridDim.x, it++) {
        T qv[D];
#pragma unroll
        for (int j = 0; j < D; j++) { qv[j] = qn[j]; qn[j] = qn2[j]; }
        if (tile + 2 * (int64_t)gridDim.x < ntiles) load_q(tile + 2 * (int64_t)gridDim.x, qn2);
        AllFkRegs<T, D> r;
        walk_chain<T, D>(qv, r);
        T* sOut = stage0 + (it % NSTAGE) * kStageElems;
        if (it >= NSTAGE) {
            if (tid == 0) bulk_wait_read_pending<NSTAGE - 1>();
            __syncthreads();
        }
        stage_row<T, N>(sOut, tid, r.v);
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_store(out + tile * kStageElems, sOut, (uint32_t)(kStageElems * sizeof(T)));
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_read_all();
}

template <typename T, int D, bool QSOA>
__global__ void __launch_bounds__(kTile) allfk_kernel(const T* __restrict__ q, int64_t B, T* __restrict__ out) {
    constexpr int N = (D + 1) * 7, NP = N + 1;  // padded row stride: odd -> conflict-free scalar stores
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sOut = reinterpret_cast<T*>(smem_raw);  // [kTile, NP]
    T* sQ = sOut + kTile * NP;
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kTile;
    const int n = (int)min((int64_t)kTile, B - base);
    const bool live = tid < n;
    T qv[D];
    if constexpr (QSOA) {
#pragma unroll
        for (int j = 0; j < D; j++) qv[j] = live ? ld_stream(q + (int64_t)j * B + base + tid) : T(0);
    } else {
        const T* src = q + base * D;
        for (int i = tid; i < n * D; i += kTile) sQ[i] = ld_stream(src + i);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < D; j++) qv[j] = live ? sQ[tid * D + j] : T(0);
    }
    AllFkVisit<T, D> v{sOut + tid * NP};
    walk_chain<T, D>(qv, v);
    __syncthreads();
    T* dst = out + base * N;
    for (int i = tid; i < n * N; i += kTile) {  // flat, fully coalesced; N is a compile-time constant
        const int r = i / N, c = i - r * N;
        dst[i] = sOut[r * NP + c];
    }
}

// ---- launchers -------------------------------------------------------------------------------------------
static int fk_variant() {  // RLITG_FK_VARIANT: 0 = tile per CTA, 1 / 2 = persistent with 1 / 2 staging buffers
    static int v = [] {
        const char* e = getenv("RLITG_FK_VARIANT");
        return e ? atoi(e) : 2;
    }();
    return v;
}

template <typename T, int D, bool QSOA, int NSTAGE>
static int launch_fk_persist(const T* q, int64_t B, int64_t ntiles, T* pos, T* rot, T* quat, T* J, cudaStream_t st) {
    const size_t smem = sizeof(T) * NSTAGE * kTile * (6 * D + 9 + 4 + 3);
    auto kern = fk_jac_persist_kernel<T, D, QSOA, NSTAGE>;
    int rc0 = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "fk_jac_persist_kernel");
    if (rc0) return rc0;
    const int num_sms = current_sm_count();
    const int64_t grid = ntiles;
    static const int dbg = [] { const char* e = getenv("RLITG_FK_DBG"); return e ? atoi(e) : 0; }();  // 1 = no stores
    // Two CTAs per SM measured best on B200 (1M configs: 45.9 us vs 48-49 us with 3-7 CTAs/SM; 8M: 6.14 TB/s):
    // the round-robin tile order keeps the set of tiles in flight -- the HBM write window -- narrow.
    static const int mult = [] { const char* e = getenv("RLITG_FK_CTAS"); return e ? atoi(e) : 2; }();
    const int64_t g2 = std::min<int64_t>(std::min<int64_t>(ntiles, (int64_t)num_sms * mult), grid);
    kern<<<(unsigned)g2, kTile, smem, st>>>(q, B, ntiles, pos, rot, quat, J, dbg);
    return check_launch("fk_jac_persist_kernel");
}

template <typename T, int D, bool QSOA, bool OSOA>
static int launch_fk_jac_t(const T* q, int64_t B, T* pos, T* rot, T* quat, T* J, cudaStream_t st) {
    const size_t smem = sizeof(T) * kTile * (6 * D + 9 + 4 + 3 + D);
    auto kern = fk_jac_kernel<T, D, QSOA, OSOA>;
    int rc0 = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "fk_jac_kernel");
    if (rc0) return rc0;
    // TMA bulk stores need 16-byte aligned global addresses; tile offsets keep the alignment of the base
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int allow_bulk = al16(pos) && al16(rot) && al16(quat) && al16(J);
    if constexpr (!OSOA) {
        const int64_t ntiles = B / kTile;
        const int variant = fk_variant();
        // (an SoA input with a partial last tile stays on the tile-per-CTA kernel: its tail cannot be re-based)
        if (allow_bulk && variant > 0 && ntiles >= 1024 && (!QSOA || ntiles * kTile == B)) {
            int rc = variant == 1 ? launch_fk_persist<T, D, QSOA, 1>(q, B, ntiles, pos, rot, quat, J, st)
                                  : launch_fk_persist<T, D, QSOA, 2>(q, B, ntiles, pos, rot, quat, J, st);
            if (rc) return rc;
            const int64_t done = ntiles * kTile;
            if (done == B) return RLITG_OK;
            kern<<<1, kTile, smem, st>>>(q + done * D, B - done, pos ? pos + done * 3 : nullptr,
                                         rot ? rot + done * 9 : nullptr, quat ? quat + done * 4 : nullptr,
                                         J ? J + done * 6 * D : nullptr, allow_bulk);
            return check_launch("fk_jac_kernel");
        }
    }
    const int64_t grid = (B + kTile - 1) / kTile;
    kern<<<(unsigned)grid, kTile, smem, st>>>(q, B, pos, rot, quat, J, allow_bulk);
    return check_launch("fk_jac_kernel");
}

template <typename T>
int launch_fk_jac(const T* q, int64_t B, int D, int flags, T* pos, T* rot, T* quat, T* J, cudaStream_t st) {
    if (B == 0) return RLITG_OK;
    const bool qs = flags & RLITG_Q_SOA, os = flags & RLITG_OUT_SOA;
#define RLITG_DISPATCH(DD)                                                                     \
    if (qs && os) return launch_fk_jac_t<T, DD, true, true>(q, B, pos, rot, quat, J, st);      \
    if (qs && !os) return launch_fk_jac_t<T, DD, true, false>(q, B, pos, rot, quat, J, st);    \
    if (!qs && os) return launch_fk_jac_t<T, DD, false, true>(q, B, pos, rot, quat, J, st);    \
    return launch_fk_jac_t<T, DD, false, false>(q, B, pos, rot, quat, J, st);
    if (D == 7) { RLITG_DISPATCH(7) }
    if (D == 8) { RLITG_DISPATCH(8) }
#undef RLITG_DISPATCH
    return set_error(RLITG_ERR_SHAPE, "D must be 7 (arm) or 8 (torso + arm)");
}

template <typename T, int D, bool QSOA>
static int launch_allfk_t(const T* q, int64_t B, T* out, cudaStream_t st) {
    const size_t smem = sizeof(T) * kTile * ((D + 1) * 7 + 1 + D);
    auto kern = allfk_kernel<T, D, QSOA>;
    int rc0 = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "allfk_kernel");
    if (rc0) return rc0;
    const int64_t ntiles = B / kTile;
    const bool al16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (al16 && ntiles >= 1024 && (!QSOA || ntiles * kTile == B)) {
        constexpr int N = (D + 1) * 7;
        // measured on B200 at 1M configs (RLITG_ALLFK_CTAS / RLITG_ALLFK_STAGES): 1 stage x 6 CTAs/SM 47.8 us, 1 x 4
        // 51.8, 2 x 4 51.7, 1 x 2 82.5; the tile-per-CTA kernel 70 us.  Eight quaternion conversions per configuration
        // make this kernel more arithmetic-heavy than fk_jac, so it wants more resident warps.
        static const int amult = [] { const char* e = getenv("RLITG_ALLFK_CTAS"); return e ? atoi(e) : 6; }();
        static const int astage = [] { const char* e = getenv("RLITG_ALLFK_STAGES"); return e ? atoi(e) : 1; }();
        auto pk = astage == 1 ? allfk_persist_kernel<T, D, QSOA, 1> : allfk_persist_kernel<T, D, QSOA, 2>;
        const size_t psmem = sizeof(T) * astage * kTile * N;
        int rc1 = ensure_dynamic_smem(reinterpret_cast<const void*>(pk), psmem, "allfk_persist_kernel");
        if (rc1) return rc1;
        const int64_t g2 = std::min<int64_t>(ntiles, (int64_t)current_sm_count() * amult);
        pk<<<(unsigned)g2, kTile, psmem, st>>>(q, B, ntiles, out);
        rc1 = check_launch("allfk_persist_kernel");
        const int64_t done = ntiles * kTile;
        if (rc1 || done == B) return rc1;
        kern<<<1, kTile, smem, st>>>(q + done * D, B - done, out + done * N);
        return check_launch("allfk_kernel");
    }
    const int64_t grid = (B + kTile - 1) / kTile;
    kern<<<(unsigned)grid, kTile, smem, st>>>(q, B, out);
    return check_launch("allfk_kernel");
}

template <typename T>
int launch_allfk(const T* q, int64_t B, int D, int flags, T* out, cudaStream_t st) {
    if (B == 0) return RLITG_OK;
    const bool qs = flags & RLITG_Q_SOA;
    if (D == 7) return qs ? launch_allfk_t<T, 7, true>(q, B, out, st) : launch_allfk_t<T, 7, false>(q, B, out, st);
    if (D == 8) return qs ? launch_allfk_t<T, 8, true>(q, B, out, st) : launch_allfk_t<T, 8, false>(q, B, out, st);
    return set_error(RLITG_ERR_SHAPE, "D must be 7 (arm) or 8 (torso + arm)");
}

template int launch_fk_jac<float>(const float*, int64_t, int, int, float*, float*, float*, float*, cudaStream_t);
template int launch_fk_jac<double>(const double*, int64_t, int, int, double*, double*, double*, double*, cudaStream_t);
template int launch_allfk<float>(const float*, int64_t, int, int, float*, cudaStream_t);
template int launch_allfk<double>(const double*, int64_t, int, int, double*, cudaStream_t);

}  // namespace rlitg
