// Scene staging: occupancy grid -> bit-packed volume + signed Euclidean distance records, sphere table upload,
// and the binary sphere-vs-grid collision test.
//
// Inputs follow the reference's formats: the pirl_gazebo occupancy grid (50 x 70 x 50 cells @ 0.02 m,
// pirl_gazebo/script/save_pointcloud.py:22-32) or any [nz][ny][nx] uint8 volume; the distance semantics are
// those of MoveIt's PropagationDistanceField as TORM builds it (torm/src/torm_optimizer.cpp:110-120: signed,
// max propagation distance 1.0 m).  The reference never connects the two (SURVEY "four
// facts" item 4): the grid -> distance-field step is this build's own, and is mirrored by orc_build_field.
//
// The EDT is the exact separable min-plus transform over integer cell offsets (bit-exact against the oracle);
// at 128^3 the three O(n) passes of brute-force line minima are 0.8 G integer ops -- a one-time setup cost.
#include <algorithm>
#include <vector>

#include "fetch_chain.cuh"
#include "rlitg_internal.h"
#include "scene.cuh"

namespace rlitg {

constexpr int32_t kEdtInf = 1 << 28;

__global__ void edt_init_kernel(const uint8_t* __restrict__ occ, int64_t n, int invert, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool seed = invert ? (occ[i] == 0) : (occ[i] != 0);
    out[i] = seed ? 0 : kEdtInf;
}

// one min-plus pass along an axis: out[i] = min_j in[j] + (i - j)^2 ; stride = element stride of that axis
__global__ void edt_pass_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int nx, int ny, int nz,
                                int axis) {
    const int64_t n = (int64_t)nx * ny * nz;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((int64_t)nx * ny));
    int len, pos;
    int64_t stride, base;
    if (axis == 0) { len = nx; pos = x; stride = 1; base = i - x; }
    else if (axis == 1) { len = ny; pos = y; stride = nx; base = i - (int64_t)y * nx; }
    else { len = nz; pos = z; stride = (int64_t)nx * ny; base = i - (int64_t)z * nx * ny; }
    int32_t best = kEdtInf;
    for (int j = 0; j < len; j++) {
        const int32_t v = in[base + j * stride] + (pos - j) * (pos - j);
        best = min(best, v);
    }
    out[i] = best;
}

__global__ void edt_signed_kernel(const int32_t* __restrict__ dso, const int32_t* __restrict__ dsf, int64_t n,
                                  float res, float cap, float* __restrict__ d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t a = dso[i], b = dsf[i];
    float v = __fmul_rn(res, __fsub_rn(__fsqrt_rn((float)a), __fsqrt_rn((float)(b >= kEdtInf ? 0 : b))));
    if (a >= kEdtInf) v = cap;
    d[i] = v < cap ? v : cap;
}

__global__ void field_record_kernel(const float* __restrict__ d, int nx, int ny, int nz, float cap, float i2r,
                                    float4* __restrict__ field) {
    const int64_t n = (int64_t)nx * ny * nz;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((int64_t)nx * ny));
    const bool border = (x < 1 || y < 1 || z < 1 || x >= nx - 1 || y >= ny - 1 || z >= nz - 1);
    float4 r;
    if (border || !(d[i] < cap)) {
        r = make_float4(1e30f, 0.f, 0.f, 0.f);
    } else {
        const int64_t sy = nx, sz = (int64_t)nx * ny;
        r.x = d[i];
        r.y = __fmul_rn(__fsub_rn(d[i + 1], d[i - 1]), i2r);
        r.z = __fmul_rn(__fsub_rn(d[i + sy], d[i - sy]), i2r);
        r.w = __fmul_rn(__fsub_rn(d[i + sz], d[i - sz]), i2r);
    }
    field[i] = r;
}

// one warp packs 32 cells into a word with a ballot
__global__ void bitpack_kernel(const uint8_t* __restrict__ occ, int nx, int ny, int nz, int wx,
                               uint32_t* __restrict__ bits) {
    const int64_t nwords = (int64_t)wx * ny * nz;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nwords) return;
    const int w = (int)(warp % wx);
    const int64_t row = warp / wx;  // z * ny + y
    const int x = w * 32 + lane;
    const bool on = (x < nx) && occ[row * nx + x] != 0;
    const uint32_t m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[warp] = m;
}

// ---- host-side scene object ------------------------------------------------------------------------------
struct Scene {
    int device = 0;
    float4* field = nullptr;
    uint32_t* bits = nullptr;
    float4* spheres = nullptr;
    int* link_start = nullptr;
    uint32_t* dil = nullptr;      // dilated bit volumes, one per distinct sphere radius
    int* sph_class = nullptr;
    int ncls = 0;
    std::vector<float> h_radius;  // host copy of the sphere radii (float32, as stored on the device)
    int S = 0;
    int nx = 0, ny = 0, nz = 0, wx = 0;
    double origin[3] = {0, 0, 0};
    double res = 0;
    std::vector<int> h_link;
    void free_grid() {
        if (field) cudaFree(field);
        if (bits) cudaFree(bits);
        free_dilated();
        field = nullptr; bits = nullptr;
        nx = ny = nz = wx = 0;
    }
    void free_dilated() {
        if (dil) cudaFree(dil);
        if (sph_class) cudaFree(sph_class);
        dil = nullptr; sph_class = nullptr; ncls = 0;
    }
    void free_spheres() {
        if (spheres) cudaFree(spheres);
        if (link_start) cudaFree(link_start);
        free_dilated();
        spheres = nullptr; link_start = nullptr; S = 0;
    }
    SceneDev dev() const {
        SceneDev d;
        d.field = field; d.bits = bits; d.spheres = spheres; d.link_start = link_start;
        d.dil = dil; d.sph_class = sph_class; d.ncls = ncls;
        d.S = S; d.nx = nx; d.ny = ny; d.nz = nz; d.wx = wx;
        d.origin[0] = origin[0]; d.origin[1] = origin[1]; d.origin[2] = origin[2];
        d.res = res;
        d.origin_f[0] = (float)origin[0]; d.origin_f[1] = (float)origin[1]; d.origin_f[2] = (float)origin[2];
        d.res_f = (float)res; d.inv_res_f = (float)(1.0 / res);
        return d;
    }
};

bool scene_ready(const Scene* s, bool need_grid, bool need_spheres) {
    if (!s) return false;
    if (need_grid && !s->field) return false;
    if (need_spheres && !s->spheres) return false;
    return true;
}
SceneDev scene_dev(const Scene* s) { return s->dev(); }

static inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// occupancy dilated by a sphere radius, bit-packed like `bits`: bit = active cell with signed distance < radius.
// One warp per 32-cell word (ballot).  These volumes turn the binary collision test into a single-bit lookup.
__global__ void dilate_kernel(const float4* __restrict__ field, int nx, int ny, int nz, int wx, float radius,
                              uint32_t* __restrict__ out) {
    const int64_t nwords = (int64_t)wx * ny * nz;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nwords) return;
    const int w = (int)(warp % wx);
    const int64_t row = warp / wx;
    const int x = w * 32 + lane;
    bool on = false;
    if (x < nx) {
        const float d = field[row * nx + x].x;
        on = d < 1e29f && d < radius;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) out[warp] = m;
}

constexpr int kMaxRadiusClasses = 8;

// (re)builds the dilated bit volumes once both the sphere table and the grid are present
static int rebuild_dilated(Scene* s, cudaStream_t st = nullptr) {
    s->free_dilated();
    if (!s->field || !s->spheres || s->S == 0) return RLITG_OK;
    std::vector<float> cls;
    std::vector<int> sc(s->S);
    for (int i = 0; i < s->S; i++) {
        int k = -1;
        for (size_t c = 0; c < cls.size(); c++)
            if (cls[c] == s->h_radius[i]) k = (int)c;
        if (k < 0) { cls.push_back(s->h_radius[i]); k = (int)cls.size() - 1; }
        sc[i] = k;
    }
    if ((int)cls.size() > kMaxRadiusClasses) return RLITG_OK;  // many distinct radii: keep the record-based test
    const size_t words = (size_t)s->wx * s->ny * s->nz;
    if (cudaMalloc(&s->dil, sizeof(uint32_t) * words * cls.size()) != cudaSuccess ||
        cudaMalloc(&s->sph_class, sizeof(int) * s->S) != cudaSuccess ||
        cudaMemcpyAsync(s->sph_class, sc.data(), sizeof(int) * s->S, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        s->free_dilated();
        return set_cuda_error("rebuild_dilated");
    }
    for (size_t c = 0; c < cls.size(); c++) {
        dilate_kernel<<<blocks_for((int64_t)words * 32, 256), 256, 0, st>>>(s->field, s->nx, s->ny, s->nz, s->wx, cls[c],
                                                                            s->dil + c * words);
        int rc = check_launch("dilate_kernel");
        if (rc) { s->free_dilated(); return rc; }
    }
    // `sc` (host) feeds the asynchronous copy above: wait for this stream only, not for the whole device
    if (cudaStreamSynchronize(st) != cudaSuccess) { s->free_dilated(); return set_cuda_error("rebuild_dilated"); }
    s->ncls = (int)cls.size();
    return RLITG_OK;
}

// ---- two-phase bit test (dilated volumes in global memory) -----------------------------------------------------------
// Phase 1: the chain walk stores, per sphere, the linear BIT index of its centre cell in the volume of its radius class
// (0xffffffff outside the grid) into the thread's own shared-memory row; phase 2 gathers the words of all spheres with
// independent loads, eight in flight, and ORs the bits.  The walk has no load in it and no branch per sphere; a waypoint
// costs one L2 round trip per eight spheres instead of one per link (the loads of a link used to be consumed before the
// walk went on: six to seven dependent round trips per thread and 52 KB of branchy code).
constexpr uint32_t kNoCell = 0xffffffffu;

template <typename T>
struct CollideIndexVisit {
    const SceneDev& sc;
    const GridMap<T>& gm;
    const float4* sph;          // shared-memory copy of the sphere table
    const int* link_start;      // shared-memory copy
    const uint32_t* cls_bits;   // shared memory: first bit of sphere s's class volume (class * words * 32)
    uint32_t* row;              // shared memory: this thread's [S] bit indices
    V3<T> off;
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        if constexpr (K < kNumLinks) {
            const int s1 = link_start[K + 1];
            for (int s = link_start[K]; s < s1; s++) {
                const float4 c = sph[s];
                const T x = off.x + f.p.x + f.c0.x * T(c.x) + f.c1.x * T(c.y) + f.c2.x * T(c.z);
                const T y = off.y + f.p.y + f.c0.y * T(c.x) + f.c1.y * T(c.y) + f.c2.y * T(c.z);
                const T z = off.z + f.p.z + f.c0.z * T(c.x) + f.c1.z * T(c.y) + f.c2.z * T(c.z);
                const int cx = gm.cell1(x, gm.ox), cy = gm.cell1(y, gm.oy), cz = gm.cell1(z, gm.oz);
                const bool in = (unsigned)cx < (unsigned)sc.nx && (unsigned)cy < (unsigned)sc.ny && (unsigned)cz < (unsigned)sc.nz;
                row[s] = in ? cls_bits[s] + (uint32_t)((cz * sc.ny + cy) * (sc.wx * 32) + cx) : kNoCell;
            }
        }
    }
};

RLITG_DEV bool any_bit_set(const uint32_t* __restrict__ dil, const uint32_t* row, int S) {
    uint32_t hit = 0u;
    for (int s0 = 0; s0 < S; s0 += 8) {
        uint32_t idx[8], w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) idx[k] = s0 + k < S ? row[s0 + k] : kNoCell;
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = idx[k] != kNoCell ? __ldg(dil + (idx[k] >> 5)) : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) hit |= w[k] >> (idx[k] & 31u);
    }
    return (hit & 1u) != 0u;
}

// the bit index is 32 bits wide: volumes beyond that (> 512 MB of dilated bits) keep the per-link test
static inline bool dil_index_fits(const Scene* s) {
    return s->dil != nullptr && (uint64_t)s->wx * s->ny * s->nz * 32ull * (uint64_t)s->ncls < 0xffffffffull;
}

// ---- binary collision: any sphere with d(centre cell) - r < 0 ---------------------------------------------
// B200 stand-in for cc_srv / TormIKSolver::collisionChecking (pirl/src/collision_checker.cpp:39-61,
// torm/src/torm_ik_solver.cpp:59-66): an approximation of FCL's mesh test, not a parity target.
template <typename T>
struct CollideVisit {
    const SceneDev& sc;
    const GridMap<T>& gm;
    const float4* sph;      // shared-memory copy of the sphere table
    const int* link_start;  // shared-memory copy
    V3<T> off;
    bool hit = false;
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        if constexpr (K < kNumLinks) {
            for (int s = link_start[K]; s < link_start[K + 1]; s++) {
                const float4 c = sph[s];
                const V3<T> x = {off.x + f.p.x + f.c0.x * T(c.x) + f.c1.x * T(c.y) + f.c2.x * T(c.z),
                                 off.y + f.p.y + f.c0.y * T(c.x) + f.c1.y * T(c.y) + f.c2.y * T(c.z),
                                 off.z + f.p.z + f.c0.z * T(c.x) + f.c1.z * T(c.y) + f.c2.z * T(c.z)};
                if (sphere_hits(sc, gm, x, s, c.w, sc.dil)) hit = true;
            }
        }
    }
};

template <typename T, bool kIdx>
__global__ void __launch_bounds__(128) sphere_collide_kernel(const T* __restrict__ q, int64_t B, SceneDev sc,
                                                             T ox, T oy, T oz,
                                                             uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t s_rows[];  // kIdx: [blockDim.x][S | 1] bit indices
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_start[kNumLinks + 1];
    __shared__ uint32_t s_cls[kMaxSpheres];
    for (int i = threadIdx.x; i < sc.S; i += blockDim.x) {
        s_sph[i] = sc.spheres[i];
        if constexpr (kIdx) s_cls[i] = (uint32_t)sc.sph_class[i] * (uint32_t)(sc.wx * sc.ny * sc.nz) * 32u;
    }
    if (threadIdx.x <= kNumLinks) s_start[threadIdx.x] = sc.link_start[threadIdx.x];
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T qv[7];
#pragma unroll
    for (int j = 0; j < 7; j++) qv[j] = q[b * 7 + j];
    const GridMap<T> gm(sc);
    if constexpr (kIdx) {
        uint32_t* row = s_rows + (size_t)threadIdx.x * (sc.S | 1);
        CollideIndexVisit<T> v{sc, gm, s_sph, s_start, s_cls, row, {ox, oy, oz}};
        walk_chain<T, 7>(qv, v);
        out[b] = any_bit_set(sc.dil, row, sc.S) ? 1 : 0;
    } else {
        CollideVisit<T> v{sc, gm, s_sph, s_start, {ox, oy, oz}};
        walk_chain<T, 7>(qv, v);
        out[b] = v.hit ? 1 : 0;
    }
}

// Persistent variant for grids whose dilated bit volumes fit in shared memory (e.g. the pirl_gazebo 50 x 70 x 50
// grid: 6 radius classes x 28 KB): each CTA stages the bit volumes once and then tests configurations from them,
// so the only global traffic is q in and one byte out per configuration.
template <typename T>
struct CollideSmemVisit {
    const SceneDev& sc;
    const GridMap<T>& gm;
    const float4* sph;
    const int* link_start;
    const int* cls;         // shared-memory copy of sph_class
    const uint32_t* dil;    // shared-memory copy of the dilated volumes
    size_t words;
    V3<T> off;
    bool hit = false;
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        if constexpr (K < kNumLinks) {
            for (int s = link_start[K]; s < link_start[K + 1]; s++) {
                const float4 c = sph[s];
                const V3<T> x = {off.x + f.p.x + f.c0.x * T(c.x) + f.c1.x * T(c.y) + f.c2.x * T(c.z),
                                 off.y + f.p.y + f.c0.y * T(c.x) + f.c1.y * T(c.y) + f.c2.y * T(c.z),
                                 off.z + f.p.z + f.c0.z * T(c.x) + f.c1.z * T(c.y) + f.c2.z * T(c.z)};
                const int cx = gm.cell1(x.x, gm.ox), cy = gm.cell1(x.y, gm.oy), cz = gm.cell1(x.z, gm.oz);
                if (cx < 0 || cy < 0 || cz < 0 || cx >= sc.nx || cy >= sc.ny || cz >= sc.nz) continue;
                const uint32_t w = dil[(size_t)cls[s] * words + ((size_t)cz * sc.ny + cy) * sc.wx + (cx >> 5)];
                if ((w >> (cx & 31)) & 1u) hit = true;
            }
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(256) sphere_collide_smem_kernel(const T* __restrict__ q, int64_t B, SceneDev sc,
                                                                  T ox, T oy, T oz,
                                                                  uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t s_dil[];
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_start[kNumLinks + 1];
    __shared__ int s_cls[kMaxSpheres];
    const size_t words = (size_t)sc.wx * sc.ny * sc.nz;
    const size_t total = words * sc.ncls;
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) s_dil[i] = sc.dil[i];
    for (int i = threadIdx.x; i < sc.S; i += blockDim.x) { s_sph[i] = sc.spheres[i]; s_cls[i] = sc.sph_class[i]; }
    if (threadIdx.x <= kNumLinks) s_start[threadIdx.x] = sc.link_start[threadIdx.x];
    __syncthreads();
    const GridMap<T> gm(sc);
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        T qv[7];
#pragma unroll
        for (int j = 0; j < 7; j++) qv[j] = q[b * 7 + j];
        CollideSmemVisit<T> v{sc, gm, s_sph, s_start, s_cls, s_dil, words, {ox, oy, oz}};
        walk_chain<T, 7>(qv, v);
        out[b] = v.hit ? 1 : 0;
    }
}

// isCurrentTrajectoryCollisionFree for N trajectories (torm/src/torm_optimizer.cpp:695-717): the rows of a trajectory
// are tested from free_vars_start_ - 1 (= row 0 here) upwards and the FIRST colliding row is collision_point_ (-1 when
// the trajectory is free).  With use_stochastic_descent_ the collision increment of the same half-iteration takes ONE
// waypoint (:1502-1515): collision_point_ when there is one, else (int)(u (free_end - free_start) + free_start) for a
// uniform u -- as a free-row index that is first_hit - 1, or (int)(u (F - 1)).  first_hit = 0 (the start row collides)
// gives index -1, which the reference writes out of bounds; the cost kernel takes no
// increment for it.  One CTA per trajectory, rows strided over the threads, block-wide minimum.
// Sphere test against the dilated bit grids with the loads of a link's spheres batched four at a time (independent
// L2 round trips instead of 2 x 24 dependent ones per waypoint: the class table sits in shared memory as a word offset)
template <typename T>
struct CollideVisitDil {
    const SceneDev& sc;
    const GridMap<T>& gm;
    const float4* sph;      // shared-memory copy of the sphere table
    const int* link_start;  // shared-memory copy
    const int* cls_off;     // shared memory: class of sphere s times the words per class
    V3<T> off;
    bool hit = false;
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<T>& f) {
        if constexpr (K < kNumLinks) {
            const int s1 = link_start[K + 1];
            for (int s0 = link_start[K]; s0 < s1; s0 += 4) {
                uint32_t w[4];
                int bit[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    w[k] = 0u;
                    bit[k] = 0;
                    const int s = s0 + k;
                    if (s < s1) {
                        const float4 c = sph[s];
                        const T x = off.x + f.p.x + f.c0.x * T(c.x) + f.c1.x * T(c.y) + f.c2.x * T(c.z);
                        const T y = off.y + f.p.y + f.c0.y * T(c.x) + f.c1.y * T(c.y) + f.c2.y * T(c.z);
                        const T z = off.z + f.p.z + f.c0.z * T(c.x) + f.c1.z * T(c.y) + f.c2.z * T(c.z);
                        const int cx = gm.cell1(x, gm.ox), cy = gm.cell1(y, gm.oy), cz = gm.cell1(z, gm.oz);
                        if (cx >= 0 && cy >= 0 && cz >= 0 && cx < sc.nx && cy < sc.ny && cz < sc.nz) {
                            w[k] = __ldg(sc.dil + (size_t)cls_off[s] + ((size_t)cz * sc.ny + cy) * sc.wx + (cx >> 5));
                            bit[k] = cx & 31;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) hit = hit || ((w[k] >> bit[k]) & 1u);
            }
        }
    }
};

// kIdx: the two-phase bit test (CollideIndexVisit; needs blockDim.x * (S | 1) words of dynamic shared memory)
template <typename T, bool kIdx>
__global__ void __launch_bounds__(256) collision_point_kernel(const T* __restrict__ traj, int Tn, SceneDev sc, T ox, T oy, T oz,
                                                             const float* __restrict__ u, int32_t* __restrict__ first_hit,
                                                             uint8_t* __restrict__ any_hit, int32_t* __restrict__ stoch_idx) {
    extern __shared__ __align__(16) uint32_t s_rows[];  // kIdx: [blockDim.x][S | 1] bit indices (odd stride: conflict free)
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_start[kNumLinks + 1];
    __shared__ int s_first;
    __shared__ int s_cls[kMaxSpheres];
    for (int i = threadIdx.x; i < sc.S; i += blockDim.x) {
        s_sph[i] = sc.spheres[i];
        s_cls[i] = sc.dil ? sc.sph_class[i] * (sc.wx * sc.ny * sc.nz) * (kIdx ? 32 : 1) : 0;
    }
    if (threadIdx.x <= kNumLinks) s_start[threadIdx.x] = sc.link_start[threadIdx.x];
    if (threadIdx.x == 0) s_first = 0x7fffffff;
    __syncthreads();
    const int64_t n = blockIdx.x;
    const T* tr = traj + n * (int64_t)Tn * 7;
    const GridMap<T> gm(sc);
    int mine = 0x7fffffff;
    for (int i = threadIdx.x; i < Tn; i += blockDim.x) {
        if (i > mine) break;
        T qv[7];
#pragma unroll
        for (int j = 0; j < 7; j++) qv[j] = tr[(size_t)i * 7 + j];
        bool hit;
        if constexpr (kIdx) {
            uint32_t* row = s_rows + (size_t)threadIdx.x * (sc.S | 1);
            CollideIndexVisit<T> v{sc, gm, s_sph, s_start, reinterpret_cast<const uint32_t*>(s_cls), row, {ox, oy, oz}};
            walk_chain<T, 7>(qv, v);
            hit = any_bit_set(sc.dil, row, sc.S);
        } else if (sc.dil) {
            CollideVisitDil<T> v{sc, gm, s_sph, s_start, s_cls, {ox, oy, oz}};
            walk_chain<T, 7>(qv, v);
            hit = v.hit;
        } else {
            CollideVisit<T> v{sc, gm, s_sph, s_start, {ox, oy, oz}};
            walk_chain<T, 7>(qv, v);
            hit = v.hit;
        }
        if (hit) mine = min(mine, i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    if ((threadIdx.x & 31) == 0 && mine != 0x7fffffff) atomicMin(&s_first, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int f = s_first == 0x7fffffff ? -1 : s_first;
        if (first_hit) first_hit[n] = f;
        if (any_hit) any_hit[n] = f >= 0;
        if (stoch_idx) {
            const int F = Tn - 1;
            int idx;
            if (f >= 0) idx = f - 1;
            else {
                idx = u ? (int)(u[n] * (float)(F - 1)) : 0;
                idx = idx < 0 ? 0 : (idx > F - 1 ? F - 1 : idx);
            }
            stoch_idx[n] = idx;
        }
    }
}

template <typename T>
int launch_sphere_collide(const T* q, int64_t B, const Scene* s, const double* base_off, uint8_t* out,
                          cudaStream_t st) {
    if (B == 0) return RLITG_OK;
    const size_t dil_bytes = sizeof(uint32_t) * (size_t)s->wx * s->ny * s->nz * s->ncls;
    if (s->dil && dil_bytes <= 200 * 1024 && B >= 4096) {
        auto kern = sphere_collide_smem_kernel<T>;
        int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), dil_bytes, "sphere_collide_smem_kernel");
        if (rc_attr) return rc_attr;
        const int num_sms = current_sm_count();
        const unsigned grid = (unsigned)std::min<int64_t>(num_sms, (B + 255) / 256);
        kern<<<grid, 256, dil_bytes, st>>>(q, B, s->dev(), (T)base_off[0], (T)base_off[1], (T)base_off[2], out);
        return check_launch("sphere_collide_smem_kernel");
    }
    static const bool two_phase = [] { const char* e = getenv("RLITG_COLLIDE_TWO_PHASE"); return e ? atoi(e) != 0 : true; }();
    if (two_phase && dil_index_fits(s)) {
        auto kern = sphere_collide_kernel<T, true>;
        const size_t smem = sizeof(uint32_t) * 128 * (size_t)(s->S | 1);
        int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "sphere_collide_kernel");
        if (rc_attr) return rc_attr;
        kern<<<blocks_for(B, 128), 128, smem, st>>>(q, B, s->dev(), (T)base_off[0], (T)base_off[1], (T)base_off[2], out);
        return check_launch("sphere_collide_kernel");
    }
    sphere_collide_kernel<T, false><<<blocks_for(B, 128), 128, 0, st>>>(q, B, s->dev(), (T)base_off[0], (T)base_off[1],
                                                                        (T)base_off[2], out);
    return check_launch("sphere_collide_kernel");
}
template int launch_sphere_collide<float>(const float*, int64_t, const Scene*, const double*, uint8_t*, cudaStream_t);
template int launch_sphere_collide<double>(const double*, int64_t, const Scene*, const double*, uint8_t*, cudaStream_t);

}  // namespace rlitg

using namespace rlitg;

#define CUDA_TRY(expr, what)                                   \
    do {                                                       \
        if ((expr) != cudaSuccess) return set_cuda_error(what); \
    } while (0)

extern "C" {

RLITG_API int rlitg_scene_create(RlitgScene** out) {
    if (!out) return set_error(RLITG_ERR_NULL, "out is NULL");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
    Scene* s = new Scene();
    s->device = dev;
    *out = reinterpret_cast<RlitgScene*>(s);
    return RLITG_OK;
}

RLITG_API int rlitg_scene_destroy(RlitgScene* h) {
    Scene* s = reinterpret_cast<Scene*>(h);
    if (!s) return RLITG_OK;
    s->free_grid();
    s->free_spheres();
    delete s;
    return RLITG_OK;
}

RLITG_API int rlitg_scene_set_spheres(RlitgScene* h, int S, const int32_t* link, const double* center,
                                      const double* radius) {
    Scene* s = reinterpret_cast<Scene*>(h);
    if (!s || !link || !center || !radius) return set_error(RLITG_ERR_NULL, "scene / link / center / radius is NULL");
    if (S < 1 || S > kMaxSpheres) return set_error(RLITG_ERR_SHAPE, "S must be in 1..64");
    std::vector<float4> sp(S);
    std::vector<int> start(kNumLinks + 1, 0);
    for (int i = 0; i < S; i++) {
        if (link[i] < 0 || link[i] >= kNumLinks) return set_error(RLITG_ERR_ARG, "sphere link index must be in 0..7");
        if (i > 0 && link[i] < link[i - 1]) return set_error(RLITG_ERR_ARG, "sphere table must be sorted by link index");
        sp[i] = make_float4((float)center[3 * i], (float)center[3 * i + 1], (float)center[3 * i + 2], (float)radius[i]);
        start[link[i] + 1]++;
    }
    for (int k = 0; k < kNumLinks; k++) start[k + 1] += start[k];
    s->free_spheres();
    CUDA_TRY(cudaMalloc(&s->spheres, sizeof(float4) * S), "cudaMalloc(spheres)");
    CUDA_TRY(cudaMalloc(&s->link_start, sizeof(int) * (kNumLinks + 1)), "cudaMalloc(link_start)");
    CUDA_TRY(cudaMemcpy(s->spheres, sp.data(), sizeof(float4) * S, cudaMemcpyHostToDevice), "H2D spheres");
    CUDA_TRY(cudaMemcpy(s->link_start, start.data(), sizeof(int) * (kNumLinks + 1), cudaMemcpyHostToDevice), "H2D link_start");
    s->S = S;
    s->h_link.assign(link, link + S);
    s->h_radius.resize(S);
    for (int i = 0; i < S; i++) s->h_radius[i] = sp[i].w;
    return rebuild_dilated(s);
}

static int scene_alloc_grid(Scene* s, int nx, int ny, int nz, const double* origin, double res) {
    if (nx < 3 || ny < 3 || nz < 3 || nx > 1024 || ny > 1024 || nz > 1024)
        return set_error(RLITG_ERR_SHAPE, "grid dimensions must be in 3..1024");
    if (!(res > 0)) return set_error(RLITG_ERR_ARG, "res must be > 0");
    s->free_grid();
    const int64_t n = (int64_t)nx * ny * nz;
    const int wx = (nx + 31) / 32;
    CUDA_TRY(cudaMalloc(&s->field, sizeof(float4) * n), "cudaMalloc(field)");
    CUDA_TRY(cudaMalloc(&s->bits, sizeof(uint32_t) * (size_t)wx * ny * nz), "cudaMalloc(bits)");
    CUDA_TRY(cudaMemset(s->bits, 0, sizeof(uint32_t) * (size_t)wx * ny * nz), "cudaMemset(bits)");
    s->nx = nx; s->ny = ny; s->nz = nz; s->wx = wx;
    s->origin[0] = origin[0]; s->origin[1] = origin[1]; s->origin[2] = origin[2];
    s->res = res;
    return RLITG_OK;
}

// frees the staging temporaries of rlitg_scene_set_occupancy on every path out of it
struct EdtTemps {
    uint8_t* occ = nullptr;
    int32_t *a = nullptr, *b = nullptr, *dso = nullptr;
    float* d = nullptr;
    ~EdtTemps() {
        if (occ) cudaFree(occ);
        if (a) cudaFree(a);
        if (b) cudaFree(b);
        if (dso) cudaFree(dso);
        if (d) cudaFree(d);
    }
};

RLITG_API int rlitg_scene_set_occupancy_stream(RlitgScene* h, const uint8_t* occ, int nx, int ny, int nz,
                                               const double* origin, double res, double max_prop, void* stream) {
    Scene* s = reinterpret_cast<Scene*>(h);
    if (!s || !occ || !origin) return set_error(RLITG_ERR_NULL, "scene / occ / origin is NULL");
    int rc = scene_alloc_grid(s, nx, ny, nz, origin, res);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t n = (int64_t)nx * ny * nz;
    EdtTemps t;
    CUDA_TRY(cudaMalloc(&t.occ, n), "cudaMalloc(occ)");
    CUDA_TRY(cudaMalloc(&t.a, sizeof(int32_t) * n), "cudaMalloc(edt)");
    CUDA_TRY(cudaMalloc(&t.b, sizeof(int32_t) * n), "cudaMalloc(edt)");
    CUDA_TRY(cudaMalloc(&t.dso, sizeof(int32_t) * n), "cudaMalloc(edt)");
    CUDA_TRY(cudaMalloc(&t.d, sizeof(float) * n), "cudaMalloc(edt)");
    CUDA_TRY(cudaMemsetAsync(s->bits, 0, sizeof(uint32_t) * (size_t)s->wx * ny * nz, st), "cudaMemset(bits)");
    CUDA_TRY(cudaMemcpyAsync(t.occ, occ, n, cudaMemcpyHostToDevice, st), "H2D occupancy");
    const int bs = 256;
    const unsigned g = blocks_for(n, bs);
    for (int pass = 0; pass < 2; pass++) {  // 0: distance to occupied, 1: distance to free
        edt_init_kernel<<<g, bs, 0, st>>>(t.occ, n, pass, t.a);
        if ((rc = check_launch("edt_init_kernel"))) return rc;
        edt_pass_kernel<<<g, bs, 0, st>>>(t.a, t.b, nx, ny, nz, 0);
        if ((rc = check_launch("edt_pass_kernel(x)"))) return rc;
        edt_pass_kernel<<<g, bs, 0, st>>>(t.b, t.a, nx, ny, nz, 1);
        if ((rc = check_launch("edt_pass_kernel(y)"))) return rc;
        edt_pass_kernel<<<g, bs, 0, st>>>(t.a, pass == 0 ? t.dso : t.b, nx, ny, nz, 2);
        if ((rc = check_launch("edt_pass_kernel(z)"))) return rc;
    }
    edt_signed_kernel<<<g, bs, 0, st>>>(t.dso, t.b, n, (float)res, (float)max_prop, t.d);
    if ((rc = check_launch("edt_signed_kernel"))) return rc;
    field_record_kernel<<<g, bs, 0, st>>>(t.d, nx, ny, nz, (float)max_prop, (float)(1.0 / (2.0 * res)), s->field);
    if ((rc = check_launch("field_record_kernel"))) return rc;
    const int64_t nwords = (int64_t)s->wx * ny * nz;
    bitpack_kernel<<<blocks_for(nwords * 32, bs), bs, 0, st>>>(t.occ, nx, ny, nz, s->wx, s->bits);
    if ((rc = check_launch("bitpack_kernel"))) return rc;
    // the temporaries are freed when this function returns: the staging kernels must be done with them (this stream only)
    CUDA_TRY(cudaStreamSynchronize(st), "scene_set_occupancy");
    return rebuild_dilated(s, st);
}

RLITG_API int rlitg_scene_set_occupancy(RlitgScene* h, const uint8_t* occ, int nx, int ny, int nz,
                                        const double* origin, double res, double max_prop) {
    return rlitg_scene_set_occupancy_stream(h, occ, nx, ny, nz, origin, res, max_prop, nullptr);
}

RLITG_API int rlitg_scene_set_field(RlitgScene* h, const float* field, int nx, int ny, int nz, const double* origin,
                                    double res) {
    Scene* s = reinterpret_cast<Scene*>(h);
    if (!s || !field || !origin) return set_error(RLITG_ERR_NULL, "scene / field / origin is NULL");
    int rc = scene_alloc_grid(s, nx, ny, nz, origin, res);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(s->field, field, sizeof(float4) * (size_t)nx * ny * nz, cudaMemcpyHostToDevice), "H2D field");
    return rebuild_dilated(s);
}

RLITG_API int rlitg_scene_get_field(const RlitgScene* h, float* field_out) {
    const Scene* s = reinterpret_cast<const Scene*>(h);
    if (!s || !field_out || !s->field) return set_error(RLITG_ERR_NULL, "scene has no grid / output is NULL");
    CUDA_TRY(cudaMemcpy(field_out, s->field, sizeof(float4) * (size_t)s->nx * s->ny * s->nz, cudaMemcpyDeviceToHost), "D2H field");
    return RLITG_OK;
}

RLITG_API int rlitg_scene_get_bits(const RlitgScene* h, uint32_t* bits_out) {
    const Scene* s = reinterpret_cast<const Scene*>(h);
    if (!s || !bits_out || !s->bits) return set_error(RLITG_ERR_NULL, "scene has no grid / output is NULL");
    CUDA_TRY(cudaMemcpy(bits_out, s->bits, sizeof(uint32_t) * (size_t)s->wx * s->ny * s->nz, cudaMemcpyDeviceToHost), "D2H bits");
    return RLITG_OK;
}

RLITG_API int rlitg_scene_get_dilated(const RlitgScene* h, int32_t* ncls_out, float* radii_out, uint32_t* bits_out) {
    const Scene* s = reinterpret_cast<const Scene*>(h);
    if (!s || !ncls_out) return set_error(RLITG_ERR_NULL, "scene / ncls_out is NULL");
    *ncls_out = s->ncls;
    if (s->ncls == 0) return RLITG_OK;
    if (radii_out) {  // class radii in first-appearance order of the sphere table
        int k = 0;
        for (int i = 0; i < s->S && k < s->ncls; i++) {
            bool seen = false;
            for (int j = 0; j < i; j++) seen = seen || (s->h_radius[j] == s->h_radius[i]);
            if (!seen) radii_out[k++] = s->h_radius[i];
        }
    }
    if (bits_out)
        CUDA_TRY(cudaMemcpy(bits_out, s->dil, sizeof(uint32_t) * (size_t)s->wx * s->ny * s->nz * s->ncls, cudaMemcpyDeviceToHost), "D2H dilated bits");
    return RLITG_OK;
}

RLITG_API int rlitg_torm_collision_point(const void* traj, int64_t N, int T, int dtype, const RlitgScene* h,
                                         const double* base_off, const float* u, int32_t* first_hit, uint8_t* any_hit,
                                         int32_t* stochastic_idx, void* stream) {
    const Scene* s = reinterpret_cast<const Scene*>(h);
    if (N < 0 || N > 0x7fffffff || T < 2) return set_error(RLITG_ERR_SHAPE, "N / T out of range");
    if (N == 0) return RLITG_OK;
    if (!traj || !base_off) return set_error(RLITG_ERR_NULL, "traj / base_off is NULL");
    if (!scene_ready(s, true, true)) return set_error(RLITG_ERR_ARG, "scene needs a grid and a sphere table");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cp_threads = std::min(256, ((T + 31) / 32) * 32);  // one waypoint per thread up to 256 rows per pass
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    static const bool two_phase = [] { const char* e = getenv("RLITG_COLLIDE_TWO_PHASE"); return e ? atoi(e) != 0 : true; }();
    const bool idx = two_phase && dil_index_fits(s);
    const size_t smem = idx ? sizeof(uint32_t) * (size_t)cp_threads * (s->S | 1) : 0;
    auto launch = [&](auto kern, auto tag) -> int {
        using TT = decltype(tag);
        int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "collision_point_kernel");
        if (rc_attr) return rc_attr;
        kern<<<(unsigned)N, cp_threads, smem, st>>>((const TT*)traj, T, s->dev(), (TT)base_off[0], (TT)base_off[1], (TT)base_off[2],
                                                   u, first_hit, any_hit, stochastic_idx);
        return check_launch("collision_point_kernel");
    };
    if (dtype == RLITG_F32) return idx ? launch(collision_point_kernel<float, true>, float()) : launch(collision_point_kernel<float, false>, float());
    return idx ? launch(collision_point_kernel<double, true>, double()) : launch(collision_point_kernel<double, false>, double());
}

RLITG_API int rlitg_sphere_grid_collide(const void* q, int64_t B, int dtype, const RlitgScene* h,
                                        const double* base_off, uint8_t* out, void* stream) {
    const Scene* s = reinterpret_cast<const Scene*>(h);
    if (B < 0) return set_error(RLITG_ERR_SHAPE, "B must be >= 0");
    if (B == 0) return RLITG_OK;
    if (!q || !out || !base_off) return set_error(RLITG_ERR_NULL, "q / out / base_off is NULL");
    if (!scene_ready(s, true, true)) return set_error(RLITG_ERR_ARG, "scene needs a grid and a sphere table");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32) return launch_sphere_collide<float>((const float*)q, B, s, base_off, out, st);
    if (dtype == RLITG_F64) return launch_sphere_collide<double>((const double*)q, B, s, base_off, out, st);
    return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
}

}  // extern "C"
