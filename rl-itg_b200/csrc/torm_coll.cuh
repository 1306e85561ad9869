// Kernel 3b: sphere potentials + CHOMP collision increments + smoothness increments (row A7).
//
// What it replaces in the reference (serial, one trajectory at a time, FP64, MoveIt + Eigen):
//   performForwardKinematics      torm/src/torm_optimizer.cpp:1302-1382   sphere positions, field queries, stencils
//   computeJointProperties        :1125-1162                              joint origins / axes in the planning frame
//   getJacobian                   :1164-1187                              3 x D positional Jacobian, ancestor-masked
//   calculateCollisionIncrements  :1484-1553 + calculatePseudoInverse :767-772, getPotential torm_optimizer.h:146-168
//   getCollisionCost              :1448-1463
//   calculateSmoothnessIncrements :1466-1473
//
// Mapping: one CTA per trajectory, one thread per waypoint (long paths in 256-row chunks with recomputed halo rows).
//   phase 1  chain walk in registers -> Jacobian table (a_j, b_j: column j = a_j x x + b_j) kept in registers, sphere
//            centres into shared memory (odd row stride: thread-per-row accesses are conflict free)
//   phase 2  the warp walks the spheres in table order: one 16-byte distance-record gather per sphere, issued two spheres
//            ahead -> potential, cost; for the active ones (potential >= 1e-4, :1526) the 7-point stencil from shared
//            memory, projector / curvature, Jacobian columns, solve, seven accumulators.  The sphere's link is uniform
//            over the warp, so the number of joints that move it is a compile-time constant of the body (switch over
//            1..7 instantiations): masked joints cost nothing.
//            Also measured in round 2 and rejected: every lane walking its own active spheres (mask first, then the
//            k-th active sphere of each lane per pass: max-over-lanes passes instead of the union, 18 vs 22) -- the second
//            pass re-gathers the record and the per-lane link defeats the compile-time joint count: 137 -> 149 us.
//            Also rejected: the trajectory split over the CTAs of a thread-block cluster (chunks of 26 - 100 rows, six
//            extra threads walking the halo rows, partial costs handed to rank 0 through DSMEM) to cut the 19 % of stall
//            samples at the end-of-CTA barrier: 192 / 182 / 125 / 147 / 118 us at 26 / 34 / 50 / 67 / 100 rows per CTA
//            against 113 us unsplit -- more, partly filled warps per trajectory and the per-CTA set-up outweigh it.
// Round 2 also built and measured the lane-per-active-(row, sphere)-pair mapping (chunks of rows per CTA in a cluster,
// compacted pair list, joint tables and contributions through shared memory): 208 - 262 us against this kernel's
// figure in profiles/README.md -- the exchange of the 42-value joint table and of the contributions through shared
// memory, the barriers between its phases and the halo rows cost more than the 35 % of idle lanes it removes
// (profiles/r02_torm_pairs_*.json).
//
// FP32 accuracy (the production build; tolerance 1e-4 of the row maximum).  tools/coll_fp32_study.py
// separates the error sources: (a) FP32 sphere centres (~1.5e-7 m after a seven-joint FP32 walk) through the
// 1 / dt^2-scaled stencil, (b) the normal-equation solve (J J^T + ridge I)^-1 at cond ~ 1e4.  So
//   (a) the FP32 build walks the chain in FP64 (lean sincos, rlitg_common.cuh; the FP64 pipe of this part runs at half
//       the FP32 rate: 36 vs 71 TFLOP/s measured) and rounds each centre once to FP32; the stencils difference against
//       the centre row before scaling (the rule rows sum to 0);
//       Since the end of round 2 the centres of the FP32 build are kept as 32-BIT FIXED POINT on a grid aligned with the
//       distance volume (res / 2^shift per unit from the grid origin: 4.8e-9 m at res 0.02, +-10 m): link origin from the
//       FP64 walk rounded to that grid + R c / res in FP32; the stencil differences are then exact integer arithmetic
//       rounded to FP32 once, relative to the DIFFERENCE instead of to the ~1 m position, and the grid cell of a centre is
//       (units + half a cell) >> shift -- no second float -> int conversion, and decided 10 x closer to the oracle.  With
//       FP32 centres 45 - 67 of the 203,776 benchmark rows were above 1e-4 (worst 3e-3: a slow sphere with a large
//       acceleration, tests/test_torm_gpu.py::fp32_centre_rounding_bound); tools/coll_fp32_study.py (FIXED=29) predicts 0 - 2.
//   (b) J^T (J J^T + ridge I)^-1 g is computed as Q1 R^-T g from the QR of [J^T ; sqrt(ridge) I] (modified Gram-Schmidt on
//       three columns of length k + 3): error ~ sqrt(cond) eps instead of cond eps, same instruction count as forming
//       J J^T, inverting it and applying J^T.
#pragma once
#include <type_traits>

namespace rlitg {

constexpr int kCollThreads = 256;

RLITG_DEV float cell_to_bits(int c, float) { return __int_as_float(c); }
RLITG_DEV double cell_to_bits(int c, double) { return __longlong_as_double((long long)c); }
RLITG_DEV float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
RLITG_DEV int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
template <typename T> RLITG_DEV int lds_cell(uint32_t a) {  // the cell index stored in a T-sized slot (low word)
    return lds_s32(a);
}
RLITG_DEV int bits_to_cell(float b) { return __float_as_int(b); }
RLITG_DEV int bits_to_cell(double b) { return (int)__double_as_longlong(b); }

// Chain visitor of phase 1.  FT = precision of the walk (double in the FP32 production build), T = storage precision.
template <typename T, typename FT, bool kJoints, int PW = 4>
struct CollVisit {
    V3<T> a[7], b[7];      // Jacobian table: column j of getJacobian (:1170-1178) = a_j x (x - o_j) = a_j x x + b_j
    T* pos_row;            // shared-memory row [S][4] of this waypoint's sphere centres (planning frame) + cell index bits
    const GridMap<T>* gm;  // nullptr: the cell index is not wanted (halo rows)
    const float4* sph;
    const int* link_start;
    V3<FT> off;
    int quirk;
    // centre + the linear index of its grid cell (-1 outside): the cell arithmetic sits here, next to the other spheres of
    // the link, instead of in front of every gather of phase 2's serial sphere loop
    // fixed-point centres: FP32 storage, chain walked in FP64, four-value rows (see the header)
    static constexpr bool kFix = !std::is_same<FT, T>::value && sizeof(T) == 4 && PW == 4;
    RLITG_DEV void put(int s, V3<T> x) const {
        T* d = pos_row + PW * s;
        d[0] = x.x; d[1] = x.y; d[2] = x.z;
        if constexpr (PW == 4) {
            if (gm != nullptr) d[3] = cell_to_bits((int)gm->cell(x), T(0));
        }
    }
    // fixed-point grid (kFix): units per metre, 2^fx_shift units per cell, grid origin (FP64)
    FT fx_scale;
    int fx_shift;
    FT fx_ox, fx_oy, fx_oz;   // -origin x fx_scale
    RLITG_DEV void put_fixed(int s, int ix, int iy, int iz) const {
        T* d = pos_row + PW * s;
        d[0] = __int_as_float(ix); d[1] = __int_as_float(iy); d[2] = __int_as_float(iz);
        if (gm != nullptr) {  // nearest cell = floor(units / 2^shift + 0.5), as GridMap::cell1
            const int h = 1 << (fx_shift - 1);
            const int cx = (ix + h) >> fx_shift, cy = (iy + h) >> fx_shift, cz = (iz + h) >> fx_shift;
            const bool out = cx < 0 || cy < 0 || cz < 0 || cx >= gm->nx || cy >= gm->ny || cz >= gm->nz;
            d[3] = cell_to_bits(out ? -1 : (cz * gm->ny + cy) * gm->nx + cx, T(0));
        }
    }
    template <int K>
    RLITG_DEV void operator()(jidx<K>, const FrameC<FT>& f) {
        const V3<FT> o = f.p + off;
        // link origin as a two-float sum (oh + ol = o to ~1e-15): the per-sphere arithmetic below stays in storage precision
        const V3<T> oh = {T(o.x), T(o.y), T(o.z)};
        if constexpr (K < 7 && kJoints) {
            // computeJointProperties :1125-1162; axis = joint_transform * axis (:1157) adds the joint position
            const V3<FT> axf = axis_of<arm_axis(K)>(f);
            V3<T> ax = {T(axf.x), T(axf.y), T(axf.z)};
            if (quirk) ax = ax + oh;
            a[K] = ax;
            b[K] = cross(oh, ax);
        }
        const int s0 = link_start[K], s1 = link_start[K + 1];
        if (s1 > s0) {
            if constexpr (std::is_same<FT, T>::value) {
                for (int s = s0; s < s1; s++) {
                    const float4 c = sph[s];
                    const V3<T> x = {o.x + f.c0.x * T(c.x) + f.c1.x * T(c.y) + f.c2.x * T(c.z),
                                     o.y + f.c0.y * T(c.x) + f.c1.y * T(c.y) + f.c2.y * T(c.z),
                                     o.z + f.c0.z * T(c.x) + f.c1.z * T(c.y) + f.c2.z * T(c.z)};
                    put(s, x);
                }
            } else {
                // R c in FP32 (|c| <= 0.2 m: error ~1e-8 m) added to the FP64 link origin, carried as oh + ol: the sum
                // oh + (R c + ol) is the FP64 sum rounded once (+ ~4e-9 m), with no FP64 instruction per sphere
                const V3<T> ol = {T(o.x - FT(oh.x)), T(o.y - FT(oh.y)), T(o.z - FT(oh.z))};
                const V3<T> r0 = {T(f.c0.x), T(f.c0.y), T(f.c0.z)}, r1 = {T(f.c1.x), T(f.c1.y), T(f.c1.z)},
                            r2 = {T(f.c2.x), T(f.c2.y), T(f.c2.z)};
                if constexpr (kFix) {
                    // the link origin on the fixed-point grid (clamped to what an int32 holds next to |R c| <= 0.2 m)
                    // (fx_o* = -origin x scale: one FP64 FMA per coordinate; the conversion saturates, the clamp is integer)
                    constexpr int lim = 2040000000;
                    const int oix = min(max(__double2int_rn(fma(o.x, fx_scale, fx_ox)), -lim), lim);
                    const int oiy = min(max(__double2int_rn(fma(o.y, fx_scale, fx_oy)), -lim), lim);
                    const int oiz = min(max(__double2int_rn(fma(o.z, fx_scale, fx_oz)), -lim), lim);
                    const T sc32 = T(fx_scale);
                    const V3<T> q0 = r0 * sc32, q1 = r1 * sc32, q2 = r2 * sc32;
                    for (int s = s0; s < s1; s++) {
                        const float4 c = sph[s];
                        put_fixed(s, oix + __float2int_rn(q0.x * c.x + q1.x * c.y + q2.x * c.z),
                                  oiy + __float2int_rn(q0.y * c.x + q1.y * c.y + q2.y * c.z),
                                  oiz + __float2int_rn(q0.z * c.x + q1.z * c.y + q2.z * c.z));
                    }
                } else {
                    for (int s = s0; s < s1; s++) {
                        const float4 c = sph[s];
                        const V3<T> x = {oh.x + ((r0.x * c.x + r1.x * c.y + r2.x * c.z) + ol.x),
                                         oh.y + ((r0.y * c.x + r1.y * c.y + r2.y * c.z) + ol.y),
                                         oh.z + ((r0.z * c.x + r1.z * c.y + r2.z * c.z) + ol.z)};
                        put(s, x);
                    }
                }
            }
        }
    }
};

// sphere centres of one halo row (chunked paths only): kept out of line so the chain walk is not inlined twice
template <typename T, typename FT, bool kFk64, int PW = 4>
__device__ __noinline__ void halo_sphere_row(const T* qrow, T* sph_row, const float4* sph, const int* link_start,
                                             FT ox, FT oy, FT oz, int quirk, FT fx_scale, int fx_shift,
                                             FT fx_ox, FT fx_oy, FT fx_oz) {
    CollVisit<T, FT, false, PW> hv;
    hv.sph = sph; hv.link_start = link_start; hv.quirk = quirk;
    hv.off = {ox, oy, oz};
    hv.gm = nullptr;
    hv.fx_scale = fx_scale; hv.fx_shift = fx_shift; hv.fx_ox = fx_ox; hv.fx_oy = fx_oy; hv.fx_oz = fx_oz;
    hv.pos_row = sph_row;
    walk_chain<FT, 7, kFk64>(qrow, hv);
}

// One active sphere of a waypoint -> its seven joint contributions z = J^T (J J^T + ridge I)^-1 g  (or J^T g).
// nrow[k]: shared-memory rows of waypoints i-3 .. i+3 (clamped at the ends of the path, torm_trajectory.cpp:81-96).
// NANC = joints that move the sphere (1..7): compile-time, so the masked joints cost nothing (the warp walks the spheres
// in table order: the link, and with it NANC, is uniform over the warp).  acc -= z.
template <typename T, int NANC, int PW = 4, bool kFix = false>
RLITG_DEV void coll_sphere(const TormP<T>& P, const float4 rec, const T pot, const T* const (&nrow)[7], const V3<T> (&ja)[7],
                           const V3<T> (&jb)[7], int s, T (&acc)[7]) {
    const T* pc = nrow[3] + PW * s;
    // sphere velocity / acceleration: 7-point stencils across waypoints (:1364-1381), differenced against the centre row
    V3<T> xc, vel = {T(0), T(0), T(0)}, ac = {T(0), T(0), T(0)};
    if constexpr (kFix) {
        // fixed-point centres: the rules are antisymmetric (velocity) / symmetric (acceleration) about the centre row, so the
        // rows are paired in exact integer arithmetic first -- x(i+k) - x(i-k) and x(i+k) + x(i-k) - 2 x(i), modulo 2^32:
        // the results are small -- and each pair is rounded to FP32 once: 6 conversions and 6 FMAs per coordinate for the
        // two stencils (the FP32 form needs 6 subtractions and 12 FMAs).  rvs / ras = the rule entries x the unit in metres.
        const unsigned cix = __float_as_uint(pc[0]), ciy = __float_as_uint(pc[1]), ciz = __float_as_uint(pc[2]);
        xc = {fma_t(T((int)cix), P.fix_unit, P.fix_org[0]), fma_t(T((int)ciy), P.fix_unit, P.fix_org[1]),
              fma_t(T((int)ciz), P.fix_unit, P.fix_org[2])};
        const unsigned m2x = 0u - 2u * cix, m2y = 0u - 2u * ciy, m2z = 0u - 2u * ciz;
#pragma unroll
        for (int k = 4; k < 7; k++) {
            const T* pp = nrow[k] + PW * s;
            const T* pm = nrow[6 - k] + PW * s;
            const unsigned px = __float_as_uint(pp[0]), py = __float_as_uint(pp[1]), pz = __float_as_uint(pp[2]);
            const unsigned qx = __float_as_uint(pm[0]), qy = __float_as_uint(pm[1]), qz = __float_as_uint(pm[2]);
            const V3<T> dv = {T((int)(px - qx)), T((int)(py - qy)), T((int)(pz - qz))};
            const V3<T> da = {T((int)(px + qx + m2x)), T((int)(py + qy + m2y)), T((int)(pz + qz + m2z))};
            vel = axpy(vel, P.rvs[k], dv);   // rule_v[k] = -rule_v[6 - k]
            ac = axpy(ac, P.ras[k], da);     // rule_a[k] =  rule_a[6 - k]
        }
    } else {
        xc = {pc[0], pc[1], pc[2]};
#pragma unroll
        for (int k = 0; k < 7; k++) {
            if (k == 3) continue;
            const T* pk = nrow[k] + PW * s;
            const V3<T> d = {pk[0] - xc.x, pk[1] - xc.y, pk[2] - xc.z};
            vel = axpy(vel, P.rv[k], d);
            ac = axpy(ac, P.ra[k], d);
        }
    }
    T vel_mag, inv_vm;
    if constexpr (sizeof(T) == 4) {
        inv_vm = rsqrtf(dot(vel, vel));
        vel_mag = dot(vel, vel) * inv_vm;
    } else {
        vel_mag = sqrt_t(dot(vel, vel));
        inv_vm = T(1) / vel_mag;
    }
    const V3<T> nv = vel * inv_vm;
    const V3<T> pg = {-T(rec.y), -T(rec.z), -T(rec.w)};
    // orthogonal projector P = I - nv nv^T ; curvature ; Cartesian gradient (:1529-1537)
    const V3<T> Pa = axpy(ac, -dot(nv, ac), nv);
    const V3<T> Pg = axpy(pg, -dot(nv, pg), nv);
    const V3<T> curv = Pa * (inv_vm * inv_vm);
    const V3<T> g = (Pg - curv * pot) * vel_mag;
    // Jacobian columns (ancestor-masked: joints >= nanc do not move this sphere, :1170-1178)
    T cx[NANC], cy[NANC], cz[NANC];
#pragma unroll
    for (int j = 0; j < NANC; j++) {
        const V3<T> a = ja[j], b = jb[j];
        cx[j] = fma_t(a.y, xc.z, fma_t(-a.z, xc.y, b.x));
        cy[j] = fma_t(a.z, xc.x, fma_t(-a.x, xc.z, b.y));
        cz[j] = fma_t(a.x, xc.y, fma_t(-a.y, xc.x, b.z));
    }
    if (!P.use_pinv) {
#pragma unroll
        for (int j = 0; j < NANC; j++) acc[j] -= cx[j] * g.x + cy[j] * g.y + cz[j] * g.z;
        return;
    }
    // calculatePseudoInverse :767-772.  A = [J^T ; sl I] (sl = sqrt(ridge)), columns a0 a1 a2 of length NANC + 3; A = Q R by
    // modified Gram-Schmidt; J^T (A^T A)^-1 g = Q1 R^-T g.  The three tail entries of each column are carried as scalars.
    const T sl = P.sqrt_ridge, lam = P.ridge_t;
    T n0 = lam, d01 = T(0), d02 = T(0);
#pragma unroll
    for (int j = 0; j < NANC; j++) { n0 += cx[j] * cx[j]; d01 += cx[j] * cy[j]; d02 += cx[j] * cz[j]; }
    const T i0 = rsqrt_t(n0);           // 1 / r00
    const T r01 = d01 * i0, r02 = d02 * i0;
    const T f1 = r01 * i0, f2 = r02 * i0;
#pragma unroll
    for (int j = 0; j < NANC; j++) { cy[j] -= f1 * cx[j]; cz[j] -= f2 * cx[j]; }
    const T t1a = -f1 * sl, t2a = -f2 * sl;  // tails: a0 = (sl, 0, 0), a1 = (t1a, sl, 0), a2 = (t2a, 0, sl)
    T n1 = t1a * t1a + lam, d12 = t1a * t2a;
#pragma unroll
    for (int j = 0; j < NANC; j++) { n1 += cy[j] * cy[j]; d12 += cy[j] * cz[j]; }
    const T i1 = rsqrt_t(n1);
    const T r12 = d12 * i1;
    const T f3 = r12 * i1;
#pragma unroll
    for (int j = 0; j < NANC; j++) cz[j] -= f3 * cy[j];
    const T t2b = t2a - f3 * t1a, t2c = -f3 * sl;  // a2 tail = (t2b, t2c, sl)
    T n2 = t2b * t2b + t2c * t2c + lam;
#pragma unroll
    for (int j = 0; j < NANC; j++) n2 += cz[j] * cz[j];
    const T i2 = rsqrt_t(n2);
    // R^T w = g
    const T w0 = g.x * i0;
    const T w1 = (g.y - r01 * w0) * i1;
    const T w2 = (g.z - r02 * w0 - r12 * w1) * i2;
    const T u0 = w0 * i0, u1 = w1 * i1, u2 = w2 * i2;
#pragma unroll
    for (int j = 0; j < NANC; j++) acc[j] -= cx[j] * u0 + cy[j] * u1 + cz[j] * u2;
}

#define TROW(row) (s_traj + (min(max((row), 0), Tn - 1) - c0 + 6) * 7)
#define PROW(row) (s_pos + (size_t)(min(max((row), 0), Tn - 1) - c0 + 3) * PS)

// kFk64: the chain walk of phase 1 runs in FP64 inside the FP32 build (see the header); always T for the FP64 build.
// Writes costs[5] and active[]; the other cost entries are zeroed when the pose / evaluator kernel does not run.
template <typename T, bool kFk64, int MAXREG, int PW>
__global__ void __launch_bounds__(kCollThreads) __maxnreg__(MAXREG)
torm_coll_kernel(TormP<T> P, const T* __restrict__ traj, SceneDev sc, int S, const T* __restrict__ Aband, int stage_mask,
                 int zero_pose_costs, const int32_t* __restrict__ stoch, T* __restrict__ inc_smooth,
                 T* __restrict__ inc_coll, T* __restrict__ costs, float* __restrict__ active,
                 const int32_t* __restrict__ traj_stage) {
    using FT = typename std::conditional<kFk64, double, T>::type;
    const int CH = blockDim.x, tid = threadIdx.x, Tn = P.Tn;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_traj = reinterpret_cast<T*>(smem_raw);  // [(CH + 12)][7]
    T* s_red = s_traj + (CH + 12) * 7;           // [8][2]
    T* s_pos = s_red + 8 * 2;                    // [(min(CH, T) + 6)][PS] rows c0-3 .. c0+CH+2, PS = 4 S padded to odd
    const int PS = (PW * S) | 1;                 // odd row stride: thread-per-row accesses are bank-conflict free
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_start[kNumLinks + 1];
    __shared__ int s_nanc[kMaxSpheres];  // joints that move sphere s: min(link + 1, 7) (torm_optimizer.cpp:263-309)

    const bool do_coll = (stage_mask & RLITG_STAGE_COLL) != 0;
    const bool do_smooth = (stage_mask & RLITG_STAGE_SMOOTH) != 0;
    if (do_coll) {
        for (int i = tid; i < S; i += CH) {
            s_sph[i] = sc.spheres[i];
            int link = 0;
            while (link < kNumLinks - 1 && i >= sc.link_start[link + 1]) link++;
            s_nanc[i] = link + 1 < 7 ? link + 1 : 7;
        }
        if (tid <= kNumLinks) s_start[tid] = sc.link_start[tid];
    }
    const int64_t n = blockIdx.x;
    if (traj_stage && traj_stage[n] != 0) {  // this trajectory takes the end-pose stage: nothing to do here
        if (tid == 0) {
            T* c = costs + n * RLITG_TORM_NCOST;
            c[5] = T(0);
            if (zero_pose_costs) {
                c[0] = c[1] = c[2] = c[3] = c[4] = T(0);
                c[6] = c[7] = c[8] = c[9] = T(0);
            }
            if (active) active[n] = 0.f;
        }
        return;
    }
    const T* tr = traj + n * (int64_t)Tn * 7;
    const int F = Tn - 1;
    T* o_smooth = inc_smooth ? inc_smooth + n * (int64_t)F * 7 : nullptr;
    T* o_coll = inc_coll ? inc_coll + n * (int64_t)F * 7 : nullptr;
    const int stoch_row = stoch ? stoch[n] + 1 : -1;  // full-trajectory row of the single stochastic waypoint
    // RLITG_STAGE_COLL_ROW_ONLY: only that waypoint's increment is wanted (the optimiser loop with use_stochastic_descent_:
    // getCollisionCost is not part of localOptimizeTSGD) -> chain walks for rows stoch_row - 3 .. + 3, potentials of one row
    const bool row_only = stoch != nullptr && (stage_mask & RLITG_STAGE_COLL_ROW_ONLY) != 0;
    const GridMap<T> gm(sc);
    T red[2] = {T(0), T(0)};  // potential sum, active sphere count

    for (int c0 = 0; c0 < Tn; c0 += CH) {
        __syncthreads();
        stage_traj_window<T>(s_traj, tr, c0, Tn, CH, tid);
        __syncthreads();
        const int i = c0 + tid;
        const bool valid = i < Tn;

        // ---- smoothness increments: -(2 A_full x) on the free rows (:1466-1473); group row g = i + 5
        if (do_smooth && valid && i >= 1 && o_smooth) {
            const T* ab = Aband + (size_t)(i + 5) * 13;
            T coef[13];
#pragma unroll
            for (int k = 0; k < 13; k++) coef[k] = ab[k];
#pragma unroll
            for (int j = 0; j < 7; j++) {
                T sacc = T(0);
#pragma unroll
                for (int k = 0; k < 13; k++) sacc += coef[k] * (T(2) * TROW(i + k - 6)[j]);
                o_smooth[(i - 1) * 7 + j] = -sacc;
            }
        }
        if (!do_coll) continue;

        // ---- phase 1: chain walk of this row -> Jacobian table in registers, sphere centres in shared memory
        //      (+ the halo rows of a chunked path)
        CollVisit<T, FT, true, PW> v;
        v.sph = s_sph; v.link_start = s_start; v.quirk = P.quirk;
        v.off = {FT(P.off64[0]), FT(P.off64[1]), FT(P.off64[2])};
        v.gm = &gm;
        v.fx_scale = FT(P.fix_scale); v.fx_shift = P.fix_shift;
        v.fx_ox = -FT(sc.origin[0]) * v.fx_scale; v.fx_oy = -FT(sc.origin[1]) * v.fx_scale; v.fx_oz = -FT(sc.origin[2]) * v.fx_scale;
        if (valid && (!row_only || abs(i - stoch_row) <= 3)) {
            v.pos_row = PROW(i);
            walk_chain<FT, 7, kFk64>(TROW(i), v);
        }
        if (tid < 6) {
            const int hr = tid < 3 ? c0 - 3 + tid : c0 + CH + (tid - 3);
            if (hr >= 0 && hr < Tn && (hr < c0 || hr >= c0 + CH) && (!row_only || abs(hr - stoch_row) <= 3))
                halo_sphere_row<T, FT, kFk64, PW>(TROW(hr), PROW(hr), s_sph, s_start, v.off.x, v.off.y, v.off.z, P.quirk, v.fx_scale,
                                                  v.fx_shift, v.fx_ox, v.fx_oy, v.fx_oz);
        }
        // the first two distance records of this row (its own cells, written above by this thread): issued before the
        // barrier so that their L2 round trip overlaps the wait for the slowest walker
        const uint32_t a_cell = smem_u32(PROW(i)) + 3 * (uint32_t)sizeof(T);
        const T* prow_own = PROW(i);
        auto fetch = [&](int s) -> float4 {
            int c;
            if constexpr (PW == 4) {
                c = lds_cell<T>(a_cell + (uint32_t)s * 4u * (uint32_t)sizeof(T));
            } else {  // three-value rows (three resident CTAs per SM): the cell index is recomputed from the centre
                const T* pp = prow_own + 3 * s;
                c = (int)gm.cell(V3<T>{pp[0], pp[1], pp[2]});
            }
            return c >= 0 ? __ldg(sc.field + c) : make_float4(1e30f, 0.f, 0.f, 0.f);
        };
        float4 r0 = make_float4(1e30f, 0.f, 0.f, 0.f), r1 = r0;
        if (valid && !(row_only && i != stoch_row)) {
            r0 = fetch(0);
            if (S > 1) r1 = fetch(1);
        }
        __syncthreads();

        // ---- phase 2: potentials and cost over all spheres, then the CHOMP increments of this row's active spheres
        if (valid && row_only && i != stoch_row) {
            if (i >= 1 && o_coll) {  // rows that take no increment are zero (:1495)
#pragma unroll
                for (int j = 0; j < 7; j++) o_coll[(i - 1) * 7 + j] = T(0);
            }
        } else if (valid) {
            const T* prow = PROW(i);
            const bool inc_row = (i >= 1) && (stoch_row < 0 || stoch_row == i) && (o_coll != nullptr);
            T acc[7];
#pragma unroll
            for (int j = 0; j < 7; j++) acc[j] = T(0);
            const T* nrow[7];  // the seven stencil rows (clamped), resolved once per waypoint
#pragma unroll
            for (int k = 0; k < 7; k++) nrow[k] = PROW(i + k - 3);
            T csum = T(0);
            int nact = 0;
            // distance-record gathers run two spheres ahead of the arithmetic (L2 latency off the critical path)
            // shared-memory addresses of the three tables the loop head reads, formed ONCE: a static / dynamic shared
            // symbol's address is (cluster CTA id << 24) + offset on this architecture, and under the register cap the
            // compiler re-read SR_CgaCtaId twice per sphere to rebuild it (7 % of the stall samples, round 2)
            // (the identity shuffle makes the value opaque to the register allocator's rematerialisation)
            const unsigned am = __activemask();
            const int ln = threadIdx.x & 31;
            const uint32_t a_sphw = __shfl_sync(am, smem_u32(s_sph) + 12u, ln);
            const uint32_t a_nanc = __shfl_sync(am, smem_u32(s_nanc), ln);
            for (int s = 0; s < S; s++) {
                const float4 rec = r0;
                r0 = r1;
                if (s + 2 < S) r1 = fetch(s + 2);
                if (!(rec.x < 1e29f)) continue;
                const T pot = get_potential(T(rec.x), T(lds_f32(a_sphw + 16u * (uint32_t)s)), P.clearance);
                csum += pot;
                if (!inc_row || pot < T(0.0001)) continue;
                nact++;
                constexpr bool FX = CollVisit<T, FT, true, PW>::kFix;
                switch (lds_s32(a_nanc + 4u * (uint32_t)s)) {  // uniform over the warp: every lane is on sphere s
                    case 1: coll_sphere<T, 1, PW, FX>(P, rec, pot, nrow, v.a, v.b, s, acc); break;
                    case 2: coll_sphere<T, 2, PW, FX>(P, rec, pot, nrow, v.a, v.b, s, acc); break;
                    case 3: coll_sphere<T, 3, PW, FX>(P, rec, pot, nrow, v.a, v.b, s, acc); break;
                    case 4: coll_sphere<T, 4, PW, FX>(P, rec, pot, nrow, v.a, v.b, s, acc); break;
                    case 5: coll_sphere<T, 5, PW, FX>(P, rec, pot, nrow, v.a, v.b, s, acc); break;
                    case 6: coll_sphere<T, 6, PW, FX>(P, rec, pot, nrow, v.a, v.b, s, acc); break;
                    default: coll_sphere<T, 7, PW, FX>(P, rec, pot, nrow, v.a, v.b, s, acc); break;
                }
            }
            red[0] += csum;
            red[1] += T(nact);
            // collision_increments_.row(i - free_vars_start_) -= ... (:1545-1550); rows that take no increment are zero (:1495)
            if (i >= 1 && o_coll) {
#pragma unroll
                for (int j = 0; j < 7; j++) o_coll[(i - 1) * 7 + j] = acc[j];
            }
        }
    }
    T tot[2];
    block_reduce<T, 2>(red, s_red, tot);
    if (tid == 0) {
        T* c = costs + n * RLITG_TORM_NCOST;
        c[5] = (do_coll && !row_only) ? tot[0] / T(Tn) : T(0);  // getCollisionCost :1448-1463: mean over the T real rows
        if (zero_pose_costs) {
            c[0] = c[1] = c[2] = c[3] = c[4] = T(0);
            c[6] = c[7] = c[8] = c[9] = T(0);
        }
        if (active) {
            const int rows = (stoch_row >= 0) ? 1 : Tn - 1;
            active[n] = do_coll ? (float)(tot[1] / T(rows)) : 0.f;
        }
    }
}
#undef TROW
#undef PROW

// ---- kernel 3c: the collision increment of ONE waypoint per trajectory (use_stochastic_descent_, :1502-1515) ----------
// With a single waypoint per trajectory the thread-per-waypoint kernel above keeps two CTAs per SM busy with one thread
// each walking 24 spheres in turn (89 us for 1,024 trajectories, measured: pure latency).  Here a WARP takes a
// trajectory and the sphere axis goes on the lanes: lanes 0..6 walk the chain of the seven stencil rows, then lane s
// takes sphere s (gather, potential, stencil, projector, Jacobian columns from the centre row's table in shared memory,
// solve), and the seven joint contributions are summed over the lanes by a fixed butterfly.  All other rows of the
// increment are written as zeros (:1495).
constexpr int kRowWarps = 4;  // trajectories per CTA

template <typename T, bool kFk64>
__global__ void __launch_bounds__(kRowWarps * 32)
torm_coll_row_kernel(TormP<T> P, const T* __restrict__ traj, SceneDev sc, int S, int zero_costs, int64_t N,
                     const int32_t* __restrict__ stoch, T* __restrict__ inc_coll, T* __restrict__ costs,
                     float* __restrict__ active, const int32_t* __restrict__ traj_stage) {
    using FT = typename std::conditional<kFk64, double, T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int PS = (4 * S) | 1, Tn = P.Tn, F = Tn - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* s_pos = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (7 * PS + 42);  // [7][PS] stencil rows
    T* s_tab = s_pos + 7 * PS;                                                // [42] (a_j, b_j) of the centre row
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_start[kNumLinks + 1];
    __shared__ int s_nanc[kMaxSpheres];
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        s_sph[i] = sc.spheres[i];
        int link = 0;
        while (link < kNumLinks - 1 && i >= sc.link_start[link + 1]) link++;
        s_nanc[i] = link + 1 < 7 ? link + 1 : 7;
    }
    if (threadIdx.x <= kNumLinks) s_start[threadIdx.x] = sc.link_start[threadIdx.x];
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * kRowWarps + warp;
    if (n >= N) return;
    T* c = costs + n * RLITG_TORM_NCOST;
    if (lane == 0 && zero_costs) {
        c[5] = T(0);
        if (zero_costs > 1) {
            c[0] = c[1] = c[2] = c[3] = c[4] = T(0);
            c[6] = c[7] = c[8] = c[9] = T(0);
        }
    }
    if (traj_stage && traj_stage[n] != 0) {  // end-pose stage: the increment rows are left untouched
        if (lane == 0 && active) active[n] = 0.f;
        return;
    }
    const T* tr = traj + n * (int64_t)Tn * 7;
    T* o_coll = inc_coll + n * (int64_t)F * 7;
    const int row = stoch[n] + 1;  // full-trajectory row; 0 = the start row (index -1): no increment is taken
    for (int k = lane; k < F * 7; k += 32) o_coll[k] = T(0);
    if (row < 1 || row > Tn - 1) {
        if (lane == 0 && active) active[n] = 0.f;
        return;
    }
    CollVisit<T, FT, true> v;
    const GridMap<T> gm(sc);
    if (lane < 7) {
        v.sph = s_sph; v.link_start = s_start; v.quirk = P.quirk;
        v.off = {FT(P.off64[0]), FT(P.off64[1]), FT(P.off64[2])};
        v.gm = lane == 3 ? &gm : nullptr;  // the centre row keeps its grid cells (decided as in kernel 3b)
        v.fx_scale = FT(P.fix_scale); v.fx_shift = P.fix_shift;
        v.fx_ox = -FT(sc.origin[0]) * v.fx_scale; v.fx_oy = -FT(sc.origin[1]) * v.fx_scale; v.fx_oz = -FT(sc.origin[2]) * v.fx_scale;
        v.pos_row = s_pos + lane * PS;
        const int r = min(max(row - 3 + lane, 0), Tn - 1);  // clamped at the ends of the path (torm_trajectory.cpp:81-96)
        walk_chain<FT, 7, kFk64>(tr + (size_t)r * 7, v);
        if (lane == 3) {
#pragma unroll
            for (int j = 0; j < 7; j++) {
                s_tab[6 * j + 0] = v.a[j].x; s_tab[6 * j + 1] = v.a[j].y; s_tab[6 * j + 2] = v.a[j].z;
                s_tab[6 * j + 3] = v.b[j].x; s_tab[6 * j + 4] = v.b[j].y; s_tab[6 * j + 5] = v.b[j].z;
            }
        }
    }
    __syncwarp();
    T acc[7];
#pragma unroll
    for (int j = 0; j < 7; j++) acc[j] = T(0);
    int nact = 0;
    for (int s = lane; s < S; s += 32) {
        const T* nrow[7];
#pragma unroll
        for (int k = 0; k < 7; k++) nrow[k] = s_pos + k * PS;
        const int cell = bits_to_cell(nrow[3][4 * s + 3]);
        if (cell < 0) continue;
        const float4 rec = __ldg(sc.field + cell);
        if (!(rec.x < 1e29f)) continue;
        const T pot = get_potential(T(rec.x), T(s_sph[s].w), P.clearance);
        if (pot < T(0.0001)) continue;
        nact++;
        V3<T> ja[7], jb[7];
#pragma unroll
        for (int j = 0; j < 7; j++) {
            ja[j] = {s_tab[6 * j], s_tab[6 * j + 1], s_tab[6 * j + 2]};
            jb[j] = {s_tab[6 * j + 3], s_tab[6 * j + 4], s_tab[6 * j + 5]};
        }
        constexpr bool FX = CollVisit<T, FT, true>::kFix;
        switch (s_nanc[s]) {
            case 1: coll_sphere<T, 1, 4, FX>(P, rec, pot, nrow, ja, jb, s, acc); break;
            case 2: coll_sphere<T, 2, 4, FX>(P, rec, pot, nrow, ja, jb, s, acc); break;
            case 3: coll_sphere<T, 3, 4, FX>(P, rec, pot, nrow, ja, jb, s, acc); break;
            case 4: coll_sphere<T, 4, 4, FX>(P, rec, pot, nrow, ja, jb, s, acc); break;
            case 5: coll_sphere<T, 5, 4, FX>(P, rec, pot, nrow, ja, jb, s, acc); break;
            case 6: coll_sphere<T, 6, 4, FX>(P, rec, pot, nrow, ja, jb, s, acc); break;
            default: coll_sphere<T, 7, 4, FX>(P, rec, pot, nrow, ja, jb, s, acc); break;
        }
    }
    // sum over the lanes: butterfly, the same order on every run (the thread-per-waypoint kernel adds the spheres in table
    // order instead: the two agree to rounding, not bit for bit)
#pragma unroll
    for (int j = 0; j < 7; j++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nact += __shfl_xor_sync(0xffffffffu, nact, o);
    __syncwarp();  // the zero fill above and the row below come from different lanes
    if (lane < 7) {
        T val = acc[0];
#pragma unroll
        for (int j = 1; j < 7; j++) val = (lane == j) ? acc[j] : val;
        o_coll[(size_t)(row - 1) * 7 + lane] = val;
    }
    if (lane == 0 && active) active[n] = (float)nact;
}

}  // namespace rlitg
