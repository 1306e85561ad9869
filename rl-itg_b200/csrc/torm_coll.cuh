// Kernel 3b: sphere potentials + CHOMP collision increments + smoothness increments (row A7),
// mapped onto the SPHERE axis.
//
// What it replaces in the reference (serial, one trajectory at a time, FP64, MoveIt + Eigen):
//   performForwardKinematics      torm/src/torm_optimizer.cpp:1302-1382   sphere positions, field queries, stencils
//   computeJointProperties        :1125-1162                              joint origins / axes in the planning frame
//   getJacobian                   :1164-1187                              3 x D positional Jacobian, ancestor-masked
//   calculateCollisionIncrements  :1484-1553 + calculatePseudoInverse :767-772, getPotential torm_optimizer.h:146-168
//   getCollisionCost              :1448-1463
//   calculateSmoothnessIncrements :1466-1473
//
// Mapping.  A trajectory is cut into chunks of R <= 64 rows; one CTA of 256 threads works on a chunk and the CTAs of
// one trajectory form a thread-block CLUSTER (<= 8 CTAs; longer paths loop), so the only cross-chunk quantities -- the
// collision-cost sum and the active-sphere count -- are reduced through distributed shared memory in a fixed order:
// deterministic, no global atomics, no scratch buffer.  Per chunk:
//   phase 1  wide, in four steps (a thread-per-row walk left 5 of 8 warps at the barrier for 60 % of the kernel):
//            1a lane per (row, joint): sin / cos;  1b lane per (row, coordinate): the chain walk -- every row of [R | p]
//            evolves independently under right-multiplication by the joint transforms, so three lanes share a
//            configuration -- writing the eight link frames;  1c lane per (row, sphere): centre (one 16-byte record
//            (x, y, z, cell) per sphere), one 4-byte distance gather, potential; lane per (row, joint): Jacobian table
//   compact  thread per row: cost sum and 64-bit mask of active spheres (potential >= 1e-4, :1526) in sphere order,
//            warp-shuffle prefix sum of the counts -> ordered list of active (row, sphere) pairs
//   phase 2  lane per ACTIVE PAIR (round 1 ran a thread per row over all 24 spheres: 22 of 24 body passes executed per
//            warp for 14.4 active spheres): 7-point stencil, projector / curvature, 3 x k Jacobian, solve, 7 joint
//            contributions into a per-batch shared buffer
//   phase 3  thread per (row, joint): sums its row's contributions in pair order (fixed order -> reproducible)
//
// FP32 accuracy (the production build; tolerance 1e-4 of the row maximum).  tools/coll_fp32_study.py
// separates the error sources: (a) FP32 sphere centres (~1.5e-7 m after a seven-joint FP32 walk) through the
// 1 / dt^2-scaled stencil, (b) the normal-equation solve (J J^T + ridge I)^-1 at cond ~ 1e4.  So
//   (a) the FP32 build walks the chain in FP64 (lean sincos, rlitg_common.cuh) and rounds each centre once to FP32; the
//       stencils difference against the centre row before scaling (the rule rows sum to 0);
//   (b) J^T (J J^T + ridge I)^-1 g is computed as Q1 R^-T g from the QR of [J^T ; sqrt(ridge) I] (modified Gram-Schmidt on
//       three columns of length k + 3): error ~ sqrt(cond) eps instead of cond eps, same instruction count as forming
//       J J^T, inverting it and applying J^T.
#pragma once
#include <cooperative_groups.h>

#include <type_traits>

namespace rlitg {

namespace cg = cooperative_groups;

constexpr int kCollThreads = 256;
constexpr int kCollMaxRows = 64;   // rows live in warps 0 and 1 of phase 1
constexpr int kCollMaxCluster = 8; // portable cluster size

template <typename T> struct Rec4 { T x, y, z, w; };  // sphere centre + cell index (bit pattern) : one LDS.128 for float
template <typename T> RLITG_DEV Rec4<T> ld_rec(const Rec4<T>* p) { return *p; }
template <> RLITG_DEV Rec4<float> ld_rec<float>(const Rec4<float>* p) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    return {v.x, v.y, v.z, v.w};
}
RLITG_DEV float cell_to_bits(int c, float) { return __int_as_float(c); }
RLITG_DEV double cell_to_bits(int c, double) { return __longlong_as_double((long long)c); }
RLITG_DEV int bits_to_cell(float b) { return __float_as_int(b); }
RLITG_DEV int bits_to_cell(double b) { return (int)__double_as_longlong(b); }

// shared-memory layout of one chunk (byte offsets, every region 16-byte aligned); host and device agree through this
// link frame of one (row, link) as phase 1b leaves it: origin in the walk's precision, rotation in storage precision
template <typename T, typename FT>
struct alignas(16) LinkFrame {
    FT p[3];
    T R[9];  // row-major: R[3 r + c]
};
struct CollLayout {
    int R, S, PS;  // rows per chunk, spheres, records per position row
    // off_frame: link frames [(R + 6)][8] (phase 1b -> 1c), later the pair list; off_z: sin / cos [(R + 6)][7][2] (1a -> 1b),
    // then potentials [R][S] (1c -> compaction), then the contributions of one batch [256][7] (phase 2 -> 3)
    unsigned off_traj, off_pos, off_joint, off_frame, off_z, off_acc, off_rowstart, total;
};
template <typename T, typename FT>
__host__ __device__ inline CollLayout coll_layout(int R, int S) {
    CollLayout L;
    L.R = R; L.S = S; L.PS = S;
    auto al = [](unsigned b) { return (b + 15u) & ~15u; };
    auto mx = [](unsigned a, unsigned b) { return a > b ? a : b; };
    unsigned o = 0;
    L.off_traj = o; o = al(o + (unsigned)sizeof(T) * (R + 12) * 7);
    L.off_pos = o; o = al(o + (unsigned)sizeof(T) * 4 * (R + 6) * L.PS);
    L.off_joint = o; o = al(o + (unsigned)sizeof(T) * 8 * 7 * R);  // per row and joint: (a.x a.y a.z -) (b.x b.y b.z -)
    L.off_frame = o; o = al(o + mx((unsigned)sizeof(LinkFrame<T, FT>) * 8 * (R + 6), 2u * R * S));
    L.off_z = o; o = al(o + mx(mx((unsigned)sizeof(T) * 7 * kCollThreads, (unsigned)sizeof(FT) * 2 * 7 * (R + 6)),
                               (unsigned)sizeof(T) * R * S));
    L.off_acc = o; o = al(o + (unsigned)sizeof(T) * 7 * R);
    L.off_rowstart = o; o = al(o + 4u * (R + 1));
    L.total = o;
    return L;
}

// One active (row, sphere) pair -> its seven joint contributions z = J^T (J J^T + ridge I)^-1 g  (or J^T g).
// prow = this row's record row; rows i-3 .. i+3 sit at fixed offsets of PS records (phase 1 materialises the clamped
// rows at the ends of the path, torm/src/torm_trajectory.cpp:81-96).
template <typename T>
RLITG_DEV void coll_pair(const TormP<T>& P, const float4* __restrict__ field, const Rec4<T>* __restrict__ prow, int PS,
                         const T* __restrict__ jrow, float radius, int nanc, int s, T (&z)[7]) {
    const Rec4<T> xc = ld_rec(prow + s);
    // distance record of the centre cell: (d, grad d); issued first, consumed after the stencil arithmetic
    const float4 rec = __ldg(field + bits_to_cell(xc.w));
    // sphere velocity / acceleration: 7-point stencils across waypoints (:1364-1381), differenced against the centre row
    V3<T> vel = {T(0), T(0), T(0)}, ac = {T(0), T(0), T(0)};
#pragma unroll
    for (int k = 0; k < 7; k++) {
        if (k == 3) continue;
        const Rec4<T> xk = ld_rec(prow + (k - 3) * PS + s);
        const V3<T> d = {xk.x - xc.x, xk.y - xc.y, xk.z - xc.z};
        vel = axpy(vel, P.rv[k], d);
        ac = axpy(ac, P.ra[k], d);
    }
    const T pot = get_potential(T(rec.x), T(radius), P.clearance);
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
    T cx[7], cy[7], cz[7];
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const Rec4<T> a = ld_rec(reinterpret_cast<const Rec4<T>*>(jrow + j * 8));
        const Rec4<T> b = ld_rec(reinterpret_cast<const Rec4<T>*>(jrow + j * 8 + 4));
        const bool on = j < nanc;
        const T vx = fma_t(a.y, xc.z, fma_t(-a.z, xc.y, b.x)), vy = fma_t(a.z, xc.x, fma_t(-a.x, xc.z, b.y)),
                vz = fma_t(a.x, xc.y, fma_t(-a.y, xc.x, b.z));
        cx[j] = on ? vx : T(0); cy[j] = on ? vy : T(0); cz[j] = on ? vz : T(0);
    }
    if (!P.use_pinv) {
#pragma unroll
        for (int j = 0; j < 7; j++) z[j] = cx[j] * g.x + cy[j] * g.y + cz[j] * g.z;
        return;
    }
    // calculatePseudoInverse :767-772.  A = [J^T ; sl I] (sl = sqrt(ridge)), columns a0 a1 a2 of length 7 + 3; A = Q R by
    // modified Gram-Schmidt; J^T (A^T A)^-1 g = Q1 R^-T g.  The three tail entries of each column are carried as scalars.
    const T sl = P.sqrt_ridge, lam = P.ridge_t;
    T n0 = lam, d01 = T(0), d02 = T(0);
#pragma unroll
    for (int j = 0; j < 7; j++) { n0 += cx[j] * cx[j]; d01 += cx[j] * cy[j]; d02 += cx[j] * cz[j]; }
    const T i0 = rsqrt_t(n0);           // 1 / r00
    const T r01 = d01 * i0, r02 = d02 * i0;
    const T f1 = r01 * i0, f2 = r02 * i0;
#pragma unroll
    for (int j = 0; j < 7; j++) { cy[j] -= f1 * cx[j]; cz[j] -= f2 * cx[j]; }
    const T t1a = -f1 * sl, t2a = -f2 * sl;  // tails: a0 = (sl, 0, 0), a1 = (t1a, sl, 0), a2 = (t2a, 0, sl)
    T n1 = t1a * t1a + lam, d12 = t1a * t2a;
#pragma unroll
    for (int j = 0; j < 7; j++) { n1 += cy[j] * cy[j]; d12 += cy[j] * cz[j]; }
    const T i1 = rsqrt_t(n1);
    const T r12 = d12 * i1;
    const T f3 = r12 * i1;
#pragma unroll
    for (int j = 0; j < 7; j++) cz[j] -= f3 * cy[j];
    const T t2b = t2a - f3 * t1a, t2c = -f3 * sl;  // a2 tail = (t2b, t2c, sl)
    T n2 = t2b * t2b + t2c * t2c + lam;
#pragma unroll
    for (int j = 0; j < 7; j++) n2 += cz[j] * cz[j];
    const T i2 = rsqrt_t(n2);
    // R^T w = g
    const T w0 = g.x * i0;
    const T w1 = (g.y - r01 * w0) * i1;
    const T w2 = (g.z - r02 * w0 - r12 * w1) * i2;
    const T u0 = w0 * i0, u1 = w1 * i1, u2 = w2 * i2;
#pragma unroll
    for (int j = 0; j < 7; j++) z[j] = cx[j] * u0 + cy[j] * u1 + cz[j] * u2;
}

// kFk64: the chain walk of phase 1 runs in FP64 inside the FP32 build (see the header); always T for the FP64 build.
template <typename T, bool kFk64, int MINB>
__global__ void __launch_bounds__(kCollThreads, MINB)
torm_coll_kernel(TormP<T> P, const T* __restrict__ traj, SceneDev sc, int S, const T* __restrict__ Aband, int stage_mask,
                 int zero_pose_costs, const int32_t* __restrict__ stoch, T* __restrict__ inc_smooth,
                 T* __restrict__ inc_coll, T* __restrict__ costs, float* __restrict__ active,
                 const int32_t* __restrict__ traj_stage, int R, int nch) {
    using FT = typename std::conditional<kFk64, double, T>::type;
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int64_t n = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Tn = P.Tn;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const CollLayout L = coll_layout<T, FT>(R, S);
    using Frame = LinkFrame<T, FT>;
    T* s_traj = reinterpret_cast<T*>(smem_raw + L.off_traj);              // [(R + 12)][7] rows c0-6 .. c0+R+5, clamped
    Rec4<T>* s_pos = reinterpret_cast<Rec4<T>*>(smem_raw + L.off_pos);     // [(R + 6)][PS]  rows c0-3 .. c0+R+2
    T* s_joint = reinterpret_cast<T*>(smem_raw + L.off_joint);            // [R][7][8]
    Frame* s_frame = reinterpret_cast<Frame*>(smem_raw + L.off_frame);    // [(R + 6)][8] link frames (1b -> 1c)
    uint16_t* s_pair = reinterpret_cast<uint16_t*>(smem_raw + L.off_frame);  // pair list, once the frames are dead
    FT* s_sc = reinterpret_cast<FT*>(smem_raw + L.off_z);                 // [(R + 6)][7][2] sin, cos (1a -> 1b)
    T* s_pot = reinterpret_cast<T*>(smem_raw + L.off_z);                  // [R][S] potentials (1c -> compaction)
    T* s_z = reinterpret_cast<T*>(smem_raw + L.off_z);                    // [256][7] contributions of one batch
    T* s_acc = reinterpret_cast<T*>(smem_raw + L.off_acc);                // [R][7] sum of contributions per row
    int* s_rowstart = reinterpret_cast<int*>(smem_raw + L.off_rowstart);  // [R + 1]
    __shared__ float4 s_sph[kMaxSpheres];
    __shared__ int s_link[kMaxSpheres];  // link (segment tip 0..7) of sphere s
    __shared__ int s_nanc[kMaxSpheres];  // joints that move sphere s: min(link + 1, 7) (torm_optimizer.cpp:263-309)
    __shared__ int s_wtot[2];
    __shared__ T s_red[8 * 2];
    __shared__ T s_part[kCollMaxCluster][2];  // rank 0's copy collects every CTA's (cost sum, active count)

    const bool do_coll = (stage_mask & RLITG_STAGE_COLL) != 0;
    const bool do_smooth = (stage_mask & RLITG_STAGE_SMOOTH) != 0;
    if (traj_stage && traj_stage[n] != 0) {  // this trajectory takes the end-pose stage: nothing to do here
        if (rank == 0 && tid == 0) {
            T* c = costs + n * RLITG_TORM_NCOST;
            c[5] = T(0);
            if (zero_pose_costs) {
                c[0] = c[1] = c[2] = c[3] = c[4] = T(0);
                c[6] = c[7] = c[8] = c[9] = T(0);
            }
            if (active) active[n] = 0.f;
        }
        return;  // the whole cluster takes this branch: no cluster barrier is left waiting
    }
    if (do_coll) {
        for (int i = tid; i < S; i += kCollThreads) {
            s_sph[i] = sc.spheres[i];
            int link = 0;
            while (link < kNumLinks - 1 && i >= sc.link_start[link + 1]) link++;
            s_link[i] = link;
            s_nanc[i] = link + 1 < 7 ? link + 1 : 7;
        }
    }
    const T* tr = traj + n * (int64_t)Tn * 7;
    const int F = Tn - 1;
    T* o_smooth = inc_smooth ? inc_smooth + n * (int64_t)F * 7 : nullptr;
    T* o_coll = inc_coll ? inc_coll + n * (int64_t)F * 7 : nullptr;
    const int stoch_row = stoch ? stoch[n] + 1 : -1;  // full-trajectory row of the single stochastic waypoint
    const GridMap<T> gm(sc);
    const int PS = L.PS;
    T red[2] = {T(0), T(0)};  // potential sum, active sphere count

    for (int c = rank; c < nch; c += CL) {
        const int c0 = c * R;
        if (c0 >= Tn) break;
        const int Rc = min(R, Tn - c0);
        __syncthreads();
        for (int idx = tid; idx < (R + 12) * 7; idx += kCollThreads) {
            const int r = idx / 7, j = idx - r * 7;
            s_traj[idx] = tr[min(max(c0 - 6 + r, 0), Tn - 1) * 7 + j];
        }
        for (int idx = tid; idx < Rc * 7; idx += kCollThreads) s_acc[idx] = T(0);
        __syncthreads();

        // ---- smoothness increments: -(2 A_full x) on the free rows (:1466-1473); group row g = i + 5.
        //      thread per (row, joint): the store order is the memory order
        if (do_smooth && o_smooth) {
            for (int item = tid; item < Rc * 7; item += kCollThreads) {
                const int rl = item / 7, j = item - rl * 7, i = c0 + rl;
                if (i < 1) continue;
                const T* ab = Aband + (size_t)(i + 5) * 13;
                const T* xr = s_traj + rl * 7 + j;  // window row (i - 6) - (c0 - 6) = rl; the window clamps at the path ends
                T sacc = T(0);
#pragma unroll
                for (int k = 0; k < 13; k++) sacc += __ldg(ab + k) * (T(2) * xr[k * 7]);
                o_smooth[(size_t)(i - 1) * 7 + j] = -sacc;
            }
        }
        if (!do_coll) continue;

        // ---- phase 1.  Chunk-local row m = 0 .. Rc+5 holds global row c0 - 3 + m (rows outside the path are skipped here and
        //      filled with copies of the end rows below); m - 3 = 0 .. Rc-1 are the chunk's own rows, the rest is halo.
        const int M = Rc + 6;
        const V3<FT> off = {FT(P.off64[0]), FT(P.off64[1]), FT(P.off64[2])};
        // 1a: lane per (row, joint): sin / cos of the joint value
        for (int item = tid; item < M * 7; item += kCollThreads) {
            const int m = item / 7, j = item - m * 7, g = c0 - 3 + m;
            if (g < 0 || g >= Tn) continue;
            const FT q = FT(s_traj[(m + 3) * 7 + j]);
            FT sn, cs;
            if constexpr (kFk64) sincos_joint(q, sn, cs, FastTrig64{});
            else sincos_joint(q, sn, cs);
            s_sc[item * 2] = sn;
            s_sc[item * 2 + 1] = cs;
        }
        __syncthreads();
        // 1b: lane per (row, coordinate r): row r of [c0 c1 c2 | p] through the seven joints + tip (fetch_chain.cuh is the
        //     whole-frame form of the same walk); link frame k = KDL segment k's tip frame
        for (int item = tid; item < M * 3; item += kCollThreads) {
            const int m = item / 3, r = item - m * 3, g = c0 - 3 + m;
            if (g < 0 || g >= Tn) continue;
            FT e0 = r == 0 ? FT(1) : FT(0), e1 = r == 1 ? FT(1) : FT(0), e2 = r == 2 ? FT(1) : FT(0), pp = FT(0);
            const FT* scv = s_sc + (size_t)m * 14;
            Frame* fr = s_frame + (size_t)m * 8;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                pp = fma_t(FT(arm_ox<double>(j)), e0, pp);
                if (arm_oz<double>(j) != 0.0) pp = fma_t(FT(arm_oz<double>(j)), e2, pp);
                const FT sn = scv[2 * j], cs = scv[2 * j + 1];
                if (arm_axis(j) == AX_Z) { const FT a = e0, b = e1; e0 = cs * a + sn * b; e1 = cs * b - sn * a; }
                else if (arm_axis(j) == AX_Y) { const FT a = e0, b = e2; e0 = cs * a - sn * b; e2 = sn * a + cs * b; }
                else { const FT a = e1, b = e2; e1 = cs * a + sn * b; e2 = cs * b - sn * a; }
                fr[j].p[r] = pp;
                fr[j].R[3 * r] = T(e0); fr[j].R[3 * r + 1] = T(e1); fr[j].R[3 * r + 2] = T(e2);
            }
            pp = fma_t(FT(tip_ox<double>()), e0, pp);
            fr[7].p[r] = pp;
            fr[7].R[3 * r] = T(e0); fr[7].R[3 * r + 1] = T(e1); fr[7].R[3 * r + 2] = T(e2);
        }
        __syncthreads();
        // 1c: lane per (row, sphere): centre = link origin + R c (R c in storage precision: |c| <= 0.2 m -> ~1e-8 m; the sum
        //     in the walk's precision, rounded once), cell index, distance gather and potential for the chunk's own rows
        for (int item = tid; item < M * S; item += kCollThreads) {
            const int m = item / S, s = item - m * S, g = c0 - 3 + m;
            if (g < 0 || g >= Tn) continue;
            const float4 c = s_sph[s];
            const Frame& f = s_frame[(size_t)m * 8 + s_link[s]];
            Rec4<T> rec;
            rec.x = T(f.p[0] + off.x + FT(f.R[0] * T(c.x) + f.R[1] * T(c.y) + f.R[2] * T(c.z)));
            rec.y = T(f.p[1] + off.y + FT(f.R[3] * T(c.x) + f.R[4] * T(c.y) + f.R[5] * T(c.z)));
            rec.z = T(f.p[2] + off.z + FT(f.R[6] * T(c.x) + f.R[7] * T(c.y) + f.R[8] * T(c.z)));
            int cell = -1;
            const int rl = m - 3;
            if (rl >= 0 && rl < Rc) {
                const V3<T> x = {rec.x, rec.y, rec.z};
                cell = (int)gm.cell(x);
                T pot = T(0);
                if (cell >= 0) {
                    const float d = __ldg(reinterpret_cast<const float*>(sc.field + cell));
                    if (d < 1e29f) pot = get_potential(T(d), T(c.w), P.clearance);
                }
                s_pot[rl * S + s] = pot;
            }
            rec.w = cell_to_bits(cell, T(0));
            s_pos[(size_t)m * PS + s] = rec;
        }
        //     lane per (own row, joint): computeJointProperties :1125-1162 -> Jacobian table.  Column j of getJacobian
        //     (:1170-1178) = a x (x - o) = a x x + b with b = o x a; axis = joint_transform * axis (:1157) adds the joint position
        for (int item = tid; item < Rc * 7; item += kCollThreads) {
            const int rl = item / 7, j = item - rl * 7;
            const Frame& f = s_frame[(size_t)(rl + 3) * 8 + j];
            const int col = arm_axis(j) == AX_Z ? 2 : (arm_axis(j) == AX_Y ? 1 : 0);
            const V3<FT> o = {f.p[0] + off.x, f.p[1] + off.y, f.p[2] + off.z};
            V3<FT> a = {FT(f.R[col]), FT(f.R[3 + col]), FT(f.R[6 + col])};
            if (P.quirk) a = a + o;
            const V3<FT> b = cross(o, a);
            T* d = s_joint + (size_t)item * 8;
            d[0] = T(a.x); d[1] = T(a.y); d[2] = T(a.z); d[3] = T(0);
            d[4] = T(b.x); d[5] = T(b.y); d[6] = T(b.z); d[7] = T(0);
        }
        __syncthreads();
        // ---- compaction: thread per own row sums its potentials and collects its active spheres in sphere order
        //      (a fixed order: the cost is reproducible), then the ordered list of active (row, sphere) pairs
        uint64_t mask = 0;
        T csum = T(0);
        if (tid < Rc) {
            const int g_row = c0 + tid;
            const bool inc_row = (g_row >= 1) && (stoch_row < 0 || stoch_row == g_row) && (o_coll != nullptr);
            const T* pr = s_pot + tid * S;
            for (int s = 0; s < S; s++) {
                const T pot = pr[s];
                csum += pot;
                if (inc_row && !(pot < T(0.0001))) mask |= 1ull << s;
            }
        }
        const int cnt = __popcll(mask);
        red[0] += csum;
        red[1] += T(cnt);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31 && warp < 2) s_wtot[warp] = incl;
        __syncthreads();  // the potentials have been read: the batch buffer may reuse their space; frames are dead too
        const int npairs = s_wtot[0] + s_wtot[1];
        if (tid < Rc) {
            int pos = incl - cnt + (warp == 1 ? s_wtot[0] : 0);
            s_rowstart[tid] = pos;
            while (mask) {
                const int s = __ffsll((long long)mask) - 1;
                mask &= mask - 1;
                s_pair[pos++] = (uint16_t)((tid << 8) | s);
            }
            if (tid == Rc - 1) s_rowstart[Rc] = npairs;
        }
        // clamped rows at the ends of the path (the padded group trajectory repeats the first / last row): copies of row 0
        // below the path, of row Tn-1 above it, so that phase 2 reads its seven stencil rows at fixed offsets
        if (c0 < 3) {
            const Rec4<T>* src = s_pos + (size_t)(0 - (c0 - 3)) * PS;
            for (int idx = tid; idx < (3 - c0) * S; idx += kCollThreads) {
                const int r = idx / S, s = idx - r * S;
                s_pos[(size_t)r * PS + s] = src[s];
            }
        }
        if (c0 + Rc + 3 > Tn) {
            const int first = Tn - (c0 - 3);  // first smem row past the path
            const Rec4<T>* src = s_pos + (size_t)(first - 1) * PS;
            const int nrows = c0 + Rc + 3 - Tn;
            for (int idx = tid; idx < nrows * S; idx += kCollThreads) {
                const int r = idx / S, s = idx - r * S;
                s_pos[(size_t)(first + r) * PS + s] = src[s];
            }
        }
        __syncthreads();

        // ---- phase 2 / 3 in batches of 256 pairs
        for (int base = 0; base < npairs; base += kCollThreads) {
            const int p = base + tid;
            if (p < npairs) {
                const int pid = s_pair[p];
                const int rl = pid >> 8, s = pid & 0xff;
                T z[7];
                coll_pair<T>(P, sc.field, s_pos + (size_t)(rl + 3) * PS, PS, s_joint + (size_t)rl * 56, s_sph[s].w,
                             s_nanc[s], s, z);
#pragma unroll
                for (int j = 0; j < 7; j++) s_z[tid * 7 + j] = z[j];
            }
            __syncthreads();
            for (int item = tid; item < Rc * 7; item += kCollThreads) {
                const int rl = item / 7, j = item - rl * 7;
                const int lo = max(s_rowstart[rl], base), hi = min(s_rowstart[rl + 1], base + kCollThreads);
                if (lo < hi) {
                    T a = s_acc[item];
                    for (int pp = lo; pp < hi; pp++) a += s_z[(pp - base) * 7 + j];
                    s_acc[item] = a;
                }
            }
            __syncthreads();
        }
        // collision_increments_.row(i - free_vars_start_) -= ... (:1545-1550): rows that take no increment are zero (:1495)
        if (o_coll) {
            for (int item = tid; item < Rc * 7; item += kCollThreads) {
                const int i = c0 + item / 7;
                if (i >= 1) o_coll[(c0 - 1) * 7 + item] = -s_acc[item];
            }
        }
    }

    // ---- per-trajectory sums: block reduce, then across the cluster through rank 0's shared memory, fixed order
    T tot[2];
    block_reduce<T, 2>(red, s_red, tot);
    if (tid == 0) {
        T* dst = cluster.map_shared_rank(&s_part[0][0], 0);
        dst[rank * 2 + 0] = tot[0];
        dst[rank * 2 + 1] = tot[1];
    }
    cluster.sync();
    if (rank == 0 && tid == 0) {
        T csum = T(0), nact = T(0);
        for (int r = 0; r < CL; r++) { csum += s_part[r][0]; nact += s_part[r][1]; }
        T* c = costs + n * RLITG_TORM_NCOST;
        c[5] = do_coll ? csum / T(Tn) : T(0);  // getCollisionCost :1448-1463: mean over the T real rows
        if (zero_pose_costs) {
            c[0] = c[1] = c[2] = c[3] = c[4] = T(0);
            c[6] = c[7] = c[8] = c[9] = T(0);
        }
        if (active) {
            const int rows = (stoch_row >= 0) ? 1 : Tn - 1;
            active[n] = do_coll ? (float)(nact / T(rows)) : 0.f;
        }
    }
}

}  // namespace rlitg
