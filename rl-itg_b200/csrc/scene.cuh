// Device-side view of a collision scene: link-sphere table + distance-record volume (+ bit-packed occupancy).
//
// The reference queries MoveIt's PropagationDistanceField per sphere (CollisionWorldHybrid::getCollisionGradients,
// torm/src/torm_optimizer.cpp:1281,1323): nearest cell value + central-difference gradient over six neighbours.
// Both depend only on the cell, so the volume stores one 16-byte record (d, gx, gy, gz) per cell: one gather per
// sphere instead of seven.  Cell k is centred at origin + k * res; border cells and cells
// at or beyond the maximum propagation distance hold d = 1e30, g = 0 ("untouched" in MoveIt's terms).
#pragma once
#include "rlitg_common.cuh"

namespace rlitg {

constexpr int kMaxSpheres = 64;
constexpr int kNumLinks = 8;  // segment tips of the 7-joint arm chain; 7 = gripper_link (fingers ride on it)

struct SceneDev {
    const float4* field;     // [nz][ny][nx]
    const uint32_t* bits;    // occupancy, bit-packed along x: word (z*ny + y)*wx + (x >> 5)
    const float4* spheres;   // [S] (cx, cy, cz, r) in the link frame, sorted by link
    const int* link_start;   // [kNumLinks + 1] sphere index ranges per link
    const uint32_t* dil;     // [ncls][nz][ny][wx] occupancy dilated by each distinct sphere radius: bit = (d < r)
    const int* sph_class;    // [S] radius class of each sphere (index into dil); nullptr when dil is not built
    int ncls;
    int S;
    int nx, ny, nz, wx;
    double origin[3];
    double res;
    // float copies made on the host: plain FP64 issues at 1/16 of the FP32 rate on this part, and converting these (plus
    // the 1 / res division) in every thread of an FP32 kernel cost ~500 FP64-pipe cycles per warp
    float origin_f[3];
    float res_f, inv_res_f;
};

template <typename T>
struct GridMap {  // world -> cell
    T ox, oy, oz, res, inv_res;
    int nx, ny, nz;
    RLITG_DEV explicit GridMap(const SceneDev& s) : nx(s.nx), ny(s.ny), nz(s.nz) {
        if constexpr (sizeof(T) == 8) {
            ox = s.origin[0]; oy = s.origin[1]; oz = s.origin[2]; res = s.res; inv_res = 1.0 / s.res;
        } else {
            ox = s.origin_f[0]; oy = s.origin_f[1]; oz = s.origin_f[2]; res = s.res_f; inv_res = s.inv_res_f;
        }
    }
    RLITG_DEV int cell1(T x, T o) const {
        if constexpr (sizeof(T) == 8) return (int)floor((x - o) / res + T(0.5));  // literal (oracle) form
        else return (int)floorf(fmaf(x - o, inv_res, 0.5f));
    }
    // linear cell index or -1 when outside the grid
    RLITG_DEV int64_t cell(V3<T> x) const {
        const int cx = cell1(x.x, ox), cy = cell1(x.y, oy), cz = cell1(x.z, oz);
        if (cx < 0 || cy < 0 || cz < 0 || cx >= nx || cy >= ny || cz >= nz) return -1;
        return ((int64_t)cz * ny + cy) * nx + cx;
    }
};

// distance record at a world point: returns false when out of grid / inactive (potential 0, gradient 0)
template <typename T>
RLITG_DEV bool field_lookup(const SceneDev& s, const GridMap<T>& gm, V3<T> x, T& d, V3<T>& g) {
    const int64_t c = gm.cell(x);
    if (c < 0) return false;
    const float4 r = __ldg(s.field + c);
    if (!(r.x < 1e29f)) return false;
    d = T(r.x);
    g = {T(r.y), T(r.z), T(r.w)};
    return true;
}

// Binary sphere test "signed distance at the centre cell < radius".  With the dilated bit volumes this is one
// bit of a 4-byte word (32 cells along x share it) instead of a 16-byte record; the predicate is the same.
template <typename T>
RLITG_DEV bool sphere_hits(const SceneDev& s, const GridMap<T>& gm, V3<T> x, int sphere, float radius,
                           const uint32_t* dil /* s.dil or a shared-memory copy */) {
    if (dil != nullptr) {
        const int cx = gm.cell1(x.x, gm.ox), cy = gm.cell1(x.y, gm.oy), cz = gm.cell1(x.z, gm.oz);
        if (cx < 0 || cy < 0 || cz < 0 || cx >= s.nx || cy >= s.ny || cz >= s.nz) return false;
        const size_t words = (size_t)s.wx * s.ny * s.nz;
        const uint32_t w = dil[(size_t)s.sph_class[sphere] * words + ((size_t)cz * s.ny + cy) * s.wx + (cx >> 5)];
        return (w >> (cx & 31)) & 1u;
    }
    T d;
    V3<T> g;
    return field_lookup(s, gm, x, d, g) && (float)d < radius;
}

}  // namespace rlitg
