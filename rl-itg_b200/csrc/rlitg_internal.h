// Internal host-side plumbing shared by the .cu translation units: error reporting, launch accounting and
// the typed launcher prototypes behind the C ABI (include/rlitg.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/rlitg.h"

namespace rlitg {

int set_error(int code, const char* msg);
int set_cuda_error(const char* where);            // captures cudaGetLastError()
int check_launch(const char* kernel_name);        // counts the launch, reports launch errors
const char* last_error_cstr();
// Thread-safe and per device: makes sure kernel `func` may be launched with `bytes` of dynamic shared memory on
// the current device (cudaFuncSetAttribute is per device and must not race).  Returns RLITG_OK or an error code.
int ensure_dynamic_smem(const void* func, size_t bytes, const char* name);
int current_sm_count();

template <typename T>
int launch_fk_jac(const T* q, int64_t B, int D, int flags, T* pos, T* rot, T* quat, T* J, cudaStream_t st);
template <typename T>
int launch_allfk(const T* q, int64_t B, int D, int flags, T* out, cudaStream_t st);

// Smoothness operator of TormCost, built on the host in FP64 (torm.cu: smooth_build) and mirrored on the device.
struct Smooth {
    int T = 0;
    double scale = 0;            // max coefficient of the unscaled inverse (the reference's `// 96.77624`)
    double mcs = 1.0;            // max_cost_scale = max(1, scale)
    std::vector<double> band;    // [(T+10)][13], scaled quad_cost_full_
    std::vector<double> inv;     // [(T-1)][(T-1)], scaled quad_cost_inv_
    std::vector<double> chol;    // [2][(T-1)][8] banded Cholesky factor of the unscaled free block: forward / backward substitution rows
    double* d_band64 = nullptr;
    float* d_band32 = nullptr;
    double* d_inv64 = nullptr;
    double* d_chol = nullptr;
    // partitioned substitution (torm_update.cu): the free rows in `parts` blocks of `rows` rows; spike[0][i][m] = response
    // of y_i (L y = b) to the m-th row before its block, spike[1][i][m] = response of x_i (L^T x = y) to the m-th row
    // behind its block; zero for the first / last block
    int parts = 1, rows = 0;
    std::vector<double> spike;   // [2][(T-1)][6]
    double* d_spike = nullptr;
};
const double* smooth_inv_host(const Smooth* s);
int smooth_T(const Smooth* s);
struct Scene;     // scene.cu
struct SceneDev;  // scene.cuh
bool scene_ready(const Scene* s, bool need_grid, bool need_spheres);
SceneDev scene_dev(const Scene* s);

}  // namespace rlitg
