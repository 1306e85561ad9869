// Internal host-side plumbing shared by the .cu translation units: error reporting, launch accounting and
// the typed launcher prototypes behind the C ABI (include/rlitg.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rlitg.h"

namespace rlitg {

int set_error(int code, const char* msg);
int set_cuda_error(const char* where);            // captures cudaGetLastError()
int check_launch(const char* kernel_name);        // counts the launch, reports launch errors
const char* last_error_cstr();

template <typename T>
int launch_fk_jac(const T* q, int64_t B, int D, int flags, T* pos, T* rot, T* quat, T* J, cudaStream_t st);
template <typename T>
int launch_allfk(const T* q, int64_t B, int D, int flags, T* out, cudaStream_t st);

struct Smooth;    // torm.cu
const double* smooth_inv_host(const Smooth* s);
int smooth_T(const Smooth* s);
struct Scene;     // scene.cu
struct SceneDev;  // scene.cuh
bool scene_ready(const Scene* s, bool need_grid, bool need_spheres);
SceneDev scene_dev(const Scene* s);

}  // namespace rlitg
