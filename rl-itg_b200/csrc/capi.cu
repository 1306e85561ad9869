// extern "C" entry points of librlitg.so (include/rlitg.h): argument validation, dtype dispatch, error state.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "rlitg_common.cuh"
#include "rlitg_internal.h"

namespace rlitg {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* msg) {
    g_err = msg ? msg : "";
    return code;
}
int set_cuda_error(const char* where) {
    cudaError_t e = cudaGetLastError();
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return RLITG_ERR_CUDA;
}
int check_launch(const char* kernel_name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_err = std::string(kernel_name) + " launch failed: " + cudaGetErrorString(e);
        return RLITG_ERR_CUDA;
    }
    return RLITG_OK;
}
const char* last_error_cstr() { return g_err.c_str(); }

int ensure_dynamic_smem(const void* func, size_t bytes, const char* name) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> granted;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error("cudaGetDevice");
    std::lock_guard<std::mutex> lk(mu);
    size_t& have = granted[std::make_pair(func, dev)];
    if (bytes <= have) return RLITG_OK;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        g_err = std::string("cudaFuncSetAttribute(") + name + "): " + cudaGetErrorString(e);
        return RLITG_ERR_CUDA;
    }
    have = bytes;
    return RLITG_OK;
}

int current_sm_count() {
    static std::mutex mu;
    static std::map<int, int> sms;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    int& n = sms[dev];
    if (n == 0) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

static int check_common(const void* q, int64_t B, int D, int dtype) {
    if (B < 0) return set_error(RLITG_ERR_SHAPE, "B must be >= 0");
    if (D != 7 && D != 8) return set_error(RLITG_ERR_SHAPE, "D must be 7 (arm) or 8 (torso + arm)");
    if (dtype != RLITG_F32 && dtype != RLITG_F64) return set_error(RLITG_ERR_DTYPE, "dtype must be RLITG_F32 or RLITG_F64");
    if (B > 0 && q == nullptr) return set_error(RLITG_ERR_NULL, "q is NULL");
    const size_t es = dtype == RLITG_F32 ? 4 : 8;
    if (reinterpret_cast<uintptr_t>(q) % es) return set_error(RLITG_ERR_ALIGN, "q is not aligned to its element size");
    return RLITG_OK;
}

// ---- pinned / device workspace for the _host entry points ------------------------------------------------
struct HostPipe {
    static constexpr int kStreams = 3;
    std::mutex mu;
    cudaStream_t st[kStreams] = {};
    void* dbuf[kStreams] = {};
    size_t dcap[kStreams] = {};
    bool init = false;
    int ensure(size_t bytes_per_slot) {
        if (!init) {
            for (int i = 0; i < kStreams; i++)
                if (cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) != cudaSuccess)
                    return set_cuda_error("cudaStreamCreate");
            init = true;
        }
        for (int i = 0; i < kStreams; i++) {
            if (dcap[i] < bytes_per_slot) {
                if (dbuf[i]) cudaFree(dbuf[i]);
                if (cudaMalloc(&dbuf[i], bytes_per_slot) != cudaSuccess) { dcap[i] = 0; return set_cuda_error("cudaMalloc(workspace)"); }
                dcap[i] = bytes_per_slot;
            }
        }
        return RLITG_OK;
    }
};
static HostPipe g_pipes[64];  // one workspace per device
static HostPipe& host_pipe() {
    int dev = 0;
    cudaGetDevice(&dev);
    return g_pipes[dev & 63];
}

template <typename T>
static int fk_jac_host_t(const T* q, int64_t B, int D, int flags, T* pos, T* rot, T* quat, T* J) {
    if (flags & (RLITG_Q_SOA | RLITG_OUT_SOA))
        return set_error(RLITG_ERR_ARG, "rlitg_fk_jac_host takes array-of-structs buffers ([B,D] in, [B,*] out)");
    HostPipe& g_pipe = host_pipe();
    std::lock_guard<std::mutex> lk(g_pipe.mu);
    // chunk so that copies of one chunk overlap the kernel / copies of its neighbours
    const int64_t chunk = 1 << 17;
    const size_t per_cfg = sizeof(T) * (size_t)(D + 3 + 9 + 4 + 6 * D);
    int rc = g_pipe.ensure(per_cfg * chunk);
    if (rc) return rc;
    // On any failure the copies already queued on the other pipe streams may still be writing into the caller's host
    // buffers: drain every stream before the error goes back (the caller may free or reuse the buffers right away).
    auto fail = [&](int code) {
        for (int i = 0; i < HostPipe::kStreams; i++) cudaStreamSynchronize(g_pipe.st[i]);
        return code;
    };
    int slot = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, slot = (slot + 1) % HostPipe::kStreams) {
        const int64_t n = (B - b0 < chunk) ? B - b0 : chunk;
        cudaStream_t st = g_pipe.st[slot];
        T* dq = static_cast<T*>(g_pipe.dbuf[slot]);
        T* dpos = dq + chunk * D;
        T* drot = dpos + chunk * 3;
        T* dquat = drot + chunk * 9;
        T* dJ = dquat + chunk * 4;
        if (cudaMemcpyAsync(dq, q + b0 * D, sizeof(T) * n * D, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return fail(set_cuda_error("H2D q"));
        rc = launch_fk_jac<T>(dq, n, D, 0, pos ? dpos : nullptr, rot ? drot : nullptr, quat ? dquat : nullptr,
                              J ? dJ : nullptr, st);
        if (rc) return fail(rc);
        cudaError_t e = cudaSuccess;
        if (pos) e = cudaMemcpyAsync(pos + b0 * 3, dpos, sizeof(T) * n * 3, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && rot) e = cudaMemcpyAsync(rot + b0 * 9, drot, sizeof(T) * n * 9, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && quat) e = cudaMemcpyAsync(quat + b0 * 4, dquat, sizeof(T) * n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && J) e = cudaMemcpyAsync(J + b0 * 6 * D, dJ, sizeof(T) * n * 6 * D, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return fail(set_cuda_error("D2H outputs"));
    }
    rc = RLITG_OK;
    for (int i = 0; i < HostPipe::kStreams; i++)   // every stream is drained even when an earlier one reports an error
        if (cudaStreamSynchronize(g_pipe.st[i]) != cudaSuccess && rc == RLITG_OK) rc = set_cuda_error("cudaStreamSynchronize");
    return rc;
}

}  // namespace rlitg

using namespace rlitg;

extern "C" {

RLITG_API int rlitg_version(void) { return RLITG_VERSION; }
RLITG_API const char* rlitg_last_error(void) { return last_error_cstr(); }
RLITG_API int64_t rlitg_launch_count(void) { return g_launches.load(); }

RLITG_API int rlitg_fk_jac(const void* q, int64_t B, int D, int dtype, int flags, void* pos, void* rot, void* quat, void* J,
                 void* stream) {
    int rc = check_common(q, B, D, dtype);
    if (rc) return rc;
    if (B == 0) return RLITG_OK;
    if (!pos && !rot && !quat && !J) return set_error(RLITG_ERR_NULL, "all outputs are NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32)
        return launch_fk_jac<float>((const float*)q, B, D, flags, (float*)pos, (float*)rot, (float*)quat, (float*)J, st);
    return launch_fk_jac<double>((const double*)q, B, D, flags, (double*)pos, (double*)rot, (double*)quat, (double*)J, st);
}

RLITG_API int rlitg_fk(const void* q, int64_t B, int D, int dtype, int flags, void* pos, void* quat, void* stream) {
    if (B > 0 && (!pos || !quat)) return set_error(RLITG_ERR_NULL, "pos / quat is NULL");
    return rlitg_fk_jac(q, B, D, dtype, flags, pos, nullptr, quat, nullptr, stream);
}

RLITG_API int rlitg_jacobian(const void* q, int64_t B, int D, int dtype, int flags, void* J, void* stream) {
    if (B > 0 && !J) return set_error(RLITG_ERR_NULL, "J is NULL");
    return rlitg_fk_jac(q, B, D, dtype, flags, nullptr, nullptr, nullptr, J, stream);
}

RLITG_API int rlitg_allfk(const void* q, int64_t B, int D, int dtype, int flags, void* out, void* stream) {
    int rc = check_common(q, B, D, dtype);
    if (rc) return rc;
    if (B == 0) return RLITG_OK;
    if (!out) return set_error(RLITG_ERR_NULL, "out is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RLITG_F32) return launch_allfk<float>((const float*)q, B, D, flags, (float*)out, st);
    return launch_allfk<double>((const double*)q, B, D, flags, (double*)out, st);
}

RLITG_API int rlitg_fk_jac_host(const void* q, int64_t B, int D, int dtype, int flags, void* pos, void* rot, void* quat,
                      void* J) {
    int rc = check_common(q, B, D, dtype);
    if (rc) return rc;
    if (B == 0) return RLITG_OK;
    if (!pos && !rot && !quat && !J) return set_error(RLITG_ERR_NULL, "all outputs are NULL");
    if (dtype == RLITG_F32)
        return fk_jac_host_t<float>((const float*)q, B, D, flags, (float*)pos, (float*)rot, (float*)quat, (float*)J);
    return fk_jac_host_t<double>((const double*)q, B, D, flags, (double*)pos, (double*)rot, (double*)quat, (double*)J);
}

}  // extern "C"

// ---- FP32 FMA-pipe probe: the denominator of the TORM kernel's roofline is measured, not nominal -----------
namespace rlitg {
__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, float* out) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = 1.0f + 1e-3f * (threadIdx.x + k);
    const float b = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) a[k] = fmaf(a[k], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) s += a[k];
    if (s == 123.456f) *out = s;  // keeps the chains alive without a store in the common case
}
}  // namespace rlitg

extern "C" RLITG_API int rlitg_fma_probe(int ctas, int iters, float* out, void* stream) {
    if (ctas < 1 || iters < 1 || !out) return rlitg::set_error(RLITG_ERR_ARG, "bad fma probe arguments");
    rlitg::fma_probe_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, out);
    return rlitg::check_launch("fma_probe_kernel");
}

// ---- FP64 FMA-pipe probe (same shape as the FP32 one): decides whether an FP64 chain walk inside an FP32 kernel is
// affordable on this part (kernel 3b, torm_coll.cuh)
namespace rlitg {
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double* out) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    const double b = 0.999, c = 1e-4;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    if (s == 123.456) *out = s;
}
}  // namespace rlitg

extern "C" RLITG_API int rlitg_dfma_probe(int ctas, int iters, double* out, void* stream) {
    if (ctas < 1 || iters < 1 || !out) return rlitg::set_error(RLITG_ERR_ARG, "bad dfma probe arguments");
    rlitg::dfma_probe_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(iters, out);
    return rlitg::check_launch("dfma_probe_kernel");
}

// ---- HBM write-stream probe: the ceiling for a kernel that, like fk_jac, writes ~8x what it reads -----------
// Persistent CTAs drain `tile_bytes` shared-memory tiles to consecutive global addresses with TMA bulk stores
// (the fk_jac drain path, no arithmetic) and optionally read `read_bytes_per_tile` first.
namespace rlitg {
__global__ void __launch_bounds__(128) write_probe_kernel(char* __restrict__ dst, const float* __restrict__ src,
                                                          int64_t ntiles, int tile_bytes, int read_floats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float acc = 0.f;
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
        if (read_floats > 0) {
            const float* s = src + t * read_floats;
            for (int i = threadIdx.x; i < read_floats; i += blockDim.x) acc += __ldg(s + i);
        }
        unsigned char* stage = smem_raw + (it & 1) * tile_bytes;
        if (it >= 2) {
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        reinterpret_cast<float*>(stage)[threadIdx.x] = acc;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            rlitg::bulk_store(dst + t * tile_bytes, stage, (uint32_t)tile_bytes);  // same evict-first stores as fk_jac
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
}  // namespace rlitg

extern "C" RLITG_API int rlitg_write_probe(void* dst, const void* src, int64_t ntiles, int tile_bytes, int read_floats,
                                           int ctas, void* stream) {
    if (!dst || ntiles < 1 || tile_bytes < 512 || (tile_bytes & 15) || tile_bytes > 100 * 1024 || ctas < 1 ||
        (read_floats > 0 && !src))
        return rlitg::set_error(RLITG_ERR_ARG, "bad write probe arguments");
    auto kern = rlitg::write_probe_kernel;
    const size_t smem = 2 * (size_t)tile_bytes;
    int rc_attr = rlitg::ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem, "write_probe_kernel");
    if (rc_attr) return rc_attr;
    kern<<<ctas, 128, smem, static_cast<cudaStream_t>(stream)>>>((char*)dst, (const float*)src, ntiles, tile_bytes,
                                                                 read_floats);
    return rlitg::check_launch("write_probe_kernel");
}

// ---- device memory helpers for hosts built without the CUDA toolkit (include/rlitg.hpp) ---------------------
extern "C" {
RLITG_API int rlitg_device_count(int* count) {
    if (!count) return rlitg::set_error(RLITG_ERR_NULL, "count is NULL");
    if (cudaGetDeviceCount(count) != cudaSuccess) return rlitg::set_cuda_error("cudaGetDeviceCount");
    return RLITG_OK;
}
RLITG_API int rlitg_set_device(int device) {
    if (cudaSetDevice(device) != cudaSuccess) return rlitg::set_cuda_error("cudaSetDevice");
    return RLITG_OK;
}
RLITG_API int rlitg_dev_alloc(void** out, uint64_t bytes) {
    if (!out) return rlitg::set_error(RLITG_ERR_NULL, "out is NULL");
    *out = nullptr;
    if (bytes == 0) return RLITG_OK;
    if (cudaMalloc(out, (size_t)bytes) != cudaSuccess) return rlitg::set_cuda_error("cudaMalloc");
    return RLITG_OK;
}
RLITG_API int rlitg_dev_free(void* ptr) {
    if (ptr && cudaFree(ptr) != cudaSuccess) return rlitg::set_cuda_error("cudaFree");
    return RLITG_OK;
}
RLITG_API int rlitg_dev_zero(void* ptr, uint64_t bytes, void* stream) {
    if (bytes == 0) return RLITG_OK;
    if (!ptr) return rlitg::set_error(RLITG_ERR_NULL, "ptr is NULL");
    if (cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream) != cudaSuccess)
        return rlitg::set_cuda_error("cudaMemsetAsync");
    return RLITG_OK;
}
static int copy_any(void* dst, const void* src, uint64_t bytes, cudaMemcpyKind kind, void* stream) {
    if (bytes == 0) return RLITG_OK;
    if (!dst || !src) return rlitg::set_error(RLITG_ERR_NULL, "copy: NULL pointer");
    if (cudaMemcpyAsync(dst, src, (size_t)bytes, kind, (cudaStream_t)stream) != cudaSuccess)
        return rlitg::set_cuda_error("cudaMemcpyAsync");
    return RLITG_OK;
}
RLITG_API int rlitg_copy_h2d(void* dst_dev, const void* src_host, uint64_t bytes, void* stream) {
    return copy_any(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream);
}
RLITG_API int rlitg_copy_d2h(void* dst_host, const void* src_dev, uint64_t bytes, void* stream) {
    return copy_any(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream);
}
RLITG_API int rlitg_copy_d2d(void* dst_dev, const void* src_dev, uint64_t bytes, void* stream) {
    return copy_any(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, stream);
}
RLITG_API int rlitg_stream_sync(void* stream) {
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return rlitg::set_cuda_error("cudaStreamSynchronize");
    return RLITG_OK;
}
}  // extern "C"
