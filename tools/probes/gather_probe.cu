// How many L2 sectors / DRAM bytes does a per-thread 32-byte gather cost on this part?  Each thread reads ONE 32-byte
// aligned row at a pseudo-random position of a 1 GB table (no reuse), in several instruction forms.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/gather_probe tools/probes/gather_probe.cu
// run:   ncu --metrics lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,gpu__time_duration.sum tools/probes/gather_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void gather(const float* __restrict__ tab, size_t rows, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t r = ((size_t)i * 2654435761ull + 12345ull) % rows;
    const float* p = tab + r * 8;
    float v[8];
    if (MODE == 0) {  // two 128-bit loads through L1
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if (MODE == 1) {  // one 256-bit load, .nc
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
    } else if (MODE == 2) {  // one 256-bit load, .cg
        asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
    } else if (MODE == 3) {  // two 128-bit loads, .cg
        const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {  // one 128-bit load only (half a row)
        const float4 a = __ldcg(reinterpret_cast<const float4*>(p));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = v[5] = v[6] = v[7] = 0.f;
    }
    out[i] = v[0] + v[1] + v[2] + v[3] + v[4] + v[5] + v[6] + v[7];
}

int main() {
    const size_t rows = (size_t)1 << 25;  // 32 M rows x 32 B = 1 GB
    const int n = 1 << 20;
    float *tab, *out;
    cudaMalloc(&tab, rows * 32);
    cudaMalloc(&out, n * 4);
    cudaMemset(tab, 0, rows * 32);
    for (int rep = 0; rep < 2; rep++) {
        gather<0><<<n / 128, 128>>>(tab, rows, out, n);
        gather<1><<<n / 128, 128>>>(tab, rows, out, n);
        gather<2><<<n / 128, 128>>>(tab, rows, out, n);
        gather<3><<<n / 128, 128>>>(tab, rows, out, n);
        gather<4><<<n / 128, 128>>>(tab, rows, out, n);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
