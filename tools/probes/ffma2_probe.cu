// Packed FP32 arithmetic on sm_100a (fma.rn.f32x2 -> SASS FFMA2, add / mul likewise): what does one issue slot buy?
// The TORM kernels are bound by issue slots at 13 - 14 resident warps per SM, not by the FMA pipe:
// if an FFMA2 costs one issue slot for two FMAs, packing the 3-vector / column-pair arithmetic frees slots for the
// integer, shared-memory and control instructions around it.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/ffma2_probe tools/probes/ffma2_probe.cu
// run:   tools/probes/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: 16 scalar FMA chains            MODE 1: 8 packed chains (the same 16 floats)
// MODE 2: 16 packed chains (32 floats)    MODE 3: 16 scalar FMA + 16 integer ops per round
// MODE 4: 8 packed FMA + 16 integer ops   MODE 5: ONE dependent scalar chain (latency)
// MODE 6: ONE dependent packed chain      MODE 7: 16 scalar FADD chains,  MODE 8: 8 packed FADD chains
template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, float* out) {
    float a[32];
    unsigned u[16];
#pragma unroll
    for (int k = 0; k < 32; k++) a[k] = 1.0f + 1e-3f * (threadIdx.x + k);
#pragma unroll
    for (int k = 0; k < 16; k++) u[k] = threadIdx.x * 7u + k;
    const float b = 0.999f, c = 1e-4f;
    const float2 b2 = {b, b}, c2 = {c, c};
    for (int i = 0; i < iters; i++) {
        if (MODE == 0 || MODE == 3) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = fmaf(a[k], b, c);
        }
        if (MODE == 1 || MODE == 4) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float2 r = __ffma2_rn(make_float2(a[2 * k], a[2 * k + 1]), b2, c2);
                a[2 * k] = r.x; a[2 * k + 1] = r.y;
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                float2 r = __ffma2_rn(make_float2(a[2 * k], a[2 * k + 1]), b2, c2);
                a[2 * k] = r.x; a[2 * k + 1] = r.y;
            }
        }
        if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int k = 0; k < 16; k++) u[k] = (u[k] ^ (unsigned)i) + 0x9e3779b9u;  // LOP3 + IADD, or one LOP3/IADD3
        }
        if (MODE == 5) a[0] = fmaf(a[0], b, c);
        if (MODE == 6) {
            float2 r = __ffma2_rn(make_float2(a[0], a[1]), b2, c2);
            a[0] = r.x; a[1] = r.y;
        }
        if (MODE == 7) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = a[k] + c;
        }
        if (MODE == 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float2 r = __fadd2_rn(make_float2(a[2 * k], a[2 * k + 1]), c2);
                a[2 * k] = r.x; a[2 * k + 1] = r.y;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k++) s += a[k];
    unsigned x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= u[k];
    if (s == 123.456f || x == 0xdeadbeefu) *out = s + x;
}

template <int MODE>
static void run(const char* name, double flop_per_thread_iter, int ctas, int threads) {
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<ctas, threads>>>(100, out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<ctas, threads>>>(iters, out);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = flop_per_thread_iter * iters * (double)ctas * threads;
    printf("%-44s %8.3f ms  %7.2f TFLOP/s  (%.2f ns per round per thread)\n", name, ms, fl / ms * 1e-9, ms * 1e6 / iters);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %d MHz\n", p.name, sms, p.clockRate / 1000);
    run<0>("16 x FFMA", 32, sms * 8, 256);
    run<1>("8 x FFMA2 (same 16 floats)", 32, sms * 8, 256);
    run<2>("16 x FFMA2 (32 floats)", 64, sms * 8, 256);
    run<3>("16 x FFMA + 16 integer", 32, sms * 8, 256);
    run<4>("8 x FFMA2 + 16 integer", 32, sms * 8, 256);
    run<7>("16 x FADD", 16, sms * 8, 256);
    run<8>("8 x FADD2", 16, sms * 8, 256);
    run<5>("dependent FFMA chain, 1 warp / SM", 2, sms, 32);
    run<6>("dependent FFMA2 chain, 1 warp / SM", 4, sms, 32);
    return 0;
}
