// Micro-probe: fp32 FMA issue rate on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2), 16 independent chains/thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ffma2_probe tools/ffma2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rc) : "l"(ra), "l"(rb), "l"(rc));
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float s0, float s1) {
    float2 acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
    float2 m = make_float2(s0, s1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) {   // 2 scalar FFMA
                acc[k].x = fmaf(acc[k].x, m.x, s1);
                acc[k].y = fmaf(acc[k].y, m.y, s0);
            } else if (MODE == 1) {   // 1 packed FFMA2
                acc[k] = ffma2(acc[k], m, make_float2(s1, s0));
            } else {   // packed with scalar-broadcast multiplicand
                acc[k] = ffma2(make_float2(s0, s0), acc[k], make_float2(s1, s0));
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * 8, iters = 20000;
    float* out;
    cudaMalloc(&out, sizeof(float) * grid * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[3] = {"FFMA x2", "FFMA2", "FFMA2 (scalar bcast)"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
            if (mode == 1) probe<1><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
            if (mode == 2) probe<2><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)grid * 256 * iters * 32.0;   // scalar FMAs
        printf("%-22s %.3f ms  %.2f TFLOP/s fp32  (%.1f FMA/clk/SM at 1.965 GHz)\n", names[mode], ms, 2 * fma / ms * 1e-9,
               fma / (ms * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
