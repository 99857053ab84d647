// Micro-probe: how fast can the GEMM epilogue's store path write a (M x N) fp32 matrix, per SM, with the tile ownership of
// gemm_tc16.cu (persistent CTA, 8 warps, warp = 32 rows x 128 columns of a 128 x 256 tile)?
//   mode 0: TMA bulk tensor stores, boxes of 32 rows x 64 B  (SWIZZLE_64B)  -- the round-1 epilogue
//   mode 1: TMA bulk tensor stores, boxes of 32 rows x 128 B (SWIZZLE_128B)
//   mode 2: st.global.v4, thread = row (32 sectors per warp instruction)     -- the pre-TMA epilogue
//   mode 3: staged through shared memory, st.global.v4 with 8 lanes per 128-byte row segment
// Ring depth R boxes per warp (TMA modes).  Prints us, GB/s and bytes/clk/SM at the measured SM clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tma_store_probe tools/tma_store_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
template <int N>
__device__ __forceinline__ void wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int MODE, int R>
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap tm, float* Y, int ld, int M, int N, int reps) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    constexpr int ROWB = (MODE == 0) ? 64 : 128;        // bytes per box row
    constexpr int BOX = 32 * ROWB;
    constexpr int COLS = ROWB / 4;                      // fp32 columns per box
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ql = warp & 3, ch = warp >> 2;
    unsigned char* ring = smem + warp * R * BOX;
    const int m_tiles = M / 128, n_tiles = N / 256, tiles = m_tiles * n_tiles;
    int bi = 0;
    for (int rep = 0; rep < reps; ++rep)
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int row_base = (t / n_tiles) * 128 + ql * 32, col_base = (t % n_tiles) * 256 + ch * 128;
            const float v = (float)(t + lane);
            if (MODE <= 1) {
                for (int c = 0; c < 128; c += COLS) {
                    if (lane == 0) wait_read<R - 1>();
                    __syncwarp();
                    unsigned char* bx = ring + bi * BOX;
                    bi = (bi + 1 == R) ? 0 : bi + 1;
                    const int sw = (MODE == 0) ? ((lane >> 1) & 3) : (lane & 7);
#pragma unroll
                    for (int j = 0; j < ROWB / 16; ++j)
                        *reinterpret_cast<float4*>(bx + lane * ROWB + ((j ^ sw) << 4)) = make_float4(v, v + 1, v + 2, v + 3);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tm, bx, col_base + c, row_base);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else if (MODE == 2) {
                float* yr = Y + (size_t)(row_base + lane) * ld + col_base;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) *reinterpret_cast<float4*>(yr + 4 * j) = make_float4(v, v + 1, v + 2, v + 3);
            } else {
                // 32 rows x 32 columns staged (XOR-swizzled 16-byte pieces), then 8 lanes write one row's 128 bytes
                unsigned char* bx = ring;
                for (int c = 0; c < 128; c += 32) {
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(bx + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v, v + 1, v + 2, v + 3);
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int r = 4 * k + (lane >> 3), j = lane & 7;
                        const float4 x = *reinterpret_cast<const float4*>(bx + r * 128 + ((j ^ (r & 7)) << 4));
                        *reinterpret_cast<float4*>(Y + (size_t)(row_base + r) * ld + col_base + c + 4 * j) = x;
                    }
                }
            }
        }
    if (MODE <= 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            printf("%s failed: %s\n", #x, cudaGetErrorString(e_));                   \
            exit(1);                                                                 \
        }                                                                            \
    } while (0)

template <int MODE, int R>
static void run(EncodeFn enc, float* Y, int M, int N, double clk_ghz, const char* name, int grid = 148) {
    CUtensorMap tm;
    const int cols = (MODE == 0) ? 16 : 32;
    cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {(cuuint32_t)cols, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     MODE == 0 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(1);
    }
    const int smem = 8 * R * 4096 + 1024;
    CK(cudaFuncSetAttribute(probe<MODE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int reps = 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) probe<MODE, R><<<grid, 256, smem>>>(tm, Y, N, M, N, reps);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe<MODE, R><<<grid, 256, smem>>>(tm, Y, N, M, N, reps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)M * N * 4 * reps;
    printf("%-46s R=%d grid=%3d  %8.1f us per pass  %7.1f GB/s  %5.1f B/clk/SM\n", name, R, grid, ms * 1e3 / reps,
           bytes / ms / 1e6, bytes / (ms * 1e-3) / grid / (clk_ghz * 1e9));
}

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)p;
    const int M = 51200, N = 512;
    float* Y;
    CK(cudaMalloc(&Y, (size_t)M * N * 4));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    printf("M=%d N=%d fp32 (%.0f MB per pass), SM clock attribute %.3f GHz\n", M, N, M * (double)N * 4 / 1e6, ghz);
    run<0, 2>(enc, Y, M, N, ghz, "TMA store, 32 x 64 B boxes");
    run<0, 4>(enc, Y, M, N, ghz, "TMA store, 32 x 64 B boxes");
    run<1, 2>(enc, Y, M, N, ghz, "TMA store, 32 x 128 B boxes");
    run<1, 4>(enc, Y, M, N, ghz, "TMA store, 32 x 128 B boxes");
    run<2, 1>(enc, Y, M, N, ghz, "st.global.v4, thread = row");
    run<3, 1>(enc, Y, M, N, ghz, "staged, st.global.v4 8 lanes per 128 B row");
    // is the ~19 B/clk/SM a per-SM or a chip-wide limit?  Same work on fewer SMs:
    for (int grid : {74, 37, 16, 4}) {
        run<0, 2>(enc, Y, M, N, ghz, "TMA store, 32 x 64 B boxes", grid);
        run<3, 1>(enc, Y, M, N, ghz, "staged, st.global.v4 8 lanes per 128 B row", grid);
    }
    return 0;
}
