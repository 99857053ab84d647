// Micro-probe: tcgen05.mma.kind::f16 (M=128, K=16, A and B from shared memory, 64B-swizzled K-major) issue-to-completion
// rate as a function of N and of how many independent TMEM accumulators the MMAs rotate over.  One CTA per SM; one thread
// issues `reps` MMAs (operands: whatever is in shared memory), commits, waits; clock64 around it.
//   same accumulator every time  = the dependent chain of a K loop
//   2 / 4 accumulators in turn   = independent chains interleaved
// Also varies how many distinct operand tiles the MMAs read (1 = same tile every time; 6 = a 3-MMA x 2-k-step stage).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_chain_probe tools/mma_chain_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}

template <int N, int n_acc, int n_tiles>
__global__ void __launch_bounds__(128, 1) probe(int reps, long long* out) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    // operand tiles: A 128 x 32 halfs (8 KB, two K=16 steps), B 256 x 32 halfs (16 KB); n_tiles of each
    for (int i = threadIdx.x; i < 6 * 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(smem);
        uint32_t phase = 0;
        for (int pass = 0; pass < 3; ++pass) {
            const long long t0 = clock64();
            const uint64_t a0 = make_desc(sa), b0 = make_desc(sa + 8192);
            for (int i = 0; i < reps; i += 12) {
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    const uint64_t off = (uint64_t)(((j % n_tiles) * 24576 + (j & 1) * 32) >> 4);   // descriptor start is addr >> 4
                    mma(tmem + (uint32_t)((j % n_acc) * N), a0 + off, b0 + off, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "W:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D;\n\t"
                "bra W;\n\t"
                "D:\n\t"
                "}" ::"r"(smem_u32(&bar)),
                "r"(phase)
                : "memory");
            phase ^= 1;
            const long long t1 = clock64();
            if (blockIdx.x == 0 && pass == 2) out[0] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, int n_acc, int n_tiles>
static void one(int grid, int smem, int reps, long long* out) {
    cudaFuncSetAttribute(probe<N, n_acc, n_tiles>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<N, n_acc, n_tiles><<<grid, 128, smem>>>(reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        exit(1);
    }
    long long clk = 0;
    cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
    printf("grid %3d  N=%3d  accumulators=%d  operand tiles=%d : %7.1f clk per MMA  (full rate %d)\n", grid, N, n_acc, n_tiles,
           (double)clk / reps, N / 2);
}

int main() {
    long long* out;
    cudaMalloc(&out, 8);
    const int smem = 6 * 24576 + 1024;
    const int reps = 1200;
    printf("tcgen05.mma kind::f16 M=128 K=16, %d MMAs per measurement, all 148 SMs busy\n", reps);
    for (int grid : {1, 148}) {
        one<64, 1, 1>(grid, smem, reps, out);
        one<64, 2, 6>(grid, smem, reps, out);
        one<64, 4, 6>(grid, smem, reps, out);
        one<128, 1, 1>(grid, smem, reps, out);
        one<128, 1, 6>(grid, smem, reps, out);
        one<128, 2, 1>(grid, smem, reps, out);
        one<128, 2, 6>(grid, smem, reps, out);
        one<128, 4, 6>(grid, smem, reps, out);
        one<192, 1, 6>(grid, smem, reps, out);
        one<192, 2, 6>(grid, smem, reps, out);
        one<256, 1, 1>(grid, smem, reps, out);
        one<256, 1, 6>(grid, smem, reps, out);
        one<256, 2, 6>(grid, smem, reps, out);
    }
    return 0;
}
