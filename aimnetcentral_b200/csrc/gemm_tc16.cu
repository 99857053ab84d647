// tcgen05 backend 2 of the per-atom MLP GEMMs: Y[M,N] = epilogue(A[M,K] @ W[N,K]^T), fp32-faithful, on the
// kind::f16 tensor pipe (twice the TF32 rate).
//
// Precision scheme ("3xFP16, row-chunk scaled"; tools/split_emulation.py is the CPU proof): fp16 carries the same 11
// significant bits as tf32, its only weakness is the 5-bit exponent.  Every K=32 chunk of an A row is therefore scaled
// by its own power of two s so that the chunk maximum lands in [2^13, 2^14); hi = rn_fp16(s x), lo = rn_fp16(s x - hi).
// With the maximum at 2^13 the fp16 subnormal spacing (2^-24) is 2^-37 of it, so the lo part never needs a second
// scale and hi*hi + hi*lo + lo*hi goes into ONE accumulator exactly like the 3xTF32 scheme (same rms error vs fp64:
// 9.9e-8 vs 9.7e-8 on GELU-like data; equal on wide-dynamic-range and outlier cases).  The weights get one power-of-two
// scale per tensor (host side, at load).  Chunks are K=32 anyway: the tensor core adds into its fp32 accumulator with
// truncation, so TMEM only ever holds one chunk and the epilogue warps add chunks in registers with round-to-nearest
// (see gemm_tc.cu); un-scaling is folded into that add (acc = fma(chunk, 1/(s_a s_w), acc)) and costs nothing.
// Fixed chunking, no atomics: bitwise run-to-run reproducible.
//
// Structure (one persistent CTA per SM, 640 threads, warp-specialised):
//   warp 18      TMA producer   per stage (K=32): A 128x32 fp32 (dense landing zone), W_hi and W_lo bn x 32 fp16
//                               (64B-swizzled K-major boxes)
//   warps 8-15   splitter       two groups of four warps, alternating stages (one stage is a serial chain of shared
//                               loads, shuffles, a barrier, converts and stores; two in flight hide that latency):
//                               landing zone -> row-chunk max (8 lanes per row, 3 shuffles) -> A_hi | A_lo fp16 tiles
//                               written over the landing zone in the 64B-swizzled UMMA layout; 1/(s_a s_w) per row to
//                               a small ring in shared memory; fence.proxy.async + mbarrier arrive
//   warp 19      MMA issuer     one elected thread: 2 k-steps x 3 tcgen05.mma.kind::f16 (M128 x N<=256 x K16) per stage,
//                               tcgen05.commit frees the stage and publishes the chunk (TMEM double-buffered, 2 x 256 cols)
//   warp 16      TMEM allocator
//   warps 0-7    epilogue       per chunk: tcgen05.ld 32x32b.x32, acc = fma(chunk, inv_scale[row], acc) in 128 fp32
//                               registers per thread (packed FFMA2); per tile: bias / GELU (+ gelu') / *aux (packed
//                               two-at-a-time math) through 64B-swizzled
//                               32x16 shared-memory boxes and TMA bulk tensor stores (loads for aux)
// Four 48 KB stages; setmaxnreg moves registers from the control/splitter warps to the epilogue warps.
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"

namespace aimnet {

namespace tc16 {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 4;
constexpr int A_LAND = BM * BK * 4;        // 16 KB fp32 landing zone, becomes A_hi (8 KB) | A_lo (8 KB) in place
constexpr int A_HALF = BM * BK * 2;        // 8 KB
constexpr int B_BYTES = BN * BK * 2;       // 16 KB
constexpr int STAGE_BYTES = A_LAND + 2 * B_BYTES;   // 48 KB
static_assert((STAGES & (STAGES - 1)) == 0, "STAGES must be a power of two");
constexpr int SCALE_SLOTS = 8;             // ring of per-row inverse scales; at most 6 chunks are ever in flight
constexpr int EPI_BOX = 2048;              // 32 rows x 16 fp32 columns per epilogue warp
constexpr int OFF_BARS = STAGES * STAGE_BYTES;
constexpr int OFF_SCALE = OFF_BARS + 2048;
constexpr int OFF_EPI = OFF_SCALE + SCALE_SLOTS * BM * 4;
constexpr int SMEM_BYTES = OFF_EPI + 8 * EPI_BOX + 1024 /*align*/;
constexpr int NUM_THREADS = 640;
// epilogue = warps 0-7, splitter groups = warps 8-11 / 12-15; the single-thread roles whose latency gates the pipeline
// get the highest warp ids (the sub-partition arbiter favours them, B300_MICROARCH.md "hi-wid-first")
constexpr int kWarpAlloc = 16, kWarpTma = 18, kWarpMma = 19;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 64B-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO | SBO=512B |
// version 1 (sm_100) | layout SWIZZLE_64B (4).  A row is 32 halfs = 64 bytes; 8-row groups are 512 bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

struct Params {
    const float* bias;
    const float* w_inv_scale;   // device scalar: 1 / s_w of the (pre-scaled) weight tensor
    const float* aux;
    int M, N, K, mode;
    int bn;      // N-tile width (multiple of 32, <= 256): N is cut into equal tiles so that no CTA gets a sliver
    unsigned long long* trace;   // debug: per-stage SM-clock stamps of CTA 0 (8 events x kTraceLen), or nullptr
};
constexpr int kTraceLen = 2048;
__device__ __forceinline__ void stamp(const Params& p, int ev, int idx) {
    if (p.trace != nullptr && blockIdx.x == 0 && idx < kTraceLen) p.trace[ev * kTraceLen + idx] = clock64();
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                 const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmY,
                 const __grid_constant__ CUtensorMap tmAux, Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* full_tma = bars;                  // [STAGES]
    uint64_t* full_split = bars + STAGES;       // [STAGES]
    uint64_t* empty = bars + 2 * STAGES;        // [STAGES]
    uint64_t* tmem_full = bars + 3 * STAGES;    // [2]
    uint64_t* tmem_empty = bars + 3 * STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
    volatile int* chunk_last = reinterpret_cast<volatile int*>(bars + 3 * STAGES + 5);   // [2] last chunk of its tile?
    uint64_t* epi_bar = bars + 20;                                                  // [8] one per epilogue warp
    float* sbias = reinterpret_cast<float*>(smem + OFF_BARS + 256);                // [256]
    float* rowscale = reinterpret_cast<float*>(smem + OFF_SCALE);                  // [SCALE_SLOTS][BM]
    unsigned char* epi_buf = smem + OFF_EPI;                                       // 8 x 2 KB, 1 KB aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + p.bn - 1) / p.bn;
    const uint32_t tx_bytes = (uint32_t)(A_LAND + 2 * p.bn * BK * 2);
    const int tiles = m_tiles * n_tiles;
    const int nk = p.K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_tma[s], 1);
            mbar_init(&full_split[s], 128);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 8);
        }
        for (int w = 0; w < 8; ++w) mbar_init(&epi_bar[w], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpAlloc) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto stage_ptr = [&](int s) { return smem + s * STAGE_BYTES; };

    if (warp == kWarpTma) {
        // ------------------------------------------------ TMA producer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int cit = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * p.bn;
                for (int ks = 0; ks < nk; ++ks, ++cit) {
                    mbar_wait(&empty[s], ph ^ 1);
                    stamp(p, 0, cit);
                    unsigned char* sp = stage_ptr(s);
                    mbar_expect_tx(&full_tma[s], tx_bytes);
                    tma_load_2d(sp, &tmA, &full_tma[s], ks * BK, m0);
                    tma_load_2d(sp + A_LAND, &tmBh, &full_tma[s], ks * BK, n0);
                    tma_load_2d(sp + A_LAND + B_BYTES, &tmBl, &full_tma[s], ks * BK, n0);
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ------------------------------------------------ MMA issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int cit = 0;   // running chunk (= stage) counter -> TMEM buffer / phase
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int n0 = (t % n_tiles) * p.bn;
                int n_tile = min(p.bn, p.N - n0);
                // kind::f16: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
                uint32_t idesc = (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
                for (int ks = 0; ks < nk; ++ks, ++cit) {
                    int b = cit & 1;
                    uint32_t aph = (uint32_t)(cit >> 1) & 1;
                    mbar_wait(&tmem_empty[b], aph ^ 1);
                    stamp(p, 1, cit);
                    mbar_wait(&full_split[s], ph);
                    stamp(p, 2, cit);
                    tc_fence_after();
                    uint32_t d_tmem = tmem_base + (uint32_t)(b * BN);
                    uint32_t sa = smem_u32(stage_ptr(s));
                    uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_HALF);
                    uint64_t b_hi = make_desc(sa + A_LAND), b_lo = make_desc(sa + A_LAND + B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        uint64_t adv = (uint64_t)(kk * 32 >> 4);   // 16 halfs = 32 bytes along K inside the swizzle atom
                        tc_mma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, kk > 0 ? 1u : 0u);
                        tc_mma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                        tc_mma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                    }
                    tc_commit(&empty[s]);   // frees the stage once these MMAs have read it
                    chunk_last[b] = (ks == nk - 1) ? 1 : 0;
                    __threadfence_block();
                    tc_commit(&tmem_full[b]);
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    } else if (warp >= 8) {
        // ------------------------------------------------ splitter: fp32 landing zone -> scaled fp16 hi | lo
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int grp = (warp - 8) >> 2;                 // group g converts the stages with (running index & 1) == g
        const int tsp = (threadIdx.x - 256) & 127;
        const float w_inv = *p.w_inv_scale;
        const int row0 = tsp >> 3, c8 = tsp & 7;   // this thread's rows are row0 + 16 r, its columns 4 c8 .. 4 c8 + 3
        const int my_tiles = tiles > (int)blockIdx.x ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        const int total = my_tiles * nk;                 // stages this CTA runs through, in order
        {
            for (int cit = grp; cit < total; cit += 2) {
                const int s = cit & (STAGES - 1);
                const uint32_t ph = (uint32_t)(cit / STAGES) & 1;
                mbar_wait(&full_tma[s], ph);
                if (tsp == 0) stamp(p, 3, cit);
                unsigned char* sp = stage_ptr(s);
                const float4* land = reinterpret_cast<const float4*>(sp);
                float4 v[8];
                float mx[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    v[r] = land[tsp + 128 * r];
                    mx[r] = fmaxf(fmaxf(fabsf(v[r].x), fabsf(v[r].y)), fmaxf(fabsf(v[r].z), fabsf(v[r].w)));
                }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
                }
                // every thread of this group has read the landing zone (named barrier 2 / 3, one per group)
                if (grp == 0)
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                else
                    asm volatile("bar.sync 3, 128;" ::: "memory");
                float* scale_slot = rowscale + (cit & (SCALE_SLOTS - 1)) * BM;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int row = row0 + 16 * r;
                    // s = 2^(13 - floor(log2 max)) from the exponent field, clamped so that s and 1/s stay normal
                    int e = (int)(__float_as_uint(mx[r]) >> 23);
                    e = min(max(e, 14), 254);
                    const float sc = __uint_as_float((uint32_t)(267 - e) << 23);
                    const float inv = __uint_as_float((uint32_t)(e - 13) << 23);
                    const __half2 h01 = __floats2half2_rn(v[r].x * sc, v[r].y * sc);
                    const __half2 h23 = __floats2half2_rn(v[r].z * sc, v[r].w * sc);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn(fmaf(v[r].x, sc, -f01.x), fmaf(v[r].y, sc, -f01.y));
                    const __half2 l23 = __floats2half2_rn(fmaf(v[r].z, sc, -f23.x), fmaf(v[r].w, sc, -f23.y));
                    // 64B swizzle: 16-byte chunk index (c8 >> 1) XOR bits [1,2] of the row
                    const int off = row * 64 + ((((c8 >> 1) ^ (row >> 1)) & 3) << 4) + ((c8 & 1) << 3);
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<const uint32_t*>(&h01);
                    hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                    lv.x = *reinterpret_cast<const uint32_t*>(&l01);
                    lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                    *reinterpret_cast<uint2*>(sp + off) = hv;
                    *reinterpret_cast<uint2*>(sp + A_HALF + off) = lv;
                    if (c8 == 0) scale_slot[row] = inv * w_inv;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&full_split[s]);
                if (tsp == 0) stamp(p, 4, cit);
            }
        }
    } else {
        // ------------------------------------------------ epilogue (warps 0-7)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        const int ql = warp & 3;            // TMEM lane quarter this warp may access
        const int ch = warp >> 2;           // column half of the 256-wide accumulator
        int cit = 0;
        uint32_t epi_phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * p.bn;
            int n_tile = min(p.bn, p.N - n0);
            float2 acc[64];   // one output row x 128 columns, as register pairs for the packed FFMA2 / FMUL2 / FADD2
#pragma unroll
            for (int k = 0; k < 64; ++k) acc[k] = make_float2(0.f, 0.f);
            if (MODE == 1 || MODE == 2) {
                // bias of this tile's columns -> shared memory (read back as warp-wide broadcasts in the epilogue)
                asm volatile("bar.sync 1, 256;");   // previous tile's readers are done
                int cb = threadIdx.x;
                sbias[cb] = (cb < n_tile) ? p.bias[n0 + cb] : 0.f;
                asm volatile("bar.sync 1, 256;");
            }
            for (int last = 0; !last; ++cit) {
                int b = cit & 1;
                uint32_t aph = (uint32_t)(cit >> 1) & 1;
                mbar_wait(&tmem_full[b], aph);
                if (threadIdx.x == 0) stamp(p, 5, cit);
                tc_fence_after();
                last = chunk_last[b];
                const float inv = rowscale[(cit & (SCALE_SLOTS - 1)) * BM + ql * 32 + lane];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    int col0 = ch * 128 + c * 32;
                    if (col0 < n_tile) {
                        uint32_t r[32];
                        uint32_t taddr = tmem_base + ((uint32_t)(ql * 32) << 16) + (uint32_t)(b * BN + col0);
                        tc_ld32(taddr, r);
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            acc[c * 16 + k] = ffma2s(inv, make_float2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])),
                                                     acc[c * 16 + k]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[b]);
                if (threadIdx.x == 0) stamp(p, 6, cit);
            }
            // ---- tile epilogue.  Each thread holds one output row (lane) x 128 columns.  Rows are 1-3 KB apart in
            // global memory, so the values go through a 64B-swizzled 32x16 shared-memory box per warp and leave (or,
            // for the aux operand of mode 3, arrive) as TMA bulk tensor copies: full 64-byte row segments, no LSU work.
            {
                unsigned char* sw = epi_buf + warp * EPI_BOX;
                uint64_t* ebar = &epi_bar[warp];
                const int row_base = m0 + ql * 32;
                const int rsw = (lane >> 1) & 3;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int col0 = ch * 128 + c * 16;
                    if (col0 < n_tile) {
                        const int col = n0 + col0;
                        if (MODE == 3) {
                            // aux block -> smem (the previous bulk store must have finished reading the buffer)
                            if (lane == 0) {
                                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                                mbar_expect_tx(ebar, EPI_BOX);
                                tma_load_2d(sw, &tmAux, ebar, col, row_base);
                            }
                            mbar_wait(ebar, epi_phase);
                            epi_phase ^= 1;
#pragma unroll
                            for (int v4 = 0; v4 < 4; ++v4) {
                                float4 g = *reinterpret_cast<const float4*>(sw + lane * 64 + ((v4 ^ rsw) << 4));
                                acc[c * 8 + 2 * v4 + 0] = fmul2(acc[c * 8 + 2 * v4 + 0], make_float2(g.x, g.y));
                                acc[c * 8 + 2 * v4 + 1] = fmul2(acc[c * 8 + 2 * v4 + 1], make_float2(g.z, g.w));
                            }
                            __syncwarp();
                        } else {
                            if (MODE == 1 || MODE == 2) {
#pragma unroll
                                for (int v4 = 0; v4 < 4; ++v4) {
                                    float4 bz = *reinterpret_cast<const float4*>(sbias + col0 + 4 * v4);
                                    acc[c * 8 + 2 * v4 + 0] = fadd2(acc[c * 8 + 2 * v4 + 0], make_float2(bz.x, bz.y));
                                    acc[c * 8 + 2 * v4 + 1] = fadd2(acc[c * 8 + 2 * v4 + 1], make_float2(bz.z, bz.w));
                                }
                            }
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            __syncwarp();
                        }
                        // y -> smem box -> global
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            float2 z0 = acc[c * 8 + 2 * v4 + 0], z1 = acc[c * 8 + 2 * v4 + 1];
                            if (MODE == 2) {
                                float2 g0, g1;
                                gelu_pair2(z0, z0, g0);
                                gelu_pair2(z1, z1, g1);
                                acc[c * 8 + 2 * v4 + 0] = g0;   // gelu' takes over the accumulator registers
                                acc[c * 8 + 2 * v4 + 1] = g1;
                            }
                            *reinterpret_cast<float4*>(sw + lane * 64 + ((v4 ^ rsw) << 4)) = make_float4(z0.x, z0.y, z1.x, z1.y);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmY, sw, col, row_base);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        if (MODE == 2 && p.aux != nullptr) {
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            __syncwarp();
#pragma unroll
                            for (int v4 = 0; v4 < 4; ++v4) {
                                float2 g0 = acc[c * 8 + 2 * v4 + 0], g1 = acc[c * 8 + 2 * v4 + 1];
                                *reinterpret_cast<float4*>(sw + lane * 64 + ((v4 ^ rsw) << 4)) = make_float4(g0.x, g0.y, g1.x, g1.y);
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmAux, sw, col, row_base);
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                        }
                    }
                }
            }
            if (threadIdx.x == 0) stamp(p, 7, cit - 1);   // end of this tile's epilogue (indexed by its last chunk)
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores retired before exit
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpAlloc) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- weight preparation on the device (operator seam / tests; the engine splits on the host at load) -------------
__global__ void absmax_kernel(const float* __restrict__ w, size_t n, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

__global__ void split_fp16_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, size_t n,
                                  const unsigned int* __restrict__ maxbits, float* __restrict__ inv_scale) {
    int e = (int)(*maxbits >> 23);
    e = min(max(e, 14), 254);
    const float sc = __uint_as_float((uint32_t)(267 - e) << 23);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *inv_scale = __uint_as_float((uint32_t)(e - 13) << 23);
    if (i >= n) return;
    float x = w[i] * sc;
    __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn(x - __half2float(h));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    });
    return fn;
}

static int make_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elem_bytes, int rows, int cols, int ld,
                    int box_rows, int box_cols, CUtensorMapSwizzle swz) {
    EncodeFn enc = get_encode();
    if (!enc) {
        set_error("gemm_tc16: cuTensorMapEncodeTiled not available");
        return AIMNET_ECUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc16: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return AIMNET_ECUDA;
    }
    return AIMNET_OK;
}

}  // namespace tc16

static unsigned long long* g_trace = nullptr;
void gemm_tc16_set_trace(unsigned long long* buf) { g_trace = buf; }

// hi / lo: (N, ldw) fp16 split of s_w * W, inv_scale: device scalar 1 / s_w
int split_fp16_device(const float* w, void* hi, void* lo, float* inv_scale, unsigned int* scratch, size_t n,
                      cudaStream_t st) {
    if (n == 0) return AIMNET_OK;
    AIM_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(unsigned int), st));
    tc16::absmax_kernel<<<148, 256, 0, st>>>(w, n, scratch);
    AIM_LAUNCH_CHECK();
    tc16::split_fp16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, (__half*)hi, (__half*)lo, n, scratch, inv_scale);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

int gemm_nt_tc16(const float* A, int lda, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw,
                 const float* bias, float* Y, int ldy, float* aux, int ldaux, int M, int N, int K, int mode,
                 cudaStream_t st) {
    using namespace tc16;
    AIM_REQUIRE(K % BK == 0 && N % 32 == 0, "gemm_tc16: K and N must be multiples of 32");
    AIM_REQUIRE(((uintptr_t)Y & 15) == 0 && ldy % 4 == 0 && (aux == nullptr || (((uintptr_t)aux & 15) == 0 && ldaux % 4 == 0)),
                "gemm_tc16: outputs must be 16-byte aligned");
    AIM_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)Whi & 15) == 0 && ((uintptr_t)Wlo & 15) == 0 && lda % 4 == 0 && ldw % 8 == 0,
                "gemm_tc16: operands must be 16-byte aligned");
    AIM_REQUIRE(w_inv_scale != nullptr, "gemm_tc16: weight scale missing");
    static bool configured = false;
    static int num_sms = 148;
    if (!configured) {
        AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        int dev = 0;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    CUtensorMap tmA, tmBh, tmBl, tmY, tmAux;
    int rc;
    if ((rc = make_map(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, K, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    int n_tiles = (N + BN - 1) / BN;
    int bn = ((N + n_tiles - 1) / n_tiles + 31) / 32 * 32;
    if ((rc = make_map(&tmBh, Whi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, N, K, ldw, bn, BK, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map(&tmBl, Wlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, N, K, ldw, bn, BK, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map(&tmY, Y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldy, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map(&tmAux, aux ? aux : Y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, aux ? ldaux : ldy, 32, 16,
                       CU_TENSOR_MAP_SWIZZLE_64B)))
        return rc;
    Params p{bias, w_inv_scale, aux, M, N, K, mode, bn, g_trace};
    int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    int grid = tiles < num_sms ? tiles : num_sms;
    switch (mode) {
        case 0: gemm_tc16_kernel<0><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmBh, tmBl, tmY, tmAux, p); break;
        case 1: gemm_tc16_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmBh, tmBl, tmY, tmAux, p); break;
        case 2: gemm_tc16_kernel<2><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmBh, tmBl, tmY, tmAux, p); break;
        default: gemm_tc16_kernel<3><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmBh, tmBl, tmY, tmAux, p); break;
    }
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

}  // namespace aimnet
