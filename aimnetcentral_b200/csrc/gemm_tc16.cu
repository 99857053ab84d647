// tcgen05 backend 2 of the per-atom MLP GEMMs: Y[M,N] = epilogue(A[M,K] @ W[N,K]^T), fp32-faithful, on the
// kind::f16 tensor pipe (twice the TF32 rate).
//
// Precision scheme ("3xFP16, row-chunk scaled"; tools/split_emulation.py is the CPU proof): fp16 carries the same 11
// significant bits as tf32, its only weakness is the 5-bit exponent.  Every K=64 chunk of an A row is therefore scaled
// by its own power of two s so that the chunk maximum lands in [2^13, 2^14); hi = rn_fp16(s x), lo = rn_fp16(s x - hi).
// With the maximum at 2^13 the fp16 subnormal spacing (2^-24) is 2^-37 of it, so the lo part never needs a second
// scale and hi*hi + hi*lo + lo*hi goes into ONE accumulator exactly like the 3xTF32 scheme (same rms error vs fp64:
// 9.9e-8 vs 9.7e-8 on GELU-like data; equal on wide-dynamic-range and outlier cases).  The weights get one power-of-two
// scale per tensor (host side, at load).  Chunks exist anyway: the tensor core adds into its fp32 accumulator with
// truncation, so TMEM only ever holds one chunk (K=64: 4 k-steps x 3 MMAs, the same number of truncating adds as the
// K=32 chunks of the 3xTF32 kernel) and the epilogue warps add chunks in registers with round-to-nearest (see
// gemm_tc.cu); un-scaling is folded into that add (acc = fma(chunk, 1/(s_a s_w), acc)) and costs nothing.  K=64 rather
// than K=32 because a TMEM buffer cycles through MMA -> commit -> drain -> release, about 1900 clk with two buffers
// (tools/gemm_trace.py): at K=32 that is 950 clk per 768 clk of MMA, at K=64 the tensor pipe is the longer leg.
// Fixed chunking, no atomics: bitwise run-to-run reproducible.
//
// Activations travel PRE-SPLIT between GEMMs ("SplitView": fp16 hi, fp16 lo, one fp32 inverse scale per row-chunk; the
// same bytes per element as fp32).  A first version split the fp32 A tile inside this kernel with dedicated warps; the
// pipeline trace (tools/gemm_trace.py) showed that those warps, which redo the conversion for every N tile, ate half of
// the SM's issue slots and set the stage period (2300 clk against 768 clk of MMA).  Now the producer of an activation
// splits it once: the epilogue of the GEMM that computes it (each thread already holds a row x 32-column group, i.e.
// exactly one row-chunk of the consumer, in registers), or presplit_kernel for tensors that come from other kernels.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 10      TMA producer   per stage (K=32): A_hi, A_lo 128 x 32 fp16 and W_hi, W_lo bn x 32 fp16, 64B-swizzled
//                               K-major boxes straight into the UMMA layout
//   warp 11      MMA issuer     one elected thread: 2 k-steps x 3 tcgen05.mma.kind::f16 (M128 x N<=256 x K16) per stage,
//                               tcgen05.commit frees the stage and publishes the chunk (TMEM double-buffered, 2 x 256 cols)
//   warp 8       TMEM allocator
//   warps 0-7    epilogue       per chunk: tcgen05.ld 32x32b.x32, acc = fma(chunk, inv_scale[row, chunk], acc) in 128
//                               fp32 registers per thread (packed FFMA2); per tile: bias / GELU (+ gelu') / *aux with
//                               packed two-at-a-time math, then either fp32 or pre-split output through 64B-swizzled
//                               2 KB shared-memory boxes and TMA bulk tensor stores (loads for aux)
// Four 48 KB stages; setmaxnreg moves registers from the control warps to the epilogue warps.
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

namespace tc16 {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 4;
constexpr int CHUNK_STAGES = 2;            // stages per TMEM chunk; one activation scale covers CHUNK_STAGES * BK = 64 columns
constexpr int A_HALF = BM * BK * 2;        // 8 KB: one fp16 A tile (hi or lo)
constexpr int B_BYTES = BN * BK * 2;       // 16 KB
constexpr int STAGE_BYTES = 2 * A_HALF + 2 * B_BYTES;   // 48 KB
constexpr int EPI_BOX = 2048;              // 32 rows x 64 bytes; two per epilogue warp
constexpr int OFF_BARS = STAGES * STAGE_BYTES;
constexpr int OFF_EPI = OFF_BARS + 2048;
constexpr int SMEM_BYTES = OFF_EPI + 16 * EPI_BOX + 1024 /*align*/;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int NUM_THREADS = 384;
// epilogue = warps 0-7; the single-thread roles whose latency gates the pipeline get the highest warp ids (the
// sub-partition arbiter favours them, B300_MICROARCH.md "hi-wid-first")
constexpr int kWarpAlloc = 8, kWarpTma = 10, kWarpMma = 11;

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
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
    const float* a_inv;         // (M, lda_inv): 1 / s_a per row-chunk of A
    const float* aux;
    float* out_inv;             // split output: (M, ld_out_inv) inverse scales per row-chunk of Y
    int lda_inv, ld_out_inv, ldaux;
    int M, N, K, mode;
    int bn;      // N-tile width (multiple of 32, <= 256): N is cut into equal tiles so that no CTA gets a sliver
    unsigned long long* trace;   // debug: per-stage SM-clock stamps of CTA 0 (8 events x kTraceLen), or nullptr
};
constexpr int kTraceLen = 2048;
__device__ __forceinline__ void stamp(const Params& p, int ev, int idx) {
    if (p.trace != nullptr && blockIdx.x == 0 && idx < kTraceLen) p.trace[ev * kTraceLen + idx] = clock64();
}

// power-of-two scale that puts m (>= 0) into [2^13, 2^14), and its inverse, straight from the exponent field; clamped
// so that both stay normal numbers (m == 0 or denormal: scale 2^126, every product is zero anyway)
__device__ __forceinline__ void chunk_scale(float m, float& sc, float& inv) {
    int e = (int)(__float_as_uint(m) >> 23);
    e = min(max(e, 14), 254);
    sc = __uint_as_float((uint32_t)(267 - e) << 23);
    inv = __uint_as_float((uint32_t)(e - 13) << 23);
}
// (x, y) * sc -> fp16 hi pair and fp16 lo pair (lo = rn(sc v - hi), exact difference)
__device__ __forceinline__ void split_pair(float2 v, float sc, uint32_t& hi, uint32_t& lo) {
    const float2 sv = fmul2(v, make_float2(sc, sc));
    const __half2 h = __floats2half2_rn(sv.x, sv.y);
    const float2 hf = __half22float2(h);
    const float2 df = ffma2(hf, make_float2(-1.0f, -1.0f), sv);   // exact: hf is sv rounded to 11 bits
    const __half2 l = __floats2half2_rn(df.x, df.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int MODE, bool SPLIT_OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc16_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                 const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                 const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmY2,
                 const __grid_constant__ CUtensorMap tmAux, Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* full = bars;                      // [STAGES]
    uint64_t* empty = bars + STAGES;            // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;    // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* sbias = reinterpret_cast<float*>(smem + OFF_BARS + 256);                // [256]
    unsigned char* epi_buf = smem + OFF_EPI;                                       // 8 x 2 x 2 KB, 1 KB aligned

    // Mode 3 runs three operand stages and uses the fourth stage's 48 KB as a ring of 2 KB boxes (three per epilogue
    // warp) into which TMA prefetches the aux operand 16 columns at a time; the other modes keep four stages.
    constexpr int kStages = (MODE == 3) ? STAGES - 1 : STAGES;
    unsigned char* aux_buf = smem + (STAGES - 1) * STAGE_BYTES;                      // mode 3 only: 24 x 2 KB
    uint64_t* aux_bar = reinterpret_cast<uint64_t*>(smem + OFF_BARS + 1280);         // [8 warps][3]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + p.bn - 1) / p.bn;
    const uint32_t tx_bytes = (uint32_t)(2 * A_HALF + 2 * p.bn * BK * 2);
    const int tiles = m_tiles * n_tiles;
    const int nk = p.K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 8);
        }
        for (int w = 0; w < 24; ++w) mbar_init(&aux_bar[w], 1);
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
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
                    mbar_expect_tx(&full[s], tx_bytes);
                    tma_load_2d(sp, &tmAh, &full[s], ks * BK, m0);
                    tma_load_2d(sp + A_HALF, &tmAl, &full[s], ks * BK, m0);
                    tma_load_2d(sp + 2 * A_HALF, &tmBh, &full[s], ks * BK, n0);
                    tma_load_2d(sp + 2 * A_HALF + B_BYTES, &tmBl, &full[s], ks * BK, n0);
                    if (++s == kStages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ------------------------------------------------ MMA issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int cit = 0;   // running chunk (= stage) counter -> TMEM buffer / phase
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int n0 = (t % n_tiles) * p.bn;
                int n_tile = min(p.bn, p.N - n0);
                // kind::f16: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
                uint32_t idesc = (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
                for (int ks = 0; ks < nk; ++cit) {
                    int b = cit & 1;
                    uint32_t aph = (uint32_t)(cit >> 1) & 1;
                    mbar_wait(&tmem_empty[b], aph ^ 1);
                    stamp(p, 1, cit);
                    uint32_t d_tmem = tmem_base + (uint32_t)(b * BN);
                    for (int j = 0; j < CHUNK_STAGES && ks < nk; ++j, ++ks) {
                        mbar_wait(&full[s], ph);
                        if (j == 0) stamp(p, 2, cit);
                        tc_fence_after();
                        uint32_t sa = smem_u32(stage_ptr(s));
                        uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_HALF);
                        uint64_t b_hi = make_desc(sa + 2 * A_HALF), b_lo = make_desc(sa + 2 * A_HALF + B_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            uint64_t adv = (uint64_t)(kk * 32 >> 4);   // 16 halfs = 32 bytes along K inside the swizzle atom
                            tc_mma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                            tc_mma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            tc_mma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                        }
                        tc_commit(&empty[s]);   // frees the stage once these MMAs have read it
                        if (++s == kStages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    tc_commit(&tmem_full[b]);
                }
            }
        }
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    } else {
        // ------------------------------------------------ epilogue (warps 0-7)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int ql = warp & 3;            // TMEM lane quarter this warp may access
        const int ch = warp >> 2;           // column half of the 256-wide accumulator
        const float w_inv = *p.w_inv_scale;
        int cit = 0;
        unsigned char* box0 = epi_buf + warp * 2 * EPI_BOX;
        unsigned char* box1 = box0 + EPI_BOX;
        unsigned char* abox = aux_buf + warp * 3 * EPI_BOX;   // mode 3: this warp's aux ring
        uint64_t* abar = aux_bar + warp * 3;
        int ag = 0;                                           // running aux step -> ring slot ag % 3, phase (ag / 3) & 1
        const int rsw = (lane >> 1) & 3;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * p.bn;
            const int n_tile = min(p.bn, p.N - n0);
            const int row_base = m0 + ql * 32;
            const int row = row_base + lane;
            const float* inv_row = p.a_inv + (size_t)min(row, p.M - 1) * p.lda_inv;
            float2 acc[64];   // one output row x 128 columns, as register pairs for the packed FFMA2 / FMUL2 / FADD2
#pragma unroll
            for (int k = 0; k < 64; ++k) acc[k] = make_float2(0.f, 0.f);
            // mode 3: number of 16-column aux steps of this warp in this tile; the first three are requested now and
            // arrive while the tile's MMAs run
            const int aux_steps = (MODE == 3) ? max(0, min(n_tile - ch * 128, 128)) / 16 : 0;
            auto aux_request = [&](int step, int slot) {   // lane 0 only
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&abar[slot], EPI_BOX);
                tma_load_2d(abox + slot * EPI_BOX, &tmAux, &abar[slot], n0 + ch * 128 + 16 * step, row_base);
            };
            if (MODE == 3 && lane == 0) {
                for (int k = 0; k < 3 && k < aux_steps; ++k) aux_request(k, (ag + k) % 3);
            }
            if (MODE == 1 || MODE == 2) {
                // bias of this tile's columns -> shared memory (read back as warp-wide broadcasts in the epilogue)
                asm volatile("bar.sync 1, 256;");   // previous tile's readers are done
                int cb = threadIdx.x;
                sbias[cb] = (cb < n_tile) ? p.bias[n0 + cb] : 0.f;
                asm volatile("bar.sync 1, 256;");
            }
            const int nchunk = (nk + CHUNK_STAGES - 1) / CHUNK_STAGES;
            float inv_next = __ldg(inv_row) * w_inv;
            int kc = 0;
            for (int last = 0; !last; ++cit, ++kc) {
                const float inv = inv_next;
                if (kc + 1 < nchunk) inv_next = __ldg(inv_row + kc + 1) * w_inv;   // in flight while we wait for the chunk
                int b = cit & 1;
                uint32_t aph = (uint32_t)(cit >> 1) & 1;
                mbar_wait(&tmem_full[b], aph);
                if (threadIdx.x == 0) stamp(p, 5, cit);
                tc_fence_after();
                last = (kc + 1 == nchunk) ? 1 : 0;   // the tile's last chunk (the issuer runs the same count)
                const uint32_t taddr = tmem_base + ((uint32_t)(ql * 32) << 16) + (uint32_t)(b * BN + ch * 128);
                // two 32-column loads in flight at a time
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int col0 = ch * 128 + c2 * 64;
                    if (col0 < n_tile) {
                        uint32_t r0[32], r1[32];
                        const bool two = col0 + 32 < n_tile;
                        tc_ld32(taddr + c2 * 64, r0);
                        if (two) tc_ld32(taddr + c2 * 64 + 32, r1);
                        tc_ld_wait();
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            acc[c2 * 32 + k] = ffma2s(inv, make_float2(__uint_as_float(r0[2 * k]), __uint_as_float(r0[2 * k + 1])),
                                                      acc[c2 * 32 + k]);
                        if (two) {
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                acc[c2 * 32 + 16 + k] = ffma2s(inv, make_float2(__uint_as_float(r1[2 * k]), __uint_as_float(r1[2 * k + 1])),
                                                               acc[c2 * 32 + 16 + k]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[b]);
                if (threadIdx.x == 0) stamp(p, 6, cit);
            }
            // ---- tile epilogue.  Each thread holds one output row (lane) x 128 columns; a pair of 32-column groups is
            // one row-chunk (K=64) of the GEMM that consumes this output.  Rows are 1-3 KB apart in global memory, so
            // outputs go through 64B-swizzled 32-row shared-memory boxes and leave as TMA bulk tensor stores: full
            // 64-byte row segments, no LSU work.  The two boxes of a warp form a ring: a box is rewritten only after the
            // store issued from it two steps earlier has read it (wait_group.read 1), so the math of one 16-column step
            // overlaps the store of the previous one.  The aux operand of mode 3 is prefetched into registers one
            // 32-column group ahead with plain 16-byte loads (a thread reads 128 contiguous bytes of its own row).
            int box_i = 0;
            auto box_acquire = [&]() -> unsigned char* {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                unsigned char* bx = box_i ? box1 : box0;
                box_i ^= 1;
                return bx;
            };
            auto box_store = [&](const CUtensorMap* map, const unsigned char* bx, int c0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(map, bx, c0, row_base);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            };
            int aux_step = 0;   // mode 3: 16-column steps consumed in this tile
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                if (ch * 128 + c2 * 64 < n_tile) {
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
                        const int col0 = ch * 128 + c2 * 64 + hc * 32;
                        if (col0 < n_tile) {
                            const int col = n0 + col0;
                            float2* v = &acc[c2 * 32 + hc * 16];
                            if (MODE == 3) {
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {
                                    const int slot = ag % 3;
                                    mbar_wait(&abar[slot], (uint32_t)(ag / 3) & 1);
                                    const unsigned char* bx = abox + slot * EPI_BOX;
#pragma unroll
                                    for (int v4 = 0; v4 < 4; ++v4) {
                                        const float4 g = *reinterpret_cast<const float4*>(bx + lane * 64 + ((v4 ^ rsw) << 4));
                                        v[hb * 8 + 2 * v4 + 0] = fmul2(v[hb * 8 + 2 * v4 + 0], make_float2(g.x, g.y));
                                        v[hb * 8 + 2 * v4 + 1] = fmul2(v[hb * 8 + 2 * v4 + 1], make_float2(g.z, g.w));
                                    }
                                    __syncwarp();   // every lane has read the box: refill it three steps ahead
                                    if (lane == 0 && aux_step + 3 < aux_steps) aux_request(aux_step + 3, slot);
                                    ++ag;
                                    ++aux_step;
                                }
                            } else if (MODE == 1 || MODE == 2) {
#pragma unroll
                                for (int v4 = 0; v4 < 8; ++v4) {
                                    float4 bz = *reinterpret_cast<const float4*>(sbias + col0 + 4 * v4);
                                    v[2 * v4 + 0] = fadd2(v[2 * v4 + 0], make_float2(bz.x, bz.y));
                                    v[2 * v4 + 1] = fadd2(v[2 * v4 + 1], make_float2(bz.z, bz.w));
                                }
                            }
                            if (MODE == 2) {
                                // y = gelu(z) stays in the accumulator registers, gelu'(z) leaves as fp32 through the boxes
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {
                                    float2 g[8];
#pragma unroll
                                    for (int k = 0; k < 8; ++k) gelu_pair2(v[hb * 8 + k], v[hb * 8 + k], g[k]);
                                    if (p.aux != nullptr) {
                                        unsigned char* bx = box_acquire();
#pragma unroll
                                        for (int v4 = 0; v4 < 4; ++v4)
                                            *reinterpret_cast<float4*>(bx + lane * 64 + ((v4 ^ rsw) << 4)) =
                                                make_float4(g[2 * v4].x, g[2 * v4].y, g[2 * v4 + 1].x, g[2 * v4 + 1].y);
                                        box_store(&tmAux, bx, col + 16 * hb);
                                    }
                                }
                            }
                            if (!SPLIT_OUT) {
#pragma unroll
                                for (int hb = 0; hb < 2; ++hb) {
                                    unsigned char* bx = box_acquire();
#pragma unroll
                                    for (int v4 = 0; v4 < 4; ++v4) {
                                        const float2 z0 = v[hb * 8 + 2 * v4 + 0], z1 = v[hb * 8 + 2 * v4 + 1];
                                        *reinterpret_cast<float4*>(bx + lane * 64 + ((v4 ^ rsw) << 4)) = make_float4(z0.x, z0.y, z1.x, z1.y);
                                    }
                                    box_store(&tmY, bx, col + 16 * hb);
                                }
                            }
                        }
                    }
                    if (SPLIT_OUT) {
                        // this thread's 64 values are one row-chunk of the consumer: scale, split, store hi | lo | 1/s
                        // (columns past n_tile were never touched and are zero)
                        float2* v = &acc[c2 * 32];
                        float m = 0.f;
#pragma unroll
                        for (int k = 0; k < 32; ++k) m = fmaxf(m, fmaxf(fabsf(v[k].x), fabsf(v[k].y)));
                        float sc, inv;
                        chunk_scale(m, sc, inv);
                        const int colp = n0 + ch * 128 + c2 * 64;
                        if (row < p.M) p.out_inv[(size_t)row * p.ld_out_inv + (colp >> 6)] = inv;
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            if (ch * 128 + c2 * 64 + hc * 32 < n_tile) {
                                uint32_t hi[16], lo[16];
#pragma unroll
                                for (int k = 0; k < 16; ++k) split_pair(v[hc * 16 + k], sc, hi[k], lo[k]);
                                unsigned char* bh = box_acquire();
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    *reinterpret_cast<uint4*>(bh + lane * 64 + ((j ^ rsw) << 4)) =
                                        make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                                box_store(&tmY, bh, colp + hc * 32);
                                unsigned char* bl = box_acquire();
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    *reinterpret_cast<uint4*>(bl + lane * 64 + ((j ^ rsw) << 4)) =
                                        make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                                box_store(&tmY2, bl, colp + hc * 32);
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

// ---- fp32 -> pre-split form for activations that do not come out of a GEMM epilogue (conv rows, head / NSE
// gradients): 16 lanes per row-chunk (one float4 each), 4 shuffles for the chunk maximum; HBM-bound ---------------
__global__ void __launch_bounds__(256) presplit_kernel(const float* __restrict__ X, int ldx, int M, int K,
                                                       __half* __restrict__ hi, __half* __restrict__ lo, int ldh,
                                                       float* __restrict__ inv_out, int ldinv) {
    const int nkc = (K + 63) >> 6;
    const long long total = (long long)M * nkc;
    const long long rc = (long long)blockIdx.x * 16 + (threadIdx.x >> 4);
    const int l16 = threadIdx.x & 15;
    const bool ok = rc < total;
    const int row = ok ? (int)(rc / nkc) : 0, kc = ok ? (int)(rc % nkc) : 0;
    const int col = kc * 64 + 4 * l16;
    const bool in = ok && col < K;   // K is a multiple of 32: the last chunk of a row may be half empty
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) v = *reinterpret_cast<const float4*>(X + (size_t)row * ldx + col);
    float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sc, inv;
    chunk_scale(m, sc, inv);
    uint2 h, l;
    split_pair(make_float2(v.x, v.y), sc, h.x, l.x);
    split_pair(make_float2(v.z, v.w), sc, h.y, l.y);
    if (!in) return;
    const size_t o = (size_t)row * ldh + col;
    *reinterpret_cast<uint2*>(hi + o) = h;
    *reinterpret_cast<uint2*>(lo + o) = l;
    if (l16 == 0) inv_out[(size_t)row * ldinv + kc] = inv;
}

// pre-split -> fp32 (test seam only)
__global__ void unsplit_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int ldh,
                               const float* __restrict__ inv, int ldinv, int M, int N, float* __restrict__ Y, int ldy) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)M * N) return;
    int row = (int)(i / N), col = (int)(i % N);
    float s = inv[(size_t)row * ldinv + (col >> 6)];
    Y[(size_t)row * ldy + col] = (__half2float(hi[(size_t)row * ldh + col]) + __half2float(lo[(size_t)row * ldh + col])) * s;
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
    float sc, inv;
    chunk_scale(__uint_as_float(*maxbits), sc, inv);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *inv_scale = inv;
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
unsigned long long* gemm_tc16_get_trace() { return g_trace; }

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

// X (M, K) fp32 -> pre-split form; K a multiple of 32
int presplit_f32(const float* X, int ldx, int M, int K, const SplitMat& out, cudaStream_t st) {
    AIM_REQUIRE(K % 32 == 0 && ldx % 4 == 0 && out.ld % 4 == 0 && out.ld >= K && out.ldinv >= K / 32, "presplit: bad layout");
    if (M == 0) return AIMNET_OK;
    long long total = (long long)M * ((K + 63) / 64);
    tc16::presplit_kernel<<<(unsigned)((total + 15) / 16), 256, 0, st>>>(X, ldx, M, K, (__half*)out.hi, (__half*)out.lo, out.ld,
                                                                         out.inv, out.ldinv);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

int unsplit_f32(const SplitMat& in, int M, int N, float* Y, int ldy, cudaStream_t st) {
    if (M == 0) return AIMNET_OK;
    size_t n = (size_t)M * N;
    tc16::unsplit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __half*)in.hi, (const __half*)in.lo, in.ld, in.inv,
                                                                       in.ldinv, M, N, Y, ldy);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

// A pre-split (M, K); output either fp32 Y (Ysplit == nullptr) or pre-split (Ysplit, for a consuming GEMM)
int gemm_nt_tc16(const SplitMat& A, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw, const float* bias,
                 float* Y, int ldy, const SplitMat* Ysplit, float* aux, int ldaux, int M, int N, int K, int mode,
                 cudaStream_t st) {
    using namespace tc16;
    AIM_REQUIRE(K % BK == 0 && N % 32 == 0, "gemm_tc16: K and N must be multiples of 32");
    AIM_REQUIRE(A.hi && A.lo && A.inv && A.ld % 8 == 0 && A.ld >= K && A.ldinv >= K / 32, "gemm_tc16: bad pre-split A operand");
    AIM_REQUIRE(((uintptr_t)A.hi & 15) == 0 && ((uintptr_t)A.lo & 15) == 0 && ((uintptr_t)Whi & 15) == 0 &&
                    ((uintptr_t)Wlo & 15) == 0 && ldw % 8 == 0,
                "gemm_tc16: operands must be 16-byte aligned");
    if (Ysplit) {
        AIM_REQUIRE(Ysplit->hi && Ysplit->lo && Ysplit->inv && Ysplit->ld % 8 == 0 && Ysplit->ld >= N && Ysplit->ldinv >= N / 32 &&
                        ((uintptr_t)Ysplit->hi & 15) == 0 && ((uintptr_t)Ysplit->lo & 15) == 0,
                    "gemm_tc16: bad pre-split output");
    } else {
        AIM_REQUIRE(Y && ((uintptr_t)Y & 15) == 0 && ldy % 4 == 0, "gemm_tc16: fp32 output must be 16-byte aligned");
    }
    AIM_REQUIRE(aux == nullptr || (((uintptr_t)aux & 15) == 0 && ldaux % 4 == 0), "gemm_tc16: aux must be 16-byte aligned");
    AIM_REQUIRE(w_inv_scale != nullptr, "gemm_tc16: weight scale missing");
    static bool configured_dev[kMaxDevices] = {};
    static int num_sms_dev[kMaxDevices] = {};
    const int dslot = current_device_slot();
    int& num_sms = num_sms_dev[dslot];
    if (!configured_dev[dslot]) {
#define AIM_TC16_ATTR(MODE)                                                                                                     \
    AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
    AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        AIM_TC16_ATTR(0)
        AIM_TC16_ATTR(1)
        AIM_TC16_ATTR(2)
        AIM_TC16_ATTR(3)
#undef AIM_TC16_ATTR
        int dev = 0;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured_dev[dslot] = true;
    }
    CUtensorMap tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux;
    const CUtensorMapDataType F16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16, F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle SW64 = CU_TENSOR_MAP_SWIZZLE_64B;
    int rc;
    if ((rc = make_map(&tmAh, A.hi, F16, 2, M, K, A.ld, BM, BK, SW64))) return rc;
    if ((rc = make_map(&tmAl, A.lo, F16, 2, M, K, A.ld, BM, BK, SW64))) return rc;
    int n_tiles = (N + BN - 1) / BN;
    int bn = ((N + n_tiles - 1) / n_tiles + 63) / 64 * 64;   // tile origins on chunk (64-column) boundaries of the output
    if ((rc = make_map(&tmBh, Whi, F16, 2, N, K, ldw, bn, BK, SW64))) return rc;
    if ((rc = make_map(&tmBl, Wlo, F16, 2, N, K, ldw, bn, BK, SW64))) return rc;
    if (Ysplit) {
        if ((rc = make_map(&tmY, Ysplit->hi, F16, 2, M, N, Ysplit->ld, 32, 32, SW64))) return rc;
        if ((rc = make_map(&tmY2, Ysplit->lo, F16, 2, M, N, Ysplit->ld, 32, 32, SW64))) return rc;
    } else {
        if ((rc = make_map(&tmY, Y, F32, 4, M, N, ldy, 32, 16, SW64))) return rc;
        tmY2 = tmY;
    }
    if (aux) {
        if ((rc = make_map(&tmAux, aux, F32, 4, M, N, ldaux, 32, 16, SW64))) return rc;
    } else {
        tmAux = tmY;
    }
    Params p{bias, w_inv_scale, A.inv, aux, Ysplit ? Ysplit->inv : nullptr, A.ldinv, Ysplit ? Ysplit->ldinv : 0, ldaux, M, N, K, mode, bn,
             g_trace};
    int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    int grid = tiles < num_sms ? tiles : num_sms;
#define AIM_TC16_LAUNCH(MODE)                                                                                            \
    if (Ysplit)                                                                                                          \
        gemm_tc16_kernel<MODE, true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux, p); \
    else                                                                                                                 \
        gemm_tc16_kernel<MODE, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux, p);
    switch (mode) {
        case 0: AIM_TC16_LAUNCH(0) break;
        case 1: AIM_TC16_LAUNCH(1) break;
        case 2: AIM_TC16_LAUNCH(2) break;
        default: AIM_TC16_LAUNCH(3) break;
    }
#undef AIM_TC16_LAUNCH
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

}  // namespace aimnet
