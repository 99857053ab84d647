// tcgen05 backend 3 (EXPERIMENTAL, reachable through the aimnet2_gemm_nt seam only): the 3xFP16 row-chunk-scaled GEMM
// of gemm_tc16.cu with the tile epilogue software-pipelined under the next tile's MMAs.
//
// STATUS: written at the end of round 1 with no GPU time left.  It compiles for sm_100a and has never run.  The engine
// does not use it; tests/test_gpu_unverified.py holds its first tests.  DESIGN.md section 8 item 1(c) is the plan it follows.
//
// Why: in gemm_tc16.cu the eight epilogue warps both drain the K=64 chunk accumulators out of TMEM and, once a tile's K
// loop is over, run its epilogue (bias / GELU / split / stores, ~13 k clk for 128x256) while the MMA issuer can only
// run two chunks ahead: the tensor pipe idles for most of that time (46 % busy).  Here
//   * tiles are 128 x 128, so TMEM holds FOUR chunk accumulators (4 x 128 columns) and the issuer can run four chunks
//     (~4 k clk) ahead of the drains;
//   * every epilogue thread owns one row x 64 columns and keeps TWO accumulator sets (2 x 64 registers instead of one
//     set of 128): while it drains the chunks of tile t+1 into one set, it runs the epilogue of tile t from the other,
//     one slice after each of the first three chunk drains (slice 0: columns 0-31 through the epilogue math, slice 1:
//     columns 32-63 + the row-chunk scale, slice 2: the pre-split stores), so the ~6.5 k clk of epilogue are spread over
//     a K loop of 11 x 1 k clk and the issuer is never held up.
// Everything else (operand layout, precision scheme, box rings, TMA stores, bitwise reproducibility) is gemm_tc16.cu's.
// The bias is read straight from global memory (warp-uniform float4 loads) instead of a shared-memory staging area,
// which removes the per-tile CTA barriers.
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

namespace tc16p {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 4, NBUF = 4;
constexpr int CHUNK_STAGES = 2;            // stages per TMEM chunk; one activation scale covers CHUNK_STAGES * BK = 64 columns
constexpr int A_HALF = BM * BK * 2;        // 8 KB: one fp16 A tile (hi or lo)
constexpr int B_BYTES = BN * BK * 2;       // 8 KB
constexpr int STAGE_BYTES = 2 * A_HALF + 2 * B_BYTES;   // 32 KB
constexpr int EPI_BOX = 2048;              // 32 rows x 64 bytes
constexpr int OFF_BARS = STAGES * STAGE_BYTES;
constexpr int OFF_EPI = OFF_BARS + 2048;             // 8 warps x 2 store boxes
constexpr int OFF_AUX = OFF_EPI + 16 * EPI_BOX;      // 8 warps x 3 aux (mode 3) boxes
constexpr int SMEM_BYTES = OFF_AUX + 24 * EPI_BOX + 1024 /*align*/;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int NUM_THREADS = 384;
constexpr int kWarpAlloc = 8, kWarpTma = 10, kWarpMma = 11;   // epilogue = warps 0-7

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

// K-major, 64B-swizzled shared-memory operand descriptor (see gemm_tc16.cu)
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
    int bn;                     // N-tile width: 64 or 128
};

__device__ __forceinline__ void chunk_scale(float m, float& sc, float& inv) {
    int e = (int)(__float_as_uint(m) >> 23);
    e = min(max(e, 14), 254);
    sc = __uint_as_float((uint32_t)(267 - e) << 23);
    inv = __uint_as_float((uint32_t)(e - 13) << 23);
}
__device__ __forceinline__ void split_pair(float2 v, float sc, uint32_t& hi, uint32_t& lo) {
    const float2 sv = fmul2(v, make_float2(sc, sc));
    const __half2 h = __floats2half2_rn(sv.x, sv.y);
    const float2 hf = __half22float2(h);
    const float2 df = ffma2(hf, make_float2(-1.0f, -1.0f), sv);
    const __half2 l = __floats2half2_rn(df.x, df.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// per-warp epilogue state: the box rings and the tile whose epilogue is in flight
struct Epi {
    const Params* p;
    const CUtensorMap* mapY;
    const CUtensorMap* mapY2;
    const CUtensorMap* mapAux;
    unsigned char* box0;
    unsigned char* box1;
    unsigned char* abox;
    uint64_t* abar;
    int box_i;        // next box of the two-box store ring
    int ag;           // running aux step: ring slot ag % 3, phase (ag / 3) & 1
    int lane, rsw, ch;
    // the finished tile
    int row_base, row, n0, n_tile, aux_steps, aux_step;
    float sc;
};

__device__ __forceinline__ unsigned char* box_acquire(Epi& c) {
    if (c.lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    unsigned char* bx = c.box_i ? c.box1 : c.box0;
    c.box_i ^= 1;
    return bx;
}
__device__ __forceinline__ void box_store(Epi& c, const CUtensorMap* map, const unsigned char* bx, int c0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (c.lane == 0) {
        tma_store_2d(map, bx, c0, c.row_base);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
__device__ __forceinline__ void aux_request(Epi& c, int step, int slot) {   // lane 0 only
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&c.abar[slot], EPI_BOX);
    tma_load_2d(c.abox + slot * EPI_BOX, c.mapAux, &c.abar[slot], c.n0 + c.ch * 64 + 16 * step, c.row_base);
}

// a tile has finished its K loop: remember where it lives and start fetching its aux operand
template <int MODE>
__device__ __forceinline__ void epi_begin(Epi& c, int m0, int n0, int n_tile, int ql) {
    c.row_base = m0 + ql * 32;
    c.row = c.row_base + c.lane;
    c.n0 = n0;
    c.n_tile = n_tile;
    c.aux_step = 0;
    c.aux_steps = (MODE == 3) ? max(0, min(n_tile - c.ch * 64, 64)) / 16 : 0;
    if (MODE == 3 && c.lane == 0) {
        for (int k = 0; k < 3 && k < c.aux_steps; ++k) aux_request(c, k, (c.ag + k) % 3);
    }
}

// epilogue math of one 32-column group (HC = 0, 1) of the finished tile; fp32 outputs leave here, pre-split ones later
template <int MODE, bool SPLIT_OUT, int HC>
__device__ __forceinline__ void epi_math(float2 (&fin)[32], Epi& c) {
    const int col0 = c.ch * 64 + HC * 32;
    if (col0 >= c.n_tile) return;
    const int col = c.n0 + col0;
    float2* v = &fin[HC * 16];
    if (MODE == 3) {
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            const int slot = c.ag % 3;
            mbar_wait(&c.abar[slot], (uint32_t)(c.ag / 3) & 1);
            const unsigned char* bx = c.abox + slot * EPI_BOX;
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
                const float4 g = *reinterpret_cast<const float4*>(bx + c.lane * 64 + ((v4 ^ c.rsw) << 4));
                v[hb * 8 + 2 * v4 + 0] = fmul2(v[hb * 8 + 2 * v4 + 0], make_float2(g.x, g.y));
                v[hb * 8 + 2 * v4 + 1] = fmul2(v[hb * 8 + 2 * v4 + 1], make_float2(g.z, g.w));
            }
            __syncwarp();   // every lane has read the box: refill it three steps ahead
            if (c.lane == 0 && c.aux_step + 3 < c.aux_steps) aux_request(c, c.aux_step + 3, slot);
            ++c.ag;
            ++c.aux_step;
        }
    } else if (MODE == 1 || MODE == 2) {
        const float4* bz4 = reinterpret_cast<const float4*>(c.p->bias + col);   // warp-uniform: one broadcast per load
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
            const float4 bz = __ldg(bz4 + v4);
            v[2 * v4 + 0] = fadd2(v[2 * v4 + 0], make_float2(bz.x, bz.y));
            v[2 * v4 + 1] = fadd2(v[2 * v4 + 1], make_float2(bz.z, bz.w));
        }
    }
    if (MODE == 2) {
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float2 g[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) gelu_pair2(v[hb * 8 + k], v[hb * 8 + k], g[k]);
            if (c.p->aux != nullptr) {
                unsigned char* bx = box_acquire(c);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    *reinterpret_cast<float4*>(bx + c.lane * 64 + ((v4 ^ c.rsw) << 4)) =
                        make_float4(g[2 * v4].x, g[2 * v4].y, g[2 * v4 + 1].x, g[2 * v4 + 1].y);
                box_store(c, c.mapAux, bx, col + 16 * hb);
            }
        }
    }
    if (!SPLIT_OUT) {
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            unsigned char* bx = box_acquire(c);
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
                const float2 z0 = v[hb * 8 + 2 * v4 + 0], z1 = v[hb * 8 + 2 * v4 + 1];
                *reinterpret_cast<float4*>(bx + c.lane * 64 + ((v4 ^ c.rsw) << 4)) = make_float4(z0.x, z0.y, z1.x, z1.y);
            }
            box_store(c, c.mapY, bx, col + 16 * hb);
        }
    }
}

// the thread's 64 values are one row-chunk (K=64) of the consuming GEMM: its scale (columns past n_tile are zero)
__device__ __forceinline__ void epi_scale(float2 (&fin)[32], Epi& c) {
    if (c.ch * 64 >= c.n_tile) return;
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) m = fmaxf(m, fmaxf(fabsf(fin[k].x), fabsf(fin[k].y)));
    float inv;
    chunk_scale(m, c.sc, inv);
    const int colp = c.n0 + c.ch * 64;
    if (c.row < c.p->M) c.p->out_inv[(size_t)c.row * c.p->ld_out_inv + (colp >> 6)] = inv;
}

// pre-split stores of one 32-column group
template <int HC>
__device__ __forceinline__ void epi_split_store(float2 (&fin)[32], Epi& c) {
    if (c.ch * 64 + HC * 32 >= c.n_tile) return;
    const int colp = c.n0 + c.ch * 64 + HC * 32;
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) split_pair(fin[HC * 16 + k], c.sc, hi[k], lo[k]);
    unsigned char* bh = box_acquire(c);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(bh + c.lane * 64 + ((j ^ c.rsw) << 4)) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
    box_store(c, c.mapY, bh, colp);
    unsigned char* bl = box_acquire(c);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(bl + c.lane * 64 + ((j ^ c.rsw) << 4)) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
    box_store(c, c.mapY2, bl, colp);
}

// the finished tile's epilogue in three slices of comparable length
template <int MODE, bool SPLIT_OUT, int SLICE>
__device__ __forceinline__ void epi_slice(float2 (&fin)[32], Epi& c) {
    if (SLICE == 0) {
        epi_math<MODE, SPLIT_OUT, 0>(fin, c);
    } else if (SLICE == 1) {
        epi_math<MODE, SPLIT_OUT, 1>(fin, c);
        if (SPLIT_OUT) {
            epi_scale(fin, c);
            epi_split_store<0>(fin, c);
        }
    } else {
        if (SPLIT_OUT) epi_split_store<1>(fin, c);
    }
}

// what the epilogue warps need to follow the MMA issuer
struct Pipe {
    uint64_t* tmem_full;
    uint64_t* tmem_empty;
    volatile int* chunk_last;
    uint32_t tmem_base;
    int cit;          // running chunk counter -> TMEM buffer cit % NBUF, phase (cit / NBUF) & 1
    int ql;
    int nchunk;
    float w_inv;
};

// K loop of tile t into `acc`, with the epilogue of the previously finished tile (`fin`) interleaved
template <int MODE, bool SPLIT_OUT>
__device__ __forceinline__ void run_tile(float2 (&acc)[32], float2 (&fin)[32], bool have_fin, int m0, int n0, int n_tile,
                                         Epi& c, Pipe& px) {
    const int row = m0 + px.ql * 32 + c.lane;
    const float* inv_row = c.p->a_inv + (size_t)min(row, c.p->M - 1) * c.p->lda_inv;
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = make_float2(0.f, 0.f);
    float inv_next = __ldg(inv_row) * px.w_inv;
    int kc = 0;
    for (int last = 0; !last; ++px.cit, ++kc) {
        const float inv = inv_next;
        if (kc + 1 < px.nchunk) inv_next = __ldg(inv_row + kc + 1) * px.w_inv;
        const int b = px.cit & (NBUF - 1);
        const uint32_t aph = (uint32_t)(px.cit / NBUF) & 1;
        mbar_wait(&px.tmem_full[b], aph);
        tc_fence_after();
        last = px.chunk_last[b];
        // 64 columns as two 32-column loads, one after the other: 32 live temporaries instead of 64 next to the two
        // accumulator sets (ptxas serialises the pair in gemm_tc16.cu anyway)
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
            if (c.ch * 64 + hc * 32 < n_tile) {
                uint32_t r[32];
                tc_ld32(px.tmem_base + ((uint32_t)(px.ql * 32) << 16) + (uint32_t)(b * BN + c.ch * 64 + hc * 32), r);
                tc_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    acc[hc * 16 + k] = ffma2s(inv, make_float2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])), acc[hc * 16 + k]);
            }
        }
        tc_fence_before();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&px.tmem_empty[b]);
        // the issuer is up to NBUF chunks ahead: time for a slice of the previous tile's epilogue
        if (have_fin) {
            if (kc == 0) epi_slice<MODE, SPLIT_OUT, 0>(fin, c);
            else if (kc == 1) epi_slice<MODE, SPLIT_OUT, 1>(fin, c);
            else if (kc == 2) epi_slice<MODE, SPLIT_OUT, 2>(fin, c);
        }
    }
    if (have_fin) {   // short K loops: whatever is left of the previous tile's epilogue
        if (kc <= 1) epi_slice<MODE, SPLIT_OUT, 1>(fin, c);
        if (kc <= 2) epi_slice<MODE, SPLIT_OUT, 2>(fin, c);
    }
    epi_begin<MODE>(c, m0, n0, n_tile, px.ql);
}

template <int MODE, bool SPLIT_OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc16p_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmY2,
                  const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* full = bars;                               // [STAGES]
    uint64_t* empty = bars + STAGES;                     // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;             // [NBUF]
    uint64_t* tmem_empty = bars + 2 * STAGES + NBUF;     // [NBUF]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NBUF);
    volatile int* chunk_last = reinterpret_cast<volatile int*>(bars + 2 * STAGES + 2 * NBUF + 1);   // [NBUF] ints
    uint64_t* aux_bar = reinterpret_cast<uint64_t*>(smem + OFF_BARS + 1024);                        // [8 warps][3]
    unsigned char* epi_buf = smem + OFF_EPI;
    unsigned char* aux_buf = smem + OFF_AUX;

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
        for (int b = 0; b < NBUF; ++b) {
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
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * p.bn;
                for (int ks = 0; ks < nk; ++ks) {
                    mbar_wait(&empty[s], ph ^ 1);
                    unsigned char* sp = stage_ptr(s);
                    mbar_expect_tx(&full[s], tx_bytes);
                    tma_load_2d(sp, &tmAh, &full[s], ks * BK, m0);
                    tma_load_2d(sp + A_HALF, &tmAl, &full[s], ks * BK, m0);
                    tma_load_2d(sp + 2 * A_HALF, &tmBh, &full[s], ks * BK, n0);
                    tma_load_2d(sp + 2 * A_HALF + B_BYTES, &tmBl, &full[s], ks * BK, n0);
                    if (++s == STAGES) {
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
            int cit = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int n0 = (t % n_tiles) * p.bn;
                int n_tile = min(p.bn, p.N - n0);
                // kind::f16: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
                uint32_t idesc = (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
                for (int ks = 0; ks < nk; ++cit) {
                    int b = cit & (NBUF - 1);
                    uint32_t aph = (uint32_t)(cit / NBUF) & 1;
                    mbar_wait(&tmem_empty[b], aph ^ 1);
                    tc_fence_after();
                    uint32_t d_tmem = tmem_base + (uint32_t)(b * BN);
                    for (int j = 0; j < CHUNK_STAGES && ks < nk; ++j, ++ks) {
                        mbar_wait(&full[s], ph);
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
                        if (++s == STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    chunk_last[b] = (ks == nk) ? 1 : 0;
                    __threadfence_block();
                    tc_commit(&tmem_full[b]);
                }
            }
        }
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    } else {
        // ------------------------------------------------ epilogue (warps 0-7): lane quarter ql, column half ch
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        Epi c;
        c.p = &p;
        c.mapY = &tmY;
        c.mapY2 = &tmY2;
        c.mapAux = &tmAux;
        c.box0 = epi_buf + warp * 2 * EPI_BOX;
        c.box1 = c.box0 + EPI_BOX;
        c.abox = aux_buf + warp * 3 * EPI_BOX;
        c.abar = aux_bar + warp * 3;
        c.box_i = 0;
        c.ag = 0;
        c.lane = lane;
        c.rsw = (lane >> 1) & 3;
        c.ch = warp >> 2;
        c.row_base = c.row = c.n0 = c.n_tile = c.aux_steps = c.aux_step = 0;
        c.sc = 1.f;
        Pipe px;
        px.tmem_full = tmem_full;
        px.tmem_empty = tmem_empty;
        px.chunk_last = chunk_last;
        px.tmem_base = tmem_base;
        px.cit = 0;
        px.ql = warp & 3;
        px.nchunk = (nk + CHUNK_STAGES - 1) / CHUNK_STAGES;
        px.w_inv = *p.w_inv_scale;
        float2 accA[32], accB[32];   // two accumulator sets: one collects the running tile, the other waits for its epilogue
        bool have_fin = false;
        int t = blockIdx.x;
        while (t < tiles) {
            {
                const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * p.bn;
                run_tile<MODE, SPLIT_OUT>(accA, accB, have_fin, m0, n0, min(p.bn, p.N - n0), c, px);
                have_fin = true;
                t += gridDim.x;
            }
            if (t >= tiles) {
                epi_slice<MODE, SPLIT_OUT, 0>(accA, c);
                epi_slice<MODE, SPLIT_OUT, 1>(accA, c);
                epi_slice<MODE, SPLIT_OUT, 2>(accA, c);
                have_fin = false;
                break;
            }
            {
                const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * p.bn;
                run_tile<MODE, SPLIT_OUT>(accB, accA, true, m0, n0, min(p.bn, p.N - n0), c, px);
                t += gridDim.x;
            }
            if (t >= tiles) {
                epi_slice<MODE, SPLIT_OUT, 0>(accB, c);
                epi_slice<MODE, SPLIT_OUT, 1>(accB, c);
                epi_slice<MODE, SPLIT_OUT, 2>(accB, c);
                have_fin = false;
                break;
            }
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
                    int box_rows, int box_cols) {
    EncodeFn enc = get_encode();
    if (!enc) {
        set_error("gemm_tc16p: cuTensorMapEncodeTiled not available");
        return AIMNET_ECUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc16p: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return AIMNET_ECUDA;
    }
    return AIMNET_OK;
}

}  // namespace tc16p

// same contract as gemm_nt_tc16 (gemm_tc16.cu)
int gemm_nt_tc16p(const SplitMat& A, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw, const float* bias,
                  float* Y, int ldy, const SplitMat* Ysplit, float* aux, int ldaux, int M, int N, int K, int mode,
                  cudaStream_t st) {
    using namespace tc16p;
    AIM_REQUIRE(K % BK == 0 && N % 32 == 0, "gemm_tc16p: K and N must be multiples of 32");
    AIM_REQUIRE(A.hi && A.lo && A.inv && A.ld % 8 == 0 && A.ld >= K && A.ldinv >= K / 32, "gemm_tc16p: bad pre-split A operand");
    AIM_REQUIRE(((uintptr_t)A.hi & 15) == 0 && ((uintptr_t)A.lo & 15) == 0 && ((uintptr_t)Whi & 15) == 0 &&
                    ((uintptr_t)Wlo & 15) == 0 && ldw % 8 == 0,
                "gemm_tc16p: operands must be 16-byte aligned");
    if (Ysplit) {
        AIM_REQUIRE(Ysplit->hi && Ysplit->lo && Ysplit->inv && Ysplit->ld % 8 == 0 && Ysplit->ld >= N && Ysplit->ldinv >= N / 32 &&
                        ((uintptr_t)Ysplit->hi & 15) == 0 && ((uintptr_t)Ysplit->lo & 15) == 0,
                    "gemm_tc16p: bad pre-split output");
    } else {
        AIM_REQUIRE(Y && ((uintptr_t)Y & 15) == 0 && ldy % 4 == 0, "gemm_tc16p: fp32 output must be 16-byte aligned");
    }
    AIM_REQUIRE(aux == nullptr || (((uintptr_t)aux & 15) == 0 && ldaux % 4 == 0), "gemm_tc16p: aux must be 16-byte aligned");
    AIM_REQUIRE(bias == nullptr || ((uintptr_t)bias & 15) == 0, "gemm_tc16p: bias must be 16-byte aligned");
    AIM_REQUIRE(w_inv_scale != nullptr, "gemm_tc16p: weight scale missing");
    if (M == 0) return AIMNET_OK;
    static bool configured_dev[kMaxDevices] = {};
    static int num_sms_dev[kMaxDevices] = {};
    const int dslot = current_device_slot();
    int& num_sms = num_sms_dev[dslot];
    if (!configured_dev[dslot]) {
#define AIM_TC16P_ATTR(MODE)                                                                                                      \
    AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16p_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
    AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16p_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        AIM_TC16P_ATTR(0)
        AIM_TC16P_ATTR(1)
        AIM_TC16P_ATTR(2)
        AIM_TC16P_ATTR(3)
#undef AIM_TC16P_ATTR
        int dev = 0;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured_dev[dslot] = true;
    }
    CUtensorMap tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux;
    const CUtensorMapDataType F16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16, F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    int rc;
    if ((rc = make_map(&tmAh, A.hi, F16, 2, M, K, A.ld, BM, BK))) return rc;
    if ((rc = make_map(&tmAl, A.lo, F16, 2, M, K, A.ld, BM, BK))) return rc;
    int n_tiles = (N + BN - 1) / BN;
    int bn = ((N + n_tiles - 1) / n_tiles + 63) / 64 * 64;   // 64 or 128: tile origins on row-chunk boundaries of the output
    if ((rc = make_map(&tmBh, Whi, F16, 2, N, K, ldw, bn, BK))) return rc;
    if ((rc = make_map(&tmBl, Wlo, F16, 2, N, K, ldw, bn, BK))) return rc;
    if (Ysplit) {
        if ((rc = make_map(&tmY, Ysplit->hi, F16, 2, M, N, Ysplit->ld, 32, 32))) return rc;
        if ((rc = make_map(&tmY2, Ysplit->lo, F16, 2, M, N, Ysplit->ld, 32, 32))) return rc;
    } else {
        if ((rc = make_map(&tmY, Y, F32, 4, M, N, ldy, 32, 16))) return rc;
        tmY2 = tmY;
    }
    if (aux) {
        if ((rc = make_map(&tmAux, aux, F32, 4, M, N, ldaux, 32, 16))) return rc;
    } else {
        tmAux = tmY;
    }
    Params p{bias, w_inv_scale, A.inv, aux, Ysplit ? Ysplit->inv : nullptr, A.ldinv, Ysplit ? Ysplit->ldinv : 0, ldaux, M, N, K, mode, bn};
    int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    int grid = tiles < num_sms ? tiles : num_sms;
#define AIM_TC16P_LAUNCH(MODE)                                                                                            \
    if (Ysplit)                                                                                                           \
        gemm_tc16p_kernel<MODE, true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux, p); \
    else                                                                                                                  \
        gemm_tc16p_kernel<MODE, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux, p);
    switch (mode) {
        case 0: AIM_TC16P_LAUNCH(0) break;
        case 1: AIM_TC16P_LAUNCH(1) break;
        case 2: AIM_TC16P_LAUNCH(2) break;
        default: AIM_TC16P_LAUNCH(3) break;
    }
#undef AIM_TC16P_LAUNCH
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

}  // namespace aimnet
