// tcgen05 backend 5 of the per-atom MLP GEMMs: the two-stream kernel of gemm_tc16d.cu on CTA PAIRS (cta_group::2).
//
// What backend 4 taught (tools/gemm_trace2.py, tools/tma_store_probe.cu): two 128 x 128 streams per SM do hide
// the tile epilogue, but each stream's K loop runs at ~1 850 clk per K=64 chunk instead of 768 — with A and W both
// streamed from L2 a 128 x 128 tile needs 85 B/clk per SM, and the three 32 KB stages a stream can afford cover ~1 150
// clk of MMA work against ~2 200 clk of TMA latency.  A CTA pair shares the W tile: the pair computes a 256 x 128 tile per
// stream, each CTA loads its own 128 rows of A and HALF of the W tile (64 rows), the leader CTA's MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256) and the accumulator rows 0-127 / 128-255 land in the TMEM of CTA 0 / CTA 1.
// Per SM that is 24 KB per stage (64 B/clk at full tensor rate) and four stages per stream.
//
//   cluster (2,1,1); per CTA: warps 0-3 / 4-7 epilogue of stream 0 / 1 (own 128 rows), warp 8 / 9 TMA producer of stream
//   0 / 1 (warp 8 also allocates TMEM, cta_group::2), warp 10 / 11 MMA issuer of stream 0 / 1 (leader CTA only).
//   full[s]       leader's barrier; both CTAs' TMA loads complete_tx on it (expect_tx covers both CTAs' bytes)
//   empty[s], tmem_full[b]   one per CTA, signalled by multicast tcgen05.commit (mask 0b11)
//   tmem_empty[b] leader's barrier, 8 arrivals (4 epilogue warps x 2 CTAs; the peer's arrive remotely)
// Arithmetic, chunking and epilogue are those of backend 2: bit-identical results.
#include "launchers.cuh"
#include "tc16_ptx.cuh"

namespace aimnet {

namespace tc16c {

using namespace tcx;

constexpr int BM = 128, BN = 128, BK = 32, NSTREAM = 2, STAGES = 4;   // BM: rows per CTA (the pair tile has 256)
constexpr int CHUNK_STAGES = 2;            // stages per TMEM chunk; one activation scale covers CHUNK_STAGES * BK = 64 columns
constexpr int A_HALF = BM * BK * 2;        // 8 KB: one fp16 A tile (hi or lo)
constexpr int B_BYTES = (BN / 2) * BK * 2; // 4 KB: this CTA's half of the W tile
constexpr int STAGE_BYTES = 2 * A_HALF + 2 * B_BYTES;   // 24 KB
constexpr int EPI_BOX = 2048;              // 32 rows x 64 bytes; two per epilogue warp
constexpr int OFF_BARS = NSTREAM * STAGES * STAGE_BYTES;   // 192 KB
constexpr int OFF_EPI = OFF_BARS + 2048;
constexpr int SMEM_BYTES = OFF_EPI + 16 * EPI_BOX + 1024 /*align*/;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int NUM_THREADS = 384;
constexpr int kWarpTma = 8, kWarpMma = 10;   // + stream
constexpr int kBarsPerStream = 12;           // full[4] empty[4] tmem_full[2] tmem_empty[2]

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a local shared-memory object) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on the barrier at cluster address `bar` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

struct Params {
    const float* bias;
    const float* w_inv_scale;   // device scalar: 1 / s_w of the (pre-scaled) weight tensor
    const float* a_inv;         // (M, lda_inv): 1 / s_a per row-chunk of A
    const float* aux;
    float* out_inv;             // split output: (M, ld_out_inv) inverse scales per row-chunk of Y
    int lda_inv, ld_out_inv, ldaux;
    int M, N, K, mode;
    int bn;      // N-tile width (multiple of 64, <= 128)
    unsigned long long* trace;   // debug: SM-clock stamps of CTA 0 (2 streams x 8 events x kTraceLen), or nullptr
};
constexpr int kTraceLen = 2048;
__device__ __forceinline__ void stamp(const Params& p, int strm, int ev, int idx) {
    if (p.trace != nullptr && blockIdx.x == 0 && idx < kTraceLen) p.trace[(strm * 8 + ev) * kTraceLen + idx] = clock64();
}

template <int MODE, bool SPLIT_OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc16c_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmY2,
                  const __grid_constant__ CUtensorMap tmAux, Params p) {
    extern __shared__ unsigned char smem_dyn[];
    // pointer arithmetic on the __shared__ array keeps the address space known to the compiler (LDS / STS, not generic)
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NSTREAM * kBarsPerStream);
    uint64_t* stagger = bars + NSTREAM * kBarsPerStream + 1;
    float* sbias_all = reinterpret_cast<float*>(smem + OFF_BARS + 256);              // [2][128]
    uint64_t* aux_bar = reinterpret_cast<uint64_t*>(smem + OFF_BARS + 1280);         // [8 warps][3]
    unsigned char* epi_buf = smem + OFF_EPI;                                         // 8 x 2 x 2 KB, 1 KB aligned

    constexpr int kStages = (MODE == 3) ? STAGES - 1 : STAGES;
    unsigned char* aux_buf = smem + NSTREAM * (STAGES - 1) * STAGE_BYTES;            // mode 3 only: 24 x 2 KB

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs)
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = (p.N + p.bn - 1) / p.bn;   // pair tiles: 256 rows
    const uint32_t tx_bytes = (uint32_t)(2 * (2 * A_HALF + 2 * (p.bn / 2) * BK * 2));       // both CTAs' loads of a stage
    const int tiles = m_tiles * n_tiles;
    const int nk = p.K / BK;
    const int nchunk = (nk + CHUNK_STAGES - 1) / CHUNK_STAGES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTREAM; ++s) {
            uint64_t* b = bars + s * kBarsPerStream;
            for (int j = 0; j < 2 * STAGES + 2; ++j) mbar_init(&b[j], 1);   // full, empty, tmem_full
            mbar_init(&b[2 * STAGES + 2], 8);                               // tmem_empty (leader's is used): one arrive
            mbar_init(&b[2 * STAGES + 3], 8);                               // per epilogue warp of the stream in both CTAs
        }
        mbar_init(stagger, 1);
        for (int w = 0; w < 24; ++w) mbar_init(&aux_bar[w], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpTma) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised and TMEM allocated before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // stream of this warp and its barriers
    const int strm = (warp < 8) ? (warp >> 2) : (warp & 1);
    uint64_t* full = bars + strm * kBarsPerStream;
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = full + 2 * STAGES;
    uint64_t* tmem_empty = full + 2 * STAGES + 2;
    unsigned char* ring = smem + strm * kStages * STAGE_BYTES;
    const uint32_t tmem_strm = tmem_base + (uint32_t)(strm * 2 * BN);

    if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");

    if (warp == kWarpTma || warp == kWarpTma + 1) {
        // ------------------------------------------------ TMA producer of stream `strm`
        if (lane == 0) {
            int sit = 0;   // running stage counter (trace only)
            int s = 0;
            uint32_t ph = 0;
            for (int t = cid + strm * ncl; t < tiles; t += NSTREAM * ncl) {
                const int m0 = (t / n_tiles) * 2 * BM + (int)rank * BM, n0 = (t % n_tiles) * p.bn;
                const int n_tile = min(p.bn, p.N - n0);
                const int nb = n0 + (int)rank * (n_tile >> 1);   // this CTA's half of the W tile
                for (int ks = 0; ks < nk; ++ks) {
                    mbar_wait(&empty[s], ph ^ 1);
                    stamp(p, strm, 0, sit);
                    unsigned char* sp = ring + s * STAGE_BYTES;
                    const uint32_t fb = map_to_cta(&full[s], 0);
                    if (rank == 0) mbar_expect_tx(&full[s], tx_bytes);
                    tma_load_2d_pair(sp, &tmAh, fb, ks * BK, m0);
                    tma_load_2d_pair(sp + A_HALF, &tmAl, fb, ks * BK, m0);
                    tma_load_2d_pair(sp + 2 * A_HALF, &tmBh, fb, ks * BK, nb);
                    tma_load_2d_pair(sp + 2 * A_HALF + B_BYTES, &tmBl, fb, ks * BK, nb);
                    stamp(p, strm, 3, sit++);
                    if (++s == kStages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == kWarpMma || warp == kWarpMma + 1) {
        // ------------------------------------------------ MMA issuer of stream `strm`
        if (lane == 0 && rank == 0) {
            int s = 0;
            uint32_t ph = 0;
            int cit = 0;   // running chunk counter -> TMEM buffer / phase
            bool first = true;
            for (int t = cid + strm * ncl; t < tiles; t += NSTREAM * ncl) {
                const int n0 = (t % n_tiles) * p.bn;
                const int n_tile = min(p.bn, p.N - n0);
                // kind::f16: D fp32 (bit 4), A/B fp16 (format 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
                const uint32_t idesc = (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
                // stream 1 starts one K loop behind stream 0: its MMAs then fall under stream 0's tile epilogue
                if (first && strm == 1) mbar_wait(stagger, 0);
                for (int ks = 0; ks < nk; ++cit) {
                    const int b = cit & 1;
                    const uint32_t aph = (uint32_t)(cit >> 1) & 1;
                    mbar_wait(&tmem_empty[b], aph ^ 1);
                    stamp(p, strm, 1, cit);
                    const uint32_t d_tmem = tmem_strm + (uint32_t)(b * BN);
                    for (int j = 0; j < CHUNK_STAGES && ks < nk; ++j, ++ks) {
                        mbar_wait(&full[s], ph);
                        if (j == 0) stamp(p, strm, 2, cit);
                        stamp(p, strm, 4, 2 * cit + j);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(ring + s * STAGE_BYTES);
                        const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_HALF);
                        const uint64_t b_hi = make_desc(sa + 2 * A_HALF), b_lo = make_desc(sa + 2 * A_HALF + B_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 32 >> 4);   // 16 halfs = 32 bytes along K inside the swizzle atom
                            tc_mma_f16_pair(d_tmem, a_lo + adv, b_hi + adv, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                            tc_mma_f16_pair(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            tc_mma_f16_pair(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                        }
                        tc_commit_pair(&empty[s]);   // frees the stage in both CTAs once these MMAs have read it
                        if (++s == kStages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    tc_commit_pair(&tmem_full[b]);
                }
                if (first && strm == 0) mbar_arrive(stagger);
                first = false;
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------ epilogue of stream `strm` (4 warps)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int ql = warp & 3;            // TMEM lane quarter this warp may access
        const float w_inv = *p.w_inv_scale;
        float* sbias = sbias_all + strm * BN;
        int cit = 0;
        unsigned char* box0 = epi_buf + warp * 2 * EPI_BOX;
        unsigned char* box1 = box0 + EPI_BOX;
        unsigned char* abox = aux_buf + warp * 3 * EPI_BOX;   // mode 3: this warp's aux ring
        uint64_t* abar = aux_bar + warp * 3;
        int ag = 0;                                           // running aux step -> ring slot ag % 3, phase (ag / 3) & 1
        const int rsw = (lane >> 1) & 3;
        const uint32_t te0 = map_to_cta(&tmem_empty[0], 0), te1 = map_to_cta(&tmem_empty[1], 0);   // the leader's barriers
        for (int t = cid + strm * ncl; t < tiles; t += NSTREAM * ncl) {
            const int m0 = (t / n_tiles) * 2 * BM + (int)rank * BM, n0 = (t % n_tiles) * p.bn;
            const int n_tile = min(p.bn, p.N - n0);
            const int row_base = m0 + ql * 32;
            const int row = row_base + lane;
            const float* inv_row = p.a_inv + (size_t)min(row, p.M - 1) * p.lda_inv;
            float2 acc[64];   // one output row x 128 columns, as register pairs for the packed FFMA2 / FMUL2 / FADD2
#pragma unroll
            for (int k = 0; k < 64; ++k) acc[k] = make_float2(0.f, 0.f);
            // mode 3: 16-column aux steps of this warp in this tile; the first three are requested now and arrive while
            // the tile's MMAs run
            const int aux_steps = (MODE == 3) ? n_tile / 16 : 0;
            auto aux_request = [&](int step, int slot) {   // lane 0 only
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&abar[slot], EPI_BOX);
                tma_load_2d(abox + slot * EPI_BOX, &tmAux, &abar[slot], n0 + 16 * step, row_base);
            };
            if (MODE == 3 && lane == 0) {
                for (int k = 0; k < 3 && k < aux_steps; ++k) aux_request(k, (ag + k) % 3);
            }
            if (MODE == 1 || MODE == 2) {
                // bias of this tile's columns -> shared memory (read back as warp-wide broadcasts in the epilogue)
                const int cb = threadIdx.x & 127;
                if (strm == 0) asm volatile("bar.sync 1, 128;"); else asm volatile("bar.sync 2, 128;");   // previous readers done
                sbias[cb] = (cb < n_tile) ? p.bias[n0 + cb] : 0.f;
                if (strm == 0) asm volatile("bar.sync 1, 128;"); else asm volatile("bar.sync 2, 128;");
            }
            float inv_next = __ldg(inv_row) * w_inv;
            for (int kc = 0; kc < nchunk; ++cit, ++kc) {
                const float inv = inv_next;
                if (kc + 1 < nchunk) inv_next = __ldg(inv_row + kc + 1) * w_inv;   // in flight while we wait for the chunk
                const int b = cit & 1;
                const uint32_t aph = (uint32_t)(cit >> 1) & 1;
                mbar_wait(&tmem_full[b], aph);
                if (ql == 0 && lane == 0) stamp(p, strm, 5, cit);
                tc_fence_after();
                const uint32_t taddr = tmem_strm + ((uint32_t)(ql * 32) << 16) + (uint32_t)(b * BN);
                // two 32-column loads in flight at a time
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int col0 = c2 * 64;
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
                if (lane == 0) mbar_arrive_cluster(b ? te1 : te0);
                if (ql == 0 && lane == 0) stamp(p, strm, 6, cit);
            }
            // ---- tile epilogue (see gemm_tc16.cu): each thread holds one output row x 128 columns; outputs leave
            // through 64B-swizzled 32-row shared-memory boxes as TMA bulk tensor stores, two boxes per warp as a ring.
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
                if (c2 * 64 < n_tile) {
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
                        const int col0 = c2 * 64 + hc * 32;
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
                                    const float4 bz = *reinterpret_cast<const float4*>(sbias + col0 + 4 * v4);
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
                        const int colp = n0 + c2 * 64;
                        if (row < p.M) p.out_inv[(size_t)row * p.ld_out_inv + (colp >> 6)] = inv;
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            if (c2 * 64 + hc * 32 < n_tile) {
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
            if (ql == 0 && lane == 0) stamp(p, strm, 7, cit - 1);   // end of this tile's epilogue (indexed by its last chunk)
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores retired before exit
    }
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal it or read its operands
    if (warp == kWarpTma) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace tc16c

// same contract as gemm_nt_tc16 (gemm_tc16.cu)
int gemm_nt_tc16c(const SplitMat& A, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw, const float* bias,
                  float* Y, int ldy, const SplitMat* Ysplit, float* aux, int ldaux, int M, int N, int K, int mode,
                  cudaStream_t st) {
    using namespace tc16c;
    AIM_REQUIRE(K % BK == 0 && N % 32 == 0, "gemm_tc16c: K and N must be multiples of 32");
    AIM_REQUIRE(A.hi && A.lo && A.inv && A.ld % 8 == 0 && A.ld >= K && A.ldinv >= K / 32, "gemm_tc16c: bad pre-split A operand");
    AIM_REQUIRE(((uintptr_t)A.hi & 15) == 0 && ((uintptr_t)A.lo & 15) == 0 && ((uintptr_t)Whi & 15) == 0 &&
                    ((uintptr_t)Wlo & 15) == 0 && ldw % 8 == 0,
                "gemm_tc16c: operands must be 16-byte aligned");
    if (Ysplit) {
        AIM_REQUIRE(Ysplit->hi && Ysplit->lo && Ysplit->inv && Ysplit->ld % 8 == 0 && Ysplit->ld >= N && Ysplit->ldinv >= N / 32 &&
                        ((uintptr_t)Ysplit->hi & 15) == 0 && ((uintptr_t)Ysplit->lo & 15) == 0,
                    "gemm_tc16c: bad pre-split output");
    } else {
        AIM_REQUIRE(Y && ((uintptr_t)Y & 15) == 0 && ldy % 4 == 0, "gemm_tc16c: fp32 output must be 16-byte aligned");
    }
    AIM_REQUIRE(aux == nullptr || (((uintptr_t)aux & 15) == 0 && ldaux % 4 == 0), "gemm_tc16c: aux must be 16-byte aligned");
    AIM_REQUIRE(w_inv_scale != nullptr, "gemm_tc16c: weight scale missing");
    static bool configured_dev[kMaxDevices] = {};
    static int num_sms_dev[kMaxDevices] = {};
    const int dslot = current_device_slot();
    int& num_sms = num_sms_dev[dslot];
    if (!configured_dev[dslot]) {
#define AIM_TC16C_ATTR(MODE)                                                                                                      \
    AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16c_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
    AIM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc16c_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        AIM_TC16C_ATTR(0)
        AIM_TC16C_ATTR(1)
        AIM_TC16C_ATTR(2)
        AIM_TC16C_ATTR(3)
#undef AIM_TC16C_ATTR
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
    const int n_tiles = (N + BN - 1) / BN;
    const int bn = ((N + n_tiles - 1) / n_tiles + 63) / 64 * 64;   // tile origins on chunk (64-column) boundaries of the output
    if ((rc = make_map(&tmBh, Whi, F16, 2, N, K, ldw, bn / 2, BK, SW64))) return rc;   // a CTA loads half of the W tile
    if ((rc = make_map(&tmBl, Wlo, F16, 2, N, K, ldw, bn / 2, BK, SW64))) return rc;
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
             gemm_tc16_get_trace()};
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + bn - 1) / bn);   // pair tiles
    const int clusters = tiles < num_sms / 2 ? tiles : num_sms / 2;
    const int grid = 2 * clusters;
#define AIM_TC16C_LAUNCH(MODE)                                                                                            \
    if (Ysplit)                                                                                                           \
        gemm_tc16c_kernel<MODE, true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux, p); \
    else                                                                                                                  \
        gemm_tc16c_kernel<MODE, false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmAh, tmAl, tmBh, tmBl, tmY, tmY2, tmAux, p);
    switch (mode) {
        case 0: AIM_TC16C_LAUNCH(0) break;
        case 1: AIM_TC16C_LAUNCH(1) break;
        case 2: AIM_TC16C_LAUNCH(2) break;
        default: AIM_TC16C_LAUNCH(3) break;
    }
#undef AIM_TC16C_LAUNCH
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

}  // namespace aimnet
