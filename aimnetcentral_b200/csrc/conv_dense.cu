// AEV + conv_sv message passing for batches of small molecules: the molecule's feature tables are staged into shared memory
// with TMA and every pair of the molecule is walked from there ("dense walk").  Forward and analytic backward.
//
// Same arithmetic and reference semantics as conv.cu (calc_distances aimnet/ops.py:37-66, AEVSV._calc_aev
// aimnet/modules/aev.py:94-110, ConvSV.forward aimnet/modules/aev.py:156-189, Warp kernels
// aimnet/kernels/conv_sv_2d_sp_wp.py:90-164).  Why a second implementation: the list kernels of conv.cu gather every
// neighbour row (1 KB of a[j], 4 KB of dS[j]) through L1 / L2, and ncu shows them waiting on those gathers (long-scoreboard
// stalls 5.5 / 3.5 per issued instruction, FMA pipe 35 / 45 %; an intermediate version that only re-mapped the threads,
// two centres per warp and all 16 channels per lane, executed 25 % fewer instructions and was SLOWER, 8.3 stalls per issue:
// profiles/r2e_conv2_*).  For a molecule of n <= 128 atoms everything a CTA gathers is one contiguous block of the feature
// table (atoms of a molecule are contiguous), so:
//
//   * one CTA works on one molecule (forward) or one (molecule, quarter of the radial shifts) item (backward) at a time;
//     a single elected thread fetches the block with cp.async.bulk.tensor (2-D tensor maps over the (N, 4 q, 16 g, 4 a) feature
//     table and the (N, 16 a, 16 g, 4 d) gradient table; the box selects the item's radial-shift columns), double-buffered
//     where it fits: the next item's block lands while this one is computed, completion through an mbarrier;
//   * every centre atom walks ALL atoms of its molecule in index order (pairs beyond the cutoff contribute exactly zero:
//     fc(d >= rc) == 0; the order equals that of the canonical sorted list rows, so the forward sums are bitwise those of the
//     list walk); the centres of a warp move in lock step, so one shared-memory read of a neighbour row serves 2 (forward) or
//     8 (backward) pairs and there is no global gather left in the pair loop;
//   * lane = (centre, radial shift g) owns all 16 feature channels of its (i, g): accumulators are register pairs over two
//     neighbouring channels fed straight by the float4 reads, the pair weight is the packed broadcast operand.
//   * backward: per atom from its OWN pairs only (gather form, no atomics, deterministic; see conv.cu); the four
//     radial-shift quarters of an atom write partial forces / charge gradients that a last kernel adds in fixed order.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {
namespace convd {

constexpr int kSlots = 16;           // forward: neighbour slots staged per centre and round
// shared-memory layouts of the forward epilogue, as in conv.cu
constexpr int kAghRow = 20;
constexpr int kSvRow = 52;
constexpr int kSvAtom = kA * kSvRow + 16;
__device__ __forceinline__ int sv_off(int a) { return a * kSvRow + ((a >> 3) << 4); }

struct PairEntry {
    float ux, uy, uz, d;
    float fc, dfc;
    int j;       // index of the neighbour inside the molecule
    float inv;   // 1/d
};

__device__ __forceinline__ float aev_exp(float x) { return __expf(x); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "CONVD_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CONVD_DONE;\n\t"
        "bra CONVD_WAIT;\n\t"
        "CONVD_DONE:\n\t"
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

__device__ __forceinline__ void mix16(const float* __restrict__ w, const float* __restrict__ s, float* t) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4* s4 = reinterpret_cast<const float4*>(s + k * kG);
        const float4 s0 = s4[0], s1 = s4[1], s2 = s4[2], s3 = s4[3];
        float2 acc = make_float2(0.f, 0.f);
        acc = ffma2(make_float2(w0.x, w0.y), make_float2(s0.x, s0.y), acc);
        acc = ffma2(make_float2(w0.z, w0.w), make_float2(s0.z, s0.w), acc);
        acc = ffma2(make_float2(w1.x, w1.y), make_float2(s1.x, s1.y), acc);
        acc = ffma2(make_float2(w1.z, w1.w), make_float2(s1.z, s1.w), acc);
        acc = ffma2(make_float2(w2.x, w2.y), make_float2(s2.x, s2.y), acc);
        acc = ffma2(make_float2(w2.z, w2.w), make_float2(s2.z, s2.w), acc);
        acc = ffma2(make_float2(w3.x, w3.y), make_float2(s3.x, s3.y), acc);
        acc = ffma2(make_float2(w3.z, w3.w), make_float2(s3.z, s3.w), acc);
        t[k] = acc.x + acc.y;
    }
}

// geometry + cutoff of the pair (centre il, neighbour k) of one molecule, coordinates from shared memory
template <bool kWithDeriv>
__device__ __forceinline__ PairEntry pair_entry(int il, bool atom_ok, int k, int n, const float4* __restrict__ xyz,
                                                const AevParams& aev) {
    const bool ok = atom_ok && k < n && k != il;
    float rx = 1.f, ry = 1.f, rz = 1.f;
    if (ok) {
        const float4 xi = xyz[il], xj = xyz[k];
        rx = xj.x - xi.x;
        ry = xj.y - xi.y;
        rz = xj.z - xi.z;
    }
    const float d = sqrtf(rx * rx + ry * ry + rz * rz);
    const float inv = 1.0f / d;
    // cosine cutoff, aimnet/ops.py:82-85
    const float dc = fminf(fmaxf(d, 1e-6f), aev.rc);
    float sn, cs;
    sincosf(dc * (kPi / aev.rc), &sn, &cs);
    PairEntry e;
    e.ux = rx * inv;
    e.uy = ry * inv;
    e.uz = rz * inv;
    e.d = d;
    e.fc = ok ? 0.5f * (cs + 1.0f) : 0.f;
    e.dfc = (kWithDeriv && ok && d > 1e-6f && d < aev.rc) ? -0.5f * (kPi / aev.rc) * sn : 0.f;
    e.j = ok ? k : (atom_ok ? il : 0);
    e.inv = inv;
    return e;
}

// channel pair ap (0..7) of a lane <-> channels a_lo = 4 (ap >> 1) + 2 (ap & 1), a_lo + 1: the (x,y) / (z,w) halves of the
// float4 of quad ap >> 1 in the gather layout
__device__ __forceinline__ int pair_lo(int ap) { return 4 * (ap >> 1) + 2 * (ap & 1); }

struct Layout {     // byte offsets into dynamic shared memory (host-computed)
    int tiles, agh, sv, xyz, q, dsq, bars, buf;   // buf is 1024-aligned
    int buf_bytes;    // one buffer (all boxes of one item)
    int nbuf;         // 1 or 2
    int total;
};

struct Params {
    int n_mol, n_atoms;
    int warps;          // warps per CTA
    int rows_box_a;     // rows (atom, quad) per box of the feature table
    int n_box_a;        // boxes per item
    int n_box_d;        // backward: boxes (256 rows of (atom, channel)) of the gradient table per item
    int max_seg;        // largest molecule of the batch
    Layout L;
};

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(512, 1) fwd_kernel(const __grid_constant__ CUtensorMap tmA, Params p,
                                                     const int32_t* __restrict__ mol_ptr, const float* __restrict__ coord,
                                                     AevParams aev, const float* __restrict__ q,
                                                     const float* __restrict__ agh_a, const float* __restrict__ agh_q,
                                                     float* __restrict__ x, int ldx, float* __restrict__ T_a,
                                                     float* __restrict__ T_q, int with_q) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    PairEntry* tiles = reinterpret_cast<PairEntry*>(smem + p.L.tiles);
    float* aghT_a = reinterpret_cast<float*>(smem + p.L.agh);   // [a][h][g], row stride kAghRow
    float* aghT_q = aghT_a + kA * kH * kAghRow;                 // [c][h][g]
    float* sv_all = reinterpret_cast<float*>(smem + p.L.sv);    // per warp: one atom's [a][k][g] (sv_off) + [c][k][g]
    float4* xyz = reinterpret_cast<float4*>(smem + p.L.xyz);
    float* q_s = reinterpret_cast<float*>(smem + p.L.q);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.L.bars);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, c = lane >> 4, g = lane & 15;
    const float shift_g = aev.shifts[g];
    for (int e = tid; e < kA * kG * kH; e += nthr) {
        int a = e / (kG * kH), gg = (e / kH) % kG, hh = e % kH;
        aghT_a[(a * kH + hh) * kAghRow + gg] = agh_a[e];
    }
    if (with_q)
        for (int e = tid; e < C * kG * kH; e += nthr) {
            int cc = e / (kG * kH), gg = (e / kH) % kG, hh = e % kH;
            aghT_q[(cc * kH + hh) * kAghRow + gg] = agh_q[e];
        }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tx_bytes = (uint32_t)p.L.buf_bytes;
    auto issue = [&](int m, int b) {   // one thread: fetch molecule m's block of the feature table into buffer b
        const int lo = mol_ptr[m];
        unsigned char* dst = smem + p.L.buf + b * p.L.buf_bytes;
        mbar_expect_tx(&full[b], tx_bytes);
        for (int k = 0; k < p.n_box_a; ++k)
            tma_load_2d(dst + (size_t)k * p.rows_box_a * 256, &tmA, &full[b], 0, 4 * lo + k * p.rows_box_a);
    };
    if (tid == 0)
        for (int b = 0; b < p.L.nbuf; ++b) {
            const int m = blockIdx.x + b * gridDim.x;
            if (m < p.n_mol) issue(m, b);
        }
    PairEntry* tile = tiles + warp * 32;
    float* svl = sv_all + warp * (kSvAtom + 2 * kSvRow);
    float* svql = svl + kSvAtom;
    int it = 0;
    for (int m = blockIdx.x; m < p.n_mol; m += gridDim.x, ++it) {
        const int b = it % p.L.nbuf;
        const uint32_t phase = (uint32_t)(it / p.L.nbuf) & 1u;
        const int lo = mol_ptr[m], n = mol_ptr[m + 1] - lo;
        for (int k = tid; k < n; k += nthr) {
            xyz[k] = make_float4(coord[3 * (lo + k)], coord[3 * (lo + k) + 1], coord[3 * (lo + k) + 2], 0.f);
            if (with_q)
                for (int cc = 0; cc < C; ++cc) q_s[k * C + cc] = q[(size_t)(lo + k) * C + cc];
        }
        mbar_wait(&full[b], phase);
        __syncthreads();
        const float4* abuf = reinterpret_cast<const float4*>(smem + p.L.buf + b * p.L.buf_bytes);   // [atom][quad][g] float4
        const int n_pairs = (n + 1) >> 1;
        for (int cp = warp; cp < n_pairs; cp += p.warps) {
            const int il = 2 * cp + c;
            const bool atom_ok = il < n;
            const int ilc = atom_ok ? il : 0;
            float2 S[8][4];   // [channel pair][d]
#pragma unroll
            for (int ap = 0; ap < 8; ++ap)
#pragma unroll
                for (int d = 0; d < 4; ++d) S[ap][d] = make_float2(0.f, 0.f);
            float Sq[C][4];
#pragma unroll
            for (int cc = 0; cc < C; ++cc)
#pragma unroll
                for (int d = 0; d < 4; ++d) Sq[cc][d] = 0.f;
            for (int k0 = 0; k0 < n; k0 += kSlots) {
                __syncwarp();
                tile[lane] = pair_entry<false>(ilc, atom_ok, k0 + (lane & 15), n, xyz, aev);
                __syncwarp();
                const int lim = min(kSlots, n - k0);
#pragma unroll 2
                for (int s = 0; s < lim; ++s) {
                    const PairEntry e = tile[c * kSlots + s];
                    const float4* row = abuf + e.j * 64 + g;
                    const float4 v0 = row[0], v1 = row[16], v2 = row[32], v3 = row[48];
                    const float xg = e.d - shift_g;
                    const float w0 = aev_exp(-aev.eta * xg * xg) * e.fc;
                    const float wv[4] = {w0, w0 * e.ux, w0 * e.uy, w0 * e.uz};
                    float2 wd[4];
#pragma unroll
                    for (int d = 0; d < 4; ++d) wd[d] = make_float2(wv[d], wv[d]);
                    const float2 av[8] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y),
                                          make_float2(v1.z, v1.w), make_float2(v2.x, v2.y), make_float2(v2.z, v2.w),
                                          make_float2(v3.x, v3.y), make_float2(v3.z, v3.w)};
#pragma unroll
                    for (int ap = 0; ap < 8; ++ap)
#pragma unroll
                        for (int d = 0; d < 4; ++d) S[ap][d] = ffma2(av[ap], wd[d], S[ap][d]);
                    if (with_q) {
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) {
                            const float qj = q_s[e.j * C + cc];
#pragma unroll
                            for (int d = 0; d < 4; ++d) Sq[cc][d] = fmaf(qj, wv[d], Sq[cc][d]);
                        }
                    }
                }
            }
            // ---- epilogue: scalar part and the atom's own features straight to x ----
            if (atom_ok) {
                const int i = lo + il;
                float* xr = x + (size_t)i * ldx;
                const float4* own = abuf + il * 64 + g;
                const float4 o0 = own[0], o1 = own[16], o2 = own[32], o3 = own[48];
                const float ov[kA] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w, o2.x, o2.y, o2.z, o2.w, o3.x, o3.y, o3.z, o3.w};
#pragma unroll
                for (int a = 0; a < kA; ++a) xr[a * kG + g] = ov[a];
#pragma unroll
                for (int ap = 0; ap < 8; ++ap) {
                    xr[kAG + pair_lo(ap) * kG + g] = S[ap][0].x;
                    xr[kAG + (pair_lo(ap) + 1) * kG + g] = S[ap][0].y;
                }
                int base = 2 * kAG + kAH;
                if (with_q) {
                    if (g < C) xr[base + g] = q_s[il * C + g];
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) xr[base + C + cc * kG + g] = Sq[cc][0];
                    base += C * (1 + kG + kH);
                }
                for (int cidx = base + g; cidx < ldx; cidx += kG) xr[cidx] = 0.f;
            }
            // ---- vector part through the warp's scratch, one centre after the other:
            //      T[a,h,k] = sum_g agh[a,g,h] * Sv[a,g,k]   (aimnet/modules/aev.py:188), mixed by the whole warp ----
#pragma unroll 1
            for (int cc2 = 0; cc2 < 2; ++cc2) {
                __syncwarp();
                if (c == cc2) {
#pragma unroll
                    for (int ap = 0; ap < 8; ++ap) {
                        const int o = sv_off(pair_lo(ap)) + g;   // channels a_lo, a_lo + 1 never straddle the pad between 7 and 8
                        svl[o] = S[ap][1].x;
                        svl[o + kG] = S[ap][2].x;
                        svl[o + 2 * kG] = S[ap][3].x;
                        svl[o + kSvRow] = S[ap][1].y;
                        svl[o + kSvRow + kG] = S[ap][2].y;
                        svl[o + kSvRow + 2 * kG] = S[ap][3].y;
                    }
                    if (with_q) {
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) {
                            svql[cc * kSvRow + g] = Sq[cc][1];
                            svql[cc * kSvRow + kG + g] = Sq[cc][2];
                            svql[cc * kSvRow + 2 * kG + g] = Sq[cc][3];
                        }
                    }
                }
                __syncwarp();
                if (2 * cp + cc2 >= n) break;
                const int ia = lo + 2 * cp + cc2;
                float* xr = x + (size_t)ia * ldx;
#pragma unroll 2
                for (int e = lane; e < kAH; e += 32) {
                    const int a = e / kH;
                    float t[3];
                    mix16(aghT_a + e * kAghRow, svl + sv_off(a), t);
                    float* To = T_a + (size_t)ia * kTA + e * 3;
                    To[0] = t[0];
                    To[1] = t[1];
                    To[2] = t[2];
                    xr[2 * kAG + e] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
                }
                if (with_q) {
                    const int base = 2 * kAG + kAH;
                    for (int e = lane; e < C * kH; e += 32) {
                        const int cq = e / kH;
                        float t[3];
                        mix16(aghT_q + e * kAghRow, svql + cq * kSvRow, t);
                        float* To = T_q + (size_t)ia * (C * kH * 3) + e * 3;
                        To[0] = t[0];
                        To[1] = t[1];
                        To[2] = t[2];
                        xr[base + C + C * kG + e] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
                    }
                }
            }
        }
        __syncthreads();   // every warp is done with buffer b, xyz and q_s
        if (tid == 0) {
            const int m2 = m + p.L.nbuf * gridDim.x;
            if (m2 < p.n_mol) issue(m2, b);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward: item = (molecule, quarter gq of the radial shifts); lane = (centre c8 of the warp's eight, g4), g = 4 gq + g4
//   grad_a[i,a,g]  = sum_j <dS[j,a,g,:], g_sv(j->i)[g,:]>            g_sv(j->i) = (gs, -gs u_{i->j})
//   F_i            = sum_j ( w(i->j) - w(j->i) )                       w = dE/dr of a pair
// ------------------------------------------------------------------------------------------------------------
template <int C, bool kGradA>
__global__ void __launch_bounds__(384, 1) bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD,
                                                     Params p, const int32_t* __restrict__ mol_ptr,
                                                     const float* __restrict__ coord, AevParams aev,
                                                     const float* __restrict__ q, const float* __restrict__ dS_q,
                                                     float* __restrict__ grad_a, float* __restrict__ gq_part,
                                                     float* __restrict__ f_part, int with_q) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    PairEntry* tiles = reinterpret_cast<PairEntry*>(smem + p.L.tiles);
    float4* xyz = reinterpret_cast<float4*>(smem + p.L.xyz);
    float* q_s = reinterpret_cast<float*>(smem + p.L.q);
    float4* dsq_s = reinterpret_cast<float4*>(smem + p.L.dsq);   // [atom][c][g4]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.L.bars);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, c8 = lane >> 2, g4 = lane & 3;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_items = 4 * p.n_mol;
    const int a_bytes = p.n_box_a * p.rows_box_a * 64;   // feature slice: rows (atom, quad) of 4 g x 4 a floats
    const uint32_t tx_bytes = (uint32_t)p.L.buf_bytes;
    auto issue = [&](int item, int b) {
        const int m = item >> 2, gq = item & 3;
        const int lo = mol_ptr[m];
        unsigned char* dst = smem + p.L.buf + b * p.L.buf_bytes;
        mbar_expect_tx(&full[b], tx_bytes);
        for (int k = 0; k < p.n_box_a; ++k)
            tma_load_2d(dst + (size_t)k * p.rows_box_a * 64, &tmA, &full[b], 16 * gq, 4 * lo + k * p.rows_box_a);
        for (int k = 0; k < p.n_box_d; ++k)
            tma_load_2d(dst + a_bytes + (size_t)k * 256 * 64, &tmD, &full[b], 16 * gq, 16 * lo + k * 256);
    };
    if (tid == 0)
        for (int b = 0; b < p.L.nbuf; ++b) {
            const int item = blockIdx.x + b * gridDim.x;
            if (item < n_items) issue(item, b);
        }
    PairEntry* tile = tiles + warp * 32;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int b = it % p.L.nbuf;
        const uint32_t phase = (uint32_t)(it / p.L.nbuf) & 1u;
        const int m = item >> 2, gq = item & 3;
        const int lo = mol_ptr[m], n = mol_ptr[m + 1] - lo;
        const int g = 4 * gq + g4;
        const float shift_g = aev.shifts[g];
        for (int k = tid; k < n; k += nthr) {
            xyz[k] = make_float4(coord[3 * (lo + k)], coord[3 * (lo + k) + 1], coord[3 * (lo + k) + 2], 0.f);
            if (with_q)
                for (int cc = 0; cc < C; ++cc) q_s[k * C + cc] = q[(size_t)(lo + k) * C + cc];
        }
        if (with_q)
            for (int e = tid; e < n * C * 4; e += nthr) {   // dS_q (N, C, 16 g) float4 -> this quarter's [atom][c][g4]
                const int k = e / (C * 4), r = e % (C * 4), cc = r >> 2, gg = r & 3;
                dsq_s[e] = reinterpret_cast<const float4*>(dS_q)[(size_t)(lo + k) * (C * kG) + cc * kG + 4 * gq + gg];
            }
        mbar_wait(&full[b], phase);
        __syncthreads();
        const float4* abuf = reinterpret_cast<const float4*>(smem + p.L.buf + b * p.L.buf_bytes);   // [atom][quad][g4] float4
        const float4* dbuf = reinterpret_cast<const float4*>(smem + p.L.buf + b * p.L.buf_bytes + a_bytes);   // [atom][a][g4] float4
        for (int base = 0; base < n; base += 8 * p.warps) {
            const int il = base + warp * 8 + c8;
            const bool atom_ok = il < n;
            const int ilc = atom_ok ? il : 0;
            // own atom: dS_i[a][g][:] as (scalar,x) / (y,z) register pairs and a_i[a][g] for all 16 channels
            float2 dSi01[kA], dSi23[kA];
            float ai[kA];
#pragma unroll
            for (int a = 0; a < kA; ++a) {
                const float4 v = dbuf[(ilc * 16 + a) * 4 + g4];
                dSi01[a] = make_float2(v.x, v.y);
                dSi23[a] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const float4 o = abuf[(ilc * 4 + qd) * 4 + g4];
                ai[4 * qd + 0] = o.x;
                ai[4 * qd + 1] = o.y;
                ai[4 * qd + 2] = o.z;
                ai[4 * qd + 3] = o.w;
            }
            float2 dSqi01[C], dSqi23[C];
            float qi[C];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                const float4 v = with_q ? dsq_s[(ilc * C + cc) * 4 + g4] : make_float4(0, 0, 0, 0);
                dSqi01[cc] = make_float2(v.x, v.y);
                dSqi23[cc] = make_float2(v.z, v.w);
                qi[cc] = with_q ? q_s[ilc * C + cc] : 0.f;
            }
            float2 ga2[kA];
#pragma unroll
            for (int a = 0; a < kA; ++a) ga2[a] = make_float2(0.f, 0.f);
            float2 gq2[C];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) gq2[cc] = make_float2(0.f, 0.f);
            float fx = 0.f, fy = 0.f, fz = 0.f;
            for (int k0 = 0; k0 < n; k0 += 4) {   // four neighbours per round: lane = (centre, slot) stages one pair
                __syncwarp();
                tile[lane] = pair_entry<true>(ilc, atom_ok, k0 + g4, n, xyz, aev);
                __syncwarp();
                const int lim = min(4, n - k0);
                for (int s = 0; s < lim; ++s) {
                    const PairEntry e = tile[c8 * 4 + s];
                    const float4* arow = abuf + e.j * 16 + g4;
                    const float4* drow = dbuf + e.j * 64 + g4;
                    const float xg = e.d - shift_g;
                    const float ex = aev_exp(-aev.eta * xg * xg);
                    const float gs = ex * e.fc;
                    const float dgs = ex * (e.dfc - 2.0f * aev.eta * xg * e.fc);
                    // g_sv(j->i)[g,:] = (gs, -gs u): grad_a[i] += <dS[j], g_sv(j->i)> as two packed FMAs per channel
                    const float2 G01 = make_float2(gs, -gs * e.ux), G23 = make_float2(-gs * e.uy, -gs * e.uz);
                    // p = contraction for the pair (i -> j), r = for the reverse pair (j -> i), over all 16 channels of (i, g)
                    float2 p01 = make_float2(0.f, 0.f), p23 = p01, r01 = p01, r23 = p01;
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const float4 av = arow[4 * qd];
                        const float aj[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int a = 4 * qd + k;
                            const float4 dj = drow[4 * a];
                            const float2 dj01 = make_float2(dj.x, dj.y), dj23 = make_float2(dj.z, dj.w);
                            if (kGradA) {
                                ga2[a] = ffma2(dj01, G01, ga2[a]);
                                ga2[a] = ffma2(dj23, G23, ga2[a]);
                            }
                            p01 = ffma2s(aj[k], dSi01[a], p01);
                            p23 = ffma2s(aj[k], dSi23[a], p23);
                            r01 = ffma2s(ai[a], dj01, r01);
                            r23 = ffma2s(ai[a], dj23, r23);
                        }
                    }
                    if (with_q) {
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) {
                            const float qj = q_s[e.j * C + cc];
                            const float4 dqj = dsq_s[(e.j * C + cc) * 4 + g4];
                            const float2 dq01 = make_float2(dqj.x, dqj.y), dq23 = make_float2(dqj.z, dqj.w);
                            if (kGradA) {
                                gq2[cc] = ffma2(dq01, G01, gq2[cc]);
                                gq2[cc] = ffma2(dq23, G23, gq2[cc]);
                            }
                            p01 = ffma2s(qj, dSqi01[cc], p01);
                            p23 = ffma2s(qj, dSqi23[cc], p23);
                            r01 = ffma2s(qi[cc], dq01, r01);
                            r23 = ffma2s(qi[cc], dq23, r23);
                        }
                    }
                    const float gsi = gs * e.inv;
                    // this thread's share of w(i->j) = u (A + C.u) + (B - u (B.u))/d
                    const float pu = p01.y * e.ux + p23.x * e.uy + p23.y * e.uz;
                    const float sc = (p01.x + pu) * dgs - pu * gsi;
                    // reverse pair (j->i): u' = -u;  w' = -u (A' - (r.u) dgs) + (B' - u (B'.u))/d
                    const float ru = r01.y * e.ux + r23.x * e.uy + r23.y * e.uz;
                    const float scr = (ru - r01.x) * dgs - ru * gsi;
                    // F_i += w - w' = u (sc - scr) + (B - B') gs/d
                    const float ds = sc - scr;
                    fx += fmaf(e.ux, ds, (p01.y - r01.y) * gsi);
                    fy += fmaf(e.uy, ds, (p23.x - r23.x) * gsi);
                    fz += fmaf(e.uz, ds, (p23.y - r23.y) * gsi);
                }
            }
            // reduce the force / grad_q shares over the four radial shifts of this quarter
            auto quad_sum = [](float v) {
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                return v;
            };
            fx = quad_sum(fx);
            fy = quad_sum(fy);
            fz = quad_sum(fz);
            float gqs[C];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) gqs[cc] = (kGradA && with_q) ? quad_sum(gq2[cc].x + gq2[cc].y) : 0.f;
            if (atom_ok) {
                const int i = lo + il;
                if (kGradA) {
#pragma unroll
                    for (int a = 0; a < kA; ++a) grad_a[(size_t)i * kAG + a * kG + g] = ga2[a].x + ga2[a].y;
                }
                if (g4 == 0) {
                    float* fp = f_part + ((size_t)gq * p.n_atoms + i) * 3;
                    fp[0] = fx;
                    fp[1] = fy;
                    fp[2] = fz;
                    if (kGradA && with_q) {
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) gq_part[((size_t)gq * p.n_atoms + i) * C + cc] = gqs[cc];
                    }
                }
            }
        }
        __syncthreads();   // every warp is done with buffer b, xyz, q_s, dsq_s
        if (tid == 0) {
            const int item2 = item + p.L.nbuf * gridDim.x;
            if (item2 < n_items) issue(item2, b);
        }
    }
}

// forces[i] += sum over the four quarters (fixed order); grad_q[i, c] = sum over the quarters
__global__ void combine_kernel(int n_atoms, int C, const float* __restrict__ f_part, const float* __restrict__ gq_part,
                               float* __restrict__ forces, float* __restrict__ grad_q) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3 * n_atoms) {
        float v = f_part[t];
        v += f_part[(size_t)3 * n_atoms + t];
        v += f_part[(size_t)6 * n_atoms + t];
        v += f_part[(size_t)9 * n_atoms + t];
        forces[t] += v;
    }
    if (grad_q != nullptr && t < C * n_atoms) {
        const size_t s = (size_t)C * n_atoms;
        grad_q[t] = ((gq_part[t] + gq_part[s + t]) + gq_part[2 * s + t]) + gq_part[3 * s + t];
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    });
    return fn;
}

// fp32 matrix (rows, 64 columns), row pitch 256 B; box = (box_cols, box_rows), no swizzle
static int make_map(CUtensorMap* m, const float* ptr, long long rows, int box_cols, int box_rows) {
    EncodeFn enc = get_encode();
    if (!enc) {
        set_error("conv_dense: cuTensorMapEncodeTiled not available");
        return AIMNET_ECUDA;
    }
    cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {256};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv_dense: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return AIMNET_ECUDA;
    }
    return AIMNET_OK;
}

constexpr int kSmemBudget = 227 * 1024 - 1024;   // dynamic shared memory per CTA, minus the alignment slack

static int up(int v, int a) { return (v + a - 1) / a * a; }

// boxes of the feature table for molecules of at most max_seg atoms: at most 256 rows (= 64 atoms) each
static void feature_boxes(int max_seg, int& rows_box, int& n_box) {
    n_box = (max_seg + 63) / 64;
    const int atoms_box = (max_seg + n_box - 1) / n_box;
    rows_box = 4 * atoms_box;
}

static bool plan_fwd(int C, int max_seg, Params& p) {
    feature_boxes(max_seg, p.rows_box_a, p.n_box_a);
    const int n_pairs = (max_seg + 1) / 2;
    const int sweeps = (n_pairs + 15) / 16;
    p.warps = (n_pairs + sweeps - 1) / sweeps;   // <= 16
    Layout& L = p.L;
    int o = 0;
    L.tiles = o, o += p.warps * 32 * (int)sizeof(PairEntry);
    L.agh = o, o += (kA + 2) * kH * kAghRow * 4;
    L.sv = o, o += p.warps * (kSvAtom + 2 * kSvRow) * 4;
    L.xyz = o = up(o, 16), o += max_seg * 16;
    L.q = o, o += up(max_seg * C * 4, 16);
    L.dsq = o;
    L.bars = o = up(o, 16), o += 64;
    L.buf = o = up(o, 1024);
    L.buf_bytes = p.n_box_a * p.rows_box_a * 256;
    L.nbuf = (L.buf + 2 * L.buf_bytes <= kSmemBudget) ? 2 : 1;
    L.total = L.buf + L.nbuf * L.buf_bytes + 1024;
    return L.buf + L.buf_bytes <= kSmemBudget;
}

static bool plan_bwd(int C, int max_seg, Params& p) {
    feature_boxes(max_seg, p.rows_box_a, p.n_box_a);
    p.n_box_d = (16 * max_seg + 255) / 256;
    p.warps = std::min(12, (max_seg + 7) / 8);
    Layout& L = p.L;
    int o = 0;
    L.tiles = o, o += p.warps * 32 * (int)sizeof(PairEntry);
    L.agh = L.sv = o;
    L.xyz = o = up(o, 16), o += max_seg * 16;
    L.q = o, o += up(max_seg * C * 4, 16);
    L.dsq = o, o += max_seg * C * 4 * 16;
    L.bars = o = up(o, 16), o += 64;
    L.buf = o = up(o, 1024);
    L.buf_bytes = p.n_box_a * p.rows_box_a * 64 + p.n_box_d * 256 * 64;
    L.nbuf = (L.buf + 2 * L.buf_bytes <= kSmemBudget) ? 2 : 1;
    L.total = L.buf + L.nbuf * L.buf_bytes + 1024;
    return L.buf + L.buf_bytes <= kSmemBudget;
}

}  // namespace convd

// largest molecule the dense walk can stage (both directions) for C charge channels
int conv_dense_max_atoms(int C) {
    static int cached[3] = {0, 0, 0};
    if (C < 1 || C > 2) return 0;
    if (cached[C]) return cached[C];
    int best = 0;
    for (int n = 2; n <= 256; ++n) {
        convd::Params p{};
        if (convd::plan_fwd(C, n, p) && convd::plan_bwd(C, n, p)) best = n;
    }
    return cached[C] = best;
}

template <int C>
static int conv_dense_fwd_launch(int n_atoms, int n_mol, int max_seg, const int32_t* mol_ptr, const float* coord,
                                 const AevParams& aev, const float* aT, const float* q, const float* agh_a,
                                 const float* agh_q, float* x, int ldx, float* T_a, float* T_q, int with_q, cudaStream_t st) {
    using namespace convd;
    Params p{};
    p.n_mol = n_mol, p.n_atoms = n_atoms, p.max_seg = max_seg;
    AIM_REQUIRE(plan_fwd(C, max_seg, p), "conv_dense_fwd: molecule too large for the shared-memory walk");
    static bool configured_dev[kMaxDevices] = {};
    static int sms_dev[kMaxDevices] = {};
    const int dslot = current_device_slot();
    if (!configured_dev[dslot]) {
        AIM_CUDA_CHECK(cudaFuncSetAttribute(fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int dev = 0;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&sms_dev[dslot], cudaDevAttrMultiProcessorCount, dev));
        configured_dev[dslot] = true;
    }
    CUtensorMap tmA;
    AIM_TRY(make_map(&tmA, aT, 4LL * n_atoms, 64, p.rows_box_a));
    const int grid = std::min(n_mol, sms_dev[dslot]);
    fwd_kernel<C><<<grid, 32 * p.warps, p.L.total, st>>>(tmA, p, mol_ptr, coord, aev, q, agh_a, agh_q, x, ldx, T_a, T_q, with_q);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

int launch_conv_dense_fwd(int C, int n_atoms, int n_mol, int max_seg, const int32_t* mol_ptr, const float* coord,
                          const AevParams& aev, const float* aT, const float* q, const float* agh_a, const float* agh_q,
                          float* x, int ldx, float* T_a, float* T_q, int with_q, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    if (C == 1)
        return conv_dense_fwd_launch<1>(n_atoms, n_mol, max_seg, mol_ptr, coord, aev, aT, q, agh_a, agh_q, x, ldx, T_a, T_q, with_q, st);
    return conv_dense_fwd_launch<2>(n_atoms, n_mol, max_seg, mol_ptr, coord, aev, aT, q, agh_a, agh_q, x, ldx, T_a, T_q, with_q, st);
}

template <int C>
static int conv_dense_bwd_launch(int n_atoms, int n_mol, int max_seg, const int32_t* mol_ptr, const float* coord,
                                 const AevParams& aev, const float* aT, const float* q, const float* dS_a, const float* dS_q,
                                 float* grad_a, float* grad_q, float* forces, float* f_part, float* gq_part, int with_q,
                                 int want_grad_a, cudaStream_t st) {
    using namespace convd;
    Params p{};
    p.n_mol = n_mol, p.n_atoms = n_atoms, p.max_seg = max_seg;
    AIM_REQUIRE(plan_bwd(C, max_seg, p), "conv_dense_bwd: molecule too large for the shared-memory walk");
    static bool configured_dev[kMaxDevices] = {};
    static int sms_dev[kMaxDevices] = {};
    const int dslot = current_device_slot();
    if (!configured_dev[dslot]) {
        AIM_CUDA_CHECK(cudaFuncSetAttribute(bwd_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AIM_CUDA_CHECK(cudaFuncSetAttribute(bwd_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int dev = 0;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&sms_dev[dslot], cudaDevAttrMultiProcessorCount, dev));
        configured_dev[dslot] = true;
    }
    CUtensorMap tmA, tmD;
    AIM_TRY(make_map(&tmA, aT, 4LL * n_atoms, 16, p.rows_box_a));
    AIM_TRY(make_map(&tmD, dS_a, 16LL * n_atoms, 16, 256));
    const int grid = std::min(4 * n_mol, sms_dev[dslot]);
    if (want_grad_a)
        bwd_kernel<C, true><<<grid, 32 * p.warps, p.L.total, st>>>(tmA, tmD, p, mol_ptr, coord, aev, q, dS_q, grad_a, gq_part, f_part, with_q);
    else
        bwd_kernel<C, false><<<grid, 32 * p.warps, p.L.total, st>>>(tmA, tmD, p, mol_ptr, coord, aev, q, dS_q, grad_a, gq_part, f_part, with_q);
    AIM_LAUNCH_CHECK();
    const int total = std::max(3, C) * n_atoms;
    combine_kernel<<<(total + 255) / 256, 256, 0, st>>>(n_atoms, C, f_part, gq_part, forces, (want_grad_a && with_q) ? grad_q : nullptr);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

// the gather half of the backward pass (dS_a / dS_q must have been produced by conv_bwd_prep); f_part (4, N, 3) and
// gq_part (4, N, C) are scratch; `forces` is accumulated with a plain += by the combining kernel
int launch_conv_dense_bwd_gather(int C, int n_atoms, int n_mol, int max_seg, const int32_t* mol_ptr, const float* coord,
                                 const AevParams& aev, const float* aT, const float* q, const float* dS_a, const float* dS_q,
                                 float* grad_a, float* grad_q, float* forces, float* f_part, float* gq_part, int with_q,
                                 int want_grad_a, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    if (C == 1)
        return conv_dense_bwd_launch<1>(n_atoms, n_mol, max_seg, mol_ptr, coord, aev, aT, q, dS_a, dS_q, grad_a, grad_q, forces,
                                        f_part, gq_part, with_q, want_grad_a, st);
    return conv_dense_bwd_launch<2>(n_atoms, n_mol, max_seg, mol_ptr, coord, aev, aT, q, dS_a, dS_q, grad_a, grad_q, forces, f_part,
                                    gq_part, with_q, want_grad_a, st);
}

}  // namespace aimnet
