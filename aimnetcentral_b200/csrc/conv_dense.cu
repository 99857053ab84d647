// AEV + conv_sv message passing for batches of small molecules: the molecule's feature tables are staged into shared memory
// with TMA and every pair of the molecule is walked from there ("dense walk").  Forward (the engine's default for such
// batches) and analytic backward (a measured alternative, slower than the list backward: see below).
//
// Same arithmetic and reference semantics as conv.cu (calc_distances aimnet/ops.py:37-66, AEVSV._calc_aev
// aimnet/modules/aev.py:94-110, ConvSV.forward aimnet/modules/aev.py:156-189, Warp kernels
// aimnet/kernels/conv_sv_2d_sp_wp.py:90-164).  Why a second implementation: the list kernels of conv.cu gather every
// neighbour row (1 KB of a[j], 4 KB of dS[j]) through L1 / L2 and wait for those gathers (ncu: long-scoreboard stalls 5.5 /
// 3.5 per issued instruction, FMA pipe 35 / 45 %).  For a molecule of n <= ~100 atoms everything a CTA gathers is one
// contiguous block of the feature table (atoms of a molecule are contiguous), so:
//
//   * one CTA works on one molecule (forward) or one (molecule, quarter of the radial shifts) item (backward) at a time;
//     a single elected thread fetches the block with cp.async.bulk.tensor (2-D tensor maps over the (N, 4 q, 16 g, 4 a) feature
//     table and the (N, 16 a, 16 g, 4 d) gradient table; the box selects the item's radial-shift columns), double-buffered
//     where it fits: the next item's block lands while this one is computed, completion through an mbarrier;
//   * every centre atom walks ALL atoms of its molecule in index order (pairs beyond the cutoff contribute exactly zero:
//     fc(d >= rc) == 0; the order equals that of the canonical sorted list rows): no global gather is left in the pair loop;
//   * what the shared-memory pipe charges is one wavefront per 128 bytes per quarter-warp WHATEVER the addresses are
//     (a 16-byte warp load is four wavefronts even when all lanes read the same row: measured, profiles/r2g_convd_*: a first
//     version whose lanes shared neighbour rows ran at 89 % of the shared-memory pipe with 44 % of its wavefronts bank
//     conflicts).  Reuse therefore has to happen in registers: every THREAD handles TWO centre atoms, so each 16-byte read
//     of a neighbour row feeds twice the FMAs, and every lane of a warp reads a distinct address (pair geometry is staged in a
//     small per-warp table read back as 8-byte broadcasts);
//   * measured on 1024 x 50 atoms (profiles/r2k_convd_*): forward 242 us vs 279 us for the list kernel.  Backward 1 012 us vs
//     646 us: both execute at ~40 % issue utilisation, and the dense backward executes 451 M warp instructions against 300 M
//     (walking all 50 atoms instead of the 42 inside the cutoff, per-pair force algebra that does not amortise over more
//     channels at 168 registers per thread, 7 warps per SM).  The engine therefore pairs the dense forward with the list
//     backward; the dense backward stays selectable (conv_impl 2) and tested.
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
constexpr int kSvRow = 52;
constexpr int kSvAtom = kA * kSvRow + 16;
__device__ __forceinline__ int sv_off(int a) { return a * kSvRow + ((a >> 3) << 4); }

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

// geometry + cutoff of the pair (centre il, neighbour k) of one molecule, coordinates from shared memory
struct PairGeom {
    float ux, uy, uz, d, fc, dfc, inv;
};
template <bool kWithDeriv>
__device__ __forceinline__ PairGeom pair_geom(int il, bool atom_ok, int k, int n, const float4* __restrict__ xyz,
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
    PairGeom e;
    e.ux = rx * inv;
    e.uy = ry * inv;
    e.uz = rz * inv;
    e.d = d;
    e.fc = ok ? 0.5f * (cs + 1.0f) : 0.f;
    e.dfc = (kWithDeriv && ok && d > 1e-6f && d < aev.rc) ? -0.5f * (kPi / aev.rc) * sn : 0.f;
    e.inv = inv;
    return e;
}

struct Layout {     // byte offsets into dynamic shared memory (host-computed)
    int ent, agh, aghq, sv, xyz, q, dsq, bars, buf;   // buf is 1024-aligned
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

constexpr int kAghPad = 196;         // agh[a] row (16 g x 12 h = 192 floats) padded: 8 consecutive rows hit disjoint banks
constexpr int kEntFwd = 5 * 16 * 2;  // forward pair table of a warp: [field d, fc, ux, uy, uz][16 slots][2 centres]
constexpr int kEntBwd = 7 * 4 * 8;   // backward: [field d, fc, ux, uy, uz, dfc, inv][4 slots][8 centres]

// ------------------------------------------------------------------------------------------------------------
// forward: one CTA = one molecule at a time; one warp = one PAIR of centre atoms, lane = (channel half h, radial shift g);
// every THREAD accumulates both centres of the pair, so each neighbour row read from shared memory (2 x 16 bytes per lane,
// every lane a distinct address: full shared-memory bandwidth) feeds 64 FMAs.
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(512, 1) fwd_kernel(const __grid_constant__ CUtensorMap tmA, Params p,
                                                     const int32_t* __restrict__ mol_ptr, const float* __restrict__ coord,
                                                     AevParams aev, const float* __restrict__ q,
                                                     const float* __restrict__ agh_a, const float* __restrict__ agh_q,
                                                     float* __restrict__ x, int ldx, float* __restrict__ T_a,
                                                     float* __restrict__ T_q, int with_q) {
    extern __shared__ unsigned char smem_dyn[];
    // 1 KB alignment by pointer arithmetic on the __shared__ array (an integer round trip would make every access a generic
    // LD / ST instead of LDS / STS)
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    float* ent_all = reinterpret_cast<float*>(smem + p.L.ent);
    float* aghS = reinterpret_cast<float*>(smem + p.L.agh);     // [a][g][h], row stride kAghPad
    float* aghqS = reinterpret_cast<float*>(smem + p.L.aghq);   // [c][g][h]
    float* sv_all = reinterpret_cast<float*>(smem + p.L.sv);    // per warp: one atom's [a][k][g] (sv_off) + [c][k][g]
    float4* xyz = reinterpret_cast<float4*>(smem + p.L.xyz);
    float* q_s = reinterpret_cast<float*>(smem + p.L.q);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.L.bars);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, h = lane >> 4, g = lane & 15;
    const float shift_g = aev.shifts[g];
    for (int e = tid; e < kA * kG * kH; e += nthr) aghS[(e / (kG * kH)) * kAghPad + e % (kG * kH)] = agh_a[e];
    if (with_q)
        for (int e = tid; e < C * kG * kH; e += nthr) aghqS[e] = agh_q[e];
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
    float* ent = ent_all + warp * kEntFwd;
    const float2* ent2 = reinterpret_cast<const float2*>(ent);
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
            const int il0 = 2 * cp;
            const bool ok1 = il0 + 1 < n;
            // S[c][channel pair][d]: channel pairs of this half are the (x,y) / (z,w) halves of quads 2h, 2h + 1
            float2 S[2][4][4];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int ap = 0; ap < 4; ++ap)
#pragma unroll
                    for (int d = 0; d < 4; ++d) S[c][ap][d] = make_float2(0.f, 0.f);
            // charge channels: with one channel both halves accumulate it (the h == 0 half stores it); with two (NSE) the half h
            // accumulates and stores channel h: half the FMAs per lane, identical instructions in both halves
            constexpr int CQ = (C == 2) ? 1 : C;
            const int cm = (C == 2) ? h : 0;
            float Sq[2][CQ][4];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int cc = 0; cc < CQ; ++cc)
#pragma unroll
                    for (int d = 0; d < 4; ++d) Sq[c][cc][d] = 0.f;
            for (int k0 = 0; k0 < n; k0 += kSlots) {
                __syncwarp();
                {   // lane = (centre, slot) stages one pair of the warp's table
                    const int c = lane >> 4, sl = lane & 15;
                    const PairGeom e = pair_geom<false>(min(il0 + c, n - 1), il0 + c < n, k0 + sl, n, xyz, aev);
                    ent[(0 * kSlots + sl) * 2 + c] = e.d;
                    ent[(1 * kSlots + sl) * 2 + c] = e.fc;
                    ent[(2 * kSlots + sl) * 2 + c] = e.ux;
                    ent[(3 * kSlots + sl) * 2 + c] = e.uy;
                    ent[(4 * kSlots + sl) * 2 + c] = e.uz;
                }
                __syncwarp();
                const int lim = min(kSlots, n - k0);
#pragma unroll 2
                for (int s = 0; s < lim; ++s) {
                    const float4* row = abuf + (k0 + s) * 64 + 32 * h + g;
                    const float4 v0 = row[0], v1 = row[16];
                    const float2 e_d = ent2[0 * kSlots + s], e_fc = ent2[1 * kSlots + s], e_ux = ent2[2 * kSlots + s],
                                 e_uy = ent2[3 * kSlots + s], e_uz = ent2[4 * kSlots + s];
                    const float2 av[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y),
                                          make_float2(v1.z, v1.w)};
                    const float dd[2] = {e_d.x, e_d.y}, fcv[2] = {e_fc.x, e_fc.y}, uxv[2] = {e_ux.x, e_ux.y},
                                uyv[2] = {e_uy.x, e_uy.y}, uzv[2] = {e_uz.x, e_uz.y};
                    float qj[CQ];
                    if (with_q) {
#pragma unroll
                        for (int cc = 0; cc < CQ; ++cc) qj[cc] = q_s[(k0 + s) * C + cm + cc];
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float xg = dd[c] - shift_g;
                        const float w0 = aev_exp(-aev.eta * xg * xg) * fcv[c];
                        const float wv[4] = {w0, w0 * uxv[c], w0 * uyv[c], w0 * uzv[c]};
                        float2 wd[4];
#pragma unroll
                        for (int d = 0; d < 4; ++d) wd[d] = make_float2(wv[d], wv[d]);
#pragma unroll
                        for (int ap = 0; ap < 4; ++ap)
#pragma unroll
                            for (int d = 0; d < 4; ++d) S[c][ap][d] = ffma2(av[ap], wd[d], S[c][ap][d]);
                        if (with_q) {
#pragma unroll
                            for (int cc = 0; cc < CQ; ++cc)
#pragma unroll
                                for (int d = 0; d < 4; ++d) Sq[c][cc][d] = fmaf(qj[cc], wv[d], Sq[c][cc][d]);
                        }
                    }
                }
            }
            // ---- epilogue, one centre after the other: scalar part and the atom's own features straight to x, vector
            //      part through the warp's scratch for T[a,h,k] = sum_g agh[a,g,h] * Sv[a,g,k]   (aimnet/modules/aev.py:188)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c == 1 && !ok1) break;
                const int il = il0 + c, i = lo + il;
                float* xr = x + (size_t)i * ldx;
                __syncwarp();   // the previous centre's mixing is done with the scratch
                {
                    const float4* own = abuf + il * 64 + 32 * h + g;
                    const float4 o0 = own[0], o1 = own[16];
                    const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) xr[(8 * h + k) * kG + g] = ov[k];
#pragma unroll
                    for (int ap = 0; ap < 4; ++ap) {
                        const int a = 8 * h + 2 * ap;
                        xr[kAG + a * kG + g] = S[c][ap][0].x;
                        xr[kAG + (a + 1) * kG + g] = S[c][ap][0].y;
                        const int o = sv_off(a) + g;   // channels a, a + 1 never straddle the pad between 7 and 8
                        svl[o] = S[c][ap][1].x;
                        svl[o + kG] = S[c][ap][2].x;
                        svl[o + 2 * kG] = S[c][ap][3].x;
                        svl[o + kSvRow] = S[c][ap][1].y;
                        svl[o + kSvRow + kG] = S[c][ap][2].y;
                        svl[o + kSvRow + 2 * kG] = S[c][ap][3].y;
                    }
                    int base = 2 * kAG + kAH;
                    if (with_q) {
                        if (h == 0 && g < C) xr[base + g] = q_s[il * C + g];
                        if (C == 2 || h == 0) {
#pragma unroll
                            for (int cc = 0; cc < CQ; ++cc) {
                                xr[base + C + (cm + cc) * kG + g] = Sq[c][cc][0];
                                svql[(cm + cc) * kSvRow + g] = Sq[c][cc][1];
                                svql[(cm + cc) * kSvRow + kG + g] = Sq[c][cc][2];
                                svql[(cm + cc) * kSvRow + 2 * kG + g] = Sq[c][cc][3];
                            }
                        }
                        base += C * (1 + kG + kH);
                    }
                    for (int cidx = base + lane; cidx < ldx; cidx += 32) xr[cidx] = 0.f;
                }
                __syncwarp();
                {   // lane = (channel a, half hh of the 12 mixed components): 6 x 3 outputs from 16 g
                    const int a = lane >> 1, hh = lane & 1;
                    float t[6][3];
#pragma unroll
                    for (int jj = 0; jj < 6; ++jj) t[jj][0] = t[jj][1] = t[jj][2] = 0.f;
                    const float* wrow = aghS + a * kAghPad + 6 * hh;
                    const float* srow = svl + sv_off(a);
#pragma unroll
                    for (int g0 = 0; g0 < kG; g0 += 4) {
                        const float4 s0 = *reinterpret_cast<const float4*>(srow + g0);
                        const float4 s1 = *reinterpret_cast<const float4*>(srow + kG + g0);
                        const float4 s2 = *reinterpret_cast<const float4*>(srow + 2 * kG + g0);
                        const float sk[4][3] = {{s0.x, s1.x, s2.x}, {s0.y, s1.y, s2.y}, {s0.z, s1.z, s2.z}, {s0.w, s1.w, s2.w}};
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            const float2* w2 = reinterpret_cast<const float2*>(wrow + (g0 + gg) * kH);
                            const float2 wa = w2[0], wb = w2[1], wc = w2[2];
                            const float wv[6] = {wa.x, wa.y, wb.x, wb.y, wc.x, wc.y};
#pragma unroll
                            for (int jj = 0; jj < 6; ++jj)
#pragma unroll
                                for (int k = 0; k < 3; ++k) t[jj][k] = fmaf(wv[jj], sk[gg][k], t[jj][k]);
                        }
                    }
                    float2* To = reinterpret_cast<float2*>(T_a + (size_t)i * kTA + a * (kH * 3) + 18 * hh);
#pragma unroll
                    for (int u = 0; u < 9; ++u) To[u] = make_float2(t[(2 * u) / 3][(2 * u) % 3], t[(2 * u + 1) / 3][(2 * u + 1) % 3]);
                    float2* xo = reinterpret_cast<float2*>(xr + 2 * kAG + a * kH + 6 * hh);
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const float n0 = t[2 * u][0] * t[2 * u][0] + t[2 * u][1] * t[2 * u][1] + t[2 * u][2] * t[2 * u][2];
                        const float n1 = t[2 * u + 1][0] * t[2 * u + 1][0] + t[2 * u + 1][1] * t[2 * u + 1][1] + t[2 * u + 1][2] * t[2 * u + 1][2];
                        xo[u] = make_float2(n0, n1);
                    }
                }
                if (with_q && lane < C * kH) {   // charge channels: lane = (c, mixed component)
                    const int cq = lane / kH, hq = lane % kH;
                    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
                    for (int gg = 0; gg < kG; ++gg) {
                        const float wq = aghqS[(cq * kG + gg) * kH + hq];
                        t0 = fmaf(wq, svql[cq * kSvRow + gg], t0);
                        t1 = fmaf(wq, svql[cq * kSvRow + kG + gg], t1);
                        t2 = fmaf(wq, svql[cq * kSvRow + 2 * kG + gg], t2);
                    }
                    float* To = T_q + (size_t)i * (C * kH * 3) + lane * 3;
                    To[0] = t0;
                    To[1] = t1;
                    To[2] = t2;
                    xr[2 * kAG + kAH + C + C * kG + lane] = t0 * t0 + t1 * t1 + t2 * t2;
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
// backward: item = (molecule, quarter gq of the radial shifts); lane = (centre pair cp of the warp's four, channel half h,
// g4), g = 4 gq + g4; every THREAD handles both centres of its pair for 8 channels: a neighbour's rows (a[j]: 2 x 16 bytes,
// dS[j]: 8 x 16 bytes per lane) feed 192 FMAs.  Channel half h owns channels 4h..4h+3 and 8+4h..8+4h+3 (feature quads h and
// h + 2); the gradient table arrives with its channel rows permuted (conv_bwd_prep, permute = 1) so that row 2 t + h is
// the t-th channel of half h: the two halves read neighbouring rows, all eight lanes of a quarter-warp distinct banks.
//   grad_a[i,a,g]  = sum_j <dS[j,a,g,:], g_sv(j->i)[g,:]>            g_sv(j->i) = (gs, -gs u_{i->j})
//   F_i            = sum_j ( w(i->j) - w(j->i) )                       w = dE/dr of a pair
// ------------------------------------------------------------------------------------------------------------
template <int C, bool kGradA>
__global__ void __launch_bounds__(320, 1) bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD,
                                                     Params p, const int32_t* __restrict__ mol_ptr,
                                                     const float* __restrict__ coord, AevParams aev,
                                                     const float* __restrict__ q, const float* __restrict__ dS_q,
                                                     float* __restrict__ grad_a, float* __restrict__ gq_part,
                                                     float* __restrict__ f_part, int with_q) {
    extern __shared__ unsigned char smem_dyn[];
    // 1 KB alignment by pointer arithmetic on the __shared__ array (an integer round trip would make every access a generic
    // LD / ST instead of LDS / STS)
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    float* ent_all = reinterpret_cast<float*>(smem + p.L.ent);
    float4* xyz = reinterpret_cast<float4*>(smem + p.L.xyz);
    float* q_s = reinterpret_cast<float*>(smem + p.L.q);
    float4* dsq_s = reinterpret_cast<float4*>(smem + p.L.dsq);   // [atom][c][g4]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.L.bars);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, cp = lane >> 3, h = (lane >> 2) & 1, g4 = lane & 3;
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
    float* ent = ent_all + warp * kEntBwd;
    const float2* ent2 = reinterpret_cast<const float2*>(ent);
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
        const float4* abuf = reinterpret_cast<const float4*>(smem + p.L.buf + b * p.L.buf_bytes);            // [atom][quad][g4]
        const float4* dbuf = reinterpret_cast<const float4*>(smem + p.L.buf + b * p.L.buf_bytes + a_bytes);  // [atom][row][g4]
        for (int base = 0; base < n; base += 8 * p.warps) {
            const int il0 = base + warp * 8 + 2 * cp;   // this thread's centres: il0, il0 + 1
            // own atoms: dS_i[t][g][:] as (scalar,x) / (y,z) register pairs and a_i[t][g], t = this half's 8 channels
            float2 dSi01[2][8], dSi23[2][8];
            float ai[2][8];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int ilc = min(il0 + c, n - 1);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float4 v = dbuf[(ilc * 16 + 2 * t + h) * 4 + g4];
                    dSi01[c][t] = make_float2(v.x, v.y);
                    dSi23[c][t] = make_float2(v.z, v.w);
                }
                const float4 o0 = abuf[(ilc * 4 + h) * 4 + g4], o1 = abuf[(ilc * 4 + h + 2) * 4 + g4];
                ai[c][0] = o0.x, ai[c][1] = o0.y, ai[c][2] = o0.z, ai[c][3] = o0.w;
                ai[c][4] = o1.x, ai[c][5] = o1.y, ai[c][6] = o1.z, ai[c][7] = o1.w;
            }
            // charge channels ride on the h == 0 half only (their contribution is added once)
            float2 dSqi01[2][C], dSqi23[2][C];
            float qi[2][C];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int ilc = min(il0 + c, n - 1);
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    const float4 v = (with_q && h == 0) ? dsq_s[(ilc * C + cc) * 4 + g4] : make_float4(0, 0, 0, 0);
                    dSqi01[c][cc] = make_float2(v.x, v.y);
                    dSqi23[c][cc] = make_float2(v.z, v.w);
                    qi[c][cc] = (with_q && h == 0) ? q_s[ilc * C + cc] : 0.f;
                }
            }
            float2 ga2[2][8];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int t = 0; t < 8; ++t) ga2[c][t] = make_float2(0.f, 0.f);
            float2 gq2[2][C];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int cc = 0; cc < C; ++cc) gq2[c][cc] = make_float2(0.f, 0.f);
            float fx[2] = {0.f, 0.f}, fy[2] = {0.f, 0.f}, fz[2] = {0.f, 0.f};
            for (int k0 = 0; k0 < n; k0 += 4) {   // four neighbours per round
                __syncwarp();
                {   // lane = (centre of the warp's eight, slot) stages one pair of the warp's table
                    const int ce = lane >> 2, sl = lane & 3;
                    const int ilx = base + warp * 8 + ce;
                    const PairGeom e = pair_geom<true>(min(ilx, n - 1), ilx < n, k0 + sl, n, xyz, aev);
                    ent[(0 * 4 + sl) * 8 + ce] = e.d;
                    ent[(1 * 4 + sl) * 8 + ce] = e.fc;
                    ent[(2 * 4 + sl) * 8 + ce] = e.ux;
                    ent[(3 * 4 + sl) * 8 + ce] = e.uy;
                    ent[(4 * 4 + sl) * 8 + ce] = e.uz;
                    ent[(5 * 4 + sl) * 8 + ce] = e.dfc;
                    ent[(6 * 4 + sl) * 8 + ce] = e.inv;
                }
                __syncwarp();
                const int lim = min(4, n - k0);
                for (int s = 0; s < lim; ++s) {
                    const int j = k0 + s;
                    const float4 a0 = abuf[(j * 4 + h) * 4 + g4], a1 = abuf[(j * 4 + h + 2) * 4 + g4];
                    const float aj[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float4* drow = dbuf + (j * 16 + h) * 4 + g4;   // row 2 t + h at drow[8 t]
                    const float2 e_d = ent2[(0 * 4 + s) * 4 + cp], e_fc = ent2[(1 * 4 + s) * 4 + cp], e_ux = ent2[(2 * 4 + s) * 4 + cp],
                                 e_uy = ent2[(3 * 4 + s) * 4 + cp], e_uz = ent2[(4 * 4 + s) * 4 + cp], e_dfc = ent2[(5 * 4 + s) * 4 + cp],
                                 e_inv = ent2[(6 * 4 + s) * 4 + cp];
                    const float dd[2] = {e_d.x, e_d.y}, fcv[2] = {e_fc.x, e_fc.y}, uxv[2] = {e_ux.x, e_ux.y}, uyv[2] = {e_uy.x, e_uy.y},
                                uzv[2] = {e_uz.x, e_uz.y}, dfcv[2] = {e_dfc.x, e_dfc.y}, invv[2] = {e_inv.x, e_inv.y};
                    float gs[2], dgs[2];
                    float2 G01[2], G23[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float xg = dd[c] - shift_g;
                        const float ex = aev_exp(-aev.eta * xg * xg);
                        gs[c] = ex * fcv[c];
                        dgs[c] = ex * (dfcv[c] - 2.0f * aev.eta * xg * fcv[c]);
                        // g_sv(j->i)[g,:] = (gs, -gs u): grad_a[i] += <dS[j], g_sv(j->i)> as two packed FMAs per channel
                        G01[c] = make_float2(gs[c], -gs[c] * uxv[c]);
                        G23[c] = make_float2(-gs[c] * uyv[c], -gs[c] * uzv[c]);
                    }
                    // p = contraction for the pair (i -> j), r = for the reverse pair (j -> i), over this half's 8 channels
                    float2 p01[2], p23[2], r01[2], r23[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c) p01[c] = p23[c] = r01[c] = r23[c] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float4 dj = drow[8 * t];
                        const float2 dj01 = make_float2(dj.x, dj.y), dj23 = make_float2(dj.z, dj.w);
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            if (kGradA) {
                                ga2[c][t] = ffma2(dj01, G01[c], ga2[c][t]);
                                ga2[c][t] = ffma2(dj23, G23[c], ga2[c][t]);
                            }
                            p01[c] = ffma2s(aj[t], dSi01[c][t], p01[c]);
                            p23[c] = ffma2s(aj[t], dSi23[c][t], p23[c]);
                            r01[c] = ffma2s(ai[c][t], dj01, r01[c]);
                            r23[c] = ffma2s(ai[c][t], dj23, r23[c]);
                        }
                    }
                    if (with_q) {
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) {
                            const float qj = q_s[j * C + cc];
                            const float4 dqj = dsq_s[(j * C + cc) * 4 + g4];
                            const float2 dq01 = make_float2(dqj.x, dqj.y), dq23 = make_float2(dqj.z, dqj.w);
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                if (kGradA) {
                                    gq2[c][cc] = ffma2(dq01, G01[c], gq2[c][cc]);
                                    gq2[c][cc] = ffma2(dq23, G23[c], gq2[c][cc]);
                                }
                                p01[c] = ffma2s(qj, dSqi01[c][cc], p01[c]);
                                p23[c] = ffma2s(qj, dSqi23[c][cc], p23[c]);
                                r01[c] = ffma2s(qi[c][cc], dq01, r01[c]);
                                r23[c] = ffma2s(qi[c][cc], dq23, r23[c]);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float gsi = gs[c] * invv[c];
                        // this thread's share of w(i->j) = u (A + C.u) + (B - u (B.u))/d
                        const float pu = p01[c].y * uxv[c] + p23[c].x * uyv[c] + p23[c].y * uzv[c];
                        const float sc = (p01[c].x + pu) * dgs[c] - pu * gsi;
                        // reverse pair (j->i): u' = -u;  w' = -u (A' - (r.u) dgs) + (B' - u (B'.u))/d
                        const float ru = r01[c].y * uxv[c] + r23[c].x * uyv[c] + r23[c].y * uzv[c];
                        const float scr = (ru - r01[c].x) * dgs[c] - ru * gsi;
                        // F_i += w - w' = u (sc - scr) + (B - B') gs/d
                        const float ds = sc - scr;
                        fx[c] += fmaf(uxv[c], ds, (p01[c].y - r01[c].y) * gsi);
                        fy[c] += fmaf(uyv[c], ds, (p23[c].x - r23[c].x) * gsi);
                        fz[c] += fmaf(uzv[c], ds, (p23[c].y - r23[c].y) * gsi);
                    }
                }
            }
            // reduce the force / grad_q shares over the eight lanes (two channel halves x four radial shifts) of the pair
            auto oct_sum = [](float v) {
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                return v;
            };
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float sx = oct_sum(fx[c]), sy = oct_sum(fy[c]), sz = oct_sum(fz[c]);
                float gqs[C];
#pragma unroll
                for (int cc = 0; cc < C; ++cc)
                    gqs[cc] = (kGradA && with_q) ? oct_sum(h == 0 ? gq2[c][cc].x + gq2[c][cc].y : 0.f) : 0.f;
                const int il = il0 + c;
                if (il < n) {
                    const int i = lo + il;
                    if (kGradA) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int a = 4 * h + (t & 3) + ((t >> 2) << 3);
                            grad_a[(size_t)i * kAG + a * kG + g] = ga2[c][t].x + ga2[c][t].y;
                        }
                    }
                    if ((lane & 7) == 0) {
                        float* fp = f_part + ((size_t)gq * p.n_atoms + i) * 3;
                        fp[0] = sx;
                        fp[1] = sy;
                        fp[2] = sz;
                        if (kGradA && with_q) {
#pragma unroll
                            for (int cc = 0; cc < C; ++cc) gq_part[((size_t)gq * p.n_atoms + i) * C + cc] = gqs[cc];
                        }
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
    L.ent = o, o += p.warps * kEntFwd * 4;
    L.agh = o = up(o, 16), o += kA * kAghPad * 4;
    L.aghq = o, o += 2 * kG * kH * 4;
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
    p.warps = std::min(10, (max_seg + 7) / 8);
    Layout& L = p.L;
    int o = 0;
    L.ent = o, o += p.warps * kEntBwd * 4;
    L.agh = L.aghq = L.sv = o;
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
