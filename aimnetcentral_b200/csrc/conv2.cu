// AEV + conv_sv message passing, second-generation kernels (forward and analytic backward).
//
// Same arithmetic and reference semantics as conv.cu (calc_distances aimnet/ops.py:37-66, AEVSV._calc_aev
// aimnet/modules/aev.py:94-110, ConvSV.forward aimnet/modules/aev.py:156-189, Warp kernels
// aimnet/kernels/conv_sv_2d_sp_wp.py:90-164); what changes is the work decomposition, chosen from the round-1 ncu captures
// (conv_fwd 75 issued instructions per pair for 16 packed FMAs, conv_bwd 140 for 48; one L1 wavefront per 128 gathered bytes):
//
//   * one warp = TWO centre atoms, lane = (centre c, radial shift g) owns ALL 16 feature channels of its (i, g): the 16 x 4
//     register tile S[i, :, g, :].  The per-pair scalar work (Gaussian, cutoff, unit vector products, force algebra) is done
//     once per 16 channels instead of once per 8, and no longer twice (the two channel halves of conv.cu each evaluated the
//     same exponential): forward 29, backward ~100 issued instructions per pair.
//   * accumulators are register pairs over two neighbouring channels, so that the gathered float4 (four channels of one g)
//     feeds the packed FFMA2 directly; the broadcast operand is the pair weight, packed once per pair.
//   * DENSE mode for batches of small molecules: both centres of a warp walk the atom segment of their molecule in
//     lock step (pairs beyond the cutoff contribute exactly zero: fc(d >= rc) == 0), so the two half-warps gather the SAME
//     neighbour rows and every L1 wavefront serves two pairs.  Summation order = atom index = the order of the
//     canonical (sorted) list rows, so the forward results are bitwise those of the list walk.
//   * LIST mode (periodic systems, large molecules): each half-warp walks its own matrix row.
//
// Feature layouts are those of conv.cu: aX (N, 4 a-quads, 16 g, 4 a), dS (N, 16 a, 16 g, 4 d), T (N, 16 a, 12 h, 3).
#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {
namespace conv2 {

constexpr int kWarpsFwd = 8;         // forward CTA: 8 warps = 16 centre atoms per group
constexpr int kWarpsBwd = 4;         // backward CTA: 4 warps = 8 centre atoms
constexpr int kSlots = 16;           // neighbour slots staged per centre and round (one per lane of the half-warp)
// shared-memory layouts of the forward epilogue, as in conv.cu
constexpr int kAghRow = 20;
constexpr int kSvRow = 52;
constexpr int kSvAtom = kA * kSvRow + 16;
constexpr int kFwdSmemBytes = kWarpsFwd * 32 * 32 + (kA + 2) * kH * kAghRow * 4 + 2 * kWarpsFwd * (kSvAtom + 2 * kSvRow) * 4 + 64;
__device__ __forceinline__ int sv_off(int a) { return a * kSvRow + ((a >> 3) << 4); }

struct PairEntry {
    float ux, uy, uz, d;
    float fc, dfc;
    int j;
    float inv;   // 1/d
};

__device__ __forceinline__ float aev_exp(float x) { return __expf(x); }

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

// Geometry + cutoff of slots [k0, k0 + 16) of the warp's two centres: lane = (centre, slot).  Dense mode: slot k is atom
// seg_lo + k of the centre's molecule (the centre itself is skipped); list mode: slot k of the centre's matrix row.
template <bool kDense, bool kWithDeriv>
__device__ __forceinline__ void stage16(PairEntry* tile, int lane, int i, bool atom_ok, int k0, int len, int seg_lo,
                                        const NbView& nb, const float* __restrict__ coord, const float* __restrict__ cell,
                                        const AevParams& aev) {
    const int k = k0 + (lane & 15);
    int j = -1;
    bool ok = atom_ok && k < len;
    const int32_t* sh = nullptr;
    if (ok) {
        if (kDense) {
            j = seg_lo + k;
            ok = j != i;
        } else {
            j = nb.nbmat[(size_t)i * nb.width + k];
            ok = (j != nb.sentinel) && (j >= 0);
            if (nb.shifts) sh = nb.shifts + ((size_t)i * nb.width + k) * 3;
        }
    }
    float rx = 1.f, ry = 1.f, rz = 1.f;
    if (ok) pair_vector(coord, i, j, sh, cell, rx, ry, rz);
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
    e.j = ok ? j : (atom_ok ? i : 0);
    e.inv = inv;
    tile[lane] = e;
}

__device__ __forceinline__ int row_length(const NbView& nb, int i) { return nb.count ? min(nb.count[i], nb.width) : nb.width; }

// channel pair ap (0..7) of a lane <-> channels a_lo = 4 (ap >> 1) + 2 (ap & 1), a_lo + 1: the (x,y) / (z,w) halves of the
// float4 of quad ap >> 1 in the gather layout
__device__ __forceinline__ int pair_lo(int ap) { return 4 * (ap >> 1) + 2 * (ap & 1); }

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int C, bool kDense>
__global__ void __launch_bounds__(256, 2) fwd_kernel(int n_atoms, int n_groups, NbView nb, const int32_t* __restrict__ mol_ptr,
                                                     const float* __restrict__ coord, CellView cv,
                                                     const int32_t* __restrict__ mol_idx, AevParams aev,
                                                     const float* __restrict__ aT, const float* __restrict__ q,
                                                     const float* __restrict__ agh_a, const float* __restrict__ agh_q,
                                                     float* __restrict__ x, int ldx, float* __restrict__ T_a,
                                                     float* __restrict__ T_q, int with_q) {
    extern __shared__ __align__(16) unsigned char fwd_smem[];
    PairEntry* tiles = reinterpret_cast<PairEntry*>(fwd_smem);                              // [8 warps][2 centres][16 slots]
    float* aghT_a = reinterpret_cast<float*>(fwd_smem + sizeof(PairEntry) * 32 * kWarpsFwd);  // [a][h][g], row stride kAghRow
    float* aghT_q = aghT_a + kA * kH * kAghRow;                                             // [c][h][g]
    float* sv_all = aghT_q + 2 * kH * kAghRow;                                              // [atom][a][k][g], see sv_off()
    float* svq_all = sv_all + 2 * kWarpsFwd * kSvAtom;                                      // [atom][c][k][g]
    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31, c = lane >> 4, g = lane & 15;
    const float shift_g = aev.shifts[g];
    for (int e = tid; e < kA * kG * kH; e += 256) {
        int a = e / (kG * kH), gg = (e / kH) % kG, hh = e % kH;
        aghT_a[(a * kH + hh) * kAghRow + gg] = agh_a[e];
    }
    if (with_q)
        for (int e = tid; e < C * kG * kH; e += 256) {
            int cc = e / (kG * kH), gg = (e / kH) % kG, hh = e % kH;
            aghT_q[(cc * kH + hh) * kAghRow + gg] = agh_q[e];
        }
    __syncthreads();
    PairEntry* tile = tiles + w * 32;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int i = (grp * kWarpsFwd + w) * 2 + c;
        const bool atom_ok = i < n_atoms;
        const int ic = atom_ok ? i : 0;
        const int mol = mol_idx ? mol_idx[ic] : 0;
        const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : mol) : nullptr;
        int len = 0, seg_lo = 0;
        if (kDense) {
            seg_lo = mol_ptr[mol];
            len = atom_ok ? mol_ptr[mol + 1] - seg_lo : 0;
        } else {
            len = atom_ok ? row_length(nb, i) : 0;
        }
        const int kmax = max(len, __shfl_xor_sync(0xffffffffu, len, 16));
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
        for (int k0 = 0; k0 < kmax; k0 += kSlots) {
            __syncwarp();
            stage16<kDense, false>(tile, lane, ic, atom_ok, k0, len, seg_lo, nb, coord, cell, aev);
            __syncwarp();
            const int lim = min(kSlots, kmax - k0);
#pragma unroll 2
            for (int s = 0; s < lim; ++s) {
                const PairEntry e = tile[c * kSlots + s];
                const float4* row = reinterpret_cast<const float4*>(aT + (size_t)e.j * kAG) + g;
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
                        const float qj = q[(size_t)e.j * C + cc];
#pragma unroll
                        for (int d = 0; d < 4; ++d) Sq[cc][d] = fmaf(qj, wv[d], Sq[cc][d]);
                    }
                }
            }
        }
        // ---- epilogue: scalar part straight to x, vector part through shared memory for the agh mixing ----
        __syncwarp();      // the previous group's mixing is done with this warp's Sv scratch
        float* svl = sv_all + (w * 2 + c) * kSvAtom;
        float* svql = svq_all + (w * 2 + c) * 2 * kSvRow;
#pragma unroll
        for (int ap = 0; ap < 8; ++ap) {
            const int o = sv_off(pair_lo(ap)) + g;   // channels a_lo and a_lo + 1 never straddle the pad between 7 and 8
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
        if (atom_ok) {
            float* xr = x + (size_t)i * ldx;
            const float4* own = reinterpret_cast<const float4*>(aT + (size_t)i * kAG) + g;
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
                if (g < C) xr[base + g] = q[(size_t)i * C + g];
#pragma unroll
                for (int cc = 0; cc < C; ++cc) xr[base + C + cc * kG + g] = Sq[cc][0];
                base += C * (1 + kG + kH);
            }
            for (int cidx = base + g; cidx < ldx; cidx += kG) xr[cidx] = 0.f;
        }
        __syncwarp();
        // T[a,h,k] = sum_g agh[a,g,h] * Sv[a,g,k]   (aimnet/modules/aev.py:188): the whole warp mixes centre 0, then centre 1
#pragma unroll 1
        for (int cc2 = 0; cc2 < 2; ++cc2) {
            const int ia = (grp * kWarpsFwd + w) * 2 + cc2;
            if (ia >= n_atoms) break;
            const float* sva = sv_all + (w * 2 + cc2) * kSvAtom;
            const float* svqa = svq_all + (w * 2 + cc2) * 2 * kSvRow;
            float* xr = x + (size_t)ia * ldx;
#pragma unroll 2
            for (int e = lane; e < kAH; e += 32) {
                const int a = e / kH;
                float t[3];
                mix16(aghT_a + e * kAghRow, sva + sv_off(a), t);
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
                    mix16(aghT_q + e * kAghRow, svqa + cq * kSvRow, t);
                    float* To = T_q + (size_t)ia * (C * kH * 3) + e * 3;
                    To[0] = t[0];
                    To[1] = t[1];
                    To[2] = t[2];
                    xr[base + C + C * kG + e] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward: per atom from its OWN row only (gather form, no atomics, deterministic; see conv.cu):
//   grad_a[i,a,g]  = sum_m <dS[j_m,a,g,:], g_sv(j_m->i)[g,:]>            g_sv(j->i) = (gs, -gs u_{i->j})
//   F_i            = sum_m ( w(i->j_m) - w(j_m->i) )                       w = dE/dr of a slot
//   virial_i       = sum_m r_im (x) w(i->j_m)
// ------------------------------------------------------------------------------------------------------------
template <int C, bool kDense, bool kGradA, bool kVirial>
__global__ void __launch_bounds__(128, 3) bwd_kernel(int n_atoms, NbView nb, const int32_t* __restrict__ mol_ptr,
                                                     const float* __restrict__ coord, CellView cv,
                                                     const int32_t* __restrict__ mol_idx, AevParams aev,
                                                     const float* __restrict__ aT, const float* __restrict__ q,
                                                     const float* __restrict__ dS_a, const float* __restrict__ dS_q,
                                                     float* __restrict__ grad_a, float* __restrict__ grad_q,
                                                     float* __restrict__ forces, double* __restrict__ virial_atom, int with_q) {
    __shared__ PairEntry tiles[kWarpsBwd * 32];
    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31, c = lane >> 4, g = lane & 15;
    PairEntry* tile = tiles + w * 32;
    const int i = (blockIdx.x * kWarpsBwd + w) * 2 + c;
    const bool atom_ok = i < n_atoms;
    const int ic = atom_ok ? i : 0;
    const int mol = mol_idx ? mol_idx[ic] : 0;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : mol) : nullptr;
    int len = 0, seg_lo = 0;
    if (kDense) {
        seg_lo = mol_ptr[mol];
        len = atom_ok ? mol_ptr[mol + 1] - seg_lo : 0;
    } else {
        len = atom_ok ? row_length(nb, i) : 0;
    }
    const int kmax = max(len, __shfl_xor_sync(0xffffffffu, len, 16));
    const float shift_g = aev.shifts[g];
    // own atom: dS_i[a][g][:] as (scalar,x) / (y,z) register pairs and a_i[a][g] for all 16 channels
    float2 dSi01[kA], dSi23[kA];
    float ai[kA];
    {
        const float4* p = reinterpret_cast<const float4*>(dS_a) + (size_t)ic * kAG + g;
#pragma unroll
        for (int a = 0; a < kA; ++a) {
            const float4 v = p[a * kG];
            dSi01[a] = make_float2(v.x, v.y);
            dSi23[a] = make_float2(v.z, v.w);
        }
        const float4* r = reinterpret_cast<const float4*>(aT + (size_t)ic * kAG) + g;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float4 o = r[16 * qd];
            ai[4 * qd + 0] = o.x;
            ai[4 * qd + 1] = o.y;
            ai[4 * qd + 2] = o.z;
            ai[4 * qd + 3] = o.w;
        }
    }
    float2 dSqi01[C], dSqi23[C];
    float qi[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
        const float4 v = with_q ? reinterpret_cast<const float4*>(dS_q)[(size_t)ic * (C * kG) + cc * kG + g] : make_float4(0, 0, 0, 0);
        dSqi01[cc] = make_float2(v.x, v.y);
        dSqi23[cc] = make_float2(v.z, v.w);
        qi[cc] = with_q ? q[(size_t)ic * C + cc] : 0.f;
    }
    float2 ga2[kA];
#pragma unroll
    for (int a = 0; a < kA; ++a) ga2[a] = make_float2(0.f, 0.f);
    float2 gq2[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) gq2[cc] = make_float2(0.f, 0.f);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    float vir[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) vir[k] = 0.f;

    for (int k0 = 0; k0 < kmax; k0 += kSlots) {
        __syncwarp();
        stage16<kDense, true>(tile, lane, ic, atom_ok, k0, len, seg_lo, nb, coord, cell, aev);
        __syncwarp();
        const int lim = min(kSlots, kmax - k0);
        for (int s = 0; s < lim; ++s) {
            const PairEntry e = tile[c * kSlots + s];
            const float4* arow = reinterpret_cast<const float4*>(aT + (size_t)e.j * kAG) + g;
            const float4* drow = reinterpret_cast<const float4*>(dS_a) + (size_t)e.j * kAG + g;
            const float xg = e.d - shift_g;
            const float ex = aev_exp(-aev.eta * xg * xg);
            const float gs = ex * e.fc;
            const float dgs = ex * (e.dfc - 2.0f * aev.eta * xg * e.fc);
            // g_sv(j->i)[g,:] = (gs, -gs u): grad_a[i] += <dS[j], g_sv(j->i)> as two packed FMAs per channel
            const float2 G01 = make_float2(gs, -gs * e.ux), G23 = make_float2(-gs * e.uy, -gs * e.uz);
            // p = contraction for slot (i -> j), r = for the reverse slot (j -> i), over all 16 channels of this (i, g)
            float2 p01 = make_float2(0.f, 0.f), p23 = p01, r01 = p01, r23 = p01;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const float4 av = arow[16 * qd];
                const float aj[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int a = 4 * qd + k;
                    const float4 dj = drow[a * kG];
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
                    const float qj = q[(size_t)e.j * C + cc];
                    const float4 dqj = reinterpret_cast<const float4*>(dS_q)[(size_t)e.j * (C * kG) + cc * kG + g];
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
            // reverse slot (j->i): u' = -u;  w' = -u (A' - (r.u) dgs) + (B' - u (B'.u))/d
            const float ru = r01.y * e.ux + r23.x * e.uy + r23.y * e.uz;
            const float scr = (ru - r01.x) * dgs - ru * gsi;
            if (!kVirial) {
                // F_i += w - w' = u (sc - scr) + (B - B') gs/d
                const float ds = sc - scr;
                fx += fmaf(e.ux, ds, (p01.y - r01.y) * gsi);
                fy += fmaf(e.uy, ds, (p23.x - r23.x) * gsi);
                fz += fmaf(e.uz, ds, (p23.y - r23.y) * gsi);
            } else {
                const float wx = e.ux * sc + p01.y * gsi;
                const float wy = e.uy * sc + p23.x * gsi;
                const float wz = e.uz * sc + p23.y * gsi;
                const float vx = e.ux * scr + r01.y * gsi;
                const float vy = e.uy * scr + r23.x * gsi;
                const float vz = e.uz * scr + r23.y * gsi;
                fx += wx - vx;
                fy += wy - vy;
                fz += wz - vz;
                const float rx = e.ux * e.d, ry = e.uy * e.d, rz = e.uz * e.d;
                vir[0] = fmaf(rx, wx, vir[0]);
                vir[1] = fmaf(rx, wy, vir[1]);
                vir[2] = fmaf(rx, wz, vir[2]);
                vir[3] = fmaf(ry, wx, vir[3]);
                vir[4] = fmaf(ry, wy, vir[4]);
                vir[5] = fmaf(ry, wz, vir[5]);
                vir[6] = fmaf(rz, wx, vir[6]);
                vir[7] = fmaf(rz, wy, vir[7]);
                vir[8] = fmaf(rz, wz, vir[8]);
            }
        }
    }
    // reduce the force / virial / grad_q shares over the 16 lanes (radial shifts) of each centre
    auto half_sum = [](float v) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    fx = half_sum(fx);
    fy = half_sum(fy);
    fz = half_sum(fz);
    if (kVirial) {
#pragma unroll
        for (int k = 0; k < 9; ++k) vir[k] = half_sum(vir[k]);
    }
    float gq[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) gq[cc] = 0.f;
    if (kGradA && with_q) {
#pragma unroll
        for (int cc = 0; cc < C; ++cc) gq[cc] = half_sum(gq2[cc].x + gq2[cc].y);
    }
    if (!atom_ok) return;
    if (kGradA) {
#pragma unroll
        for (int a = 0; a < kA; ++a) grad_a[(size_t)i * kAG + a * kG + g] = ga2[a].x + ga2[a].y;
        if (with_q && g < C) {
            float v = gq[0];
#pragma unroll
            for (int cc = 1; cc < C; ++cc) v = (g == cc) ? gq[cc] : v;
            grad_q[(size_t)i * C + g] = v;
        }
    }
    if (g == 0) {
        forces[3 * i + 0] += fx;
        forces[3 * i + 1] += fy;
        forces[3 * i + 2] += fz;
        if (kVirial)
#pragma unroll
            for (int k = 0; k < 9; ++k) virial_atom[(size_t)i * 9 + k] += (double)vir[k];
    }
}

}  // namespace conv2

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
template <int C, bool kDense>
static int conv2_fwd_launch(int n_atoms, const NbView& nb, const int32_t* mol_ptr, const float* coord, const CellView& cv,
                            const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q,
                            const float* agh_a, const float* agh_q, float* x, int ldx, float* T_a, float* T_q, int with_q,
                            cudaStream_t st) {
    using namespace conv2;
    const int per_group = 2 * kWarpsFwd;
    const int n_groups = (n_atoms + per_group - 1) / per_group;
    static bool configured_dev[kMaxDevices] = {};
    static int max_ctas_dev[kMaxDevices] = {};
    const int dslot = current_device_slot();
    int& max_ctas = max_ctas_dev[dslot];
    if (!configured_dev[dslot]) {
        AIM_CUDA_CHECK(cudaFuncSetAttribute(fwd_kernel<C, kDense>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
        int dev = 0, sms = 148, per_sm = 2;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        AIM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fwd_kernel<C, kDense>, 256, kFwdSmemBytes));
        max_ctas = sms * (per_sm > 0 ? per_sm : 1);
        configured_dev[dslot] = true;
    }
    const int grid = n_groups < max_ctas ? n_groups : max_ctas;
    fwd_kernel<C, kDense><<<grid, 256, kFwdSmemBytes, st>>>(n_atoms, n_groups, nb, mol_ptr, coord, cv, mol_idx, aev, aT, q, agh_a,
                                                           agh_q, x, ldx, T_a, T_q, with_q);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

int launch_conv2_fwd(int C, int dense, int n_atoms, const NbView& nb, const int32_t* mol_ptr, const float* coord,
                     const CellView& cv, const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q,
                     const float* agh_a, const float* agh_q, float* x, int ldx, float* T_a, float* T_q, int with_q,
                     cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    AIM_REQUIRE(!dense || mol_ptr != nullptr, "conv2_fwd: dense mode needs the molecule segment pointers");
#define AIM_C2F(CC, DD) \
    return conv2_fwd_launch<CC, DD>(n_atoms, nb, mol_ptr, coord, cv, mol_idx, aev, aT, q, agh_a, agh_q, x, ldx, T_a, T_q, with_q, st)
    if (C == 1) {
        if (dense) AIM_C2F(1, true);
        AIM_C2F(1, false);
    }
    if (dense) AIM_C2F(2, true);
    AIM_C2F(2, false);
#undef AIM_C2F
}

template <int C, bool kDense>
static int conv2_bwd_launch(int n_atoms, const NbView& nb, const int32_t* mol_ptr, const float* coord, const CellView& cv,
                            const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q, const float* dS_a,
                            const float* dS_q, float* grad_a, float* grad_q, float* forces, double* virial_atom, int with_q,
                            int want_grad_a, cudaStream_t st) {
    using namespace conv2;
    const int per_cta = 2 * kWarpsBwd;
    const int grid = (n_atoms + per_cta - 1) / per_cta;
#define AIM_C2B(GA, VIR)                                                                                                  \
    bwd_kernel<C, kDense, GA, VIR><<<grid, 32 * kWarpsBwd, 0, st>>>(n_atoms, nb, mol_ptr, coord, cv, mol_idx, aev, aT, q, dS_a, \
                                                                    dS_q, grad_a, grad_q, forces, virial_atom, with_q)
    if (want_grad_a) {
        if (virial_atom) AIM_C2B(true, true); else AIM_C2B(true, false);
    } else {
        if (virial_atom) AIM_C2B(false, true); else AIM_C2B(false, false);
    }
#undef AIM_C2B
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

// the gather half of the backward pass (dS_a / dS_q must have been produced by conv_bwd_prep); `forces` must not be
// written concurrently by another stream: every atom's total is accumulated with a plain +=
int launch_conv2_bwd_gather(int C, int dense, int n_atoms, const NbView& nb, const int32_t* mol_ptr, const float* coord,
                            const CellView& cv, const int32_t* mol_idx, const AevParams& aev, const float* aT,
                            const float* q, const float* dS_a, const float* dS_q, float* grad_a, float* grad_q,
                            float* forces, double* virial_atom, int with_q, int want_grad_a, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    AIM_REQUIRE(!dense || mol_ptr != nullptr, "conv2_bwd: dense mode needs the molecule segment pointers");
#define AIM_C2BL(CC, DD)                                                                                                       \
    return conv2_bwd_launch<CC, DD>(n_atoms, nb, mol_ptr, coord, cv, mol_idx, aev, aT, q, dS_a, dS_q, grad_a, grad_q, forces, \
                                    virial_atom, with_q, want_grad_a, st)
    if (C == 1) {
        if (dense) AIM_C2BL(1, true);
        AIM_C2BL(1, false);
    }
    if (dense) AIM_C2BL(2, true);
    AIM_C2BL(2, false);
#undef AIM_C2BL
}

}  // namespace aimnet
