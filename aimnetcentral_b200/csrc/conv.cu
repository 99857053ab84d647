// AEV + conv_sv message passing, forward and analytic backward.
//
// Reference semantics: calc_distances (aimnet/ops.py:37-66), AEVSV._calc_aev (aimnet/modules/aev.py:94-110),
// ConvSV.forward (aimnet/modules/aev.py:156-189), Warp kernels aimnet/kernels/conv_sv_2d_sp_wp.py:90-164.
// The reference materialises g_sv (N,M,16,4) in HBM (256 B per pair) and re-reads it for every convolution and for
// the autograd pass; here the radial basis is recomputed from a small per-tile pair table in shared memory.
//
// Thread mapping (forward and backward): one warp = one centre atom; lane = (half h, radial shift g) owns the
// 8(a) x 4(d) register tile of S[i, 8h..8h+7, g, :].  Per neighbour it loads two float4 of a[j] and does an 8x4
// outer-product update: 32 FMAs for 2 vector loads, no cross-thread traffic; 8 atoms (warps) per CTA.  Features are kept in a gather
// layout aX (N, 4 a-quads, 16 g, 4 a) so that the 16 g-lanes of an atom read 256 contiguous bytes per float4 load
// (index(a,g) = ((a>>2)*16 + g)*4 + (a&3)); dS (N, 16 a, 16 g, 4 d) is lane-contiguous as is.
//
// Backward: the neighbour matrix is full (both directions), so everything atom i needs is in its own row:
//   grad_a[i,a,g]  = sum_m <dS[j_m,a,g,:], g_sv(j_m->i)[g,:]>            g_sv(j->i) = (gs, -gs u_{i->j})
//   F_i            = sum_m ( w(i->j_m) - w(j_m->i) )                       w = dE/dr of a slot
//   virial_i       = sum_m r_im (x) w(i->j_m)
// w is linear in the per-g partial contractions, so every thread accumulates its own g-share of the force over all
// pairs and the 16 lanes of an atom are reduced once at the end: no per-pair reductions, no atomics, deterministic
// (the reference's backward kernel scatters with wp.atomic_add, conv_sv_2d_sp_wp.py:115-136).
//
// Layout of one MLP input row x (ld = ldx, zero padded):
//   [0,256)   a[i] flat (a*16+g)            [256,512) S_s[a,g]          [512,704) avf_v[a,h]
//   pass>0:   [704,704+C) q[i,c]            [704+C, 704+17C) Sq_s[c,g]  [704+17C, 704+29C) avfq_v[c,h]
#include <cstdlib>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

constexpr int kAtomsPerCta = 8;     // one warp per atom
constexpr int kSlotsPerTile = 32;   // neighbour slots staged per tile and atom (one per lane)
constexpr int kHalfA = 8;           // feature channels per thread (two half-warps split the 16 channels)
// shared-memory layouts of the forward epilogue (bank-conflict-free for the access patterns described there)
constexpr int kAghRow = 20;                       // agh^T row: 16 g + 4 pad -> 8 consecutive rows cover all 32 banks once
constexpr int kSvRow = 52;                        // Sv of one channel: 3 k x 16 g + 4 pad
constexpr int kSvAtom = kA * kSvRow + 16;         // + 16 between channels 7 and 8 (the two half-warps write different banks)
constexpr int kFwdSmemBytes = 256 * 32 + (kA + 2) * kH * kAghRow * 4 + kAtomsPerCta * (kSvAtom + 2 * kSvRow) * 4 + 64;
__device__ __forceinline__ int sv_off(int a) { return a * kSvRow + ((a >> 3) << 4); }
// t[k] = sum_g w[g] * s[k][g], w: 16 floats, s: 3 rows of 16 floats (16-byte aligned)
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

struct PairEntry {
    float ux, uy, uz, d;
    float fc, dfc;
    int j;
    float inv;   // 1/d
};

// Gaussian of the radial basis, aimnet/ops.py:93-96 (exp_expand).  ex2.approx of x*log2(e): relative error <= 2^-22 + |x| 2^-24
// (x in [-20, 0] wherever the result matters), far inside the 1e-4 parity budget; forward and backward use the same
// function, so the forces stay the exact gradient of the energy that is returned.
__device__ __forceinline__ float aev_exp(float x) { return __expf(x); }

// stage geometry + cutoff for slots [m0, m0+32) of the CTA's 8 atoms: thread (atom = warp, slot = lane)
template <bool kWithDeriv>
__device__ __forceinline__ void stage_pairs(PairEntry* tile, int i, bool atom_ok, int m0, int row_len, const NbView& nb,
                                            const float* __restrict__ coord, const float* __restrict__ cell,
                                            const AevParams& aev) {
    int tid = threadIdx.x;
    int m = m0 + (tid & 31);
    int j = nb.sentinel;
    if (atom_ok && m < row_len) j = nb.nbmat[(size_t)i * nb.width + m];
    bool ok = atom_ok && (j != nb.sentinel) && (j >= 0);
    float rx = 1.f, ry = 1.f, rz = 1.f;
    if (ok) {
        const int32_t* sh = nb.shifts ? nb.shifts + ((size_t)i * nb.width + m) * 3 : nullptr;
        pair_vector(coord, i, j, sh, cell, rx, ry, rz);
    }
    float d = sqrtf(rx * rx + ry * ry + rz * rz);
    float inv = 1.0f / d;
    // cosine cutoff, aimnet/ops.py:82-85
    float dc = fminf(fmaxf(d, 1e-6f), aev.rc);
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
    tile[tid] = e;
}

__device__ __forceinline__ int row_length(const NbView& nb, int i) {
    return nb.count ? min(nb.count[i], nb.width) : nb.width;
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256, 3) conv_fwd_kernel(int n_atoms, int n_groups, NbView nb, const float* __restrict__ coord,
                                                          CellView cv, const int32_t* __restrict__ mol_idx,
                                                          AevParams aev, const float* __restrict__ aT,
                                                          const float* __restrict__ q, const float* __restrict__ agh_a,
                                                          const float* __restrict__ agh_q, float* __restrict__ x,
                                                          int ldx, float* __restrict__ T_a, float* __restrict__ T_q,
                                                          int with_q) {
    // dynamic shared memory (kFwdSmemBytes): pair tile | agh^T tables | per-atom vector parts
    extern __shared__ __align__(16) unsigned char fwd_smem[];
    PairEntry* tile = reinterpret_cast<PairEntry*>(fwd_smem);                        // 8 atoms x 32 slots
    float* aghT_a = reinterpret_cast<float*>(fwd_smem + sizeof(PairEntry) * 256);    // [a][h][g], row stride kAghRow
    float* aghT_q = aghT_a + kA * kH * kAghRow;                                      // [c][h][g]
    float* sv_all = aghT_q + 2 * kH * kAghRow;                                       // [atom][a][k][g], see sv_off()
    float* svq_all = sv_all + kAtomsPerCta * kSvAtom;                                // [atom][c][k][g]
    const int tid = threadIdx.x;
    const int al = tid >> 5, lane = tid & 31, g = lane & 15, h = lane >> 4;
    const float shift_g = aev.shifts[g];
    // agh (a,g,h) -> shared memory transposed to (a,h,g): the mixing epilogue reads 16 contiguous g per (a,h).  Staged
    // once: the CTA is persistent and walks groups of 8 atoms; after this barrier the warps run independently.
    for (int e = tid; e < kA * kG * kH; e += 256) {
        int a = e / (kG * kH), gg = (e / kH) % kG, hh = e % kH;
        aghT_a[(a * kH + hh) * kAghRow + gg] = agh_a[e];
    }
    if (with_q)
        for (int e = tid; e < C * kG * kH; e += 256) {
            int c = e / (kG * kH), gg = (e / kH) % kG, hh = e % kH;
            aghT_q[(c * kH + hh) * kAghRow + gg] = agh_q[e];
        }
    __syncthreads();
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int i = grp * kAtomsPerCta + al;
        const bool atom_ok = i < n_atoms;
        const int ic = atom_ok ? i : 0;
        const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (mol_idx ? mol_idx[ic] : 0)) : nullptr;
        const int len = atom_ok ? row_length(nb, i) : 0;
        // accumulators as (scalar, x) / (y, z) register pairs: the 8x4 outer-product update is 16 packed FFMA2 per pair
        float2 S01[kHalfA], S23[kHalfA];
#pragma unroll
        for (int a = 0; a < kHalfA; ++a) S01[a] = S23[a] = make_float2(0.f, 0.f);
        float2 Sq01[C], Sq23[C];
#pragma unroll
        for (int c = 0; c < C; ++c) Sq01[c] = Sq23[c] = make_float2(0.f, 0.f);
        // every warp stages the 32 slots of its own atom (thread = slot) and reads only those: warp-level sync is enough
        for (int m0 = 0; m0 < len; m0 += kSlotsPerTile) {
            __syncwarp();
            stage_pairs<false>(tile, ic, atom_ok, m0, len, nb, coord, cell, aev);
            __syncwarp();
            int lim = min(kSlotsPerTile, len - m0);
#pragma unroll 2
            for (int s = 0; s < lim; ++s) {
                const PairEntry e = tile[al * 32 + s];
                const float4* row = reinterpret_cast<const float4*>(aT + (size_t)e.j * kAG) + g + 32 * h;
                float4 v0 = row[0], v1 = row[16];
                float xg = e.d - shift_g;
                float w0 = aev_exp(-aev.eta * xg * xg) * e.fc;
                const float2 w01 = make_float2(w0, w0 * e.ux), w23 = make_float2(w0 * e.uy, w0 * e.uz);
                float av[kHalfA] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int a = 0; a < kHalfA; ++a) {
                    S01[a] = ffma2s(av[a], w01, S01[a]);
                    S23[a] = ffma2s(av[a], w23, S23[a]);
                }
                if (with_q) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float qj = q[(size_t)e.j * C + c];
                        Sq01[c] = ffma2s(qj, w01, Sq01[c]);
                        Sq23[c] = ffma2s(qj, w23, Sq23[c]);
                    }
                }
            }
        }
        // ---- epilogue: scalar part straight to x, vector part through shared memory for the agh mixing ----
        __syncwarp();      // the previous atom's mixing is done with this warp's Sv scratch
        float* svl = sv_all + al * kSvAtom;
#pragma unroll
        for (int a = 0; a < kHalfA; ++a) {
            const int o = sv_off(kHalfA * h + a) + g;
            svl[o] = S01[a].y;
            svl[o + kG] = S23[a].x;
            svl[o + 2 * kG] = S23[a].y;
        }
        float* svql = svq_all + al * 2 * kSvRow;
        if (with_q && h == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                svql[c * kSvRow + g] = Sq01[c].y;
                svql[c * kSvRow + kG + g] = Sq23[c].x;
                svql[c * kSvRow + 2 * kG + g] = Sq23[c].y;
            }
        }
        if (atom_ok) {
            float* xr = x + (size_t)i * ldx;
            const float4* own = reinterpret_cast<const float4*>(aT + (size_t)i * kAG) + g + 32 * h;
            float4 o0 = own[0], o1 = own[16];
            float ov[kHalfA] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
            for (int a = 0; a < kHalfA; ++a) {
                int aa = kHalfA * h + a;
                xr[aa * kG + g] = ov[a];
                xr[kAG + aa * kG + g] = S01[a].x;
            }
            int base = 2 * kAG + kAH;
            if (with_q) {
                if (lane < C) xr[base + lane] = q[(size_t)i * C + lane];
                if (h == 0) {
#pragma unroll
                    for (int c = 0; c < C; ++c) xr[base + C + c * kG + g] = Sq01[c].x;
                }
                base += C * (1 + kG + kH);
            }
            for (int c = base + lane; c < ldx; c += 32) xr[c] = 0.f;
        }
        __syncwarp();
        // T[a,h,k] = sum_g agh[a,g,h] * Sv[a,g,k]   (aimnet/modules/aev.py:188); each lane handles 6 (a,h) pairs: the 16
        // weights and the three 16-vectors of Sv come in as 16-byte shared loads, the dot products run as packed FFMA2
        if (atom_ok) {
            float* xr = x + (size_t)i * ldx;
#pragma unroll 2
            for (int e = lane; e < kAH; e += 32) {
                const int a = e / kH;
                float t[3];
                mix16(aghT_a + e * kAghRow, svl + sv_off(a), t);
                float* To = T_a + (size_t)i * kTA + e * 3;
                To[0] = t[0];
                To[1] = t[1];
                To[2] = t[2];
                xr[2 * kAG + e] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
            }
            if (with_q) {
                int base = 2 * kAG + kAH;
                for (int e = lane; e < C * kH; e += 32) {
                    const int c = e / kH;
                    float t[3];
                    mix16(aghT_q + e * kAghRow, svql + c * kSvRow, t);
                    float* To = T_q + (size_t)i * (C * kH * 3) + e * 3;
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
// backward step 1: per atom, turn d(loss)/d(x row) into d(loss)/dS^a and d(loss)/dS^q, stored transposed for the
// gather: dS_a (N,16 a,16 g,4), dS_q (N,C,16 g,4)
//   d avf_v[a,h] -> dT[a,h,d] = 2 T[a,h,d] * d avf_v[a,h] -> dSv[a,g,d] = sum_h agh[a,g,h] dT[a,h,d]
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) conv_bwd_prep_kernel(int n_atoms, int n_groups, const float* __restrict__ dx, int ldx,
                                                            const float* __restrict__ T_a,
                                                            const float* __restrict__ T_q,
                                                            const float* __restrict__ agh_a,
                                                            const float* __restrict__ agh_q, float* __restrict__ dS_a,
                                                            float* __restrict__ dS_q, int with_q, int permute) {
    // one warp per atom, 8 atoms per block.  agh is staged once per block in its natural (a,g,h) layout: a lane reads
    // the 12 weights of its (a,g) as three 16-byte loads (row stride 12 words: 8 consecutive rows cover all banks);
    // dT[a] (36 floats, row stride 36) comes in as nine 16-byte broadcast loads.
    __shared__ __align__(16) float agh_s[kA * kG * kH];          // 12 KB
    __shared__ __align__(16) float aghq_s[2 * kG * kH];
    __shared__ __align__(16) float dT_s[kAtomsPerCta][kTA];
    __shared__ __align__(16) float dTq_s[kAtomsPerCta][2 * kH * 3];
    for (int e = threadIdx.x; e < kA * kG * kH / 4; e += 256)
        reinterpret_cast<float4*>(agh_s)[e] = reinterpret_cast<const float4*>(agh_a)[e];
    if (with_q)
        for (int e = threadIdx.x; e < C * kG * kH; e += 256) aghq_s[e] = agh_q[e];
    const int al = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* dT = dT_s[al];
    float* dTq = dTq_s[al];
    const int base = 2 * kAG + kAH;
    __syncthreads();   // agh staged; from here on the warps of this persistent CTA run independently
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int i = grp * kAtomsPerCta + al;
    if (i >= n_atoms) continue;
    const float* dxr = dx + (size_t)i * ldx;
    __syncwarp();
    for (int e = lane; e < kTA; e += 32) dT[e] = 2.0f * T_a[(size_t)i * kTA + e] * dxr[2 * kAG + e / 3];
    if (with_q)
        for (int e = lane; e < C * kH * 3; e += 32)
            dTq[e] = 2.0f * T_q[(size_t)i * (C * kH * 3) + e] * dxr[base + C + C * kG + e / 3];
    __syncwarp();
#pragma unroll 2
    for (int k = 0; k < kAG / 32; ++k) {
        const int e = lane + 32 * k;
        const int aa = e >> 4;
        const float4* w4 = reinterpret_cast<const float4*>(agh_s + e * kH);
        const float4* t4 = reinterpret_cast<const float4*>(dT + aa * (kH * 3));
        const float4 wa = w4[0], wb = w4[1], wc = w4[2];
        const float w[kH] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w};
        float tv[kH * 3];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float4 v = t4[j];
            tv[4 * j + 0] = v.x;
            tv[4 * j + 1] = v.y;
            tv[4 * j + 2] = v.z;
            tv[4 * j + 3] = v.w;
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int h = 0; h < kH; ++h) {
            s0 = fmaf(w[h], tv[3 * h + 0], s0);
            s1 = fmaf(w[h], tv[3 * h + 1], s1);
            s2 = fmaf(w[h], tv[3 * h + 2], s2);
        }
        // permute: channel rows in the order the dense walk reads them (conv_dense.cu: the two channel halves of a warp read
        // neighbouring rows, i.e. different shared-memory banks): row = 2 t + h with h = bit 2 of a, t = a & 3 | (a >> 3) << 2
        const int row = permute ? 2 * ((aa & 3) + ((aa >> 3) << 2)) + ((aa >> 2) & 1) : aa;
        reinterpret_cast<float4*>(dS_a)[(size_t)i * kAG + row * kG + (e & 15)] = make_float4(dxr[kAG + e], s0, s1, s2);
    }
    if (with_q && lane < C * kG) {
        int cc = lane >> 4, gq = lane & 15;
        float4 oq;
        oq.x = dxr[base + C + lane];
        float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int h = 0; h < kH; ++h) {
            float w = aghq_s[(cc * kG + gq) * kH + h];
            q0 = fmaf(w, dTq[(cc * kH + h) * 3 + 0], q0);
            q1 = fmaf(w, dTq[(cc * kH + h) * 3 + 1], q1);
            q2 = fmaf(w, dTq[(cc * kH + h) * 3 + 2], q2);
        }
        oq.y = q0;
        oq.z = q1;
        oq.w = q2;
        reinterpret_cast<float4*>(dS_q)[(size_t)i * (C * kG) + lane] = oq;
    }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward step 2 (see the header comment)
// ------------------------------------------------------------------------------------------------------------
template <int C, bool kGradA, bool kVirial>
__global__ void __launch_bounds__(256, 2) conv_bwd_kernel(int n_atoms, NbView nb, const float* __restrict__ coord,
                                                          CellView cv, const int32_t* __restrict__ mol_idx,
                                                          AevParams aev, const float* __restrict__ aT,
                                                          const float* __restrict__ q, const float* __restrict__ dS_a,
                                                          const float* __restrict__ dS_q, float* __restrict__ grad_a,
                                                          float* __restrict__ grad_q, float* __restrict__ forces,
                                                          double* __restrict__ virial_atom, int with_q,
                                                          const int* __restrict__ skip_if) {
    __shared__ PairEntry tile[256];
    if (skip_if != nullptr && *skip_if != 0) return;   // the by-species kernels below have done this pass (uniform)
    const int tid = threadIdx.x;
    const int al = tid >> 5, lane = tid & 31, g = lane & 15, h = lane >> 4;
    const int i = blockIdx.x * kAtomsPerCta + al;
    const bool atom_ok = i < n_atoms;
    const int ic = atom_ok ? i : 0;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (mol_idx ? mol_idx[ic] : 0)) : nullptr;
    const int len = atom_ok ? row_length(nb, i) : 0;
    const float shift_g = aev.shifts[g];
    // own atom: dS_i[a][g][d] (as (scalar,x) / (y,z) register pairs) and a_i[a][g] for this thread's 8 channels
    float2 dSi01[kHalfA], dSi23[kHalfA];
    float ai[kHalfA];
    {
        const float4* p = reinterpret_cast<const float4*>(dS_a) + (size_t)ic * kAG + (kHalfA * h) * kG + g;
#pragma unroll
        for (int a = 0; a < kHalfA; ++a) {
            float4 v = p[a * kG];
            dSi01[a] = make_float2(v.x, v.y);
            dSi23[a] = make_float2(v.z, v.w);
        }
        const float4* r = reinterpret_cast<const float4*>(aT + (size_t)ic * kAG) + g + 32 * h;
        float4 o0 = r[0], o1 = r[16];
        float ov[kHalfA] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
        for (int a = 0; a < kHalfA; ++a) ai[a] = ov[a];
    }
    // charge channels (their partial sums are added once).  One channel: the h == 0 half handles it.  Two channels (NSE):
    // half h handles channel h -- the same instructions for both halves, only the addresses differ, so neither half idles
    // while the other works (slot 0 of the arrays below then holds "this lane's channel").
    const bool qhalf = with_q && (C == 2 || h == 0);
    const int cm = (C == 2) ? h : 0;         // channel of slot 0
    constexpr int CQ = (C == 2) ? 1 : C;     // channels per lane
    float2 dSqi01[CQ], dSqi23[CQ];
    float qi[CQ];
#pragma unroll
    for (int c = 0; c < CQ; ++c) {
        float4 v = qhalf ? reinterpret_cast<const float4*>(dS_q)[(size_t)ic * (C * kG) + (cm + c) * kG + g] : make_float4(0, 0, 0, 0);
        dSqi01[c] = make_float2(v.x, v.y);
        dSqi23[c] = make_float2(v.z, v.w);
        qi[c] = qhalf ? q[(size_t)ic * C + cm + c] : 0.f;
    }
    // grad_a / grad_q as two partial sums each (the halves of one packed accumulator), added at the end
    float2 ga2[kHalfA];
#pragma unroll
    for (int a = 0; a < kHalfA; ++a) ga2[a] = make_float2(0.f, 0.f);
    float2 gq2[CQ];
#pragma unroll
    for (int c = 0; c < CQ; ++c) gq2[c] = make_float2(0.f, 0.f);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    float vir[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) vir[k] = 0.f;

    for (int m0 = 0; m0 < len; m0 += kSlotsPerTile) {   // per-warp staging: see conv_fwd_kernel
        __syncwarp();
        stage_pairs<true>(tile, ic, atom_ok, m0, len, nb, coord, cell, aev);
        __syncwarp();
        int lim = min(kSlotsPerTile, len - m0);
        for (int s = 0; s < lim; ++s) {
            const PairEntry e = tile[al * 32 + s];
            const float4* arow = reinterpret_cast<const float4*>(aT + (size_t)e.j * kAG) + g + 32 * h;
            const float4* drow = reinterpret_cast<const float4*>(dS_a) + (size_t)e.j * kAG + (kHalfA * h) * kG + g;
            float4 v0 = arow[0], v1 = arow[16];
            float aj[kHalfA] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            float xg = e.d - shift_g;
            float ex = aev_exp(-aev.eta * xg * xg);
            float gs = ex * e.fc;
            float dgs = ex * (e.dfc - 2.0f * aev.eta * xg * e.fc);
            // g_sv(j->i)[g,:] = (gs, -gs u): grad_a[i] += <dS[j], g_sv(j->i)> as two packed FMAs per channel
            const float2 G01 = make_float2(gs, -gs * e.ux), G23 = make_float2(-gs * e.uy, -gs * e.uz);
            // p = contraction for slot (i -> j), r = for the reverse slot (j -> i); partial over this thread's channels
            float2 p01 = make_float2(0.f, 0.f), p23 = p01, r01 = p01, r23 = p01;
#pragma unroll
            for (int a = 0; a < kHalfA; ++a) {
                float4 dj = drow[a * kG];
                const float2 dj01 = make_float2(dj.x, dj.y), dj23 = make_float2(dj.z, dj.w);
                if (kGradA) {
                    ga2[a] = ffma2(dj01, G01, ga2[a]);
                    ga2[a] = ffma2(dj23, G23, ga2[a]);
                }
                p01 = ffma2s(aj[a], dSi01[a], p01);
                p23 = ffma2s(aj[a], dSi23[a], p23);
                r01 = ffma2s(ai[a], dj01, r01);
                r23 = ffma2s(ai[a], dj23, r23);
            }
            if (qhalf) {
#pragma unroll
                for (int c = 0; c < CQ; ++c) {
                    float qj = q[(size_t)e.j * C + cm + c];
                    float4 dqj = reinterpret_cast<const float4*>(dS_q)[(size_t)e.j * (C * kG) + (cm + c) * kG + g];
                    const float2 dq01 = make_float2(dqj.x, dqj.y), dq23 = make_float2(dqj.z, dqj.w);
                    if (kGradA) {
                        gq2[c] = ffma2(dq01, G01, gq2[c]);
                        gq2[c] = ffma2(dq23, G23, gq2[c]);
                    }
                    p01 = ffma2s(qj, dSqi01[c], p01);
                    p23 = ffma2s(qj, dSqi23[c], p23);
                    r01 = ffma2s(qi[c], dq01, r01);
                    r23 = ffma2s(qi[c], dq23, r23);
                }
            }
            const float gsi = gs * e.inv;
            // this thread's share of w(i->j) = u (A + C.u) + (B - u (B.u))/d
            float pu = p01.y * e.ux + p23.x * e.uy + p23.y * e.uz;
            float sc = (p01.x + pu) * dgs - pu * gsi;
            // reverse slot (j->i): u' = -u;  w' = -u (A' - (r.u) dgs) + (B' - u (B'.u))/d
            float ru = r01.y * e.ux + r23.x * e.uy + r23.y * e.uz;
            float scr = (ru - r01.x) * dgs - ru * gsi;
            if (!kVirial) {
                // F_i += w - w' = u (sc - scr) + (B - B') gs/d
                float ds = sc - scr;
                fx += fmaf(e.ux, ds, (p01.y - r01.y) * gsi);
                fy += fmaf(e.uy, ds, (p23.x - r23.x) * gsi);
                fz += fmaf(e.uz, ds, (p23.y - r23.y) * gsi);
            } else {
                float wx = e.ux * sc + p01.y * gsi;
                float wy = e.uy * sc + p23.x * gsi;
                float wz = e.uz * sc + p23.y * gsi;
                float vx = e.ux * scr + r01.y * gsi;
                float vy = e.uy * scr + r23.x * gsi;
                float vz = e.uz * scr + r23.y * gsi;
                fx += wx - vx;
                fy += wy - vy;
                fz += wz - vz;
                float rx = e.ux * e.d, ry = e.uy * e.d, rz = e.uz * e.d;
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
    // reduce the force / virial / grad_q shares over the 32 lanes of the atom
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (kVirial) {
#pragma unroll
        for (int k = 0; k < 9; ++k) vir[k] = warp_sum(vir[k]);
    }
    float gq[C];
#pragma unroll
    for (int c = 0; c < C; ++c) gq[c] = 0.f;
    if (kGradA && with_q) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float mine = (C == 2) ? ((h == c) ? gq2[0].x + gq2[0].y : 0.f) : gq2[C == 2 ? 0 : c].x + gq2[C == 2 ? 0 : c].y;
            gq[c] = warp_sum(mine);
        }
    }
    if (!atom_ok) return;
    if (kGradA) {
#pragma unroll
        for (int a = 0; a < kHalfA; ++a) grad_a[(size_t)i * kAG + (kHalfA * h + a) * kG + g] = ga2[a].x + ga2[a].y;
        if (with_q && lane < C) {
            float v = gq[0];
#pragma unroll
            for (int c = 1; c < C; ++c) v = (lane == c) ? gq[c] : v;
            grad_q[(size_t)i * C + lane] = v;
        }
    }
    if (lane == 0) {
        forces[3 * i + 0] += fx;
        forces[3 * i + 1] += fy;
        forces[3 * i + 2] += fz;
        if (kVirial)
#pragma unroll
            for (int k = 0; k < 9; ++k) virial_atom[(size_t)i * 9 + k] += (double)vir[k];
    }
}

// ------------------------------------------------------------------------------------------------------------
// operator seam: conv_sv_2d_sp with an explicit g tensor (aimnet/kernels/conv_sv_2d_sp_wp.py:90-164)
// ------------------------------------------------------------------------------------------------------------
__global__ void conv_op_fwd_kernel(const float* __restrict__ a, const int32_t* __restrict__ idx,
                                   const float* __restrict__ g, float* __restrict__ out, int B, int A, int G, int M) {
    int b = blockIdx.x;
    int AG = A * G;
    for (int e = threadIdx.x; e < AG; e += blockDim.x) {
        int gg = e % G;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B - 1) {
            for (int m = 0; m < M; ++m) {
                int j = idx[(size_t)b * M + m];
                if (j >= B - 1 || j < 0) continue;
                float av = a[(size_t)j * AG + e];
                float4 gv = reinterpret_cast<const float4*>(g)[((size_t)b * M + m) * G + gg];
                acc.x += av * gv.x;
                acc.y += av * gv.y;
                acc.z += av * gv.z;
                acc.w += av * gv.w;
            }
        }
        reinterpret_cast<float4*>(out)[(size_t)b * AG + e] = acc;
    }
}

// grad_g[b,m,g,:] = sum_a a[idx[b,m],a,g] * grad_out[b,a,g,:]
__global__ void conv_op_bwd_g_kernel(const float* __restrict__ grad_out, const float* __restrict__ a,
                                     const int32_t* __restrict__ idx, float* __restrict__ grad_g, int B, int A, int G,
                                     int M) {
    int b = blockIdx.x;
    for (int e = threadIdx.x; e < M * G; e += blockDim.x) {
        int m = e / G, gg = e % G;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = idx[(size_t)b * M + m];
        if (b < B - 1 && j < B - 1 && j >= 0) {
            for (int aa = 0; aa < A; ++aa) {
                float av = a[((size_t)j * A + aa) * G + gg];
                float4 go = reinterpret_cast<const float4*>(grad_out)[((size_t)b * A + aa) * G + gg];
                acc.x += av * go.x;
                acc.y += av * go.y;
                acc.z += av * go.z;
                acc.w += av * go.w;
            }
        }
        reinterpret_cast<float4*>(grad_g)[((size_t)b * M + m) * G + gg] = acc;
    }
}

// grad_a[idx[b,m],a,g] += <grad_out[b,a,g,:], g[b,m,g,:]>  — arbitrary (possibly asymmetric, repeated) idx, so this
// seam keeps the scatter form; red.global.add.f32 per element.
__global__ void conv_op_bwd_a_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
                                     const float* __restrict__ g, float* __restrict__ grad_a, int B, int A, int G,
                                     int M) {
    int b = blockIdx.x;
    if (b >= B - 1) return;
    int AG = A * G;
    for (int e = threadIdx.x; e < AG; e += blockDim.x) {
        int gg = e % G;
        float4 go = reinterpret_cast<const float4*>(grad_out)[(size_t)b * AG + e];
        for (int m = 0; m < M; ++m) {
            int j = idx[(size_t)b * M + m];
            if (j >= B - 1 || j < 0) continue;
            float4 gv = reinterpret_cast<const float4*>(g)[((size_t)b * M + m) * G + gg];
            atomicAdd(&grad_a[(size_t)j * AG + e], go.x * gv.x + go.y * gv.y + go.z * gv.z + go.w * gv.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
template <int C>
static int conv_fwd_launch(int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                           const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q,
                           const float* agh_a, const float* agh_q, float* x, int ldx, float* T_a, float* T_q,
                           int with_q, cudaStream_t st) {
    const int n_groups = (n_atoms + kAtomsPerCta - 1) / kAtomsPerCta;
    static bool configured_dev[kMaxDevices] = {};
    static int max_ctas_dev[kMaxDevices] = {};
    const int dslot = current_device_slot();
    int& max_ctas = max_ctas_dev[dslot];
    if (!configured_dev[dslot]) {
        AIM_CUDA_CHECK(cudaFuncSetAttribute(conv_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
        int dev = 0, sms = 148, per_sm = 3;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        AIM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_fwd_kernel<C>, 256, kFwdSmemBytes));
        max_ctas = sms * (per_sm > 0 ? per_sm : 1);
        configured_dev[dslot] = true;
    }
    // persistent CTAs (one resident wave): the agh tables are staged once per CTA, not once per 8 atoms
    const int grid = n_groups < max_ctas ? n_groups : max_ctas;
    conv_fwd_kernel<C><<<grid, 256, kFwdSmemBytes, st>>>(n_atoms, n_groups, nb, coord, cv, mol_idx, aev, aT, q, agh_a, agh_q,
                                                        x, ldx, T_a, T_q, with_q);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

int launch_conv_fwd(int C, int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                    const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q, const float* agh_a,
                    const float* agh_q, float* x, int ldx, float* T_a, float* T_q, int with_q, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    if (C == 1)
        return conv_fwd_launch<1>(n_atoms, nb, coord, cv, mol_idx, aev, aT, q, agh_a, agh_q, x, ldx, T_a, T_q, with_q, st);
    return conv_fwd_launch<2>(n_atoms, nb, coord, cv, mol_idx, aev, aT, q, agh_a, agh_q, x, ldx, T_a, T_q, with_q, st);
}

template <int C>
static int conv_bwd_prep_launch(int n_atoms, const float* dx, int ldx, const float* T_a, const float* T_q, const float* agh_a,
                                const float* agh_q, float* dS_a, float* dS_q, int with_q, int permute, cudaStream_t st) {
    const int n_groups = (n_atoms + kAtomsPerCta - 1) / kAtomsPerCta;
    static int prep_ctas_dev[kMaxDevices] = {};
    int& prep_ctas = prep_ctas_dev[current_device_slot()];
    if (prep_ctas == 0) {
        int dev = 0, sms = 148, per_sm = 4;
        AIM_CUDA_CHECK(cudaGetDevice(&dev));
        AIM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        AIM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_bwd_prep_kernel<C>, 256, 0));
        prep_ctas = sms * (per_sm > 0 ? per_sm : 1);
    }
    conv_bwd_prep_kernel<C><<<n_groups < prep_ctas ? n_groups : prep_ctas, 256, 0, st>>>(n_atoms, n_groups, dx, ldx, T_a, T_q,
                                                                                          agh_a, agh_q, dS_a, dS_q, with_q, permute);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

// backward step 1 alone (the gather step is launch_conv_bwd's second half or conv2.cu's launch_conv2_bwd_gather)
int launch_conv_bwd_prep(int C, int n_atoms, const float* dx, int ldx, const float* T_a, const float* T_q, const float* agh_a,
                         const float* agh_q, float* dS_a, float* dS_q, int with_q, int permute, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    if (C == 1) return conv_bwd_prep_launch<1>(n_atoms, dx, ldx, T_a, T_q, agh_a, agh_q, dS_a, dS_q, with_q, permute, st);
    return conv_bwd_prep_launch<2>(n_atoms, dx, ldx, T_a, T_q, agh_a, agh_q, dS_a, dS_q, with_q, permute, st);
}

template <int C>
static int conv_bwd_launch(int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                           const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q,
                           const float* dx, int ldx, const float* T_a, const float* T_q, const float* agh_a,
                           const float* agh_q, float* dS_a, float* dS_q, float* grad_a, float* grad_q, float* forces,
                           double* virial_atom, int with_q, int want_grad_a, const int* skip_if, bool prep, cudaStream_t st) {
    const int n_groups = (n_atoms + kAtomsPerCta - 1) / kAtomsPerCta;
    if (prep) AIM_TRY(conv_bwd_prep_launch<C>(n_atoms, dx, ldx, T_a, T_q, agh_a, agh_q, dS_a, dS_q, with_q, 0, st));
    int grid = n_groups;
#define AIM_CONV_BWD(GA, VIR)                                                                                       \
    conv_bwd_kernel<C, GA, VIR><<<grid, 256, 0, st>>>(n_atoms, nb, coord, cv, mol_idx, aev, aT, q, dS_a, dS_q, grad_a, \
                                                      grad_q, forces, virial_atom, with_q, skip_if)
    if (want_grad_a) {
        if (virial_atom) AIM_CONV_BWD(true, true); else AIM_CONV_BWD(true, false);
    } else {
        if (virial_atom) AIM_CONV_BWD(false, true); else AIM_CONV_BWD(false, false);
    }
#undef AIM_CONV_BWD
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

// `forces` must not be written concurrently by another stream: every atom's total is accumulated with a plain +=
int launch_conv_bwd(int C, int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                    const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q, const float* dx,
                    int ldx, const float* T_a, const float* T_q, const float* agh_a, const float* agh_q, float* dS_a,
                    float* dS_q, float* grad_a, float* grad_q, float* forces, double* virial_atom, int with_q,
                    int want_grad_a, cudaStream_t st, const int* skip_if, bool prep) {
    if (n_atoms == 0) return AIMNET_OK;
    if (C == 1)
        return conv_bwd_launch<1>(n_atoms, nb, coord, cv, mol_idx, aev, aT, q, dx, ldx, T_a, T_q, agh_a, agh_q, dS_a, dS_q,
                                  grad_a, grad_q, forces, virial_atom, with_q, want_grad_a, skip_if, prep, st);
    return conv_bwd_launch<2>(n_atoms, nb, coord, cv, mol_idx, aev, aT, q, dx, ldx, T_a, T_q, agh_a, agh_q, dS_a, dS_q,
                              grad_a, grad_q, forces, virial_atom, with_q, want_grad_a, skip_if, prep, st);
}

// ------------------------------------------------------------------------------------------------------------
// Backward of the FIRST convolution by species.  In pass 0 the convolved features are the embedding, a[j] = afv[Z_j]
// (aimnet/models/aimnet2.py:144-147): they depend on the neighbour only through its species, they need no gradient, and
// there are no charge channels yet.  The two 1 024-FMA contractions the generic kernel does per pair,
//     p[g,d] = sum_a a[j,a,g] dS[i,a,g,d]        r[g,d] = sum_a a[i,a,g] dS[j,a,g,d],
// are therefore table look-ups  p = P[i][slot(Z_j)],  r = P[j][slot(Z_i)]  into
//     P[i][s][g][d] = sum_a afv[z_s][a,g] dS[i,a,g,d]      (one contraction per atom and PRESENT species),
// and the pair kernel reads 256 bytes of the neighbour's table instead of its 1 KB of features and 4 KB of dS: half a
// warp per pair, lane = g.  Species slots are assigned on the device (no host round trip); with more than kMaxSlots
// species in one evaluation the flag stays 0, these kernels return at once and the generic kernel runs instead (the engine
// launches it behind them, with skip_if, only for models that implement more than kMaxSlots species).
// ------------------------------------------------------------------------------------------------------------
constexpr int kMaxSlots = 16;
// info: [0] 1 = by-species pass valid, [1] number of slots, [2 .. 2 + kMaxSlots) atomic number of a slot,
//       [18 .. 82) slot of an atomic number, [82], [83] presence mask of the atomic numbers, [84] block counter
constexpr int kInfoSlotOfZ = 2 + kMaxSlots, kInfoMask = kInfoSlotOfZ + 64, kInfoCount = kInfoMask + 2;
constexpr int kSpeciesInfoInts = kInfoCount + 1;
__global__ void __launch_bounds__(256) species_mask_kernel(int n, const int32_t* __restrict__ numbers, int* __restrict__ info) {
    // presence mask over all atoms; the block that finishes last turns it into the slot tables
    __shared__ unsigned int mask[2];
    __shared__ bool last;
    if (threadIdx.x < 2) mask[threadIdx.x] = 0u;
    __syncthreads();
    unsigned int m0 = 0u, m1 = 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int z = numbers[i];
        z = (z < 0 || z > 63) ? 0 : z;
        if (z < 32) m0 |= 1u << z; else m1 |= 1u << (z - 32);
    }
    m0 = __reduce_or_sync(0xffffffffu, m0);
    m1 = __reduce_or_sync(0xffffffffu, m1);
    if ((threadIdx.x & 31) == 0) {
        if (m0) atomicOr(&mask[0], m0);
        if (m1) atomicOr(&mask[1], m1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mask[0]) atomicOr(reinterpret_cast<unsigned int*>(info) + kInfoMask, mask[0]);
        if (mask[1]) atomicOr(reinterpret_cast<unsigned int*>(info) + kInfoMask + 1, mask[1]);
        __threadfence();
        last = atomicAdd(info + kInfoCount, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    const unsigned int g0 = atomicOr(reinterpret_cast<unsigned int*>(info) + kInfoMask, 0u);
    const unsigned int g1 = atomicOr(reinterpret_cast<unsigned int*>(info) + kInfoMask + 1, 0u);
    int ns = 0;
    for (int z = 0; z < 64; ++z) {
        const bool present = ((z < 32 ? (g0 >> z) : (g1 >> (z - 32))) & 1u) != 0u;
        int slot = 0;
        if (present) {
            if (ns < kMaxSlots) {
                slot = ns;
                info[2 + ns] = z;
            }
            ++ns;
        }
        info[kInfoSlotOfZ + z] = slot;
    }
    info[1] = ns <= kMaxSlots ? ns : kMaxSlots;
    info[0] = (ns <= kMaxSlots) ? 1 : 0;
}
__global__ void __launch_bounds__(256) species_slot_kernel(int n, const int32_t* __restrict__ numbers, const int* __restrict__ info,
                                                           uint8_t* __restrict__ atom_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int z = numbers[i];
    z = (z < 0 || z > 63) ? 0 : z;
    atom_slot[i] = (uint8_t)info[kInfoSlotOfZ + z];
}

// P[i][s][g][:] for the present species; warp = atom, lane = (channel half h, g) as in conv_bwd_kernel
__global__ void __launch_bounds__(256) conv0_table_kernel(int n_atoms, const int* __restrict__ info,
                                                          const float* __restrict__ afvT, const float* __restrict__ dS_a,
                                                          float4* __restrict__ P) {
    if (info[0] == 0) return;
    const int al = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane & 15, h = lane >> 4;
    const int i = blockIdx.x * kAtomsPerCta + al;
    if (i >= n_atoms) return;
    const int n_slots = info[1];
    float2 d01[kHalfA], d23[kHalfA];
    const float4* p = reinterpret_cast<const float4*>(dS_a) + (size_t)i * kAG + (kHalfA * h) * kG + g;
#pragma unroll
    for (int a = 0; a < kHalfA; ++a) {
        const float4 v = p[a * kG];
        d01[a] = make_float2(v.x, v.y);
        d23[a] = make_float2(v.z, v.w);
    }
    for (int s = 0; s < n_slots; ++s) {
        const float4* r = reinterpret_cast<const float4*>(afvT + (size_t)info[2 + s] * kAG) + g + 32 * h;
        const float4 o0 = r[0], o1 = r[16];
        const float av[kHalfA] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        float2 p01 = make_float2(0.f, 0.f), p23 = p01;
#pragma unroll
        for (int a = 0; a < kHalfA; ++a) {
            p01 = ffma2s(av[a], d01[a], p01);
            p23 = ffma2s(av[a], d23[a], p23);
        }
        // channels 0-7 (h = 0) + channels 8-15 (h = 1), in this fixed order
        const float x0 = __shfl_xor_sync(0xffffffffu, p01.x, 16), x1 = __shfl_xor_sync(0xffffffffu, p01.y, 16);
        const float x2 = __shfl_xor_sync(0xffffffffu, p23.x, 16), x3 = __shfl_xor_sync(0xffffffffu, p23.y, 16);
        if (h == 0) P[((size_t)i * kMaxSlots + s) * kG + g] = make_float4(p01.x + x0, p01.y + x1, p23.x + x2, p23.y + x3);
    }
}

template <bool kVirial>
__global__ void __launch_bounds__(256) conv0_force_kernel(int n_atoms, NbView nb, const float* __restrict__ coord, CellView cv,
                                                          const int32_t* __restrict__ mol_idx, AevParams aev,
                                                          const int* __restrict__ info,
                                                          const uint8_t* __restrict__ atom_slot,
                                                          const float4* __restrict__ P, float* __restrict__ forces,
                                                          double* __restrict__ virial_atom) {
    __shared__ PairEntry tile[256];
    __shared__ float4 own[kAtomsPerCta][kMaxSlots * kG];   // the centre's own table: 4 KB per warp
    if (info[0] == 0) return;
    const int tid = threadIdx.x;
    const int al = tid >> 5, lane = tid & 31, g = lane & 15, hw = lane >> 4;
    const int i = blockIdx.x * kAtomsPerCta + al;
    const bool atom_ok = i < n_atoms;
    const int ic = atom_ok ? i : 0;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (mol_idx ? mol_idx[ic] : 0)) : nullptr;
    const int len = atom_ok ? row_length(nb, i) : 0;
    const float shift_g = aev.shifts[g];
    const int n_slots = info[1];
    for (int k = lane; k < n_slots * kG; k += 32) own[al][k] = P[(size_t)ic * kMaxSlots * kG + k];
    const int zi = atom_slot[ic];
    float fx = 0.f, fy = 0.f, fz = 0.f;
    float vir[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) vir[k] = 0.f;
    for (int m0 = 0; m0 < len; m0 += kSlotsPerTile) {
        __syncwarp();
        stage_pairs<true>(tile, ic, atom_ok, m0, len, nb, coord, cell, aev);
        __syncwarp();
        const int lim = min(kSlotsPerTile, len - m0);
        for (int s = hw; s < lim; s += 2) {   // half a warp per pair
            const PairEntry e = tile[al * 32 + s];
            const float4 r = __ldg(P + ((size_t)e.j * kMaxSlots + zi) * kG + g);
            const float4 p = own[al][(int)atom_slot[e.j] * kG + g];
            const float xg = e.d - shift_g;
            const float ex = aev_exp(-aev.eta * xg * xg);
            const float gs = ex * e.fc;
            const float dgs = ex * (e.dfc - 2.0f * aev.eta * xg * e.fc);
            const float gsi = gs * e.inv;
            // w(i->j) = u (A + C.u) + (B - u (B.u))/d  and the reverse slot with u' = -u: see conv_bwd_kernel
            const float pu = p.y * e.ux + p.z * e.uy + p.w * e.uz;
            const float sc = (p.x + pu) * dgs - pu * gsi;
            const float ru = r.y * e.ux + r.z * e.uy + r.w * e.uz;
            const float scr = (ru - r.x) * dgs - ru * gsi;
            if (!kVirial) {
                const float ds = sc - scr;
                fx += fmaf(e.ux, ds, (p.y - r.y) * gsi);
                fy += fmaf(e.uy, ds, (p.z - r.z) * gsi);
                fz += fmaf(e.uz, ds, (p.w - r.w) * gsi);
            } else {
                const float wx = e.ux * sc + p.y * gsi, wy = e.uy * sc + p.z * gsi, wz = e.uz * sc + p.w * gsi;
                const float vx = e.ux * scr + r.y * gsi, vy = e.uy * scr + r.z * gsi, vz = e.uz * scr + r.w * gsi;
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
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (kVirial) {
#pragma unroll
        for (int k = 0; k < 9; ++k) vir[k] = warp_sum(vir[k]);
    }
    if (!atom_ok || lane != 0) return;
    forces[3 * i + 0] += fx;
    forces[3 * i + 1] += fy;
    forces[3 * i + 2] += fz;
    if (kVirial)
#pragma unroll
        for (int k = 0; k < 9; ++k) virial_atom[(size_t)i * 9 + k] += (double)vir[k];
}

int launch_species_scan(int n_atoms, const int32_t* numbers, int* info, uint8_t* atom_slot, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    AIM_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int) * kSpeciesInfoInts, st));
    int blocks = (n_atoms + 1023) / 1024;   // four atoms per thread
    if (blocks > 296) blocks = 296;
    species_mask_kernel<<<blocks, 256, 0, st>>>(n_atoms, numbers, info);
    AIM_LAUNCH_CHECK();
    species_slot_kernel<<<(n_atoms + 255) / 256, 256, 0, st>>>(n_atoms, numbers, info, atom_slot);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

// force (and virial) of the first convolution's backward from dS_a (already prepared), by species; returns at once on the
// device when `info` says that the evaluation has too many species (the caller then also launches the generic kernel with
// skip_if = info, which returns at once in the other case)
int launch_conv0_bwd_species(int n_atoms, const NbView& nb, const float* coord, const CellView& cv, const int32_t* mol_idx,
                             const AevParams& aev, const int* info, const uint8_t* atom_slot, const float* afvT,
                             const float* dS_a, float* P, float* forces, double* virial_atom, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    const int n_groups = (n_atoms + kAtomsPerCta - 1) / kAtomsPerCta;
    conv0_table_kernel<<<n_groups, 256, 0, st>>>(n_atoms, info, afvT, dS_a, reinterpret_cast<float4*>(P));
    AIM_LAUNCH_CHECK();
    if (virial_atom)
        conv0_force_kernel<true><<<n_groups, 256, 0, st>>>(n_atoms, nb, coord, cv, mol_idx, aev, info, atom_slot,
                                                          reinterpret_cast<const float4*>(P), forces, virial_atom);
    else
        conv0_force_kernel<false><<<n_groups, 256, 0, st>>>(n_atoms, nb, coord, cv, mol_idx, aev, info, atom_slot,
                                                           reinterpret_cast<const float4*>(P), forces, virial_atom);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}
int conv0_species_bytes_per_atom() { return kMaxSlots * kG * 16; }
int conv0_species_info_ints() { return kSpeciesInfoInts; }
int conv0_species_max_slots() { return kMaxSlots; }

}  // namespace aimnet

extern "C" int aimnet2_conv_sv_2d_sp_fwd(const float* a, const int32_t* idx, const float* g, float* out, int B, int A,
                                         int G, int M, void* stream) {
    using namespace aimnet;
    AIM_REQUIRE(a && idx && g && out, "conv_sv_2d_sp_fwd: null pointer");
    AIM_REQUIRE(B >= 1 && A >= 1 && G >= 1 && M >= 0, "conv_sv_2d_sp_fwd: bad shape");
    conv_op_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(a, idx, g, out, B, A, G, M);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

extern "C" int aimnet2_conv_sv_2d_sp_bwd(const float* grad_out, const float* a, const int32_t* idx, const float* g,
                                         float* grad_a, float* grad_g, int B, int A, int G, int M, void* stream) {
    using namespace aimnet;
    AIM_REQUIRE(grad_out && a && idx && g && grad_a && grad_g, "conv_sv_2d_sp_bwd: null pointer");
    AIM_REQUIRE(B >= 1 && A >= 1 && G >= 1 && M >= 0, "conv_sv_2d_sp_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    AIM_CUDA_CHECK(cudaMemsetAsync(grad_a, 0, sizeof(float) * (size_t)B * A * G, st));
    conv_op_bwd_a_kernel<<<B, 256, 0, st>>>(grad_out, idx, g, grad_a, B, A, G, M);
    AIM_LAUNCH_CHECK();
    if (M > 0) {
        conv_op_bwd_g_kernel<<<B, 256, 0, st>>>(grad_out, a, idx, grad_g, B, A, G, M);
        AIM_LAUNCH_CHECK();
    }
    return AIMNET_OK;
}
