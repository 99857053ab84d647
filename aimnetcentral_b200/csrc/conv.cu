// AEV + conv_sv message passing, forward and analytic backward.
//
// Reference semantics: calc_distances (aimnet/ops.py:37-66), AEVSV._calc_aev (aimnet/modules/aev.py:94-110),
// ConvSV.forward (aimnet/modules/aev.py:156-189), Warp kernels aimnet/kernels/conv_sv_2d_sp_wp.py:90-164.
// The reference materialises g_sv (N,M,16,4) in HBM (256 B per pair) and re-reads it for every convolution and for
// the autograd pass; here the radial basis is recomputed per neighbour tile in shared memory and never stored.
//
// Layout of one MLP input row x (ld = ldx, zero padded):
//   [0,256)   a[i] flat (a*16+g)            [256,512) S_s[a,g]          [512,704) avf_v[a,h]
//   pass>0:   [704,704+C) q[i,c]            [704+C, 704+17C) Sq_s[c,g]  [704+17C, 704+29C) avfq_v[c,h]
#include "common.cuh"

namespace aimnet {

constexpr int kTile = 32;   // neighbour slots staged per shared-memory tile

struct PairTile {
    int j[kTile];          // neighbour index (0 when slot invalid)
    float valid[kTile];    // 1 / 0
    float d[kTile];
    float u[kTile][3];
    float gs[kTile][kG];   // radial basis * cutoff   (0 for invalid slots)
    float dgs[kTile][kG];  // d gs / d d
};

// stage geometry + radial basis for slots [m0, m0+kTile) of centre atom i; all 256 threads participate
template <bool kWithDeriv>
__device__ __forceinline__ void stage_tile(PairTile& t, int i, int m0, int row_len, const NbView& nb,
                                           const float* __restrict__ coord, const float* __restrict__ cell,
                                           const AevParams& aev) {
    int tid = threadIdx.x;
    if (tid < kTile) {
        int m = m0 + tid;
        int j = nb.sentinel;
        if (m < row_len) j = nb.nbmat[(size_t)i * nb.width + m];
        bool ok = (j != nb.sentinel) && (j >= 0);
        float rx = 1.f, ry = 1.f, rz = 1.f;
        if (ok) {
            const int32_t* sh = nb.shifts ? nb.shifts + ((size_t)i * nb.width + m) * 3 : nullptr;
            pair_vector(coord, i, j, sh, cell, rx, ry, rz);
        }
        float d = sqrtf(rx * rx + ry * ry + rz * rz);
        float inv = 1.0f / d;
        t.j[tid] = ok ? j : 0;
        t.valid[tid] = ok ? 1.f : 0.f;
        t.d[tid] = d;
        t.u[tid][0] = rx * inv;
        t.u[tid][1] = ry * inv;
        t.u[tid][2] = rz * inv;
    }
    __syncthreads();
    // 32 slots x 16 shifts = 512 basis values, 2 per thread
#pragma unroll
    for (int k = 0; k < (kTile * kG) / 256; ++k) {
        int e = tid + 256 * k;
        int s = e >> 4, g = e & 15;
        float d = t.d[s];
        float v = t.valid[s];
        // cosine cutoff, aimnet/ops.py:82-85
        float dc = fminf(fmaxf(d, 1e-6f), aev.rc);
        float arg = dc * (kPi / aev.rc);
        float sn, cs;
        sincosf(arg, &sn, &cs);
        float fc = 0.5f * (cs + 1.0f) * v;
        float x = d - aev.shifts[g];
        float ex = expf(-aev.eta * x * x);
        t.gs[s][g] = ex * fc;
        if (kWithDeriv) {
            float dfc = (d > 1e-6f && d < aev.rc) ? (-0.5f * (kPi / aev.rc) * sn * v) : 0.f;
            t.dgs[s][g] = ex * (dfc - 2.0f * aev.eta * x * fc);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int row_length(const NbView& nb, int i) {
    return nb.count ? min(nb.count[i], nb.width) : nb.width;
}

// ------------------------------------------------------------------------------------------------------------
// forward: one block (256 threads = (a,g)) per atom
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) conv_fwd_kernel(int n_atoms, NbView nb, const float* __restrict__ coord,
                                                       CellView cv, const int32_t* __restrict__ mol_idx,
                                                       AevParams aev, const float* __restrict__ a,
                                                       const float* __restrict__ q, const float* __restrict__ agh_a,
                                                       const float* __restrict__ agh_q, float* __restrict__ x,
                                                       int ldx, float* __restrict__ T_a, float* __restrict__ T_q,
                                                       int with_q) {
    __shared__ PairTile tile;
    __shared__ float sv[kAG][3];       // vector part of S^a
    __shared__ float svq[2 * kG][3];   // vector part of S^q
    __shared__ float sT[kTA];
    __shared__ float sTq[2 * kH * 3];
    int i = blockIdx.x;
    int tid = threadIdx.x;
    int g = tid & 15;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (mol_idx ? mol_idx[i] : 0)) : nullptr;
    int len = row_length(nb, i);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    float qa0 = 0.f, qa1 = 0.f, qa2 = 0.f, qa3 = 0.f;
    bool qthread = with_q && tid < C * kG;
    int qc = tid >> 4;
    for (int m0 = 0; m0 < len; m0 += kTile) {
        stage_tile<false>(tile, i, m0, len, nb, coord, cell, aev);
        int lim = min(kTile, len - m0);
#pragma unroll 4
        for (int s = 0; s < lim; ++s) {
            int j = tile.j[s];
            float w = tile.gs[s][g];
            float aj = a[(size_t)j * kAG + tid];
            float p = aj * w;
            acc0 += p;
            acc1 += p * tile.u[s][0];
            acc2 += p * tile.u[s][1];
            acc3 += p * tile.u[s][2];
            if (qthread) {
                float pq = q[(size_t)j * C + qc] * w;
                qa0 += pq;
                qa1 += pq * tile.u[s][0];
                qa2 += pq * tile.u[s][1];
                qa3 += pq * tile.u[s][2];
            }
        }
        __syncthreads();
    }
    sv[tid][0] = acc1;
    sv[tid][1] = acc2;
    sv[tid][2] = acc3;
    if (qthread) {
        svq[tid][0] = qa1;
        svq[tid][1] = qa2;
        svq[tid][2] = qa3;
    }
    __syncthreads();
    // T[a,h,d] = sum_g agh[a,g,h] * Sv[a,g,d]      (aimnet/modules/aev.py:188)
    for (int e = tid; e < kTA; e += 256) {
        int aa = e / (kH * 3), rem = e % (kH * 3);
        int h = rem / 3, d = rem % 3;
        float s = 0.f;
#pragma unroll
        for (int gg = 0; gg < kG; ++gg) s += agh_a[(aa * kG + gg) * kH + h] * sv[aa * kG + gg][d];
        sT[e] = s;
        T_a[(size_t)i * kTA + e] = s;
    }
    if (with_q) {
        for (int e = tid; e < C * kH * 3; e += 256) {
            int cc = e / (kH * 3), rem = e % (kH * 3);
            int h = rem / 3, d = rem % 3;
            float s = 0.f;
#pragma unroll
            for (int gg = 0; gg < kG; ++gg) s += agh_q[(cc * kG + gg) * kH + h] * svq[cc * kG + gg][d];
            sTq[e] = s;
            T_q[(size_t)i * (C * kH * 3) + e] = s;
        }
    }
    __syncthreads();
    float* xr = x + (size_t)i * ldx;
    xr[tid] = a[(size_t)i * kAG + tid];
    xr[kAG + tid] = acc0;
    if (tid < kAH) {
        float t0 = sT[tid * 3], t1 = sT[tid * 3 + 1], t2 = sT[tid * 3 + 2];
        xr[2 * kAG + tid] = t0 * t0 + t1 * t1 + t2 * t2;
    }
    int base = 2 * kAG + kAH;   // 704
    if (with_q) {
        if (tid < C) xr[base + tid] = q[(size_t)i * C + tid];
        if (qthread) xr[base + C + tid] = qa0;
        if (tid < C * kH) {
            float t0 = sTq[tid * 3], t1 = sTq[tid * 3 + 1], t2 = sTq[tid * 3 + 2];
            xr[base + C + C * kG + tid] = t0 * t0 + t1 * t1 + t2 * t2;
        }
        base += C * (1 + kG + kH);
    }
    for (int c = base + tid; c < ldx; c += 256) xr[c] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// backward step 1: per atom, turn d(loss)/d(x row) into d(loss)/dS^a (16,16,4) and d(loss)/dS^q (C,16,4)
//   d avf_v[a,h] -> dT[a,h,d] = 2 T[a,h,d] * d avf_v[a,h] -> dSv[a,g,d] = sum_h agh[a,g,h] dT[a,h,d]
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) conv_bwd_prep_kernel(int n_atoms, const float* __restrict__ dx, int ldx,
                                                            const float* __restrict__ T_a,
                                                            const float* __restrict__ T_q,
                                                            const float* __restrict__ agh_a,
                                                            const float* __restrict__ agh_q, float* __restrict__ dS_a,
                                                            float* __restrict__ dS_q, int with_q) {
    __shared__ float dT[kTA];
    __shared__ float dTq[2 * kH * 3];
    int i = blockIdx.x, tid = threadIdx.x;
    const float* dxr = dx + (size_t)i * ldx;
    for (int e = tid; e < kTA; e += 256) dT[e] = 2.0f * T_a[(size_t)i * kTA + e] * dxr[2 * kAG + e / 3];
    int base = 2 * kAG + kAH;
    if (with_q)
        for (int e = tid; e < C * kH * 3; e += 256)
            dTq[e] = 2.0f * T_q[(size_t)i * (C * kH * 3) + e] * dxr[base + C + C * kG + e / 3];
    __syncthreads();
    int aa = tid >> 4, g = tid & 15;
    float4 o;
    o.x = dxr[kAG + tid];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int h = 0; h < kH; ++h) {
        float w = agh_a[(aa * kG + g) * kH + h];
        s0 += w * dT[(aa * kH + h) * 3 + 0];
        s1 += w * dT[(aa * kH + h) * 3 + 1];
        s2 += w * dT[(aa * kH + h) * 3 + 2];
    }
    o.y = s0;
    o.z = s1;
    o.w = s2;
    reinterpret_cast<float4*>(dS_a)[(size_t)i * kAG + tid] = o;
    if (with_q && tid < C * kG) {
        int cc = tid >> 4;
        float4 oq;
        oq.x = dxr[base + C + tid];
        float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int h = 0; h < kH; ++h) {
            float w = agh_q[(cc * kG + g) * kH + h];
            q0 += w * dTq[(cc * kH + h) * 3 + 0];
            q1 += w * dTq[(cc * kH + h) * 3 + 1];
            q2 += w * dTq[(cc * kH + h) * 3 + 2];
        }
        oq.y = q0;
        oq.z = q1;
        oq.w = q2;
        reinterpret_cast<float4*>(dS_q)[(size_t)i * (C * kG) + tid] = oq;
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward step 2: one block per atom i.  Over the (symmetric, full) neighbour row of i:
//   (1) grad_a[i,a,g]  = sum_m  <dS^a[j_m,a,g,:], g_sv(j_m -> i)[g,:]>     g_sv(j->i) = (gs, -gs*u_{i->j})
//   (2) grad_q[i,c]    = sum_m sum_g <dS^q[j_m,c,g,:], g_sv(j_m -> i)[g,:]>
//   (3) w_im = dE/dr_im through g_sv(i -> j_m): P[g,d] = sum_a a[j,a,g] dS^a[i,a,g,d] + sum_c q[j,c] dS^q[i,c,g,d]
//       dE/dr = u*(A + C.u) + (B - u (B.u))/d,  A = sum_g P[g,0] gs'_g, B_k = sum_g P[g,1+k] gs_g, C_k = sum_g P[g,1+k] gs'_g
//   forces: F_i += w_im, F_j -= w_im  (r_ij = x_j + s.cell - x_i);  virial_i += r_im (x) w_im
// (1),(2) are the gather form of the reference's atomic scatter kernel conv_sv_2d_sp_wp.py:115-136: with a full
// list every pair appears in both rows, so the contribution of centre j to neighbour i can be evaluated from i's
// own row -> no atomics, deterministic.
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) conv_bwd_kernel(int n_atoms, NbView nb, const float* __restrict__ coord,
                                                       CellView cv, const int32_t* __restrict__ mol_idx,
                                                       AevParams aev, const float* __restrict__ a,
                                                       const float* __restrict__ q, const float* __restrict__ dS_a,
                                                       const float* __restrict__ dS_q, float* __restrict__ grad_a,
                                                       float* __restrict__ grad_q, float* __restrict__ forces,
                                                       double* __restrict__ virial_atom, int with_q,
                                                       int want_grad_a) {
    __shared__ PairTile tile;
    __shared__ float red[8][kTile][8];   // per-warp partials of (A, B0..2, C0..2) per slot
    int i = blockIdx.x, tid = threadIdx.x;
    int g = tid & 15, lane = tid & 31, warp = tid >> 5;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (mol_idx ? mol_idx[i] : 0)) : nullptr;
    int len = row_length(nb, i);
    float4 dSi = reinterpret_cast<const float4*>(dS_a)[(size_t)i * kAG + tid];
    bool qthread = with_q && tid < C * kG;
    int qc = tid >> 4;
    float4 dSqi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qthread) dSqi = reinterpret_cast<const float4*>(dS_q)[(size_t)i * (C * kG) + tid];
    float ga = 0.f, gq = 0.f;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    double vir[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) vir[k] = 0.0;
    for (int m0 = 0; m0 < len; m0 += kTile) {
        stage_tile<true>(tile, i, m0, len, nb, coord, cell, aev);
        int lim = min(kTile, len - m0);
        for (int s = 0; s < lim; ++s) {
            int j = tile.j[s];
            float gsv = tile.gs[s][g], dg = tile.dgs[s][g];
            float ux = tile.u[s][0], uy = tile.u[s][1], uz = tile.u[s][2];
            float aj = a[(size_t)j * kAG + tid];
            if (want_grad_a) {
                float4 dj = reinterpret_cast<const float4*>(dS_a)[(size_t)j * kAG + tid];
                ga += gsv * (dj.x - (dj.y * ux + dj.z * uy + dj.w * uz));
            }
            float p0 = aj * dSi.x, p1 = aj * dSi.y, p2 = aj * dSi.z, p3 = aj * dSi.w;
            if (qthread) {
                float qj = q[(size_t)j * C + qc];
                p0 += qj * dSqi.x;
                p1 += qj * dSqi.y;
                p2 += qj * dSqi.z;
                p3 += qj * dSqi.w;
                if (want_grad_a) {
                    float4 dqj = reinterpret_cast<const float4*>(dS_q)[(size_t)j * (C * kG) + tid];
                    gq += gsv * (dqj.x - (dqj.y * ux + dqj.z * uy + dqj.w * uz));
                }
            }
            float vA = p0 * dg, vB0 = p1 * gsv, vB1 = p2 * gsv, vB2 = p3 * gsv, vC0 = p1 * dg, vC1 = p2 * dg,
                  vC2 = p3 * dg;
            vA = warp_sum(vA);
            vB0 = warp_sum(vB0);
            vB1 = warp_sum(vB1);
            vB2 = warp_sum(vB2);
            vC0 = warp_sum(vC0);
            vC1 = warp_sum(vC1);
            vC2 = warp_sum(vC2);
            if (lane == 0) {
                red[warp][s][0] = vA;
                red[warp][s][1] = vB0;
                red[warp][s][2] = vB1;
                red[warp][s][3] = vB2;
                red[warp][s][4] = vC0;
                red[warp][s][5] = vC1;
                red[warp][s][6] = vC2;
            }
        }
        __syncthreads();
        if (warp == 0) {
            // kTile == 32: slot s <-> lane s of warp 0
            int s = lane;
            bool act = s < lim;
            float A = 0.f, B0 = 0.f, B1 = 0.f, B2 = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
            if (act) {
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    A += red[w][s][0];
                    B0 += red[w][s][1];
                    B1 += red[w][s][2];
                    B2 += red[w][s][3];
                    C0 += red[w][s][4];
                    C1 += red[w][s][5];
                    C2 += red[w][s][6];
                }
            }
            float ux = tile.u[s][0], uy = tile.u[s][1], uz = tile.u[s][2];
            float d = tile.d[s], v = act ? tile.valid[s] : 0.f;
            float cu = C0 * ux + C1 * uy + C2 * uz;
            float bu = B0 * ux + B1 * uy + B2 * uz;
            float inv = 1.0f / d;
            float wx = (ux * (A + cu) + (B0 - ux * bu) * inv) * v;
            float wy = (uy * (A + cu) + (B1 - uy * bu) * inv) * v;
            float wz = (uz * (A + cu) + (B2 - uz * bu) * inv) * v;
            if (v != 0.f) {
                int j = tile.j[s];
                atomicAdd(&forces[3 * j + 0], -wx);
                atomicAdd(&forces[3 * j + 1], -wy);
                atomicAdd(&forces[3 * j + 2], -wz);
            }
            fx += warp_sum(wx);
            fy += warp_sum(wy);
            fz += warp_sum(wz);
            if (virial_atom) {
                float rx = ux * d, ry = uy * d, rz = uz * d;
                vir[0] += warp_sum((double)(rx * wx));
                vir[1] += warp_sum((double)(rx * wy));
                vir[2] += warp_sum((double)(rx * wz));
                vir[3] += warp_sum((double)(ry * wx));
                vir[4] += warp_sum((double)(ry * wy));
                vir[5] += warp_sum((double)(ry * wz));
                vir[6] += warp_sum((double)(rz * wx));
                vir[7] += warp_sum((double)(rz * wy));
                vir[8] += warp_sum((double)(rz * wz));
            }
        }
        __syncthreads();
    }
    if (want_grad_a) {
        grad_a[(size_t)i * kAG + tid] = ga;
        if (with_q) {
            // reduce gq over g (16 lanes) then over the two half-warps of each charge channel
            float v = qthread ? gq : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (qthread && g == 0) grad_q[(size_t)i * C + qc] = v;
        }
    }
    if (tid == 0) {
        atomicAdd(&forces[3 * i + 0], fx);
        atomicAdd(&forces[3 * i + 1], fy);
        atomicAdd(&forces[3 * i + 2], fz);
        if (virial_atom)
#pragma unroll
            for (int k = 0; k < 9; ++k) virial_atom[(size_t)i * 9 + k] += vir[k];
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
int launch_conv_fwd(int C, int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                    const int32_t* mol_idx, const AevParams& aev, const float* a, const float* q, const float* agh_a,
                    const float* agh_q, float* x, int ldx, float* T_a, float* T_q, int with_q, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    if (C == 1)
        conv_fwd_kernel<1><<<n_atoms, 256, 0, st>>>(n_atoms, nb, coord, cv, mol_idx, aev, a, q, agh_a, agh_q, x, ldx,
                                                   T_a, T_q, with_q);
    else
        conv_fwd_kernel<2><<<n_atoms, 256, 0, st>>>(n_atoms, nb, coord, cv, mol_idx, aev, a, q, agh_a, agh_q, x, ldx,
                                                   T_a, T_q, with_q);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

int launch_conv_bwd(int C, int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                    const int32_t* mol_idx, const AevParams& aev, const float* a, const float* q, const float* dx,
                    int ldx, const float* T_a, const float* T_q, const float* agh_a, const float* agh_q, float* dS_a,
                    float* dS_q, float* grad_a, float* grad_q, float* forces, double* virial_atom, int with_q,
                    int want_grad_a, cudaStream_t st) {
    if (n_atoms == 0) return AIMNET_OK;
    if (C == 1) {
        conv_bwd_prep_kernel<1><<<n_atoms, 256, 0, st>>>(n_atoms, dx, ldx, T_a, T_q, agh_a, agh_q, dS_a, dS_q, with_q);
        AIM_LAUNCH_CHECK();
        conv_bwd_kernel<1><<<n_atoms, 256, 0, st>>>(n_atoms, nb, coord, cv, mol_idx, aev, a, q, dS_a, dS_q, grad_a,
                                                   grad_q, forces, virial_atom, with_q, want_grad_a);
    } else {
        conv_bwd_prep_kernel<2><<<n_atoms, 256, 0, st>>>(n_atoms, dx, ldx, T_a, T_q, agh_a, agh_q, dS_a, dS_q, with_q);
        AIM_LAUNCH_CHECK();
        conv_bwd_kernel<2><<<n_atoms, 256, 0, st>>>(n_atoms, nb, coord, cv, mol_idx, aev, a, q, dS_a, dS_q, grad_a,
                                                   grad_q, forces, virial_atom, with_q, want_grad_a);
    }
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

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
