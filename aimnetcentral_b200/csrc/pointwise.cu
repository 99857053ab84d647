// Small per-atom / per-molecule kernels around the GEMMs: embedding, charge equilibration (NSE) forward and
// backward, energy head tail, fp64 per-molecule reductions.
#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

// gather layout of the features (see conv.cu): slot t of a row holds (a, g) with
//   a = (t >> 6) * 4 + (t & 3),  g = (t >> 2) & 15 ;  canonical flat index = a*16 + g
__device__ __forceinline__ int gather_slot_to_canonical(int t) { return (((t >> 6) << 2) + (t & 3)) * kG + ((t >> 2) & 15); }

// a0[i] = afv[Z_i]   (aimnet/models/aimnet2.py:144-147), stored in the gather layout
// 64 threads per atom (one float4 of the gather layout each), 4 atoms per block
__global__ void __launch_bounds__(256) embed_kernel(int n, const int32_t* __restrict__ numbers, const float* __restrict__ afv,
                                                    float* __restrict__ aT0) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 6), u = threadIdx.x & 63;
    if (i >= n) return;
    int z = numbers[i];
    z = (z < 0 || z > 63) ? 0 : z;
    const float* src = afv + (size_t)z * kAG + ((u >> 4) << 2) * kG + (u & 15);   // canonical (4 aq + k) * 16 + g
    reinterpret_cast<float4*>(aT0)[(size_t)i * (kAG / 4) + u] = make_float4(src[0], src[kG], src[2 * kG], src[3 * kG]);
}

// molecule segment pointers from sorted mol_idx (nullptr = one molecule)
__global__ void mol_ptr_kernel(const int32_t* __restrict__ mol_idx, int n, int n_mol, int32_t* __restrict__ ptr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (mol_idx == nullptr) {
        if (i == 0) {
            ptr[0] = 0;
            for (int s = 1; s <= n_mol; ++s) ptr[s] = n;
        }
        return;
    }
    int prev = (i == 0) ? -1 : mol_idx[i - 1];
    int cur = (i == n) ? n_mol : mol_idx[i];
    for (int s = prev + 1; s <= cur && s <= n_mol; ++s) ptr[s] = i;
}

// largest molecule (atoms) of the batch, for the engine's choice of the dense conv walk
__global__ void max_segment_kernel(const int32_t* __restrict__ ptr, int n_mol, int32_t* __restrict__ out) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int len = (m < n_mol) ? ptr[m + 1] - ptr[m] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    T r = 0;
    int nw = (blockDim.x + 31) >> 5;
    for (int k = 0; k < nw; ++k) r += smem[k];
    return r;
}

__device__ __forceinline__ float target_charge(int C, int c, const float* charge, const float* mult, int m) {
    // NSE split (aimnet/models/aimnet2.py:94-100)
    if (C == 1) return charge[m];
    float hs = 0.5f * ((mult ? mult[m] : 1.0f) - 1.0f), hq = 0.5f * charge[m];
    return c == 0 ? hq + hs : hq - hs;
}

// per molecule: sumq[m,c] = sum_i q_u, sumf[m,c] = sum_i f_raw^2   (aimnet/ops.py:99-145, nbops.py:309-377)
__global__ void __launch_bounds__(256) nse_reduce_fwd_kernel(int C, const int32_t* __restrict__ mol_ptr,
                                                             const float* __restrict__ y, int ldy,
                                                             const float* __restrict__ q_prev,
                                                             float* __restrict__ sumq, float* __restrict__ sumf) {
    __shared__ float red[8];
    int m = blockIdx.x;
    int s0 = mol_ptr[m], s1 = mol_ptr[m + 1];
    for (int c = 0; c < C; ++c) {
        float aq = 0.f, af = 0.f;
        for (int i = s0 + threadIdx.x; i < s1; i += blockDim.x) {
            float dq = y[(size_t)i * ldy + c];
            float fr = y[(size_t)i * ldy + C + c];
            aq += (q_prev ? q_prev[(size_t)i * C + c] : 0.f) + dq;
            af += fr * fr;
        }
        aq = block_sum(aq, red);
        af = block_sum(af, red);
        if (threadIdx.x == 0) {
            sumq[m * C + c] = aq;
            sumf[m * C + c] = af;
        }
    }
}

// per atom: q <- q_u + f (Q - sumq)/(sumf + eps) ; a <- a + delta_a     (aimnet/models/aimnet2.py:122-139)
__global__ void __launch_bounds__(256) nse_apply_fwd_kernel(int C, int n, const int32_t* __restrict__ mol_idx,
                                                            const float* __restrict__ charge,
                                                            const float* __restrict__ mult,
                                                            const float* __restrict__ y, int ldy,
                                                            const float* __restrict__ q_prev,
                                                            const float* __restrict__ sumq,
                                                            const float* __restrict__ sumf,
                                                            const float* __restrict__ a_old, float* __restrict__ a_new,
                                                            float* __restrict__ q_new) {
    // features live in the gather layout: 64 threads per atom (one float4 = channels 4 aq .. 4 aq + 3 of one g), 4 atoms per block
    const int i = blockIdx.x * 4 + (threadIdx.x >> 6), t = threadIdx.x & 63;
    if (i >= n) return;
    {
        const float* yr = y + (size_t)i * ldy + 2 * C + ((t >> 4) << 2) * kG + (t & 15);
        float4 a = reinterpret_cast<const float4*>(a_old)[(size_t)i * (kAG / 4) + t];
        a.x += yr[0];
        a.y += yr[kG];
        a.z += yr[2 * kG];
        a.w += yr[3 * kG];
        reinterpret_cast<float4*>(a_new)[(size_t)i * (kAG / 4) + t] = a;
    }
    if (t < C) {
        int m = mol_idx ? mol_idx[i] : 0;
        float qu = (q_prev ? q_prev[(size_t)i * C + t] : 0.f) + y[(size_t)i * ldy + t];
        float fr = y[(size_t)i * ldy + C + t];
        float f = fr * fr;
        float Q = target_charge(C, t, charge, mult, m);
        float F = sumf[m * C + t] + 1.0e-6f;
        float dQ = Q - sumq[m * C + t];
        q_new[(size_t)i * C + t] = qu + f / F * dQ;
    }
}

// backward of the NSE update.  With q_i = qu_i + f_i D/G (D = Q - sum qu, G = sum f + eps) and g_i = dE/dq_i:
//   S1 = sum_i g_i f_i ;  h_i = g_i - S1/G ;  dE/dqu_i = h_i ;  dE/df_raw_i = 2 f_raw_i (D/G) h_i
__global__ void __launch_bounds__(256) nse_reduce_bwd_kernel(int C, const int32_t* __restrict__ mol_ptr,
                                                             const float* __restrict__ y, int ldy,
                                                             const float* __restrict__ gq, float* __restrict__ s1) {
    __shared__ float red[8];
    int m = blockIdx.x;
    int p0 = mol_ptr[m], p1 = mol_ptr[m + 1];
    for (int c = 0; c < C; ++c) {
        float acc = 0.f;
        for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
            float fr = y[(size_t)i * ldy + C + c];
            acc += gq[(size_t)i * C + c] * fr * fr;
        }
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) s1[m * C + c] = acc;
    }
}

// writes dz[i, :] (gradient w.r.t. the pre-activation of the pass MLP's last Linear) and dq_prev (dE/dq of the
// previous pass through qu = q_prev + dq)
__global__ void __launch_bounds__(288) nse_apply_bwd_kernel(int C, int n, const int32_t* __restrict__ mol_idx,
                                                            const float* __restrict__ charge,
                                                            const float* __restrict__ mult,
                                                            const float* __restrict__ y, int ldy,
                                                            const float* __restrict__ gq,
                                                            const float* __restrict__ sumq,
                                                            const float* __restrict__ sumf,
                                                            const float* __restrict__ s1,
                                                            const float* __restrict__ da_tot,
                                                            const float* __restrict__ gp_last, int ldgp,
                                                            float* __restrict__ dz, int lddz,
                                                            float* __restrict__ dq_prev) {
    // 72 threads per atom, four consecutive columns each (float4 stores / gelu' loads), 4 atoms per block
    const int i = blockIdx.x * 4 + threadIdx.x / 72, u = threadIdx.x % 72;
    if (i >= n || 4 * u >= lddz) return;
    const int m = mol_idx ? mol_idx[i] : 0;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t = 4 * u + k;
        v[k] = 0.f;
        if (t < 2 * C) {
            int c = (t < C) ? t : t - C;
            float G = sumf[m * C + c] + 1.0e-6f;
            float h = gq[(size_t)i * C + c] - s1[m * C + c] / G;
            if (t < C) {
                v[k] = h;
                if (dq_prev) dq_prev[(size_t)i * C + c] = h;
            } else {
                float D = target_charge(C, c, charge, mult, m) - sumq[m * C + c];
                v[k] = 2.0f * y[(size_t)i * ldy + C + c] * (D / G) * h;
            }
        } else if (t < 2 * C + kAG) {
            v[k] = da_tot[(size_t)i * kAG + (t - 2 * C)];
        }
    }
    if (gp_last != nullptr) {   // columns past 2C + 256 hold v = 0 (and gelu' = gelu'(0) there)
        const float4 g = *reinterpret_cast<const float4*>(gp_last + (size_t)i * ldgp + 4 * u);
        v[0] *= g.x;
        v[1] *= g.y;
        v[2] *= g.z;
        v[3] *= g.w;
    }
    *reinterpret_cast<float4*>(dz + (size_t)i * lddz + 4 * u) = make_float4(v[0], v[1], v[2], v[3]);
}

// da_tot (+)= dx[:, :256] + grad_a ;  dq = base_q + dx[:, 704+c] + grad_q
__global__ void __launch_bounds__(256) accum_grads_kernel(int C, int n, const float* __restrict__ dx, int ldx,
                                                          const float* __restrict__ grad_a,
                                                          const float* __restrict__ grad_q,
                                                          const float* __restrict__ base_q, int base_q_stride,
                                                          float* __restrict__ da_tot, int accumulate,
                                                          float* __restrict__ dq) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 6), t = threadIdx.x & 63;   // 64 threads (float4) per atom, 4 atoms per block
    if (i >= n) return;
    const float4 d = *reinterpret_cast<const float4*>(dx + (size_t)i * ldx + 4 * t);
    const float4 ga = reinterpret_cast<const float4*>(grad_a)[(size_t)i * (kAG / 4) + t];
    float4 v = make_float4(d.x + ga.x, d.y + ga.y, d.z + ga.z, d.w + ga.w);
    if (accumulate) {
        const float4 o = reinterpret_cast<const float4*>(da_tot)[(size_t)i * (kAG / 4) + t];
        v = make_float4(v.x + o.x, v.y + o.y, v.z + o.z, v.w + o.w);
    }
    reinterpret_cast<float4*>(da_tot)[(size_t)i * (kAG / 4) + t] = v;
    if (t < C) {
        float b = base_q[(size_t)i * base_q_stride + (base_q_stride == 1 ? 0 : t)];
        dq[(size_t)i * C + t] = b + dx[(size_t)i * ldx + (2 * kAG + kAH) + t] + grad_q[(size_t)i * C + t];
    }
}

// energy head tail: e_i = w3 . h2_i + b3 (+ SAE in fp64); seeds the backward pass with dz2 = w3 * gelu'(z2)
// (aimnet/modules/core.py:114-132, 71-97).  One warp per atom.
__global__ void __launch_bounds__(256) head_tail_kernel(int n, const float* __restrict__ h2, int ldh,
                                                        const float* __restrict__ gp2, const float* __restrict__ w3,
                                                        float b3, const int32_t* __restrict__ numbers,
                                                        const double* __restrict__ sae, double* __restrict__ e_atom,
                                                        float* __restrict__ dz2) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    int i = warp;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int c = lane + 32 * k;
        float w = w3[c];
        acc += w * h2[(size_t)i * ldh + c];
        dz2[(size_t)i * ldh + c] = w * gp2[(size_t)i * ldh + c];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        int z = numbers[i];
        z = (z < 0 || z > 63) ? 0 : z;
        e_atom[i] = (double)(acc + b3) + sae[z];
    }
}

// energy[m] = sum_{i in m} (e_nn + e_sr + e_lr + e_d3)   fp64 (aimnet/modules/core.py:100-111)
__global__ void __launch_bounds__(256) energy_reduce_kernel(const int32_t* __restrict__ mol_ptr,
                                                            const double* __restrict__ e0,
                                                            const double* __restrict__ e1,
                                                            const double* __restrict__ e2,
                                                            const double* __restrict__ e3,
                                                            double* __restrict__ energy) {
    __shared__ double red[8];
    int m = blockIdx.x;
    int p0 = mol_ptr[m], p1 = mol_ptr[m + 1];
    double acc = 0.0;
    for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
        double v = e0[i];
        if (e1) v += e1[i];
        if (e2) v += e2[i];
        if (e3) v += e3[i];
        acc += v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) energy[m] = acc;
}

// stress[s] = sum_{i in s} virial_atom[i] / |det cell_s|   (aimnet/calculators/derivatives.py:122-137)
__global__ void __launch_bounds__(256) stress_reduce_kernel(const int32_t* __restrict__ mol_ptr, int n_cells, int n,
                                                            const double* __restrict__ virial_atom,
                                                            const float* __restrict__ cell, float* __restrict__ stress) {
    __shared__ double red[8];
    int s = blockIdx.x;
    int p0 = (n_cells == 1) ? 0 : mol_ptr[s], p1 = (n_cells == 1) ? n : mol_ptr[s + 1];
    const float* c = cell + 9 * s;
    double det = (double)c[0] * ((double)c[4] * c[8] - (double)c[5] * c[7]) -
                 (double)c[1] * ((double)c[3] * c[8] - (double)c[5] * c[6]) +
                 (double)c[2] * ((double)c[3] * c[7] - (double)c[4] * c[6]);
    double vol = fabs(det);
    for (int k = 0; k < 9; ++k) {
        double acc = 0.0;
        for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) acc += virial_atom[(size_t)i * 9 + k];
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) stress[9 * s + k] = (float)(acc / vol);
    }
}

// final charges: C==1 copy; C==2: charges = qa+qb, spin = qa-qb (aimnet/models/aimnet2.py:102-106)
__global__ void charges_out_kernel(int C, int n, const float* __restrict__ q, float* __restrict__ charges,
                                   float* __restrict__ spin) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (C == 1) {
        charges[i] = q[i];
    } else {
        float a = q[2 * i], b = q[2 * i + 1];
        charges[i] = a + b;
        if (spin) spin[i] = a - b;
    }
}

// ------------------------------------------------------------------------------------------------------------
// ---- Verlet-skin bookkeeping (engine.cu): has any atom moved by more than skin/2 since the lists were built? ----
__global__ void skin_check_kernel(int n, const float* __restrict__ x, const float* __restrict__ ref, float thr2,
                                  const int32_t* __restrict__ mol, const int32_t* __restrict__ mol_ref,
                                  int32_t* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved = false;
    if (i < n) {
        float dx = x[3 * i] - ref[3 * i], dy = x[3 * i + 1] - ref[3 * i + 1], dz = x[3 * i + 2] - ref[3 * i + 2];
        float d2 = dx * dx + dy * dy + dz * dz;
        moved = !(d2 <= thr2);                             // NaN counts as moved
        if (mol != nullptr) moved |= mol[i] != mol_ref[i];  // same atoms, another partition into molecules
    }
    if (__any_sync(0xffffffffu, moved) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
// ref = x; off = wrapped - x (the lattice vector the wrap added), or 0 without a cell; mol_ref = mol
__global__ void skin_save_kernel(int n, const float* __restrict__ x, const float* __restrict__ wrapped,
                                 float* __restrict__ ref, float* __restrict__ off, const int32_t* __restrict__ mol,
                                 int32_t* __restrict__ mol_ref) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && mol != nullptr) mol_ref[i] = mol[i];
    if (i >= 3 * n) return;
    ref[i] = x[i];
    off[i] = wrapped ? wrapped[i] - x[i] : 0.f;
}
__global__ void skin_apply_kernel(int n, const float* __restrict__ x, const float* __restrict__ off, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n) out[i] = x[i] + off[i];
}

#define AIM_K(...)          \
    do {                    \
        __VA_ARGS__;        \
        AIM_LAUNCH_CHECK(); \
    } while (0)

int launch_embed(int n, const int32_t* numbers, const float* afv, float* a0, cudaStream_t st) {
    if (n) AIM_K(embed_kernel<<<(n + 3) / 4, 256, 0, st>>>(n, numbers, afv, a0));
    return AIMNET_OK;
}
int launch_mol_ptr(const int32_t* mol_idx, int n, int n_mol, int32_t* ptr, int32_t* max_segment, cudaStream_t st) {
    AIM_K(mol_ptr_kernel<<<(n + 256) / 256, 256, 0, st>>>(mol_idx, n, n_mol, ptr));
    if (max_segment != nullptr) {
        AIM_CUDA_CHECK(cudaMemsetAsync(max_segment, 0, sizeof(int32_t), st));
        AIM_K(max_segment_kernel<<<(n_mol + 255) / 256, 256, 0, st>>>(ptr, n_mol, max_segment));
    }
    return AIMNET_OK;
}
int launch_nse_fwd(int C, int n, int n_mol, const int32_t* mol_idx, const int32_t* mol_ptr, const float* charge,
                   const float* mult, const float* y, int ldy, const float* q_prev, float* sumq, float* sumf,
                   const float* a_old, float* a_new, float* q_new, cudaStream_t st) {
    if (!n) return AIMNET_OK;
    AIM_K(nse_reduce_fwd_kernel<<<n_mol, 256, 0, st>>>(C, mol_ptr, y, ldy, q_prev, sumq, sumf));
    AIM_K(nse_apply_fwd_kernel<<<(n + 3) / 4, 256, 0, st>>>(C, n, mol_idx, charge, mult, y, ldy, q_prev, sumq, sumf, a_old, a_new,
                                                 q_new));
    return AIMNET_OK;
}
int launch_nse_bwd(int C, int n, int n_mol, const int32_t* mol_idx, const int32_t* mol_ptr, const float* charge,
                   const float* mult, const float* y, int ldy, const float* gq, const float* sumq, const float* sumf,
                   float* s1, const float* da_tot, const float* gp_last, int ldgp, float* dz, int lddz, float* dq_prev,
                   cudaStream_t st) {
    if (!n) return AIMNET_OK;
    if (lddz > 288 || lddz % 4 != 0 || (gp_last != nullptr && ldgp % 4 != 0)) {
        set_error("nse_bwd: lddz must be a multiple of 4 and at most 288");
        return AIMNET_EINVAL;
    }
    AIM_K(nse_reduce_bwd_kernel<<<n_mol, 256, 0, st>>>(C, mol_ptr, y, ldy, gq, s1));
    AIM_K(nse_apply_bwd_kernel<<<(n + 3) / 4, 288, 0, st>>>(C, n, mol_idx, charge, mult, y, ldy, gq, sumq, sumf, s1, da_tot,
                                                 gp_last, ldgp, dz, lddz, dq_prev));
    return AIMNET_OK;
}
int launch_accum_grads(int C, int n, const float* dx, int ldx, const float* grad_a, const float* grad_q,
                       const float* base_q, int base_q_stride, float* da_tot, int accumulate, float* dq,
                       cudaStream_t st) {
    if (n) AIM_K(accum_grads_kernel<<<(n + 3) / 4, 256, 0, st>>>(C, n, dx, ldx, grad_a, grad_q, base_q, base_q_stride, da_tot,
                                                      accumulate, dq));
    return AIMNET_OK;
}
int launch_head_tail(int n, const float* h2, int ldh, const float* gp2, const float* w3, float b3,
                     const int32_t* numbers, const double* sae, double* e_atom, float* dz2, cudaStream_t st) {
    if (n) AIM_K(head_tail_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, h2, ldh, gp2, w3, b3, numbers, sae, e_atom, dz2));
    return AIMNET_OK;
}
int launch_energy_reduce(int n_mol, const int32_t* mol_ptr, const double* e0, const double* e1, const double* e2,
                         const double* e3, double* energy, cudaStream_t st) {
    AIM_K(energy_reduce_kernel<<<n_mol, 256, 0, st>>>(mol_ptr, e0, e1, e2, e3, energy));
    return AIMNET_OK;
}
int launch_stress_reduce(const int32_t* mol_ptr, int n_cells, int n, const double* virial_atom, const float* cell,
                         float* stress, cudaStream_t st) {
    AIM_K(stress_reduce_kernel<<<n_cells, 256, 0, st>>>(mol_ptr, n_cells, n, virial_atom, cell, stress));
    return AIMNET_OK;
}
int launch_skin_check(int n, const float* x, const float* ref, float thr2, const int32_t* mol, const int32_t* mol_ref,
                      int32_t* flag, cudaStream_t st) {
    AIM_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    if (n) AIM_K(skin_check_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, x, ref, thr2, mol, mol_ref, flag));
    return AIMNET_OK;
}
int launch_skin_save(int n, const float* x, const float* wrapped, float* ref, float* off, const int32_t* mol,
                     int32_t* mol_ref, cudaStream_t st) {
    if (n) AIM_K(skin_save_kernel<<<(3 * n + 255) / 256, 256, 0, st>>>(n, x, wrapped, ref, off, mol, mol_ref));
    return AIMNET_OK;
}
int launch_skin_apply(int n, const float* x, const float* off, float* out, cudaStream_t st) {
    if (n) AIM_K(skin_apply_kernel<<<(3 * n + 255) / 256, 256, 0, st>>>(n, x, off, out));
    return AIMNET_OK;
}
int launch_charges_out(int C, int n, const float* q, float* charges, float* spin, cudaStream_t st) {
    if (n) AIM_K(charges_out_kernel<<<(n + 255) / 256, 256, 0, st>>>(C, n, q, charges, spin));
    return AIMNET_OK;
}

}  // namespace aimnet
