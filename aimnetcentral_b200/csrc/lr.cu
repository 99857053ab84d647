// Pair-potential walkers: embedded short-range Coulomb, external Coulomb (simple / DSF) and DFT-D3(BJ)
// (SURVEY.md §8a rows a13-a15, a17).
//
// Reference semantics: _calc_coulomb_sr (aimnet/modules/lr.py:21-62), LRCoulomb.coul_simple (:311-331),
// LRCoulomb._coul_dsf_torch (:559-615), DFTD3._compute_energy_torch and helpers (:1580-1657).
//
// All of these are symmetric pair terms over a full (both-direction) neighbour set:  E = c * sum_{ordered (i,m)} e(d_im).
// With a symmetric list everything atom i needs is in its own row:
//     dE/dq_i = 2c sum_m de/dq_i ,   F_i = 2c sum_m e'(d) u_im ,   virial_i = c sum_m e'(d) d u (x) u
// so there are no atomics and the result is run-to-run deterministic (the reference routes these through
// atomics-based nvalchemiops kernels unless deterministic=True, calculator.py:76-84).
//
// Neighbour source: either a neighbour matrix row (periodic / cutoff lists) or, for isolated molecules with the
// all-pairs "simple" Coulomb (calculator.py:1635-1636 asks for max_neighbors = N there), the molecule's own atom
// segment [mol_ptr[m], mol_ptr[m+1]) — the N x N list is never materialised.
#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

// The pair walkers divide a lot (1/d, damping denominators, switches).  IEEE division costs ~12 instructions plus a slow
// path per site; MUFU.RCP based division is within 2 ulp, three orders of magnitude inside the parity budget.
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float frcp(float b) { return __fdividef(1.0f, b); }

__device__ __forceinline__ void row_range(const PairSource& ps, int i, int& begin, int& end) {
    if (ps.nb.nbmat) {
        begin = 0;
        end = ps.nb.count ? min(ps.nb.count[i], ps.nb.width) : ps.nb.width;
    } else {
        int m = ps.mol_idx ? ps.mol_idx[i] : 0;
        begin = ps.mol_ptr[m];
        end = ps.mol_ptr[m + 1];
    }
}

// returns false when the slot is empty
__device__ __forceinline__ bool slot_geometry(const PairSource& ps, const float* __restrict__ coord,
                                              const float* __restrict__ cell, int i, int m, int& j, float& rx,
                                              float& ry, float& rz) {
    if (ps.nb.nbmat) {
        j = ps.nb.nbmat[(size_t)i * ps.nb.width + m];
        if (j == ps.nb.sentinel || j < 0) return false;
        const int32_t* sh = ps.nb.shifts ? ps.nb.shifts + ((size_t)i * ps.nb.width + m) * 3 : nullptr;
        pair_vector(coord, i, j, sh, cell, rx, ry, rz);
    } else {
        j = m;
        if (j == i) return false;
        pair_vector(coord, i, j, nullptr, nullptr, rx, ry, rz);
        if (ps.seg_cut2 > 0.f && rx * rx + ry * ry + rz * rz >= ps.seg_cut2) return false;
    }
    return true;
}

template <int MODE>
__device__ __forceinline__ void pair_phi(float d, const CoulombParams& p, float& phi, float& dphi, float& inv) {
    inv = frcp(d);
    if (MODE == PAIR_SIMPLE) {
        phi = inv;
        dphi = -inv * inv;
    } else if (MODE == PAIR_SR_EXP) {
        // exp_cutoff, aimnet/ops.py:88-90
        float t = fdiv(d, p.rc);
        bool clamped = t >= 1.0f - 1e-6f;
        t = fminf(fmaxf(t, 0.f), 1.0f - 1e-6f);
        float om = 1.0f - t * t;
        float fc = expf(-frcp(om)) * 2.718281828459045f;
        float dfc = clamped ? 0.f : fdiv(-fc * 2.0f * t, om * om * p.rc);
        phi = fc * inv;
        dphi = dfc * inv - fc * inv * inv;
    } else if (MODE == PAIR_SR_COS) {
        float dc = fminf(fmaxf(d, 1e-6f), p.rc);
        float sn, cs;
        sincosf(dc * (kPi / p.rc), &sn, &cs);
        float fc = 0.5f * (cs + 1.0f);
        float dfc = (d > 1e-6f && d < p.rc) ? -0.5f * (kPi / p.rc) * sn : 0.f;
        phi = fc * inv;
        dphi = dfc * inv - fc * inv * inv;
    } else if (MODE == PAIR_EWALD) {   // real-space Ewald term
        if (d < p.rc) {
            float ec = erfcf(p.alpha * d);
            phi = ec * inv;
            dphi = -ec * inv * inv - 1.1283791670955126f * p.alpha * expf(-p.alpha * p.alpha * d * d) * inv;
        } else {
            phi = 0.f;
            dphi = 0.f;
        }
    } else {   // DSF, lr.py:594-600
        if (d < p.rc) {
            float ec = erfcf(p.alpha * d);
            phi = ec * inv - p.shift_val + (d - p.rc) * p.shift_slope;
            dphi = -ec * inv * inv - 1.1283791670955126f * p.alpha * expf(-p.alpha * p.alpha * d * d) * inv +
                   p.shift_slope;
        } else {
            phi = 0.f;
            dphi = 0.f;
        }
    }
}

// one warp per atom
template <int MODE>
__global__ void __launch_bounds__(256) coulomb_pair_kernel(int n, PairSource ps, const float* __restrict__ coord,
                                                           CellView cv, const float* __restrict__ q,
                                                           CoulombParams p, double* __restrict__ e_atom,
                                                           float* __restrict__ gq, float* __restrict__ forces,
                                                           double* __restrict__ virial_atom, int accumulate_e,
                                                           int atom_lo) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    int i = atom_lo + warp;   // atoms [atom_lo, atom_lo + n): one periodic system of a batch (per-system Ewald parameters)
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (ps.mol_idx ? ps.mol_idx[i] : 0)) : nullptr;
    int b, e;
    row_range(ps, i, b, e);
    float qi = q[i];
    double esum = 0.0;
    float gsum = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
    float vir[9];   // per-lane fp32 partial sums (a lane sees 1/32 of the row), reduced and accumulated in fp64 below
#pragma unroll
    for (int k = 0; k < 9; ++k) vir[k] = 0.f;
    for (int m = b + lane; m < e; m += 32) {
        int j;
        float rx, ry, rz;
        if (!slot_geometry(ps, coord, cell, i, m, j, rx, ry, rz)) continue;
        float d = sqrtf(rx * rx + ry * ry + rz * rz);
        float phi, dphi, inv;
        pair_phi<MODE>(d, p, phi, dphi, inv);
        float qj = q[j];
        esum += (double)(qi * qj * phi);
        gsum += qj * phi;
        float w = qi * qj * dphi * inv;   // e'(d) / d  -> times r gives e' u
        fx += w * rx;
        fy += w * ry;
        fz += w * rz;
        if (virial_atom) {
            vir[0] += w * rx * rx;
            vir[1] += w * rx * ry;
            vir[2] += w * rx * rz;
            vir[3] += w * ry * rx;
            vir[4] += w * ry * ry;
            vir[5] += w * ry * rz;
            vir[6] += w * rz * rx;
            vir[7] += w * rz * ry;
            vir[8] += w * rz * rz;
        }
    }
    esum = warp_sum(esum);
    gsum = warp_sum(gsum);
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    double dvir[9];
    if (virial_atom)
#pragma unroll
        for (int k = 0; k < 9; ++k) dvir[k] = warp_sum((double)vir[k]);
    if (lane == 0) {
        double c = p.factor;
        double ei = c * esum;
        float gi = (float)(2.0 * c) * gsum;
        if (MODE == PAIR_DSF) {   // self term, lr.py:601-611 (full k_e = 2k)
            ei += 2.0 * c * (double)(p.self_coeff * qi * qi);
            gi += (float)(4.0 * c) * p.self_coeff * qi;
        }
        e_atom[i] = accumulate_e ? e_atom[i] + ei : ei;
        gq[i] += gi;
        if (forces) {
            float c2 = (float)(2.0 * c);
            forces[3 * i + 0] += c2 * fx;
            forces[3 * i + 1] += c2 * fy;
            forces[3 * i + 2] += c2 * fz;
        }
        if (virial_atom)
#pragma unroll
            for (int k = 0; k < 9; ++k) virial_atom[(size_t)i * 9 + k] += c * dvir[k];
    }
}

// ------------------------------------------------------------------------------------------------------------
// DFT-D3(BJ), two-body, Bohr/Hartree internally (lr.py:1580-1657)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampz(int z) { return (z < 0 || z > 94) ? 0 : z; }

// CN_i = sum_j sigmoid(16 ((rcov_i + rcov_j)/d - 1))     (lr.py:1595-1603)
__global__ void __launch_bounds__(256) d3_cn_kernel(int n, PairSource ps, const float* __restrict__ coord, CellView cv,
                                                    const int32_t* __restrict__ numbers, D3Params p,
                                                    float* __restrict__ cn) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    int i = warp;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (ps.mol_idx ? ps.mol_idx[i] : 0)) : nullptr;
    int b, e;
    row_range(ps, i, b, e);
    float rci = p.rcov[clampz(numbers[i])];
    float acc = 0.f;
    for (int m = b + lane; m < e; m += 32) {
        int j;
        float rx, ry, rz;
        if (!slot_geometry(ps, coord, cell, i, m, j, rx, ry, rz)) continue;
        float db = fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-12f) * (float)(1.0 / kBohr);
        if (db >= p.r_off) continue;   // the list may reach further (other long-range term, Verlet skin)
        float arg = 16.0f * (fdiv(rci + p.rcov[clampz(numbers[j])], db) - 1.0f);
        acc += frcp(1.0f + expf(-arg));
    }
    acc = warp_sum(acc);
    if (lane == 0) cn[i] = acc;
}

// C6 interpolation with the reference's max-shifted, thresholded Gaussian weights (lr.py:1605-1624).
// The validity mask of the reference table is separable (c6ref[zi,zj,a,b] != 0  <=>  a valid for zi and b valid for zj;
// verified for every element pair when the table is packed), so max_ab(-4(da^2+db^2)) = m_i + m_j and the weight
// factorises: exp(shifted_ab) = exp(s_i[a]) exp(s_j[b]) with s = -4 d^2 - m <= 0.  The five (s, w = exp(s),
// dw = w * (-8 d)) triples are computed once per atom; a pair then needs 25 multiply-adds and no exponential.
__global__ void d3_weights_kernel(int n, const int32_t* __restrict__ numbers, D3Params p, const float* __restrict__ cn,
                                  float* __restrict__ wtab) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int z = clampz(numbers[i]);
    float c = cn[i];
    float d[5], arg[5];
    float mx = -INFINITY;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        float ref = p.cnref[z * 5 + a];
        d[a] = c - ref;
        arg[a] = (ref >= 0.f) ? -4.0f * d[a] * d[a] : -INFINITY;
        mx = fmaxf(mx, arg[a]);
    }
    float* o = wtab + (size_t)i * 16;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        bool ok = arg[a] > -INFINITY;
        float sft = ok ? arg[a] - mx : -1.0e30f;
        float w = ok ? expf(sft) : 0.f;
        o[a] = sft;
        o[5 + a] = w;
        o[10 + a] = w * (-8.0f * d[a]);
    }
    o[15] = 0.f;
}

// c6ref rows are padded to 28 floats (kC6Row) at upload and the weight table rows are 16 floats, so a pair reads its
// 25 reference values and the neighbour's 10 weights as ten 16-byte loads instead of 35 scalar ones.
__device__ __forceinline__ void d3_c6(const D3Params& p, int zi, int zj, const float* __restrict__ wi,
                                      const float* __restrict__ wj, float& c6, float& dc6_dcni) {
    const float4* cr4 = reinterpret_cast<const float4*>(p.c6ref + ((size_t)zi * 95 + zj) * kC6Row);
    float cr[28];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float4 v = __ldg(cr4 + k);
        cr[4 * k + 0] = v.x;
        cr[4 * k + 1] = v.y;
        cr[4 * k + 2] = v.z;
        cr[4 * k + 3] = v.w;
    }
    const float4* wj4 = reinterpret_cast<const float4*>(wj);
    const float4 j0 = wj4[0], j1 = wj4[1], j2 = wj4[2];
    const float sj[5] = {j0.x, j0.y, j0.z, j0.w, j1.x};
    const float ej[5] = {j1.y, j1.z, j1.w, j2.x, j2.y};
    float si[5], ei[5], di[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        si[a] = wi[a];
        ei[a] = wi[5 + a];
        di[a] = wi[10 + a];
    }
    // sum_ab c_ab w_a w_b [s_a + s_b >= -12] with the b sums hoisted: four instructions per (a, b) instead of eight
    float tj[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) tj[b] = -12.0f - sj[b];
    float wsum = 0.f, csum = 0.f, dwsum = 0.f, dcsum = 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        float kb = 0.f, kc = 0.f;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            float keep = (si[a] >= tj[b]) ? ej[b] : 0.f;   // invalid references carry w = 0 and s = -1e30
            kb += keep;
            kc = fmaf(cr[a * 5 + b], keep, kc);
        }
        wsum = fmaf(ei[a], kb, wsum);
        csum = fmaf(ei[a], kc, csum);
        dwsum = fmaf(di[a], kb, dwsum);
        dcsum = fmaf(di[a], kc, dcsum);
    }
    if (wsum > 1e-12f) {
        float inv = frcp(fmaxf(wsum, 1e-12f));
        c6 = csum * inv;
        dc6_dcni = (dcsum - c6 * dwsum) * inv;
    } else {
        c6 = 0.f;
        dc6_dcni = 0.f;
    }
}

// E_i = c sum_m e(d), e = -C6 * damp * sw ; dEdCN_i = 2c sum_m -(dC6/dCN_i) damp sw ; direct pair force + virial
__global__ void __launch_bounds__(256) d3_energy_kernel(int n, PairSource ps, const float* __restrict__ coord,
                                                        CellView cv, const int32_t* __restrict__ numbers, D3Params p,
                                                        const float* __restrict__ wtab, double* __restrict__ e_atom,
                                                        float* __restrict__ dEdCN, float* __restrict__ forces,
                                                        double* __restrict__ virial_atom) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    int i = warp;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (ps.mol_idx ? ps.mol_idx[i] : 0)) : nullptr;
    int b, e;
    row_range(ps, i, b, e);
    int zi = clampz(numbers[i]);
    float r4i = p.r4r2[zi];
    float wi[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) wi[k] = wtab[(size_t)i * 16 + k];
    const float ib = (float)(1.0 / kBohr);
    double esum = 0.0;
    float gsum = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
    float vir[9];   // per-lane fp32 partial sums (a lane sees 1/32 of the row), reduced and accumulated in fp64 below
#pragma unroll
    for (int k = 0; k < 9; ++k) vir[k] = 0.f;
    for (int m = b + lane; m < e; m += 32) {
        int j;
        float rx, ry, rz;
        if (!slot_geometry(ps, coord, cell, i, m, j, rx, ry, rz)) continue;
        float dA = fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-12f);
        float d = dA * ib;
        // switch (lr.py:1580-1593)
        float sw = 1.f, dsw = 0.f;
        if (d > p.r_on) {
            const float iw = frcp(p.r_off - p.r_on);
            float t = fminf(fmaxf((d - p.r_on) * iw, 0.f), 1.f);
            float t2 = t * t;
            sw = 1.0f - t2 * t * (10.0f - 15.0f * t + 6.0f * t2);
            dsw = (t < 1.f) ? -30.0f * t2 * (1.0f - t) * (1.0f - t) * iw : 0.f;
        }
        if (sw == 0.f && dsw == 0.f) continue;
        int zj = clampz(numbers[j]);
        float c6, dc6;
        d3_c6(p, zi, zj, wi, wtab + (size_t)j * 16, c6, dc6);
        float rr = 3.0f * r4i * p.r4r2[zj];
        float r0 = p.a1 * sqrtf(rr) + p.a2;
        float d2 = d * d, d4 = d2 * d2, d6 = d4 * d2, d8 = d4 * d4;
        float r02 = r0 * r0, r04 = r02 * r02, r06 = r04 * r02, r08 = r04 * r04;
        float i6 = frcp(d6 + r06), i8 = frcp(d8 + r08);
        float damp = p.s6 * i6 + p.s8 * rr * i8;
        float ddamp = -(p.s6 * 6.0f * d4 * d * i6 * i6 + p.s8 * rr * 8.0f * d6 * d * i8 * i8);
        float eij = -c6 * damp * sw;
        esum += (double)eij;
        gsum += -dc6 * damp * sw;
        float de = -c6 * (ddamp * sw + damp * dsw) * ib;   // d e / d d_Angstrom
        float w = fdiv(de, dA);
        fx += w * rx;
        fy += w * ry;
        fz += w * rz;
        if (virial_atom) {
            vir[0] += w * rx * rx;
            vir[1] += w * rx * ry;
            vir[2] += w * rx * rz;
            vir[3] += w * ry * rx;
            vir[4] += w * ry * ry;
            vir[5] += w * ry * rz;
            vir[6] += w * rz * rx;
            vir[7] += w * rz * ry;
            vir[8] += w * rz * rz;
        }
    }
    esum = warp_sum(esum);
    gsum = warp_sum(gsum);
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    double dvir[9];
    if (virial_atom)
#pragma unroll
        for (int k = 0; k < 9; ++k) dvir[k] = warp_sum((double)vir[k]);
    if (lane == 0) {
        const double c = 0.5 * kHartree;
        e_atom[i] = c * esum;
        dEdCN[i] = (float)(2.0 * c) * gsum;
        if (forces) {
            float c2 = (float)(2.0 * c);
            forces[3 * i + 0] += c2 * fx;
            forces[3 * i + 1] += c2 * fy;
            forces[3 * i + 2] += c2 * fz;
        }
        if (virial_atom)
#pragma unroll
            for (int k = 0; k < 9; ++k) virial_atom[(size_t)i * 9 + k] += c * dvir[k];
    }
}

// coordination-number chain: F_i += sum_m (dEdCN_i + dEdCN_j) cn'(d) u ;  virial_i += sum_m dEdCN_i cn'(d) d u (x) u
__global__ void __launch_bounds__(256) d3_cn_force_kernel(int n, PairSource ps, const float* __restrict__ coord,
                                                          CellView cv, const int32_t* __restrict__ numbers,
                                                          D3Params p, const float* __restrict__ dEdCN,
                                                          float* __restrict__ forces,
                                                          double* __restrict__ virial_atom) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    int i = warp;
    const float* cell = cv.cell ? cv.cell + 9 * (cv.n_cells == 1 ? 0 : (ps.mol_idx ? ps.mol_idx[i] : 0)) : nullptr;
    int b, e;
    row_range(ps, i, b, e);
    float rci = p.rcov[clampz(numbers[i])];
    float gi = dEdCN[i];
    const float ib = (float)(1.0 / kBohr);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    float vir[9];   // per-lane fp32 partial sums (a lane sees 1/32 of the row), reduced and accumulated in fp64 below
#pragma unroll
    for (int k = 0; k < 9; ++k) vir[k] = 0.f;
    for (int m = b + lane; m < e; m += 32) {
        int j;
        float rx, ry, rz;
        if (!slot_geometry(ps, coord, cell, i, m, j, rx, ry, rz)) continue;
        float dA = fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-12f);
        float d = dA * ib;
        if (d >= p.r_off) continue;
        float R = rci + p.rcov[clampz(numbers[j])];
        float s = frcp(1.0f + expf(-16.0f * (fdiv(R, d) - 1.0f)));
        float dcn = s * (1.0f - s) * 16.0f * (-R / (d * d)) * ib;   // d cn / d d_Angstrom
        const float idA = frcp(dA);
        float w = (gi + dEdCN[j]) * dcn * idA;
        fx += w * rx;
        fy += w * ry;
        fz += w * rz;
        if (virial_atom) {
            float wv = gi * dcn * idA;
            vir[0] += wv * rx * rx;
            vir[1] += wv * rx * ry;
            vir[2] += wv * rx * rz;
            vir[3] += wv * ry * rx;
            vir[4] += wv * ry * ry;
            vir[5] += wv * ry * rz;
            vir[6] += wv * rz * rx;
            vir[7] += wv * rz * ry;
            vir[8] += wv * rz * rz;
        }
    }
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    double dvir[9];
    if (virial_atom)
#pragma unroll
        for (int k = 0; k < 9; ++k) dvir[k] = warp_sum((double)vir[k]);
    if (lane == 0) {
        forces[3 * i + 0] += fx;
        forces[3 * i + 1] += fy;
        forces[3 * i + 2] += fz;
        if (virial_atom)
#pragma unroll
            for (int k = 0; k < 9; ++k) virial_atom[(size_t)i * 9 + k] += dvir[k];
    }
}

// ------------------------------------------------------------------------------------------------------------
int launch_coulomb(int mode, int n, const PairSource& ps, const float* coord, const CellView& cv, const float* q,
                   const CoulombParams& p, double* e_atom, float* gq, float* forces, double* virial_atom,
                   int accumulate_e, cudaStream_t st, int atom_lo) {
    if (n == 0) return AIMNET_OK;
    dim3 grid((n + 7) / 8);
    switch (mode) {
        case PAIR_SR_EXP:
            coulomb_pair_kernel<PAIR_SR_EXP><<<grid, 256, 0, st>>>(n, ps, coord, cv, q, p, e_atom, gq, forces, virial_atom, accumulate_e, atom_lo);
            break;
        case PAIR_SR_COS:
            coulomb_pair_kernel<PAIR_SR_COS><<<grid, 256, 0, st>>>(n, ps, coord, cv, q, p, e_atom, gq, forces, virial_atom, accumulate_e, atom_lo);
            break;
        case PAIR_SIMPLE:
            coulomb_pair_kernel<PAIR_SIMPLE><<<grid, 256, 0, st>>>(n, ps, coord, cv, q, p, e_atom, gq, forces, virial_atom, accumulate_e, atom_lo);
            break;
        case PAIR_EWALD:
            coulomb_pair_kernel<PAIR_EWALD><<<grid, 256, 0, st>>>(n, ps, coord, cv, q, p, e_atom, gq, forces, virial_atom, accumulate_e, atom_lo);
            break;
        default:
            coulomb_pair_kernel<PAIR_DSF><<<grid, 256, 0, st>>>(n, ps, coord, cv, q, p, e_atom, gq, forces, virial_atom, accumulate_e, atom_lo);
            break;
    }
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

int launch_d3(int n, const PairSource& ps, const float* coord, const CellView& cv, const int32_t* numbers,
              const D3Params& p, float* cn, float* wtab, float* dEdCN, double* e_atom, float* forces, double* virial_atom,
              cudaStream_t st) {
    if (n == 0) return AIMNET_OK;
    dim3 grid((n + 7) / 8);
    d3_cn_kernel<<<grid, 256, 0, st>>>(n, ps, coord, cv, numbers, p, cn);
    AIM_LAUNCH_CHECK();
    d3_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, numbers, p, cn, wtab);
    AIM_LAUNCH_CHECK();
    d3_energy_kernel<<<grid, 256, 0, st>>>(n, ps, coord, cv, numbers, p, wtab, e_atom, dEdCN, forces, virial_atom);
    AIM_LAUNCH_CHECK();
    if (forces) {
        d3_cn_force_kernel<<<grid, 256, 0, st>>>(n, ps, coord, cv, numbers, p, dEdCN, forces, virial_atom);
        AIM_LAUNCH_CHECK();
    }
    return AIMNET_OK;
}

}  // namespace aimnet
