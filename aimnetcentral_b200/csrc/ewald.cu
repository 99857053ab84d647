// Ewald summation, reciprocal-space part.  The real-space part runs in the pair walker of
// lr.cu (PAIR_EWALD: erfc(alpha d)/d over the long-range list).
//
// Reference: LRCoulomb._coul_nvalchemi(backend="ewald") (aimnet/modules/lr.py:617-707) calls the un-vendored
// nvalchemiops.ewald_summation; its parameters follow Kolafa-Perram (aimnet/calculators/calculator.py:663-666):
//   eta = (V^2/N)^(1/6)/sqrt(2 pi), alpha = 1/(sqrt(2) eta), r_c = t eta, k_c = t/eta, t = sqrt(-2 ln accuracy).
// The Ewald energy does not depend on the split, so the engine caps r_c at 15 A (the list it already builds for
// DFT-D3) and scales alpha / k_c consistently.  Parity is UNPINNED upstream (no known answer in the reference); the CPU
// oracle is the textbook sum restated after aimnet/ops.py:196-273 and validated on the rock-salt Madelung constant.
//
//   E_rec  = k_e (4 pi / V) sum_{k in half space, |k| <= k_c} c_k |S(k)|^2 ,  c_k = exp(-k^2/4 alpha^2)/k^2 ,
//   S(k)   = sum_i q_i exp(i k.r_i)
//   dE/dq_i = k_e (8 pi / V) sum c_k (Re S cos(k.r_i) + Im S sin(k.r_i))
//   F_i     = k_e (8 pi q_i / V) sum c_k k (Re S sin(k.r_i) - Im S cos(k.r_i))
//   dE/deps_ab = k_e (4 pi / V) sum c_k |S|^2 [ -delta_ab + 2 k_a k_b (1/(4 alpha^2) + 1/k^2) ]
//   E_self = -k_e alpha/sqrt(pi) sum q_i^2 ,  E_bg = -k_e pi Q^2 / (2 V alpha^2)
// Phases: k.r = 2 pi (h f1 + k f2 + l f3) with f the fractional coordinates.  f is quantised once per evaluation to
// 32-bit fixed point (2^32 per period, 1.5e-9 rad), the integer combination h F1 + k F2 + l F3 wraps modulo one period
// by construction, and the fp32 sincospi sees an argument in [-1, 1): exact range reduction with three integer
// multiply-adds instead of an fp64 dot product + rint per (atom, k) pair.  Sums run in fp32 per lane and are flushed into
// fp64 every 16 terms.
#include <cmath>
#include <vector>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

constexpr int kEwaldRun = 8;                              // list entries per (h, k, l0 .. l0 + 7) group
constexpr float kPhaseScale = 1.4629180792671596e-9f;     // pi / 2^31: fixed-point phase (int32) -> radians in [-pi, pi)

// (uint32_t)h * F1 + k * F2 + l * F3 is the phase modulo 2^32 = one period; as int32 times pi / 2^31 it is an argument in
// [-pi, pi): the SFU sine / cosine (abs error 4e-7 there) replace the ~30-instruction sincospi polynomial.
// per atom: fractional coordinates in 32-bit fixed point, F_j = frac(sum_c r_c inv[3c + j]) * 2^32, and the charge,
// packed into one 16-byte word (the structure-factor kernel reads one word per (k, atom) pair)
__global__ void __launch_bounds__(256) ewald_frac_kernel(int n, const float* __restrict__ coord, const float* __restrict__ q,
                                                         double i0, double i1, double i2, double i3, double i4, double i5,
                                                         double i6, double i7, double i8, uint4* __restrict__ F) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = coord[3 * i], y = coord[3 * i + 1], z = coord[3 * i + 2];
    const double f[3] = {x * i0 + y * i3 + z * i6, x * i1 + y * i4 + z * i7, x * i2 + y * i5 + z * i8};
    uint32_t o[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double t = f[j] - floor(f[j]);
        o[j] = (uint32_t)(unsigned long long)llrint(t * 4294967296.0);   // 2^32 wraps to 0: the same phase
    }
    F[i] = make_uint4(o[0], o[1], o[2], __float_as_uint(q[i]));
}

// S(k) = sum_i q_i exp(i k.r_i): one warp per group of kSfBlock k vectors, lanes stride over the atoms.  With one k vector per
// warp every warp streamed all atom words through L1 / L2 (1.3 MB per k vector, 42 GB per call at cfg-5: bandwidth-bound at
// 6.8 ms); eight k vectors per loaded word leave the SFU sine / cosine as the limit.
constexpr int kSfBlock = kEwaldRun;
__global__ void __launch_bounds__(256) ewald_sf_kernel(int n, int nk, const uint4* __restrict__ F,
                                                       const int32_t* __restrict__ hkl, double* __restrict__ S) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int k0 = w * kSfBlock;
    if (k0 >= nk) return;   // nk is a multiple of kSfBlock (ewald_prepare pads every (h, k) row)
    const uint32_t H = (uint32_t)hkl[3 * k0], K = (uint32_t)hkl[3 * k0 + 1], L0 = (uint32_t)hkl[3 * k0 + 2];
    double re[kSfBlock], im[kSfBlock];
    float pre[kSfBlock], pim[kSfBlock];
#pragma unroll
    for (int t = 0; t < kSfBlock; ++t) re[t] = im[t] = 0.0, pre[t] = pim[t] = 0.f;
    int cnt = 0;
    // two atoms per lane and iteration: the rotations of a run are a dependent chain, two chains keep the FMA pipe busier
    for (int i = lane; i < n; i += 64) {
        const uint4 a = __ldg(F + i);
        const bool two = i + 32 < n;
        const uint4 b = two ? __ldg(F + i + 32) : make_uint4(0u, 0u, 0u, 0u);   // charge bits 0 = 0.0f: adds nothing
        const float qa = __uint_as_float(a.w), qb = __uint_as_float(b.w);
        float sa, ca, swa, cwa, sb, cb, swb, cwb;
        __sincosf((float)(int32_t)(H * a.x + K * a.y + L0 * a.z) * kPhaseScale, &sa, &ca);   // phase of (h, k, l0)
        __sincosf((float)(int32_t)a.z * kPhaseScale, &swa, &cwa);                            // one step in l
        __sincosf((float)(int32_t)(H * b.x + K * b.y + L0 * b.z) * kPhaseScale, &sb, &cb);
        __sincosf((float)(int32_t)b.z * kPhaseScale, &swb, &cwb);
#pragma unroll
        for (int t = 0; t < kSfBlock; ++t) {
            pre[t] = fmaf(qa, ca, fmaf(qb, cb, pre[t]));
            pim[t] = fmaf(qa, sa, fmaf(qb, sb, pim[t]));
            const float ca2 = fmaf(ca, cwa, -sa * swa), cb2 = fmaf(cb, cwb, -sb * swb);
            sa = fmaf(sa, cwa, ca * swa);
            sb = fmaf(sb, cwb, cb * swb);
            ca = ca2;
            cb = cb2;
        }
        if (++cnt == 8) {   // 16 terms per fp32 partial sum
#pragma unroll
            for (int t = 0; t < kSfBlock; ++t) {
                re[t] += (double)pre[t];
                im[t] += (double)pim[t];
                pre[t] = pim[t] = 0.f;
            }
            cnt = 0;
        }
    }
#pragma unroll
    for (int t = 0; t < kSfBlock; ++t) {
        const double r = warp_sum(re[t] + (double)pre[t]), m = warp_sum(im[t] + (double)pim[t]);
        if (lane == 0) {
            S[2 * (k0 + t)] = r;
            S[2 * (k0 + t) + 1] = m;
        }
    }
}

// total charge of the (single) system, fp64
__global__ void __launch_bounds__(256) ewald_qsum_kernel(int n, const float* __restrict__ q, double* __restrict__ out) {
    __shared__ double red[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)q[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[0] = t;
    }
}

// per k vector, after S(k) is known: one 32-byte record (c_k Re S, c_k Im S, k_x, k_y | k_z, h, k, l) so that the atom
// kernel reads two 16-byte words per (atom, k) pair instead of 60 bytes of fp64 / int arrays
__global__ void __launch_bounds__(256) ewald_pack_kernel(int nk, const int32_t* __restrict__ hkl,
                                                         const double* __restrict__ kvec, const double* __restrict__ ck,
                                                         const double* __restrict__ S, float4* __restrict__ rec) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk) return;
    const double c = ck[k];
    rec[2 * k] = make_float4((float)(c * S[2 * k]), (float)(c * S[2 * k + 1]), (float)kvec[3 * k], (float)kvec[3 * k + 1]);
    rec[2 * k + 1] = make_float4((float)kvec[3 * k + 2], __int_as_float(hkl[3 * k]), __int_as_float(hkl[3 * k + 1]),
                                 __int_as_float(hkl[3 * k + 2]));
}

// dE/dq_i and F_i: thread = atom, every thread of a warp walks the SAME k records (uniform addresses: one broadcast
// transaction per load), blockIdx.y selects a slice of the run list so that small systems still fill the machine; the slices'
// partial sums are added in fixed order by ewald_atom_finish_kernel.  (Round-2 history: one atom per warp with the lanes over
// the records streamed 1 MB of records per atom, 84 GB per call at cfg-5; eight atoms per warp with a run of records per lane
// read them with a 256-byte lane stride and sat at 75 % of the L1 throughput.)
__global__ void __launch_bounds__(128) ewald_atom_kernel(int n, int n_runs, int runs_per_slice, const uint4* __restrict__ F,
                                                         const float4* __restrict__ rec, double4* __restrict__ part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 fa = F[min(i, n - 1)];
    float cw, sw;
    __sincosf((float)(int32_t)fa.z * kPhaseScale, &sw, &cw);   // the atom's phase step per unit of l
    double g = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    float pg = 0.f, px = 0.f, py = 0.f, pz = 0.f;
    const int r0i = blockIdx.y * runs_per_slice, r1i = min(n_runs, r0i + runs_per_slice);
    for (int r = r0i; r < r1i; ++r) {
        const float4* rr = rec + 2 * (size_t)r * kEwaldRun;
        const float4 h1 = __ldg(rr + 1);
        const uint32_t h = (uint32_t)__float_as_int(h1.y), kk = (uint32_t)__float_as_int(h1.z), l = (uint32_t)__float_as_int(h1.w);
        float c, s;
        __sincosf((float)(int32_t)(h * fa.x + kk * fa.y + l * fa.z) * kPhaseScale, &s, &c);
#pragma unroll
        for (int t = 0; t < kEwaldRun; ++t) {
            const float4 a = __ldg(rr + 2 * t), b = __ldg(rr + 2 * t + 1);
            pg = fmaf(a.x, c, fmaf(a.y, s, pg));
            const float tt = fmaf(a.x, s, -a.y * c);
            px = fmaf(tt, a.z, px);
            py = fmaf(tt, a.w, py);
            pz = fmaf(tt, b.x, pz);
            const float c2 = fmaf(c, cw, -s * sw);
            s = fmaf(s, cw, c * sw);
            c = c2;
        }
        if ((r - r0i) & 1) {   // fp32 partial sums hold two runs (16 terms) between flushes into fp64
            g += (double)pg;
            fx += (double)px;
            fy += (double)py;
            fz += (double)pz;
            pg = px = py = pz = 0.f;
        }
    }
    if (i < n) part[(size_t)blockIdx.y * n + i] = make_double4(g + (double)pg, fx + (double)px, fy + (double)py, fz + (double)pz);
}

__global__ void __launch_bounds__(256) ewald_atom_finish_kernel(int n, int n_slices, const double4* __restrict__ part,
                                                                const float* __restrict__ q, double pref, double self_coeff,
                                                                double bg_unit, const double* __restrict__ qsum,
                                                                double* __restrict__ e_atom, float* __restrict__ gq,
                                                                float* __restrict__ forces) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double gs = 0.0, xs = 0.0, ys = 0.0, zs = 0.0;
    for (int sl = 0; sl < n_slices; ++sl) {
        const double4 v = part[(size_t)sl * n + i];
        gs += v.x, xs += v.y, ys += v.z, zs += v.w;
    }
    const double qi = (double)q[i];
    // 2*pref = k_e 8 pi / V ; self: E = self_coeff q^2 ; background: dE/dq = bg_coeff (already times Q)
    gq[i] += (float)(2.0 * pref * gs + 2.0 * self_coeff * qi + 2.0 * bg_unit * qsum[0]);
    e_atom[i] += self_coeff * qi * qi;
    if (forces) {
        forces[3 * i + 0] += (float)(2.0 * pref * qi * xs);
        forces[3 * i + 1] += (float)(2.0 * pref * qi * ys);
        forces[3 * i + 2] += (float)(2.0 * pref * qi * zs);
    }
}

// slices of the run list per atom: enough threads (atoms x slices) to fill the machine
static int ewald_atom_slices(int n_atoms, int n_runs) {
    int s = (148 * 2048 + n_atoms - 1) / std::max(1, n_atoms);
    return std::max(1, std::min(std::min(s, 64), std::max(1, n_runs)));
}

// single block: reciprocal energy and its strain derivative; added to atom 0's per-atom accumulators
__global__ void __launch_bounds__(256) ewald_energy_kernel(int nk, const double* __restrict__ kvec,
                                                           const double* __restrict__ ck, const double* __restrict__ S,
                                                           double pref, double inv4a2, double bg_unit,
                                                           const double* __restrict__ qsum,
                                                           double* __restrict__ e_atom, double* __restrict__ virial_atom) {
    __shared__ double red[8];
    const double e_bg = bg_unit * qsum[0] * qsum[0];   // E_bg = -k_e pi Q^2 / (2 V alpha^2)
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
        double kx = kvec[3 * k], ky = kvec[3 * k + 1], kz = kvec[3 * k + 2];
        double s2 = S[2 * k] * S[2 * k] + S[2 * k + 1] * S[2 * k + 1];
        double e = ck[k] * s2;
        double k2 = kx * kx + ky * ky + kz * kz;
        double f = 2.0 * (inv4a2 + 1.0 / k2) * e;
        acc[0] += e;
        acc[1] += f * kx * kx - e;
        acc[2] += f * kx * ky;
        acc[3] += f * kx * kz;
        acc[4] += f * ky * kx;
        acc[5] += f * ky * ky - e;
        acc[6] += f * ky * kz;
        acc[7] += f * kz * kx;
        acc[8] += f * kz * ky;
        acc[9] += f * kz * kz - e;
    }
    for (int k = 0; k < 10; ++k) {
        double v = warp_sum(acc[k]);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            if (k == 0) {
                e_atom[0] += pref * t + e_bg;
            } else if (virial_atom) {
                double bg = ((k == 1 || k == 5 || k == 9) ? -e_bg : 0.0);   // d E_bg / d eps_aa = -E_bg
                virial_atom[k - 1] += pref * t + bg;
            }
        }
    }
}


void ewald_parameters(const float* host_cell, int n_atoms, double accuracy, double rc_cap, double& alpha, double& rc,
                      double& kc, double& volume) {
    double a[9];
    for (int k = 0; k < 9; ++k) a[k] = host_cell[k];
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    volume = std::fabs(det);
    double t = std::sqrt(-2.0 * std::log(accuracy));
    double eta = std::pow(volume * volume / std::max(1, n_atoms), 1.0 / 6.0) / std::sqrt(2.0 * M_PI);
    rc = t * eta;
    if (rc_cap > 0 && rc > rc_cap) rc = rc_cap;
    eta = rc / t;
    alpha = 1.0 / (std::sqrt(2.0) * eta);
    kc = t / eta;
}

int ewald_prepare(EwaldPlan& pl, const float* host_cell, int n_atoms, double accuracy, double rc_cap, cudaStream_t st) {
    bool same = pl.n_atoms == n_atoms && pl.accuracy == accuracy && pl.rc_cap == rc_cap && pl.nk > 0;
    for (int k = 0; k < 9 && same; ++k) same = pl.cell[k] == (double)host_cell[k];
    if (same) return AIMNET_OK;
    for (int k = 0; k < 9; ++k) pl.cell[k] = host_cell[k];
    pl.n_atoms = n_atoms;
    pl.accuracy = accuracy;
    pl.rc_cap = rc_cap;
    ewald_parameters(host_cell, n_atoms, accuracy, rc_cap, pl.alpha, pl.rc, pl.kc, pl.volume);
    const double* a = pl.cell;
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    AIM_REQUIRE(std::fabs(det) > 1e-9, "ewald: singular cell");
    double inv[9];
    inv[0] = (a[4] * a[8] - a[5] * a[7]) / det;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) / det;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) / det;
    inv[3] = (a[5] * a[6] - a[3] * a[8]) / det;
    inv[4] = (a[0] * a[8] - a[2] * a[6]) / det;
    inv[5] = (a[2] * a[3] - a[0] * a[5]) / det;
    inv[6] = (a[3] * a[7] - a[4] * a[6]) / det;
    inv[7] = (a[1] * a[6] - a[0] * a[7]) / det;
    inv[8] = (a[0] * a[4] - a[1] * a[3]) / det;
    // reciprocal vectors b_j = 2 pi * column j of inv(cell)
    double b[3][3];
    for (int j = 0; j < 3; ++j)
        for (int c = 0; c < 3; ++c) b[j][c] = 2.0 * M_PI * inv[3 * c + j];
    int nmax[3];
    for (int j = 0; j < 3; ++j) {
        double an = std::sqrt(a[3 * j] * a[3 * j] + a[3 * j + 1] * a[3 * j + 1] + a[3 * j + 2] * a[3 * j + 2]);
        nmax[j] = (int)std::ceil(pl.kc * an / (2.0 * M_PI));
    }
    for (int c = 0; c < 9; ++c) pl.inv[c] = inv[c];
    std::vector<double> kv, ck;
    std::vector<int32_t> hk;
    double kc2 = pl.kc * pl.kc, inv4a2 = 1.0 / (4.0 * pl.alpha * pl.alpha);
    // Rows of fixed (h, k) hold consecutive l (the part of the row inside the sphere is an interval); every row is padded to a
    // multiple of kEwaldRun vectors with weight c_k = 0, so that each aligned group of kEwaldRun list entries shares (h, k)
    // and steps l by one: the kernels evaluate one sine / cosine per group and rotate by the atom's own l-step phase.
    for (int h = 0; h <= nmax[0]; ++h)
        for (int k = (h == 0 ? 0 : -nmax[1]); k <= nmax[1]; ++k) {
            int in_row = 0, l_next = 0;
            for (int l = ((h == 0 && k == 0) ? 1 : -nmax[2]); l <= nmax[2]; ++l) {
                double kx = h * b[0][0] + k * b[1][0] + l * b[2][0];
                double ky = h * b[0][1] + k * b[1][1] + l * b[2][1];
                double kz = h * b[0][2] + k * b[1][2] + l * b[2][2];
                double k2 = kx * kx + ky * ky + kz * kz;
                if (k2 > kc2) {
                    if (in_row > 0) break;   // past the interval
                    continue;
                }
                kv.push_back(kx);
                kv.push_back(ky);
                kv.push_back(kz);
                ck.push_back(std::exp(-k2 * inv4a2) / k2);
                hk.push_back(h);
                hk.push_back(k);
                hk.push_back(l);
                ++in_row;
                l_next = l + 1;
            }
            for (; in_row % kEwaldRun != 0; ++in_row, ++l_next) {   // padding: outside the sphere, weight zero
                kv.push_back(h * b[0][0] + k * b[1][0] + l_next * b[2][0]);
                kv.push_back(h * b[0][1] + k * b[1][1] + l_next * b[2][1]);
                kv.push_back(h * b[0][2] + k * b[1][2] + l_next * b[2][2]);
                ck.push_back(0.0);
                hk.push_back(h);
                hk.push_back(k);
                hk.push_back(l_next);
            }
        }
    pl.nk = (int)ck.size();
    if (pl.nk > pl.cap) {
        if (pl.d_kvec) cudaFree(pl.d_kvec);
        if (pl.d_hkl) cudaFree(pl.d_hkl);
        pl.cap = pl.nk + pl.nk / 4 + 64;
        AIM_CUDA_CHECK(cudaMalloc((void**)&pl.d_kvec, sizeof(double) * (6 * pl.cap + 2)));
        AIM_CUDA_CHECK(cudaMalloc((void**)&pl.d_hkl, sizeof(int32_t) * (4 + 8) * pl.cap));   // hkl (padded to 4) | records
        pl.d_ck = pl.d_kvec + 3 * pl.cap;
        pl.d_S = pl.d_ck + pl.cap;
    }
    if (n_atoms > pl.frac_cap) {
        if (pl.d_frac) cudaFree(pl.d_frac);
        pl.frac_cap = n_atoms + n_atoms / 8 + 64;
        AIM_CUDA_CHECK(cudaMalloc((void**)&pl.d_frac, sizeof(uint32_t) * 4 * pl.frac_cap));
    }
    {
        const size_t need = (size_t)ewald_atom_slices(n_atoms, pl.nk / kEwaldRun) * n_atoms;
        if (need > pl.part_cap) {
            if (pl.d_part) cudaFree(pl.d_part);
            pl.part_cap = need + need / 8 + 64;
            AIM_CUDA_CHECK(cudaMalloc((void**)&pl.d_part, sizeof(double) * 4 * pl.part_cap));
        }
    }
    if (pl.nk > 0) {
        AIM_CUDA_CHECK(cudaMemcpyAsync(pl.d_kvec, kv.data(), sizeof(double) * 3 * pl.nk, cudaMemcpyHostToDevice, st));
        AIM_CUDA_CHECK(cudaMemcpyAsync(pl.d_ck, ck.data(), sizeof(double) * pl.nk, cudaMemcpyHostToDevice, st));
        AIM_CUDA_CHECK(cudaMemcpyAsync(pl.d_hkl, hk.data(), sizeof(int32_t) * 3 * pl.nk, cudaMemcpyHostToDevice, st));
        AIM_CUDA_CHECK(cudaStreamSynchronize(st));   // kv / ck are stack-scoped host vectors
    }
    return AIMNET_OK;
}

void ewald_release(EwaldPlan& pl) {
    if (pl.d_kvec) cudaFree(pl.d_kvec);
    if (pl.d_hkl) cudaFree(pl.d_hkl);
    if (pl.d_frac) cudaFree(pl.d_frac);
    if (pl.d_part) cudaFree(pl.d_part);
    pl.d_part = nullptr;
    pl.part_cap = 0;
    pl.d_kvec = pl.d_ck = pl.d_S = nullptr;
    pl.d_hkl = nullptr;
    pl.d_frac = nullptr;
    pl.cap = pl.nk = pl.frac_cap = 0;
}

// adds the reciprocal, self and background terms to e_atom / gq / forces / virial_atom
int launch_ewald_recip(const EwaldPlan& pl, int n, const float* coord, const float* q, double* e_atom, float* gq,
                       float* forces, double* virial_atom, cudaStream_t st, double ke) {
    if (n == 0) return AIMNET_OK;
    const double pref = ke * 4.0 * M_PI / pl.volume;
    const double self_coeff = -ke * pl.alpha / std::sqrt(M_PI);
    const double bg_unit = -ke * M_PI / (2.0 * pl.volume * pl.alpha * pl.alpha);
    double* d_q = pl.d_S + 2 * (size_t)pl.cap - 0;   // one spare double behind S (see ewald_prepare)
    ewald_qsum_kernel<<<1, 256, 0, st>>>(n, q, d_q);
    AIM_LAUNCH_CHECK();
    AIM_REQUIRE(n <= pl.frac_cap, "ewald: plan prepared for fewer atoms");
    const double* iv = pl.inv;
    uint4* fq = reinterpret_cast<uint4*>(pl.d_frac);
    ewald_frac_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, coord, q, iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7],
                                                       iv[8], fq);
    AIM_LAUNCH_CHECK();
    if (pl.nk > 0) {
        ewald_sf_kernel<<<(pl.nk / kSfBlock + 7) / 8, 256, 0, st>>>(n, pl.nk, fq, pl.d_hkl, pl.d_S);
        AIM_LAUNCH_CHECK();
    }
    float4* rec = reinterpret_cast<float4*>(pl.d_hkl + 4 * (size_t)pl.cap);   // 16-byte aligned: cap * 16 bytes in
    if (pl.nk > 0) {
        ewald_pack_kernel<<<(pl.nk + 255) / 256, 256, 0, st>>>(pl.nk, pl.d_hkl, pl.d_kvec, pl.d_ck, pl.d_S, rec);
        AIM_LAUNCH_CHECK();
    }
    const int n_runs = pl.nk / kEwaldRun, slices = ewald_atom_slices(n, n_runs);
    AIM_REQUIRE((size_t)slices * n <= pl.part_cap, "ewald: plan prepared for fewer atoms");
    double4* part = reinterpret_cast<double4*>(pl.d_part);
    if (n_runs > 0) {
        const int per = (n_runs + slices - 1) / slices;
        ewald_atom_kernel<<<dim3((n + 127) / 128, slices), 128, 0, st>>>(n, n_runs, per, fq, rec, part);
        AIM_LAUNCH_CHECK();
    }
    ewald_atom_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, n_runs > 0 ? slices : 0, part, q, pref, self_coeff, bg_unit, d_q,
                                                             e_atom, gq, forces);
    AIM_LAUNCH_CHECK();
    ewald_energy_kernel<<<1, 256, 0, st>>>(pl.nk, pl.d_kvec, pl.d_ck, pl.d_S, pref, 1.0 / (4.0 * pl.alpha * pl.alpha), bg_unit,
                                           d_q, e_atom, virial_atom);
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

}  // namespace aimnet
