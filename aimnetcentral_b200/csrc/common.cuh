// Shared helpers for the AIMNet2 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/aimnet2_b200.h"

namespace aimnet {

constexpr int kA = 16;   // nfeature        (aimnet/models/aimnet2.yaml)
constexpr int kG = 16;   // nshifts_s
constexpr int kH = 12;   // ncomb_v
constexpr int kAG = kA * kG;       // 256
constexpr int kAH = kA * kH;       // 192
constexpr int kTA = kA * kH * 3;   // 576  saved vector mixing T[a,h,d]
constexpr float kPi = 3.14159265358979323846f;
constexpr double kHartree = 27.211386024367243;   // aimnet/constants.py:6
constexpr double kBohr = 0.5291772105638411;      // aimnet/constants.py:8

void set_error(const std::string& msg);
extern thread_local int g_launch_count;

#define AIM_CUDA_CHECK(expr)                                                                        \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::aimnet::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                \
            cudaGetLastError(); /* reported here: do not leave it for the next launch check */      \
            return AIMNET_ECUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define AIM_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        ::aimnet::g_launch_count++;                                                                 \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            ::aimnet::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) +   \
                                " at " + __FILE__ + ":" + std::to_string(__LINE__));                \
            return AIMNET_ECUDA;                                                                    \
        }                                                                                           \
    } while (0)

// propagate a non-zero status of a launcher
#define AIM_TRY(expr)                      \
    do {                                   \
        int _rc = (expr);                  \
        if (_rc != AIMNET_OK) return _rc;  \
    } while (0)

#define AIM_REQUIRE(cond, msg)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            ::aimnet::set_error(std::string("invalid argument: ") + msg);                          \
            return AIMNET_EINVAL;                                                                   \
        }                                                                                           \
    } while (0)

// Launch configuration caches are per device: cudaFuncSetAttribute and the SM count belong to a device, not to the
// process (one process may own engines on several GPUs).  Entries are idempotent, so unsynchronised writers agree.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
    return d;
}

struct AevParams {
    float shifts[kG];
    float eta;
    float rc;
};

// Where the geometry of neighbor slot (i, m) comes from.
struct NbView {
    const int32_t* nbmat;    // (n_atoms[+1], width) or nullptr in segment mode
    const int32_t* shifts;   // (rows, width, 3) or nullptr
    const int32_t* count;    // (n_atoms) valid-prefix length per row, or nullptr = scan full width
    int width;
    int sentinel;            // n_atoms
};

struct CellView {
    const float* cell;       // (n_cells,3,3) or nullptr
    int n_cells;             // 0, 1, or n_mol
};


// an activation matrix in the pre-split form the 3xFP16 GEMM consumes and produces (gemm_tc16.cu): fp16 hi and lo of
// s * x with one power-of-two scale per (row, 32-column chunk); inv holds 1 / s
struct SplitMat {
    void* hi;     // (rows, ld) __half
    void* lo;     // (rows, ld) __half
    float* inv;   // (rows, ldinv) fp32, ldinv >= cols / 32
    int ld, ldinv;
};

// Ewald reciprocal-space plan of one periodic system (ewald.cu): k vectors and coefficients for the current cell, device
// scratch for the structure factors and the per-atom fixed-point phases.  Owned by the engine, rebuilt when the cell, the
// atom count or the accuracy changes.
struct EwaldPlan {
    double cell[9] = {0};
    double accuracy = 0, rc_cap = 0;
    int n_atoms = 0;
    double alpha = 0, rc = 0, kc = 0, volume = 0;
    int nk = 0;
    double* d_kvec = nullptr;
    double* d_ck = nullptr;
    double* d_S = nullptr;
    int cap = 0;
    int32_t* d_hkl = nullptr;    // (cap, 3) integer reciprocal-lattice indices of the k vectors, then (cap, 8) fp32 records
    uint32_t* d_frac = nullptr;  // (frac_cap, 4) fixed-point fractional coordinates of the current positions + charge bits
    int frac_cap = 0;
    double* d_part = nullptr;    // (slices, n_atoms, 4) partial sums of the per-atom reciprocal-space terms
    size_t part_cap = 0;         // in (atom, slice) entries
    double inv[9] = {0};         // inverse cell
};
int ewald_prepare(EwaldPlan& pl, const float* host_cell, int n_atoms, double accuracy, double rc_cap, cudaStream_t st);
// Kolafa-Perram style parameters for a target accuracy (aimnet/calculators/calculator.py:663-666): eta from V and N,
// r_c = sqrt(-2 ln eps) eta (capped by rc_cap when > 0), k_c = sqrt(-2 ln eps) / eta, alpha = 1 / (sqrt(2) eta)
void ewald_parameters(const float* host_cell, int n_atoms, double accuracy, double rc_cap, double& alpha, double& rc,
                      double& kc, double& volume);
void ewald_release(EwaldPlan& pl);

// at or below this many rows (atoms) the per-atom MLPs run on the small-M fp32 SIMT kernel: the tensor-core pipelines are
// latency-bound there (gemm.cu)
constexpr int kSmallM = 512;

// one Linear's weight matrix (N, ldw) in the forms the GEMM backends consume
struct WeightView {
    const float* W;            // fp32 (SIMT backend 0)
    const float* Whi;          // tf32 hi / lo split (tcgen05 3xTF32 backend 1)
    const float* Wlo;
    const void* Wh16;          // fp16 hi / lo split of s_w * W (tcgen05 3xFP16 backend 2)
    const void* Wl16;
    const float* inv_scale16;  // device scalar 1 / s_w
    int ldw;
};

// neighbour source of the pair-potential walkers (lr.cu): matrix row, or the molecule's own atom segment
struct PairSource {
    NbView nb;                   // nb.nbmat == nullptr -> segment mode
    const int32_t* mol_idx;
    const int32_t* mol_ptr;
    float seg_cut2;              // segment mode only: skip pairs with d2 >= seg_cut2 (<= 0: no cutoff)
};

struct CoulombParams {
    float rc;         // SR envelope radius or DSF cutoff
    float alpha;      // DSF
    float shift_val, shift_slope, self_coeff;   // DSF constants (aimnet/modules/lr.py:594-606)
    double factor;    // signed prefactor c (eV*A): -k for the embedded SR term, +k for the external term
};

constexpr int kC6Row = 28;   // c6ref rows (25 values) padded to a multiple of four floats for 16-byte loads
struct D3Params {
    const float* c6ref;   // (95,95,28): (95,95,5,5) with every 25-value row padded to kC6Row
    const float* cnref;   // (95,5)
    const float* rcov;    // (95)
    const float* r4r2;    // (95)
    float s6, s8, a1, a2;
    float r_on, r_off;    // Bohr
};

enum { PAIR_SR_EXP = 0, PAIR_SR_COS = 1, PAIR_SIMPLE = 2, PAIR_DSF = 3, PAIR_EWALD = 4 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Packed fp32 FMA (sm_100 FFMA2, PTX fma.rn.f32x2): two independent IEEE fp32 FMAs per issued instruction; bitwise
// identical to two fmaf() calls.  ptxas folds the mov.b64 packs away when the halves sit in an aligned register pair
// (float4 loads, accumulators) and takes a plain 32-bit register as a broadcast multiplicand (ffma2s).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rc) : "l"(ra), "l"(rb), "l"(rc));
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}
__device__ __forceinline__ float2 ffma2s(float a, float2 b, float2 c) { return ffma2(make_float2(a, a), b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}

// exact-erf GELU and its derivative (nn.GELU() default, aimnet/modules/core.py:11-46), library form
__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float z) {
    return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
}

// GELU and GELU' together from ONE exponential: with x = |z|/sqrt(2), t = 1/(1 + 0.47 x),
//   erfc(x) = t * P8(t) * exp(-x^2)   (degree-8 fit of erfcx(x)/t on t in (0,1], |error| < 4e-9 in double),
//   Phi(z)  = z > 0 ? 1 - erfc/2 : erfc/2,   gelu = z Phi,   gelu' = Phi + z exp(-z^2/2)/sqrt(2 pi).
// fp32 evaluation: max |gelu - exact| = 3.8e-7 on [-9, 9] (torch's own fp32 GELU: 1.2e-6), max |gelu' - exact| = 2.9e-7.
__device__ __forceinline__ void gelu_pair(float z, float& y, float& gp) {
    const float ax = fabsf(z) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.47f, ax, 1.0f));
    float P = -0.019820483937064207f;
    P = fmaf(P, t, 0.14386611213498762f);
    P = fmaf(P, t, -0.3281399463335257f);
    P = fmaf(P, t, 0.22202211138586878f);
    P = fmaf(P, t, 0.025256349473553902f);
    P = fmaf(P, t, 0.19197703572753022f);
    P = fmaf(P, t, 0.23444773423159065f);
    P = fmaf(P, t, 0.2652225546919865f);
    P = fmaf(P, t, 0.26516877756878143f);
    const float E = __expf(-0.5f * z * z);
    const float e = t * P * E;
    const float Phi = z > 0.f ? fmaf(-0.5f, e, 1.0f) : 0.5f * e;
    y = z * Phi;
    gp = fmaf(z * 0.3989422804014327f, E, Phi);
}

// gelu_pair for two values at once: the polynomial, the products and the final combinations run as packed FFMA2 /
// FMUL2 (about 15 issued instructions per value instead of 26); same formulas, same error bounds.
__device__ __forceinline__ void gelu_pair2(float2 z, float2& y, float2& gp) {
    const float2 az = make_float2(fabsf(z.x), fabsf(z.y));
    const float2 den = ffma2s(0.47f * 0.70710678118654752f, az, make_float2(1.0f, 1.0f));
    float2 t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
    float2 P = make_float2(-0.019820483937064207f, -0.019820483937064207f);
    P = ffma2(P, t, make_float2(0.14386611213498762f, 0.14386611213498762f));
    P = ffma2(P, t, make_float2(-0.3281399463335257f, -0.3281399463335257f));
    P = ffma2(P, t, make_float2(0.22202211138586878f, 0.22202211138586878f));
    P = ffma2(P, t, make_float2(0.025256349473553902f, 0.025256349473553902f));
    P = ffma2(P, t, make_float2(0.19197703572753022f, 0.19197703572753022f));
    P = ffma2(P, t, make_float2(0.23444773423159065f, 0.23444773423159065f));
    P = ffma2(P, t, make_float2(0.2652225546919865f, 0.2652225546919865f));
    P = ffma2(P, t, make_float2(0.26516877756878143f, 0.26516877756878143f));
    // E = exp(-z^2/2) = 2^(-0.5 log2(e) z^2)
    const float2 arg = fmul2(fmul2(z, z), make_float2(-0.72134752044448170f, -0.72134752044448170f));
    float2 E;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E.x) : "f"(arg.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E.y) : "f"(arg.y));
    // h = erfc(|z|/sqrt 2) / 2;  Phi = 1/2 + sign(z) (1/2 - h)   (branch-free form of z > 0 ? 1 - h : h)
    const float2 d = ffma2(fmul2(t, P), fmul2(E, make_float2(-0.5f, -0.5f)), make_float2(0.5f, 0.5f));
    const float2 sd = make_float2(copysignf(d.x, z.x), copysignf(d.y, z.y));
    const float2 Phi = fadd2(sd, make_float2(0.5f, 0.5f));
    y = fmul2(z, Phi);
    gp = ffma2(fmul2(z, make_float2(0.3989422804014327f, 0.3989422804014327f)), E, Phi);
}

// r_ij = x_j + s @ cell - x_i   (aimnet/ops.py:37-66)
__device__ __forceinline__ void pair_vector(const float* __restrict__ coord, int i, int j, const int32_t* sh,
                                            const float* __restrict__ cell, float& rx, float& ry, float& rz) {
    float xj = coord[3 * j + 0], yj = coord[3 * j + 1], zj = coord[3 * j + 2];
    if (sh != nullptr) {
        float s0 = (float)sh[0], s1 = (float)sh[1], s2 = (float)sh[2];
        xj += s0 * cell[0] + s1 * cell[3] + s2 * cell[6];
        yj += s0 * cell[1] + s1 * cell[4] + s2 * cell[7];
        zj += s0 * cell[2] + s1 * cell[5] + s2 * cell[8];
    }
    rx = xj - coord[3 * i + 0];
    ry = yj - coord[3 * i + 1];
    rz = zj - coord[3 * i + 2];
}

}  // namespace aimnet
