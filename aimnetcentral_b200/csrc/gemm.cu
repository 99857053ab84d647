// Per-atom MLP GEMMs: Y[M,N] = epilogue(A[M,K] @ W[N,K]^T)   (aimnet/modules/core.py:11-46)
//
// Both operands are K-major ("NT"), which is how torch stores Linear weights (out,in) and how the activations are
// laid out, and it is the native operand layout of tcgen05.mma.  Epilogue modes:
//   0  plain store                       (input-gradient GEMM of the first layer of a stack)
//   1  + bias                            (last Linear without activation)
//   2  + bias, exact-erf GELU; also stores gelu'(z) to aux so the backward pass never recomputes erf
//   3  * aux[M,N]                        (input-gradient GEMM fused with the GELU derivative of the layer below)
//
// Backend 0 (this file): fp32 SIMT, 128x64x16 tiles, 8x4 register micro-tiles — the bit-faithful baseline.
// Backend 1 (gemm_tc.cu): tcgen05 3xTF32 with TMEM accumulators.
// Backend 2 (gemm_tc16.cu): tcgen05 3xFP16 with per-(row, K-chunk) power-of-two scaling — the default.
#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

constexpr int BM = 128, BN = 64, BK = 16;

template <int MODE>
__global__ void __launch_bounds__(256) gemm_nt_simt_kernel(const float* __restrict__ A, int lda,
                                                           const float* __restrict__ W, int ldw,
                                                           const float* __restrict__ bias, float* __restrict__ Y,
                                                           int ldy, float* __restrict__ aux, int ldaux, int M, int N,
                                                           int K) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Ws[2][BK][BN + 4];
    int tid = threadIdx.x;
    int tx = tid & 15, ty = tid >> 4;
    int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // global->smem mapping: A tile 128 rows x 16 k = 512 float4 (2 per thread); W tile 64 x 16 = 256 float4
    int ar = tid >> 2, ak = (tid & 3) * 4;          // rows ar and ar+64
    int wr = tid >> 2, wk = (tid & 3) * 4;
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    float4 ra0, ra1, rw;
    auto gload = [&](int k0) {
        int r0 = m0 + ar, r1 = m0 + ar + 64;
        ra0 = (r0 < M) ? *reinterpret_cast<const float4*>(A + (size_t)r0 * lda + k0 + ak) : make_float4(0, 0, 0, 0);
        ra1 = (r1 < M) ? *reinterpret_cast<const float4*>(A + (size_t)r1 * lda + k0 + ak) : make_float4(0, 0, 0, 0);
        int n = n0 + wr;
        rw = (n < N) ? *reinterpret_cast<const float4*>(W + (size_t)n * ldw + k0 + wk) : make_float4(0, 0, 0, 0);
    };
    auto sstore = [&](int buf) {
        As[buf][ak + 0][ar] = ra0.x;
        As[buf][ak + 1][ar] = ra0.y;
        As[buf][ak + 2][ar] = ra0.z;
        As[buf][ak + 3][ar] = ra0.w;
        As[buf][ak + 0][ar + 64] = ra1.x;
        As[buf][ak + 1][ar + 64] = ra1.y;
        As[buf][ak + 2][ar + 64] = ra1.z;
        As[buf][ak + 3][ar + 64] = ra1.w;
        Ws[buf][wk + 0][wr] = rw.x;
        Ws[buf][wk + 1][wr] = rw.y;
        Ws[buf][wk + 2][wr] = rw.z;
        Ws[buf][wk + 3][wr] = rw.w;
    };
    gload(0);
    sstore(0);
    __syncthreads();
    int nk = K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
    int col = n0 + tx * 4;
    if (col >= N) return;
    float4 bz = make_float4(0, 0, 0, 0);
    if (MODE == 1 || MODE == 2) bz = *reinterpret_cast<const float4*>(bias + col);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int row = m0 + ty * 8 + r;
        if (row >= M) continue;
        float4 z = make_float4(acc[r][0] + bz.x, acc[r][1] + bz.y, acc[r][2] + bz.z, acc[r][3] + bz.w);
        if (MODE == 2) {
            float4 gp;
            gelu_pair(z.x, z.x, gp.x);
            gelu_pair(z.y, z.y, gp.y);
            gelu_pair(z.z, z.z, gp.z);
            gelu_pair(z.w, z.w, gp.w);
            if (aux != nullptr) *reinterpret_cast<float4*>(aux + (size_t)row * ldaux + col) = gp;
        } else if (MODE == 3) {
            float4 gp = *reinterpret_cast<const float4*>(aux + (size_t)row * ldaux + col);
            z = make_float4(z.x * gp.x, z.y * gp.y, z.z * gp.z, z.w * gp.w);
        }
        *reinterpret_cast<float4*>(Y + (size_t)row * ldy + col) = z;
    }
}

// Small-M variant of the fp32 SIMT kernel: 16 x 16 output tiles; 256 threads = four groups of 64 (one row x four columns
// each), every group accumulates a quarter of the K chunks (chunks of 32 with register prefetch, own double buffer), the
// partial sums are added in fixed order at the end.  For a single molecule (M = number of atoms, ~100) the tensor-core
// kernels are latency-bound (TMEM allocation, 22 dependent TMA stages, one or two CTAs busy): ~17 us per layer.  This
// kernel spreads the same layer over (M/16) x (N/16) small CTAs (256 for 113 x 512) so that every SM works on it, and the
// four-way split of K cuts the dependent chain of a layer (22 chunks for K = 704) to a quarter: the single-molecule step
// is a chain of ~60 such kernels (profiles/r2_cfg1_*).  It is what the engine uses at or below kSmallM rows.
constexpr int SBM = 16, SBN = 16, SBK = 32, SKG = 4;

template <int MODE>
__global__ void __launch_bounds__(64 * SKG) gemm_nt_small_kernel(const float* __restrict__ A, int lda,
                                                                 const float* __restrict__ W, int ldw,
                                                                 const float* __restrict__ bias, float* __restrict__ Y,
                                                                 int ldy, float* __restrict__ aux, int ldaux, int M, int N,
                                                                 int K) {
    __shared__ float As[SKG][2][SBK][SBM + 1];
    __shared__ __align__(16) float Ws[SKG][2][SBK][SBN + 4];
    const int grp = threadIdx.x >> 6, tid = threadIdx.x & 63;
    const int tx = tid & 3, ty = tid >> 2;          // output: row ty, columns 4 tx .. 4 tx + 3
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    // loads: tile rows lr and lr + 8, k offset lk (two float4 of A and of W per thread)
    const int lr = tid >> 3, lk = (tid & 7) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float4 ra[2], rw[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = m0 + lr + 8 * u, n = n0 + lr + 8 * u;
            ra[u] = (r < M) ? *reinterpret_cast<const float4*>(A + (size_t)r * lda + k0 + lk) : make_float4(0, 0, 0, 0);
            rw[u] = (n < N) ? *reinterpret_cast<const float4*>(W + (size_t)n * ldw + k0 + lk) : make_float4(0, 0, 0, 0);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = lr + 8 * u;
            As[grp][buf][lk + 0][r] = ra[u].x;
            As[grp][buf][lk + 1][r] = ra[u].y;
            As[grp][buf][lk + 2][r] = ra[u].z;
            As[grp][buf][lk + 3][r] = ra[u].w;
            Ws[grp][buf][lk + 0][r] = rw[u].x;
            Ws[grp][buf][lk + 1][r] = rw[u].y;
            Ws[grp][buf][lk + 2][r] = rw[u].z;
            Ws[grp][buf][lk + 3][r] = rw[u].w;
        }
    };
    // group g takes chunks g, g + 4, ...; every group runs the same number of rounds (empty chunks load nothing), so the
    // block-wide barriers stay uniform
    const int nk = K / SBK;
    const int rounds = (nk + SKG - 1) / SKG;
    if (grp < nk) gload(grp * SBK);
    if (grp < nk) sstore(0);
    __syncthreads();
    for (int it = 0; it < rounds; ++it) {
        const int buf = it & 1;
        const int kt = it * SKG + grp, ktn = kt + SKG;
        if (ktn < nk) gload(ktn * SBK);
        if (kt < nk) {
#pragma unroll
            for (int k = 0; k < SBK; ++k) {
                const float a = As[grp][buf][k][ty];
                const float4 b = *reinterpret_cast<const float4*>(&Ws[grp][buf][k][tx * 4]);
                acc[0] = fmaf(a, b.x, acc[0]);
                acc[1] = fmaf(a, b.y, acc[1]);
                acc[2] = fmaf(a, b.z, acc[2]);
                acc[3] = fmaf(a, b.w, acc[3]);
            }
        }
        if (it + 1 < rounds) {
            if (ktn < nk) sstore(buf ^ 1);
            __syncthreads();
        }
    }
    // add the four partial sums in fixed order (group 0 + 1 + 2 + 3): deterministic
    __syncthreads();
    float4* part = reinterpret_cast<float4*>(&Ws[0][0][0][0]);   // 3 x 64 float4 = 3 KB of the (now idle) operand buffers
    if (grp > 0) part[(grp - 1) * 64 + tid] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (grp > 0) return;
#pragma unroll
    for (int g2 = 0; g2 < SKG - 1; ++g2) {
        const float4 pv = part[g2 * 64 + tid];
        acc[0] += pv.x;
        acc[1] += pv.y;
        acc[2] += pv.z;
        acc[3] += pv.w;
    }
    const int row = m0 + ty, col = n0 + tx * 4;
    if (row >= M || col >= N) return;
    float4 bz = make_float4(0, 0, 0, 0);
    if (MODE == 1 || MODE == 2) bz = *reinterpret_cast<const float4*>(bias + col);
    float4 z = make_float4(acc[0] + bz.x, acc[1] + bz.y, acc[2] + bz.z, acc[3] + bz.w);
    if (MODE == 2) {
        float4 gp;
        gelu_pair(z.x, z.x, gp.x);
        gelu_pair(z.y, z.y, gp.y);
        gelu_pair(z.z, z.z, gp.z);
        gelu_pair(z.w, z.w, gp.w);
        if (aux != nullptr) *reinterpret_cast<float4*>(aux + (size_t)row * ldaux + col) = gp;
    } else if (MODE == 3) {
        const float4 gp = *reinterpret_cast<const float4*>(aux + (size_t)row * ldaux + col);
        z = make_float4(z.x * gp.x, z.y * gp.y, z.z * gp.z, z.w * gp.w);
    }
    *reinterpret_cast<float4*>(Y + (size_t)row * ldy + col) = z;
}


int gemm_nt(const float* A, int lda, const WeightView& w, const float* bias, float* Y, int ldy, float* aux, int ldaux, int M,
            int N, int K, int mode, int backend, cudaStream_t st) {
    AIM_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad sizes");
    AIM_REQUIRE(K % BK == 0, "gemm: K must be a multiple of 16 (pad the operands)");
    AIM_REQUIRE(N % 4 == 0 && lda % 4 == 0 && w.ldw % 4 == 0 && ldy % 4 == 0, "gemm: N and leading dims must be multiples of 4");
    AIM_REQUIRE(mode >= 0 && mode <= 3, "gemm: bad epilogue mode");
    AIM_REQUIRE(mode != 3 || aux != nullptr, "gemm: mode 3 needs aux");
    AIM_REQUIRE((mode != 1 && mode != 2) || bias != nullptr, "gemm: bias required");
    if (M == 0) return AIMNET_OK;
    AIM_REQUIRE(backend != 2, "gemm: the 3xFP16 backend takes pre-split activations (gemm_nt_split)");
    if (backend == 1) {
        AIM_REQUIRE(w.Whi != nullptr && w.Wlo != nullptr, "gemm: 3xTF32 backend needs the tf32 split weights");
        return gemm_nt_tc(A, lda, w.Whi, w.Wlo, w.ldw, bias, Y, ldy, aux, ldaux, M, N, K, mode, st);
    }
    AIM_REQUIRE(w.W != nullptr, "gemm: SIMT backend needs the fp32 weights");
    const float* W = w.W;
    const int ldw = w.ldw;
    if (M <= kSmallM && K % SBK == 0) {
        dim3 sgrid((N + SBN - 1) / SBN, (M + SBM - 1) / SBM);
        switch (mode) {
            case 0: gemm_nt_small_kernel<0><<<sgrid, 64 * SKG, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
            case 1: gemm_nt_small_kernel<1><<<sgrid, 64 * SKG, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
            case 2: gemm_nt_small_kernel<2><<<sgrid, 64 * SKG, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
            default: gemm_nt_small_kernel<3><<<sgrid, 64 * SKG, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
        }
        AIM_LAUNCH_CHECK();
        return AIMNET_OK;
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    switch (mode) {
        case 0: gemm_nt_simt_kernel<0><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
        case 1: gemm_nt_simt_kernel<1><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
        case 2: gemm_nt_simt_kernel<2><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
        default: gemm_nt_simt_kernel<3><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, Y, ldy, aux, ldaux, M, N, K); break;
    }
    AIM_LAUNCH_CHECK();
    return AIMNET_OK;
}

// 3xFP16 backend: A pre-split; output fp32 (Ysplit == nullptr) or pre-split for a consuming GEMM
int gemm_nt_split(const SplitMat& A, const WeightView& w, const float* bias, float* Y, int ldy, const SplitMat* Ysplit,
                  float* aux, int ldaux, int M, int N, int K, int mode, int variant, cudaStream_t st) {
    AIM_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad sizes");
    AIM_REQUIRE(mode >= 0 && mode <= 3, "gemm: bad epilogue mode");
    AIM_REQUIRE(mode != 3 || aux != nullptr, "gemm: mode 3 needs aux");
    AIM_REQUIRE((mode != 1 && mode != 2) || bias != nullptr, "gemm: bias required");
    AIM_REQUIRE(w.Wh16 != nullptr && w.Wl16 != nullptr, "gemm: 3xFP16 backend needs the fp16 split weights");
    if (M == 0) return AIMNET_OK;
    if (variant == 3)   // backend 5: two tile streams per SM on CTA pairs (gemm_tc16c.cu)
        return gemm_nt_tc16c(A, w.Wh16, w.Wl16, w.inv_scale16, w.ldw, bias, Y, ldy, Ysplit, aux, ldaux, M, N, K, mode, st);
    if (variant == 2)   // backend 4: two tile streams per SM (gemm_tc16d.cu)
        return gemm_nt_tc16d(A, w.Wh16, w.Wl16, w.inv_scale16, w.ldw, bias, Y, ldy, Ysplit, aux, ldaux, M, N, K, mode, st);
    if (variant == 1)   // experimental backend 3 (gemm_tc16p.cu)
        return gemm_nt_tc16p(A, w.Wh16, w.Wl16, w.inv_scale16, w.ldw, bias, Y, ldy, Ysplit, aux, ldaux, M, N, K, mode, st);
    return gemm_nt_tc16(A, w.Wh16, w.Wl16, w.inv_scale16, w.ldw, bias, Y, ldy, Ysplit, aux, ldaux, M, N, K, mode, st);
}

}  // namespace aimnet

// Debug: per-stage SM-clock stamps of CTA 0 of every following backend-2 launch (8 events x 2048 stages, device buffer
// of 16384 uint64), nullptr to switch off.  Used by tools/gemm_trace.py only.
extern "C" int aimnet2_gemm_set_trace(void* device_buf) {
    aimnet::gemm_tc16_set_trace(reinterpret_cast<unsigned long long*>(device_buf));
    return AIMNET_OK;
}

// Operator seam for tests / tools: fp32 operands in; weights (and, for backend 2, activations) are split on the device.
// Backends 2, 3 and 4 (3 = the experimental pipelined-epilogue kernel of gemm_tc16p.cu; 4 = gemm_tc16d.cu): mode | 16
// makes the kernel write its output pre-split (the GEMM -> GEMM path of the engine), which is
// then expanded back to fp32 into Y so that the caller can check it.
extern "C" int aimnet2_gemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float* Y, int ldy,
                               float* aux, int ldaux, int M, int N, int K, int mode, int backend, void* stream) {
    using namespace aimnet;
    cudaStream_t st = (cudaStream_t)stream;
    AIM_REQUIRE(backend >= 0 && backend <= 5,
                "gemm: backend must be 0 (SIMT), 1 (3xTF32), 2 (3xFP16), 3 (3xFP16, pipelined epilogue: experimental), 4 "
                "(3xFP16, two tile streams per SM) or 5 (two tile streams on CTA pairs, cta_group::2)");
    const bool split_out = (mode & 16) != 0;
    mode &= 15;
    AIM_REQUIRE(!split_out || backend >= 2, "gemm: pre-split output exists only for the 3xFP16 backends");
    WeightView wv{W, nullptr, nullptr, nullptr, nullptr, nullptr, ldw};
    if (backend == 0) return gemm_nt(A, lda, wv, bias, Y, ldy, aux, ldaux, M, N, K, mode, 0, st);
    AIM_REQUIRE(gemm_tc_available(), "gemm: tcgen05 backends not available");
    AIM_REQUIRE(N > 0 && ldw > 0 && M >= 0 && K > 0, "gemm: bad sizes");
    size_t n = (size_t)N * ldw;
    if (backend == 1) {
        float* buf = nullptr;
        AIM_CUDA_CHECK(cudaMallocAsync(&buf, sizeof(float) * n * 2, st));
        wv.Whi = buf;
        wv.Wlo = buf + n;
        int rc = split_tf32(W, buf, buf + n, n, st);
        if (rc == AIMNET_OK) rc = gemm_nt(A, lda, wv, bias, Y, ldy, aux, ldaux, M, N, K, mode, 1, st);
        cudaFreeAsync(buf, st);
        return rc;
    }
    AIM_REQUIRE(K % 32 == 0 && N % 32 == 0, "gemm: backend 2 needs K and N to be multiples of 32");
    // one allocation: [W hi | W lo | scale, scratch | A hi | A lo | A inv | Y hi | Y lo | Y inv]
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t rows = (size_t)(M > 0 ? M : 1);
    size_t o_wh = 0, o_wl = o_wh + up(n * 2), o_sc = o_wl + up(n * 2), o_ah = o_sc + 256, o_al = o_ah + up(rows * K * 2),
           o_ai = o_al + up(rows * K * 2), o_yh = o_ai + up(rows * (K / 32) * 4), o_yl = o_yh + up(rows * N * 2),
           o_yi = o_yl + up(rows * N * 2), total = o_yi + up(rows * (N / 32) * 4);
    char* buf = nullptr;
    AIM_CUDA_CHECK(cudaMallocAsync(&buf, total, st));
    wv.Wh16 = buf + o_wh;
    wv.Wl16 = buf + o_wl;
    float* inv = reinterpret_cast<float*>(buf + o_sc);
    wv.inv_scale16 = inv;
    SplitMat As{buf + o_ah, buf + o_al, reinterpret_cast<float*>(buf + o_ai), K, K / 32};
    SplitMat Ys{buf + o_yh, buf + o_yl, reinterpret_cast<float*>(buf + o_yi), N, N / 32};
    int rc = split_fp16_device(W, buf + o_wh, buf + o_wl, inv, reinterpret_cast<unsigned int*>(inv + 16), n, st);
    if (rc == AIMNET_OK) rc = presplit_f32(A, lda, M, K, As, st);
    if (rc == AIMNET_OK)
        rc = gemm_nt_split(As, wv, bias, Y, ldy, split_out ? &Ys : nullptr, aux, ldaux, M, N, K, mode, backend - 2, st);
    if (rc == AIMNET_OK && split_out) rc = unsplit_f32(Ys, M, N, Y, ldy, st);
    cudaFreeAsync(buf, st);
    return rc;
}
