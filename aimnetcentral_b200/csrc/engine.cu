// Engine: weights resident in HBM + one-call E/F/stress evaluation (C ABI in include/aimnet2_b200.h).
// Orchestrates the kernels of nblist.cu / conv.cu / gemm*.cu / pointwise.cu / lr.cu in the order of
// AIMNet2Calculator.eval (aimnet/calculators/calculator.py:879-947) and AIMNet2.forward
// (aimnet/models/aimnet2.py:141-187); the reverse pass replaces the reference's single torch.autograd.grad call
// (aimnet/calculators/derivatives.py:96-146) with analytic kernels.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

static thread_local std::string g_error;
thread_local int g_launch_count = 0;
void set_error(const std::string& msg) { g_error = msg; }

static inline int pad32(int x) { return (x + 31) / 32 * 32; }
static inline int round16(int x) { return (x + 15) / 16 * 16; }

struct Region {   // one carved workspace buffer (debug layout)
    const char* name;
    size_t off, bytes;
};

struct Linear {
    int in = 0, out = 0, in_pad = 0, out_pad = 0;
    float* W = nullptr;    // (out_pad, in_pad)
    float* Wt = nullptr;   // (in_pad, out_pad)
    float* b = nullptr;    // (out_pad)
    // tf32 hi/lo splits of W and Wt for the tcgen05 3xTF32 backend
    float *Whi = nullptr, *Wlo = nullptr, *Wthi = nullptr, *Wtlo = nullptr;
    // fp16 hi/lo splits of s_w * W and s_w * Wt (one power-of-two scale per tensor) for the 3xFP16 backend
    __half *Wh16 = nullptr, *Wl16 = nullptr, *Wth16 = nullptr, *Wtl16 = nullptr;
    float* inv_scale16 = nullptr;   // device scalar 1 / s_w
    WeightView fwd() const { return WeightView{W, Whi, Wlo, Wh16, Wl16, inv_scale16, in_pad}; }
    WeightView bwd() const { return WeightView{Wt, Wthi, Wtlo, Wth16, Wtl16, inv_scale16, out_pad}; }
};

}  // namespace aimnet

using namespace aimnet;

struct aimnet2_engine {
    int device = 0;
    int C = 1;
    AevParams aev{};
    float *afv = nullptr, *agh_a = nullptr, *agh_q = nullptr, *w3 = nullptr;
    float* afvT = nullptr;            // the embedding table in the conv kernels' gather layout (first-pass backward by species)
    int species_pass0 = 1;            // 1 = backward of the first convolution by species tables (conv.cu), 0 = generic kernel
    int n_impl_species = 0;           // embedding rows that are not NaN (exported models mark unimplemented species that way)
    float b3 = 0.f;
    double* sae = nullptr;
    std::vector<Linear> mlp[3];
    Linear head[2];
    float sr_rc = 4.6f;
    int sr_envelope = 0;
    float *d3_c6ref = nullptr, *d3_cnref = nullptr, *d3_rcov = nullptr, *d3_r4r2 = nullptr;
    aimnet2_options_t opt{};
    int gemm_backend = 0;
    int poison = -1;              // test seam: byte written over the workspace before every evaluation (-1 = off)
    int conv_impl = 1;            // 0 = list kernels (conv.cu) always; for batches of small molecules: 1 = dense shared-memory forward (conv_dense.cu) + list backward, 2 = dense forward and backward
    int dense_min_mol = 64;       // the dense walk needs enough molecules to give every SM one (a CTA works on one molecule at a time)
    bool dense_now = false;       // the evaluation in flight walks molecule segments instead of matrix rows
    int last_max_seg = 0;         // largest molecule (atoms) of the last batch whose lists the engine built
    int small_m_rows = kSmallM;   // at or below this many atoms the MLPs run on the small-M fp32 SIMT kernel (0 = never)
    int backend_now = 0;    // backend of the evaluation in flight (gemm_backend or 0 for small systems)
    // workspace (grow-only)
    char* ws = nullptr;
    size_t ws_bytes = 0;
    int sr_cap = 64, lr_cap = 256;   // row capacities of the neighbor matrices: grow on overflow, shrink with hysteresis
    int sr_cap_next = 0, lr_cap_next = 0;   // shrunk capacities, adopted at the start of the next evaluation
    int ws_slack_evals = 0;          // consecutive evaluations that needed less than half of the workspace
    int last_sr_width = 0, last_lr_width = 0;
    int last_launches = 0;
    // host staging for eval_host
    char* stage = nullptr;
    size_t stage_bytes = 0;
    cudaStream_t own_stream = nullptr;
    int32_t* pinned_int = nullptr;   // page-locked scratch for small device->host read-backs
    // timing
    int timing = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms[5] = {0, 0, 0, 0, 0};
    // per-launch events of the two dominant kernel classes (timing level 2): pairs (begin, end) in launch order;
    // class 0 = per-atom MLP GEMMs (+ presplit), class 1 = AEV / conv_sv forward and backward
    std::vector<cudaEvent_t> gemm_ev;
    std::vector<int> gemm_ev_class;
    int gemm_ev_used = 0;
    float last_gemm_ms = 0.f, last_conv_ms = 0.f;
    int last_gemm_launches = 0, last_conv_launches = 0;
    std::vector<EwaldPlan> ewald;   // one reciprocal-space plan per periodic system of the batch
    std::vector<int32_t> sys_lo;    // host copy of the molecule segment pointers (batched Ewald)
    // Verlet-skin reuse of the neighbor lists (options.neighbor_skin > 0): what the lists in the workspace were built for
    struct {
        bool valid = false;
        int N = 0, B = 0, n_cells = 0, sr_cap = 0, lr_cap = 0;
        bool need_lr = false, dense = false;
        float sr_cut = 0.f, lr_cut = 0.f, skin = 0.f;
        const char* ws = nullptr;
        bool has_mol = false;
        std::vector<float> cell;
        std::vector<uint8_t> pbc;
        int builds = 0, reuses = 0;
    } skin;
    // CUDA-graph replay of the fixed-shape step: the kernels of one evaluation captured once per
    // (sizes, flags, options, staging buffers) signature and relaunched as one graph
    struct GraphState {
        bool enabled = false, capturing = false, valid = false, pending = false;
        cudaGraphExec_t exec = nullptr;
        std::vector<unsigned char> sig, pending_sig;
        int launches = 0, captures = 0, fallbacks = 0;
    } graph;
    std::vector<void*> owned;
    std::vector<aimnet::Region> layout;   // names / offsets of the workspace buffers of the last evaluation (debug)
    bool deterministic = false;           // recorded only: every kernel of the engine is run-to-run reproducible
};

namespace aimnet {

template <typename T>
static int upload(aimnet2_engine* e, T** dst, const T* src, size_t count) {
    AIM_CUDA_CHECK(cudaMalloc((void**)dst, sizeof(T) * count));
    e->owned.push_back(*dst);
    AIM_CUDA_CHECK(cudaMemcpy(*dst, src, sizeof(T) * count, cudaMemcpyHostToDevice));
    return AIMNET_OK;
}

static int make_linear(aimnet2_engine* e, Linear& L, const float* w, const float* b, int in, int out) {
    L.in = in;
    L.out = out;
    L.in_pad = pad32(in);
    L.out_pad = pad32(out);
    std::vector<float> W((size_t)L.out_pad * L.in_pad, 0.f), Wt((size_t)L.in_pad * L.out_pad, 0.f), B(L.out_pad, 0.f);
    for (int o = 0; o < out; ++o) {
        for (int i = 0; i < in; ++i) {
            float v = w[(size_t)o * in + i];
            W[(size_t)o * L.in_pad + i] = v;
            Wt[(size_t)i * L.out_pad + o] = v;
        }
        B[o] = b[o];
    }
    int rc;
    if ((rc = upload(e, &L.W, W.data(), W.size()))) return rc;
    if ((rc = upload(e, &L.Wt, Wt.data(), Wt.size()))) return rc;
    if ((rc = upload(e, &L.b, B.data(), B.size()))) return rc;
    auto split = [](const std::vector<float>& src, std::vector<float>& hi, std::vector<float>& lo) {
        hi.resize(src.size());
        lo.resize(src.size());
        for (size_t k = 0; k < src.size(); ++k) {
            // round-to-nearest (ties away) to tf32, like cvt.rna.tf32.f32
            uint32_t v, h, l;
            std::memcpy(&v, &src[k], 4);
            h = (v + 0x1000u) & 0xffffe000u;
            float hf, lf;
            std::memcpy(&hf, &h, 4);
            lf = src[k] - hf;
            std::memcpy(&l, &lf, 4);
            l = (l + 0x1000u) & 0xffffe000u;
            std::memcpy(&lf, &l, 4);
            hi[k] = hf;
            lo[k] = lf;
        }
    };
    std::vector<float> hi, lo;
    split(W, hi, lo);
    if ((rc = upload(e, &L.Whi, hi.data(), hi.size()))) return rc;
    if ((rc = upload(e, &L.Wlo, lo.data(), lo.size()))) return rc;
    split(Wt, hi, lo);
    if ((rc = upload(e, &L.Wthi, hi.data(), hi.size()))) return rc;
    if ((rc = upload(e, &L.Wtlo, lo.data(), lo.size()))) return rc;
    // 3xFP16: scale the tensor so that max|W| lands in [2^13, 2^14) (gemm_tc16.cu), then hi = rn(sW), lo = rn(sW - hi)
    float wmax = 0.f;
    for (float v : W) wmax = std::max(wmax, std::fabs(v));
    int ex = 0;
    if (wmax > 0.f) std::frexp(wmax, &ex);   // wmax = f * 2^ex, f in [0.5, 1)  ->  floor(log2) = ex - 1
    ex = std::min(std::max(ex - 1, -113), 127);
    const float sw = std::ldexp(1.0f, 13 - ex), inv_sw = std::ldexp(1.0f, ex - 13);
    auto split16 = [sw](const std::vector<float>& src, std::vector<__half>& h, std::vector<__half>& l) {
        h.resize(src.size());
        l.resize(src.size());
        for (size_t k = 0; k < src.size(); ++k) {
            float x = src[k] * sw;
            h[k] = __float2half_rn(x);
            l[k] = __float2half_rn(x - __half2float(h[k]));
        }
    };
    std::vector<__half> h16, l16;
    split16(W, h16, l16);
    if ((rc = upload(e, &L.Wh16, h16.data(), h16.size()))) return rc;
    if ((rc = upload(e, &L.Wl16, l16.data(), l16.size()))) return rc;
    split16(Wt, h16, l16);
    if ((rc = upload(e, &L.Wth16, h16.data(), h16.size()))) return rc;
    if ((rc = upload(e, &L.Wtl16, l16.data(), l16.size()))) return rc;
    if ((rc = upload(e, &L.inv_scale16, &inv_sw, 1))) return rc;
    return AIMNET_OK;
}

struct Bump {
    char* base;
    size_t off = 0;
    std::vector<Region>* map = nullptr;   // debug: names / offsets of the carved buffers (aimnet2_engine_debug_layout)
    template <typename T>
    T* take(size_t count, const char* name = nullptr) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        if (map && name) map->push_back(Region{name, off, sizeof(T) * count});
        off += sizeof(T) * count;
        return p;
    }
};

struct Buffers {
    float* coord_w;
    int32_t* mol_ptr;
    int32_t* nb_scratch;
    int32_t *nb_sr, *sh_sr, *cnt_sr, *nb_lr, *sh_lr, *cnt_lr;
    float* a[3];
    float* q[2];
    float* x;
    float *hA, *hB;
    float* y[2];
    float* aim;
    float* gp[3][4];
    float *h1, *h2, *gp_h1, *gp_h2;
    float* T_a[3];
    float* T_q[3];
    float *sumq[2], *sumf[2], *s1;
    double *e_nn, *e_sr, *e_lr, *e_d3;
    float *gq, *cn, *dEdCN, *d3w;
    float *dzA, *dzB, *dx, *dS_a, *dS_q, *grad_a, *grad_q, *da_tot, *dq, *dq_base;
    double* virial_atom;
    float* forces_tmp;
    // pre-split activations of the 3xFP16 GEMM backend.  h16 / aim16 / h1_16 / d16 live in the memory of hA,hB / aim /
    // h1 / dzA,dzB (fp16 hi + fp16 lo = the bytes of the fp32 matrix they replace); only x16, dz32 and the scales are extra.
    SplitMat x16, h16[2], aim16, h1_16, d16[2];
    float* dz32;
    float *dense_fpart, *dense_gqpart;   // dense conv walk: per-quarter partial forces (4, N, 3) / charge gradients (4, N, C)
    float *coord_ref, *wrap_off;   // Verlet skin: positions at list-build time, lattice offset applied by the wrap
    int32_t *skin_flag, *mol_ref;
    int32_t* graph_counts;   // graph replay: widest short-range / long-range row of the replayed build (overflow check afterwards)
    float* sp_table;         // first-pass backward by species: P[i][slot][g][4]
    int32_t* sp_info;        // ... flag, number of slots, atomic number per slot
    uint8_t* sp_slot;        // ... slot of every atom
};

static SplitMat alias_split(float* base, size_t n, int width, float* inv) {
    return SplitMat{base, reinterpret_cast<char*>(base) + n * width * 2, inv, width, width / 32};
}
static SplitMat with_ld(SplitMat m, int ld) {
    m.ld = ld;
    return m;
}

// Row capacity of a neighbor matrix after a successful build whose widest row holds `widest` entries: the reference's
// AdaptiveNeighborList rule (aimnet/calculators/neighbors.py:135-140) -- shrink to widest / 0.75 once the widest row falls
// below 2/3 of the 75 % target, i.e. below half of the capacity, so that small fluctuations do not thrash.  The new
// capacity (0 = keep) applies from the next evaluation: this one's lists are already built in the old layout.  The floor is
// the capacity a fresh engine starts with (rows that narrow cost nothing, and a fresh engine must not rebuild a skin list
// just because its first system was small).
static int shrink_cap(int cap, int widest, int floor_cap) {
    if (2 * widest >= cap) return 0;
    const int want = std::max(floor_cap, round16((widest * 4 + 2) / 3));
    return want < cap ? want : 0;
}

static void carve(aimnet2_engine* e, Bump& bp, Buffers& b, int N, int B, int sr_cap, int lr_cap, bool pbc,
                  bool need_lr, int ldx) {
    int C = e->C;
    size_t n = (size_t)std::max(N, 1);
    b.coord_w = bp.take<float>(n * 3, "coord_w");
    b.mol_ptr = bp.take<int32_t>(B + 2, "mol_ptr");
    b.nb_scratch = bp.take<int32_t>((size_t)B * 4 + 16, "nb_scratch");
    b.nb_sr = bp.take<int32_t>(n * sr_cap, "nb_sr");
    b.sh_sr = pbc ? bp.take<int32_t>(n * sr_cap * 3, "sh_sr") : nullptr;
    b.cnt_sr = bp.take<int32_t>(n, "cnt_sr");
    b.nb_lr = need_lr ? bp.take<int32_t>(n * lr_cap, "nb_lr") : nullptr;
    b.sh_lr = need_lr ? bp.take<int32_t>(n * lr_cap * 3, "sh_lr") : nullptr;
    b.cnt_lr = bp.take<int32_t>(n, "cnt_lr");
    static const char* const kNameA[3] = {"a0", "a1", "a2"};
    static const char* const kNameQ[2] = {"q0", "q1"};
    static const char* const kNameY[2] = {"y0", "y1"};
    static const char* const kNameGp[3][4] = {{"gp00", "gp01", "gp02", "gp03"}, {"gp10", "gp11", "gp12", "gp13"}, {"gp20", "gp21", "gp22", "gp23"}};
    static const char* const kNameTa[3] = {"T_a0", "T_a1", "T_a2"};
    static const char* const kNameTq[3] = {"T_q0", "T_q1", "T_q2"};
    for (int p = 0; p < 3; ++p) b.a[p] = bp.take<float>(n * kAG, kNameA[p]);
    for (int p = 0; p < 2; ++p) b.q[p] = bp.take<float>(n * C, kNameQ[p]);
    b.x = bp.take<float>(n * ldx, "x");
    b.hA = bp.take<float>(n * 512, "hA");
    b.hB = bp.take<float>(n * 512, "hB");
    for (int p = 0; p < 2; ++p) b.y[p] = bp.take<float>(n * 288, kNameY[p]);
    b.aim = bp.take<float>(n * 256, "aim");
    for (int p = 0; p < 3; ++p)
        for (int l = 0; l < 4; ++l)
            b.gp[p][l] = (l < (int)e->mlp[p].size()) ? bp.take<float>(n * e->mlp[p][l].out_pad, kNameGp[p][l]) : nullptr;
    b.h1 = bp.take<float>(n * 128, "h1");
    b.h2 = bp.take<float>(n * 128, "h2");
    b.gp_h1 = bp.take<float>(n * 128, "gp_h1");
    b.gp_h2 = bp.take<float>(n * 128, "gp_h2");
    for (int p = 0; p < 3; ++p) {
        b.T_a[p] = bp.take<float>(n * kTA, kNameTa[p]);
        b.T_q[p] = bp.take<float>(n * C * kH * 3, kNameTq[p]);
    }
    for (int p = 0; p < 2; ++p) {
        b.sumq[p] = bp.take<float>((size_t)B * C, p ? "sumq1" : "sumq0");
        b.sumf[p] = bp.take<float>((size_t)B * C, p ? "sumf1" : "sumf0");
    }
    b.s1 = bp.take<float>((size_t)B * C, "s1");
    b.e_nn = bp.take<double>(n, "e_nn");
    b.e_sr = bp.take<double>(n, "e_sr");
    b.e_lr = bp.take<double>(n, "e_lr");
    b.e_d3 = bp.take<double>(n, "e_d3");
    b.gq = bp.take<float>(n, "gq");
    b.cn = bp.take<float>(n, "cn");
    b.dEdCN = bp.take<float>(n, "dEdCN");
    b.d3w = bp.take<float>(n * 16, "d3w");
    b.dzA = bp.take<float>(n * 512, "dzA");
    b.dzB = bp.take<float>(n * 512, "dzB");
    b.dx = bp.take<float>(n * ldx, "dx");
    b.dS_a = bp.take<float>(n * kAG * 4, "dS_a");
    b.dS_q = bp.take<float>(n * C * kG * 4, "dS_q");
    b.grad_a = bp.take<float>(n * kAG, "grad_a");
    b.grad_q = bp.take<float>(n * C, "grad_q");
    b.da_tot = bp.take<float>(n * kAG, "da_tot");
    b.dq = bp.take<float>(n * C, "dq");
    b.dq_base = bp.take<float>(n * C, "dq_base");
    b.virial_atom = bp.take<double>(n * 9, "virial_atom");
    b.forces_tmp = bp.take<float>(n * 3, "forces_tmp");
    b.sp_table = bp.take<float>(n * (size_t)(conv0_species_bytes_per_atom() / 4), "sp_table");
    b.sp_info = bp.take<int32_t>(conv0_species_info_ints(), "sp_info");
    b.sp_slot = bp.take<uint8_t>(n, "sp_slot");
    b.dense_fpart = bp.take<float>(n * 12, "dense_fpart");
    b.dense_gqpart = bp.take<float>(n * 4 * C, "dense_gqpart");
    b.x16 = SplitMat{bp.take<__half>(n * ldx, "x16_hi"), bp.take<__half>(n * ldx, "x16_lo"), bp.take<float>(n * (ldx / 32), "x16_inv"), ldx, ldx / 32};
    b.h16[0] = alias_split(b.hA, n, 512, bp.take<float>(n * 16, "hA_inv"));
    b.h16[1] = alias_split(b.hB, n, 512, bp.take<float>(n * 16, "hB_inv"));
    b.aim16 = alias_split(b.aim, n, 256, bp.take<float>(n * 8, "aim_inv"));
    b.h1_16 = alias_split(b.h1, n, 128, bp.take<float>(n * 4, "h1_inv"));
    b.d16[0] = alias_split(b.dzA, n, 512, bp.take<float>(n * 16, "dzA_inv"));
    b.d16[1] = alias_split(b.dzB, n, 512, bp.take<float>(n * 16, "dzB_inv"));
    b.dz32 = bp.take<float>(n * 288, "dz32");
    b.coord_ref = bp.take<float>(n * 3, "coord_ref");
    b.wrap_off = bp.take<float>(n * 3, "wrap_off");
    b.skin_flag = bp.take<int32_t>(4, "skin_flag");
    b.mol_ref = bp.take<int32_t>(n, "mol_ref");
    b.graph_counts = bp.take<int32_t>(4, "graph_counts");
}

static void class_mark(aimnet2_engine* e, int cls, cudaStream_t st) {
    if (e->timing < 2) return;
    if (e->gemm_ev_used == (int)e->gemm_ev.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->gemm_ev.push_back(ev);
        e->gemm_ev_class.push_back(0);
    }
    e->gemm_ev_class[e->gemm_ev_used] = cls;
    cudaEventRecord(e->gemm_ev[e->gemm_ev_used++], st);
}
static void gemm_mark(aimnet2_engine* e, cudaStream_t st) { class_mark(e, 0, st); }

static int linear_fwd(aimnet2_engine* e, const Linear& L, const float* X, int ldx, int K, float* Y, float* gp, bool act,
                      int M, cudaStream_t st) {
    gemm_mark(e, st);
    int rc = gemm_nt(X, ldx, L.fwd(), L.b, Y, L.out_pad, gp, L.out_pad, M, L.out_pad, K, act ? 2 : 1, e->backend_now, st);
    gemm_mark(e, st);
    return rc;
}
// the transposed weights from input feature `row0` on (rows of Wt are input features): an input-gradient GEMM that only
// needs the columns row0 .. in_pad of dX
static WeightView rows_from(const WeightView& w, int row0) {
    const size_t o = (size_t)row0 * w.ldw;
    return WeightView{w.W ? w.W + o : nullptr, w.Whi ? w.Whi + o : nullptr, w.Wlo ? w.Wlo + o : nullptr,
                      w.Wh16 ? static_cast<const char*>(w.Wh16) + 2 * o : nullptr,
                      w.Wl16 ? static_cast<const char*>(w.Wl16) + 2 * o : nullptr, w.inv_scale16, w.ldw};
}
// dX[M,in_pad] = dZ[M,out_pad] @ W  (* gp_prev); col0 > 0: only the columns col0 .. in_pad of dX are computed
static int linear_bwd(aimnet2_engine* e, const Linear& L, const float* dZ, float* dX, int lddx, const float* gp_prev,
                      int ldgp, int M, cudaStream_t st, int col0 = 0) {
    gemm_mark(e, st);
    int rc = gemm_nt(dZ, L.out_pad, rows_from(L.bwd(), col0), nullptr, dX + col0, lddx, const_cast<float*>(gp_prev), ldgp, M,
                     L.in_pad - col0, L.out_pad, gp_prev ? 3 : 0, e->backend_now, st);
    gemm_mark(e, st);
    return rc;
}

// 3xFP16 backend: activations pre-split; out16 != nullptr writes the output pre-split for the next GEMM
static int linear_fwd16(aimnet2_engine* e, const Linear& L, const SplitMat& X, float* Y32, const SplitMat* Y16, float* gp,
                        bool act, int M, cudaStream_t st) {
    gemm_mark(e, st);
    int rc = gemm_nt_split(X, L.fwd(), L.b, Y32, L.out_pad, Y16, gp, L.out_pad, M, L.out_pad, L.in_pad, act ? 2 : 1,
                           e->backend_now - 2, st);
    gemm_mark(e, st);
    return rc;
}
static int linear_bwd16(aimnet2_engine* e, const Linear& L, const SplitMat& dZ, float* dX32, int lddx, const SplitMat* dX16,
                        const float* gp_prev, int ldgp, int M, cudaStream_t st, int col0 = 0) {
    gemm_mark(e, st);
    int rc = gemm_nt_split(dZ, rows_from(L.bwd(), col0), nullptr, dX32 ? dX32 + col0 : nullptr, lddx, dX16,
                           const_cast<float*>(gp_prev), ldgp, M, L.in_pad - col0, L.out_pad, gp_prev ? 3 : 0, e->backend_now - 2, st);
    gemm_mark(e, st);
    return rc;
}
// fp32 -> pre-split (counted with the GEMM time: it is overhead of the precision scheme)
static int presplit(aimnet2_engine* e, const float* X, int ldx, int M, int K, const SplitMat& out, cudaStream_t st) {
    gemm_mark(e, st);
    int rc = presplit_f32(X, ldx, M, K, out, st);
    gemm_mark(e, st);
    return rc;
}

static void collect_timing(aimnet2_engine* e) {
    for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&e->last_ms[k], e->ev[k], e->ev[k + 1]);
    cudaEventElapsedTime(&e->last_ms[4], e->ev[0], e->ev[4]);
    e->last_gemm_ms = e->last_conv_ms = 0.f;
    e->last_gemm_launches = e->last_conv_launches = 0;
    for (int k = 0; k + 1 < e->gemm_ev_used; k += 2) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e->gemm_ev[k], e->gemm_ev[k + 1]);
        if (e->gemm_ev_class[k] == 0) {
            e->last_gemm_ms += ms;
            e->last_gemm_launches++;
        } else {
            e->last_conv_ms += ms;
            e->last_conv_launches++;
        }
    }
}

static int build_list(aimnet2_engine* e, const float* coord, int N, float cutoff, const aimnet2_system_t* sys,
                      const int32_t* mol_idx, int sorted, int cap, int32_t* nb, int32_t* sh, int32_t* cnt, int* maxc,
                      int32_t* scratch, cudaStream_t st) {
    return neighbor_matrix_impl(coord, N, cutoff, sys->cell, sys->host_cell, sys->pbc_host, sys->n_cells, mol_idx,
                                sys->n_mol, cap, N, sorted, nb, sh, cnt, maxc, st, true, scratch, e->pinned_int);
}

static int eval_impl(aimnet2_engine* e, const aimnet2_system_t* sys, const aimnet2_result_t* res, int flags,
                     cudaStream_t st) {
    AIM_REQUIRE(e && sys && res, "engine_eval: null argument");
    const int N = sys->n_atoms, B = sys->n_mol, C = e->C;
    AIM_REQUIRE(N >= 0 && B >= 1, "engine_eval: need n_atoms >= 0 and n_mol >= 1");
    AIM_REQUIRE(sys->coord && sys->numbers && sys->charge, "engine_eval: coord, numbers and charge are required");
    AIM_REQUIRE(res->energy && res->charges, "engine_eval: energy and charges outputs are required");
    AIM_REQUIRE(B == 1 || sys->mol_idx != nullptr, "engine_eval: mol_idx required for more than one molecule");
    AIM_REQUIRE((sys->cell == nullptr) == (sys->n_cells == 0), "engine_eval: cell / n_cells mismatch");
    AIM_REQUIRE(sys->n_cells == 0 || sys->n_cells == 1 || sys->n_cells == B, "engine_eval: n_cells must be 0, 1 or n_mol");
    AIM_REQUIRE(sys->cell == nullptr || sys->host_cell != nullptr, "engine_eval: host_cell required with cell");
    const bool want_f = (flags & AIMNET_WANT_FORCES) != 0, want_s = (flags & AIMNET_WANT_STRESS) != 0;
    AIM_REQUIRE(!want_f || res->forces, "engine_eval: forces requested without an output buffer");
    AIM_REQUIRE(!want_s || (res->stress && sys->cell), "engine_eval: stress needs a cell and an output buffer");
    const aimnet2_options_t& o = e->opt;
    const bool ewald = o.coulomb_method == AIMNET_COULOMB_EWALD;
    if (ewald) {
        AIM_REQUIRE(sys->cell != nullptr && sys->n_cells == B, "engine_eval: Ewald needs a cell for every system of the batch");
        if (sys->pbc_host)
            for (int c = 0; c < 3 * sys->n_cells; ++c)
                AIM_REQUIRE(sys->pbc_host[c], "engine_eval: Ewald needs pbc on all three axes");
    }
    AIM_REQUIRE(!o.dispersion || e->d3_c6ref, "engine_eval: dispersion requested but no D3 tables were loaded");
    g_launch_count = 0;
    e->gemm_ev_used = 0;
    const bool pbc = sys->cell != nullptr;
    const bool backward = want_f || want_s;
    // small systems: the tensor-core pipelines are latency-bound below a few hundred rows, the fp32 SIMT small-M kernel wins
    const int backend_eff = (e->gemm_backend != 0 && N <= e->small_m_rows) ? 0 : e->gemm_backend;
    const bool tc16 = backend_eff >= 2;   // 3xFP16 kernels: 2 one tile stream, 3 experimental pipelined epilogue, 4 two tile streams
    e->backend_now = backend_eff;
    const int ldx = pad32(2 * kAG + kAH + C * (1 + kG + kH));
    const bool need_lr_terms = (o.coulomb_method == AIMNET_COULOMB_SIMPLE || o.coulomb_method == AIMNET_COULOMB_DSF ||
                                ewald || o.dispersion);
    const bool need_lr_list = need_lr_terms && pbc;
    AIM_REQUIRE(!(pbc && o.coulomb_method == AIMNET_COULOMB_SIMPLE),
                "engine_eval: 'simple' Coulomb is not defined for periodic systems (host switches to DSF)");
    float lr_cut = 0.f;
    if (o.coulomb_method == AIMNET_COULOMB_DSF) lr_cut = std::max(lr_cut, o.dsf_rc);
    if (o.dispersion) lr_cut = std::max(lr_cut, o.d3_cutoff);
    if (ewald) {
        // per-system splitting parameters need the atom counts of the systems: one segment-pointer read-back for a batch
        e->sys_lo.assign(B + 1, 0);
        e->sys_lo[B] = N;
        if (B > 1) {
            std::vector<int32_t> mi(N);
            AIM_CUDA_CHECK(cudaMemcpyAsync(mi.data(), sys->mol_idx, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
            AIM_CUDA_CHECK(cudaStreamSynchronize(st));
            int s0 = 0;
            for (int i = 0; i < N; ++i)
                while (s0 < mi[i] && s0 < B) e->sys_lo[++s0] = i;
            while (s0 < B) e->sys_lo[++s0] = N;
        }
        if ((int)e->ewald.size() < B) e->ewald.resize(B);
        for (int s0 = 0; s0 < B; ++s0) {
            const int ns = e->sys_lo[s0 + 1] - e->sys_lo[s0];
            AIM_REQUIRE(ns >= 1, "engine_eval: Ewald system without atoms");
            AIM_TRY(ewald_prepare(e->ewald[s0], sys->host_cell + 9 * s0, ns, o.ewald_accuracy, 15.0, st));
            lr_cut = std::max(lr_cut, (float)e->ewald[s0].rc);
        }
    }
    const bool own_sr = sys->nbmat == nullptr;
    if (e->timing) cudaEventRecord(e->ev[0], st);

    Buffers b;
    const float skin = own_sr ? std::max(0.f, o.neighbor_skin) : 0.f;
    auto skin_matches = [&]() {
        auto& k = e->skin;
        if (!k.valid || k.N != N || k.B != B || k.n_cells != sys->n_cells || k.sr_cap != e->sr_cap || k.lr_cap != e->lr_cap ||
            k.need_lr != need_lr_list || k.sr_cut != o.sr_cutoff || k.lr_cut != lr_cut || k.skin != skin || k.ws != e->ws ||
            k.has_mol != (sys->mol_idx != nullptr))
            return false;
        for (int c = 0; c < 9 * sys->n_cells; ++c)
            if (k.cell[c] != sys->host_cell[c]) return false;
        for (int c = 0; c < 3 * sys->n_cells; ++c)
            if (k.pbc[c] != (sys->pbc_host ? sys->pbc_host[c] : (uint8_t)1)) return false;
        return true;
    };
    if (e->sr_cap_next > 0) e->sr_cap = e->sr_cap_next;
    if (e->lr_cap_next > 0) e->lr_cap = e->lr_cap_next;
    e->sr_cap_next = e->lr_cap_next = 0;
    for (int attempt = 0;; ++attempt) {
        AIM_REQUIRE(attempt < 8, "engine_eval: neighbor buffers failed to converge");
        Bump probe{nullptr};
        carve(e, probe, b, N, B, e->sr_cap, e->lr_cap, pbc, need_lr_list, ldx);
        size_t need = probe.off + 1024;
        const bool capturing = e->graph.capturing;
        AIM_REQUIRE(!capturing || need <= e->ws_bytes, "engine_eval: workspace changed during graph capture");
        // the workspace follows the demand down as well: after 32 evaluations in a row that needed less than half of it
        e->ws_slack_evals = (!capturing && attempt == 0 && 2 * need < e->ws_bytes) ? e->ws_slack_evals + 1 : (attempt == 0 ? 0 : e->ws_slack_evals);
        const bool trim = e->ws_slack_evals >= 32 && !e->graph.enabled;
        if (need > e->ws_bytes || trim) {
            AIM_CUDA_CHECK(cudaStreamSynchronize(st));
            if (e->ws) AIM_CUDA_CHECK(cudaFree(e->ws));
            e->ws = nullptr;
            e->ws_bytes = 0;
            e->ws_slack_evals = 0;
            e->skin.valid = false;
            size_t want = need + need / 8;
            AIM_CUDA_CHECK(cudaMalloc((void**)&e->ws, want));
            e->ws_bytes = want;
        }
        Bump bp{e->ws};
        e->layout.clear();
        bp.map = &e->layout;
        carve(e, bp, b, N, B, e->sr_cap, e->lr_cap, pbc, need_lr_list, ldx);
        if (N == 0) break;
        if (e->poison >= 0) {
            AIM_CUDA_CHECK(cudaMemsetAsync(e->ws, e->poison, e->ws_bytes, st));
            e->skin.valid = false;
        }
        // molecule segment pointers; the largest segment lands in nb_scratch[1] and comes back with the list builder's
        // overflow read-back (no extra synchronisation)
        AIM_TRY(launch_mol_ptr(sys->mol_idx, N, B, b.mol_ptr, b.nb_scratch + 1, st));
        if (skin > 0.f && attempt == 0 && skin_matches()) {
            // lists built at cutoff + skin are still complete if no atom has moved by more than skin / 2 since the build
            AIM_TRY(launch_skin_check(N, sys->coord, b.coord_ref, 0.25f * skin * skin, sys->mol_idx, b.mol_ref, b.skin_flag, st));
            AIM_CUDA_CHECK(cudaMemcpyAsync(e->pinned_int + 8, b.skin_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            AIM_CUDA_CHECK(cudaStreamSynchronize(st));
            if (e->pinned_int[8] == 0) {
                // same lattice offsets as at build time: the stored shifts refer to those images
                if (pbc) AIM_TRY(launch_skin_apply(N, sys->coord, b.wrap_off, b.coord_w, st));
                e->skin.reuses++;
                e->dense_now = e->skin.dense;
                break;
            }
        }
        e->skin.valid = false;
        const float* coord = sys->coord;
        if (pbc) {
            AIM_TRY(wrap_positions_impl(sys->coord, b.coord_w, N, sys->cell, sys->n_cells, sys->pbc_host, sys->mol_idx, st));
            coord = b.coord_w;
        }
        bool retry = false;
        if (capturing) {
            // replayed build: no read-back inside the graph; the widest rows are kept on the device and checked after the replay
            AIM_REQUIRE(own_sr, "engine_eval: graph capture needs the engine's own lists");
            AIM_TRY(build_list(e, coord, N, o.sr_cutoff, sys, sys->mol_idx, 1, e->sr_cap, b.nb_sr, b.sh_sr, b.cnt_sr, nullptr,
                               b.nb_scratch, st));
            AIM_CUDA_CHECK(cudaMemcpyAsync(b.graph_counts, b.nb_scratch, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
            AIM_CUDA_CHECK(cudaMemsetAsync(b.graph_counts + 1, 0, sizeof(int32_t), st));
            if (need_lr_list) {
                AIM_TRY(build_list(e, coord, N, lr_cut, sys, sys->mol_idx, 0, e->lr_cap, b.nb_lr, b.sh_lr, b.cnt_lr, nullptr,
                                   b.nb_scratch, st));
                AIM_CUDA_CHECK(cudaMemcpyAsync(b.graph_counts + 1, b.nb_scratch, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
            }
            break;   // caps, widths and the dense / list choice are those of the eager evaluation that preceded the capture
        }
        if (own_sr) {
            int maxc = 0;
            int rc = build_list(e, coord, N, o.sr_cutoff + skin, sys, sys->mol_idx, 1, e->sr_cap, b.nb_sr, b.sh_sr, b.cnt_sr,
                                &maxc, b.nb_scratch, st);
            if (rc == AIMNET_NEIGHBOR_OVERFLOW) {
                e->sr_cap = round16(maxc + maxc / 4 + 1);
                retry = true;
            } else if (rc != AIMNET_OK)
                return rc;
            e->last_sr_width = std::max(1, maxc);
            e->last_max_seg = e->pinned_int[1];
            if (rc == AIMNET_OK && !e->graph.enabled) e->sr_cap_next = shrink_cap(e->sr_cap, maxc, 64);
        }
        if (!retry && need_lr_list) {
            int maxc = 0;
            int rc = build_list(e, coord, N, lr_cut + skin, sys, sys->mol_idx, 0, e->lr_cap, b.nb_lr, b.sh_lr, b.cnt_lr, &maxc,
                                b.nb_scratch, st);
            if (rc == AIMNET_NEIGHBOR_OVERFLOW) {
                e->lr_cap = round16(maxc + maxc / 8 + 1);
                retry = true;
            } else if (rc != AIMNET_OK)
                return rc;
            e->last_lr_width = std::max(1, maxc);
            if (rc == AIMNET_OK && !e->graph.enabled) e->lr_cap_next = shrink_cap(e->lr_cap, maxc, 256);
        }
        if (!retry) {
            // Dense conv walk (conv_dense.cu): the molecule's feature tables staged in shared memory, every centre walks all
            // atoms of its molecule.  Worth it when most atoms of a molecule are inside the cutoff anyway: small molecules, or
            // at least half of the molecule in the widest row.  Not with a caller-supplied matrix or a cell.
            // ... and when there are enough molecules to give every SM one (a CTA works on one molecule at a time)
            e->dense_now = e->conv_impl >= 1 && own_sr && !pbc && B >= e->dense_min_mol && e->last_max_seg >= 2 &&
                           e->last_max_seg <= conv_dense_max_atoms(C) &&
                           (e->last_max_seg <= 64 || 2 * e->last_sr_width >= e->last_max_seg);
            if (skin > 0.f) {
                AIM_TRY(launch_skin_save(N, sys->coord, pbc ? b.coord_w : nullptr, b.coord_ref, b.wrap_off, sys->mol_idx, b.mol_ref, st));
                auto& k = e->skin;
                k.N = N, k.B = B, k.n_cells = sys->n_cells, k.sr_cap = e->sr_cap, k.lr_cap = e->lr_cap;
                k.need_lr = need_lr_list, k.sr_cut = o.sr_cutoff, k.lr_cut = lr_cut, k.skin = skin;
                k.dense = e->dense_now;
                k.ws = e->ws, k.has_mol = sys->mol_idx != nullptr;
                k.cell.assign(sys->host_cell ? sys->host_cell : nullptr, sys->host_cell ? sys->host_cell + 9 * sys->n_cells : nullptr);
                k.pbc.assign((size_t)3 * sys->n_cells, (uint8_t)1);
                if (sys->pbc_host) k.pbc.assign(sys->pbc_host, sys->pbc_host + 3 * sys->n_cells);
                k.valid = true;
                k.builds++;
            }
            break;
        }
    }
    if (e->timing) cudaEventRecord(e->ev[1], st);
    // The engine's own lists refer to the wrapped copy.  A caller-supplied nbmat / shifts pair refers to the caller's
    // positions as given (the reference wraps only inside make_nbmat, which is skipped when 'nbmat' is in the data:
    // calculator.py:1071, 1521-1529), so the short-range kernels then read sys->coord.
    const float* coord_lr = pbc ? b.coord_w : sys->coord;
    const float* coord = (pbc && own_sr) ? b.coord_w : sys->coord;
    if (!own_sr || N == 0) e->dense_now = false;
    const bool dense = e->dense_now;

    NbView sr;
    if (own_sr) {
        sr = NbView{b.nb_sr, b.sh_sr, b.cnt_sr, e->sr_cap, N};
    } else {
        AIM_REQUIRE(sys->nb_width >= 1, "engine_eval: nb_width must be >= 1 with a caller-supplied nbmat");
        AIM_REQUIRE(!pbc || sys->shifts, "engine_eval: shifts required with a caller-supplied nbmat and a cell");
        sr = NbView{sys->nbmat, pbc ? sys->shifts : nullptr, nullptr, sys->nb_width, N};
        e->last_sr_width = sys->nb_width;
    }
    CellView cv{sys->cell, sys->n_cells};

    // ---------------- forward (aimnet/models/aimnet2.py:141-187) ----------------
    AIM_TRY(launch_embed(N, sys->numbers, e->afv, b.a[0], st));
    for (int p = 0; p < 3; ++p) {
        const std::vector<Linear>& L = e->mlp[p];
        const int nl = (int)L.size();
        const float* qin = (p == 0) ? nullptr : b.q[p - 1];
        class_mark(e, 1, st);
        if (dense)
            AIM_TRY(launch_conv_dense_fwd(C, N, B, e->last_max_seg, b.mol_ptr, coord, e->aev, b.a[p], qin, e->agh_a, e->agh_q, b.x,
                                          ldx, b.T_a[p], b.T_q[p], p > 0, st));
        else
            AIM_TRY(launch_conv_fwd(C, N, sr, coord, cv, sys->mol_idx, e->aev, b.a[p], qin, e->agh_a, e->agh_q, b.x, ldx,
                                    b.T_a[p], b.T_q[p], p > 0, st));
        class_mark(e, 1, st);
        if (tc16) {
            AIM_TRY(presplit(e, b.x, ldx, N, L[0].in_pad, b.x16, st));
            SplitMat in = with_ld(b.x16, ldx);
            for (int l = 0; l < nl; ++l) {
                bool last = (l == nl - 1);
                bool act = !last || p > 0;   // last_linear only for pass 0 (aimnet2.py:56,65,75)
                SplitMat out16 = with_ld(last ? b.aim16 : b.h16[l & 1], L[l].out_pad);
                bool split_out = !last || p == 2;   // y of passes 0/1 feeds the charge equilibration in fp32
                AIM_TRY(linear_fwd16(e, L[l], in, split_out ? nullptr : b.y[p], split_out ? &out16 : nullptr,
                                     act ? b.gp[p][l] : nullptr, act, N, st));
                in = out16;
            }
        } else {
            const float* in = b.x;
            int ldin = ldx;
            float* bufs[2] = {b.hA, b.hB};
            for (int l = 0; l < nl; ++l) {
                bool last = (l == nl - 1);
                bool act = !last || p > 0;   // last_linear only for pass 0 (aimnet2.py:56,65,75)
                float* out = last ? (p < 2 ? b.y[p] : b.aim) : bufs[l & 1];
                AIM_TRY(linear_fwd(e, L[l], in, ldin, L[l].in_pad, out, act ? b.gp[p][l] : nullptr, act, N, st));
                in = out;
                ldin = L[l].out_pad;
            }
        }
        if (p < 2) {
            AIM_TRY(launch_nse_fwd(C, N, B, sys->mol_idx, b.mol_ptr, sys->charge, sys->mult, b.y[p], 288, qin,
                                   b.sumq[p], b.sumf[p], b.a[p], b.a[p + 1], b.q[p], st));
        }
    }
    // energy head (aimnet/modules/core.py:114-132) + SAE (core.py:71-97)
    if (tc16) {
        AIM_TRY(linear_fwd16(e, e->head[0], b.aim16, nullptr, &b.h1_16, b.gp_h1, true, N, st));
        AIM_TRY(linear_fwd16(e, e->head[1], b.h1_16, b.h2, nullptr, b.gp_h2, true, N, st));
    } else {
        AIM_TRY(linear_fwd(e, e->head[0], b.aim, 256, 256, b.h1, b.gp_h1, true, N, st));
        AIM_TRY(linear_fwd(e, e->head[1], b.h1, 128, 128, b.h2, b.gp_h2, true, N, st));
    }
    AIM_TRY(launch_head_tail(N, b.h2, 128, b.gp_h2, e->w3, e->b3, sys->numbers, e->sae, b.e_nn, tc16 ? b.dz32 : b.dzA, st));
    AIM_TRY(launch_charges_out(C, N, b.q[1], res->charges, res->spin_charges, st));
    if (e->timing) cudaEventRecord(e->ev[2], st);

    // ---------------- pair terms on the final charges ----------------
    float* F = want_f ? res->forces : (backward ? b.forces_tmp : nullptr);
    double* vir = want_s ? b.virial_atom : nullptr;
    if (N > 0) {
        if (F) AIM_CUDA_CHECK(cudaMemsetAsync(F, 0, sizeof(float) * 3 * N, st));
        if (vir) AIM_CUDA_CHECK(cudaMemsetAsync(vir, 0, sizeof(double) * 9 * N, st));
        AIM_CUDA_CHECK(cudaMemsetAsync(b.gq, 0, sizeof(float) * N, st));
    }
    const double k = 0.5 * kHartree * kBohr;
    const float* qfin = res->charges;   // total charges (qa + qb for NSE)
    {   // embedded SRCoulomb: E -= k sum fc q_i q_j / d (lr.py:21-62, 986-1032)
        PairSource ps{sr, sys->mol_idx, b.mol_ptr, 0.f};
        CoulombParams cp{e->sr_rc, 0.f, 0.f, 0.f, 0.f, -k};
        AIM_TRY(launch_coulomb(e->sr_envelope == 0 ? PAIR_SR_EXP : PAIR_SR_COS, N, ps, coord, cv, qfin, cp, b.e_sr, b.gq,
                               backward ? F : nullptr, vir, 0, st));
    }
    PairSource lrs;
    if (need_lr_list)
        lrs = PairSource{NbView{b.nb_lr, b.sh_lr, b.cnt_lr, e->lr_cap, N}, sys->mol_idx, b.mol_ptr, 0.f};
    else
        lrs = PairSource{NbView{nullptr, nullptr, nullptr, 0, N}, sys->mol_idx, b.mol_ptr,
                         o.coulomb_method == AIMNET_COULOMB_SIMPLE ? 0.f : lr_cut * lr_cut};
    bool have_lr = false, have_d3 = false;
    if (o.coulomb_method == AIMNET_COULOMB_SIMPLE) {
        CoulombParams cp{0.f, 0.f, 0.f, 0.f, 0.f, k};
        AIM_TRY(launch_coulomb(PAIR_SIMPLE, N, lrs, coord_lr, cv, qfin, cp, b.e_lr, b.gq, backward ? F : nullptr, vir, 0, st));
        have_lr = true;
    } else if (o.coulomb_method == AIMNET_COULOMB_DSF) {
        double a = o.dsf_alpha, R = o.dsf_rc;
        double erfc_rc = std::erfc(a * R);
        CoulombParams cp;
        cp.rc = o.dsf_rc;
        cp.alpha = o.dsf_alpha;
        cp.shift_val = (float)(erfc_rc / R);
        cp.shift_slope = (float)(erfc_rc / (R * R) + 2.0 * a / std::sqrt(M_PI) * std::exp(-a * a * R * R) / R);
        cp.self_coeff = (float)(-(erfc_rc / R / 2.0 + a / std::sqrt(M_PI)));
        cp.factor = k;
        AIM_TRY(launch_coulomb(PAIR_DSF, N, lrs, coord_lr, cv, qfin, cp, b.e_lr, b.gq, backward ? F : nullptr, vir, 0, st));
        have_lr = true;
    }
    else if (ewald) {
        // every periodic system with its own alpha / r_c / k vectors (lr.py:687-696 carries batch_idx); the real-space part
        // walks the shared long-range list, the reciprocal part only sees the system's own atoms
        for (int s0 = 0; s0 < B; ++s0) {
            const EwaldPlan& pl = e->ewald[s0];
            const int lo = e->sys_lo[s0], ns = e->sys_lo[s0 + 1] - lo;
            CoulombParams cp{(float)pl.rc, (float)pl.alpha, 0.f, 0.f, 0.f, k};
            AIM_TRY(launch_coulomb(PAIR_EWALD, ns, lrs, coord_lr, cv, qfin, cp, b.e_lr, b.gq, backward ? F : nullptr, vir, 0, st, lo));
            AIM_TRY(launch_ewald_recip(pl, ns, coord_lr + 3 * (size_t)lo, qfin + lo, b.e_lr + lo, b.gq + lo,
                                       backward ? F + 3 * (size_t)lo : nullptr, vir ? vir + 9 * (size_t)lo : nullptr, st));
        }
        have_lr = true;
    }
    if (o.dispersion) {
        D3Params dp{e->d3_c6ref, e->d3_cnref, e->d3_rcov, e->d3_r4r2, o.d3_s6, o.d3_s8, o.d3_a1, o.d3_a2,
                    (float)(o.d3_cutoff * (1.0 - o.d3_smoothing) / kBohr), (float)(o.d3_cutoff / kBohr)};
        AIM_TRY(launch_d3(N, lrs, coord_lr, cv, sys->numbers, dp, b.cn, b.d3w, b.dEdCN, b.e_d3, backward ? F : nullptr, vir, st));
        have_d3 = true;
    }
    AIM_TRY(launch_energy_reduce(B, b.mol_ptr, b.e_nn, b.e_sr, have_lr ? b.e_lr : nullptr, have_d3 ? b.e_d3 : nullptr,
                                 res->energy, st));
    if (e->timing) cudaEventRecord(e->ev[3], st);

    // ---------------- analytic reverse pass ----------------
    if (backward && N > 0) {
        // head: dzA (dz32 for the pre-split backend) holds dz2 = w3 * gelu'(z2)
        float* cur = b.dzA;   // gradient w.r.t. pre-activation of the last Linear of pass 2
        float* other = b.dzB;
        int c16 = 0;          // pre-split backend: index of the d16 buffer holding that gradient
        if (tc16) {
            AIM_TRY(presplit(e, b.dz32, 128, N, 128, with_ld(b.d16[0], 128), st));
            SplitMat o1 = with_ld(b.d16[1], 128), o0 = with_ld(b.d16[0], 256);
            AIM_TRY(linear_bwd16(e, e->head[1], with_ld(b.d16[0], 128), nullptr, 0, &o1, b.gp_h1, 128, N, st));
            AIM_TRY(linear_bwd16(e, e->head[0], o1, nullptr, 0, &o0, b.gp[2][e->mlp[2].size() - 1], 256, N, st));
        } else {
            AIM_TRY(linear_bwd(e, e->head[1], b.dzA, b.dzB, 128, b.gp_h1, 128, N, st));
            AIM_TRY(linear_bwd(e, e->head[0], b.dzB, b.dzA, 256, b.gp[2][e->mlp[2].size() - 1], 256, N, st));
        }
        for (int p = 2; p >= 0; --p) {
            const std::vector<Linear>& L = e->mlp[p];
            const int nl = (int)L.size();
            for (int l = nl - 1; l >= 0; --l) {
                if (tc16) {
                    SplitMat dz = with_ld(b.d16[c16], L[l].out_pad);
                    if (l > 0) {
                        SplitMat o = with_ld(b.d16[c16 ^ 1], L[l].in_pad);
                        AIM_TRY(linear_bwd16(e, L[l], dz, nullptr, 0, &o, b.gp[p][l - 1], L[l - 1].out_pad, N, st));
                        c16 ^= 1;
                    } else {
                        // pass 0: the first 256 inputs are the atom's own embedding, nothing consumes their gradient
                        AIM_TRY(linear_bwd16(e, L[0], dz, b.dx, ldx, nullptr, nullptr, 0, N, st, p == 0 ? kAG : 0));
                    }
                } else if (l > 0) {
                    AIM_TRY(linear_bwd(e, L[l], cur, other, L[l].in_pad, b.gp[p][l - 1], L[l - 1].out_pad, N, st));
                    std::swap(cur, other);
                } else {
                    AIM_TRY(linear_bwd(e, L[0], cur, b.dx, ldx, nullptr, 0, N, st, p == 0 ? kAG : 0));
                }
            }
            const float* qin = (p == 0) ? nullptr : b.q[p - 1];
            class_mark(e, 1, st);
            if (dense && e->conv_impl == 2) {
                AIM_TRY(launch_conv_bwd_prep(C, N, b.dx, ldx, b.T_a[p], b.T_q[p], e->agh_a, e->agh_q, b.dS_a, b.dS_q, p > 0, 1, st));
                AIM_TRY(launch_conv_dense_bwd_gather(C, N, B, e->last_max_seg, b.mol_ptr, coord, e->aev, b.a[p], qin, b.dS_a, b.dS_q,
                                                     b.grad_a, b.grad_q, F, b.dense_fpart, b.dense_gqpart, p > 0, p > 0, st));
            } else if (p == 0 && e->species_pass0) {
                // the convolved features of pass 0 are the embedding: contraction tables per species instead of per pair
                // (conv.cu); the generic kernel follows with skip_if and only runs if there were too many species
                AIM_TRY(launch_conv_bwd_prep(C, N, b.dx, ldx, b.T_a[0], b.T_q[0], e->agh_a, e->agh_q, b.dS_a, b.dS_q, 0, 0, st));
                AIM_TRY(launch_species_scan(N, sys->numbers, b.sp_info, b.sp_slot, st));
                AIM_TRY(launch_conv0_bwd_species(N, sr, coord, cv, sys->mol_idx, e->aev, b.sp_info, b.sp_slot, e->afvT, b.dS_a,
                                                 b.sp_table, F, vir, st));
                if (e->n_impl_species > conv0_species_max_slots())   // only then can an evaluation have too many species
                    AIM_TRY(launch_conv_bwd(C, N, sr, coord, cv, sys->mol_idx, e->aev, b.a[0], nullptr, b.dx, ldx, b.T_a[0], b.T_q[0],
                                            e->agh_a, e->agh_q, b.dS_a, b.dS_q, b.grad_a, b.grad_q, F, vir, 0, 0, st, b.sp_info, false));
            } else {
                AIM_TRY(launch_conv_bwd(C, N, sr, coord, cv, sys->mol_idx, e->aev, b.a[p], qin, b.dx, ldx, b.T_a[p], b.T_q[p],
                                        e->agh_a, e->agh_q, b.dS_a, b.dS_q, b.grad_a, b.grad_q, F, vir, p > 0, p > 0, st));
            }
            class_mark(e, 1, st);
            if (p == 0) break;
            // dE/da_p and dE/dq_{p-1}
            if (p == 2)
                AIM_TRY(launch_accum_grads(C, N, b.dx, ldx, b.grad_a, b.grad_q, b.gq, 1, b.da_tot, 0, b.dq, st));
            else
                AIM_TRY(launch_accum_grads(C, N, b.dx, ldx, b.grad_a, b.grad_q, b.dq_base, C, b.da_tot, 1, b.dq, st));
            // NSE backward of pass p-1 -> dz of the last Linear of pass p-1
            int pp = p - 1;
            cur = tc16 ? b.dz32 : b.dzA;
            other = b.dzB;
            AIM_TRY(launch_nse_bwd(C, N, B, sys->mol_idx, b.mol_ptr, sys->charge, sys->mult, b.y[pp], 288, b.dq,
                                   b.sumq[pp], b.sumf[pp], b.s1, b.da_tot, pp > 0 ? b.gp[pp][e->mlp[pp].size() - 1] : nullptr,
                                   288, cur, 288, pp > 0 ? b.dq_base : nullptr, st));
            if (tc16) {
                AIM_TRY(presplit(e, b.dz32, 288, N, 288, with_ld(b.d16[0], 288), st));
                c16 = 0;
            }
        }
        if (want_s)
            AIM_TRY(launch_stress_reduce(b.mol_ptr, sys->n_cells, N, b.virial_atom, sys->cell, res->stress, st));
    }
    if (e->timing) cudaEventRecord(e->ev[4], st);

    // optional: hand the short-range matrix back in the reference layout (padding row appended)
    if (res->nbmat_out && own_sr && N > 0) {
        int w = res->nbmat_out_width;
        AIM_REQUIRE(w >= 1 && w <= e->sr_cap, "engine_eval: nbmat_out_width out of range");
        AIM_CUDA_CHECK(cudaMemcpy2DAsync(res->nbmat_out, sizeof(int32_t) * w, b.nb_sr, sizeof(int32_t) * e->sr_cap,
                                         sizeof(int32_t) * w, N, cudaMemcpyDeviceToDevice, st));
        if (res->shifts_out && pbc)
            AIM_CUDA_CHECK(cudaMemcpy2DAsync(res->shifts_out, sizeof(int32_t) * 3 * w, b.sh_sr,
                                             sizeof(int32_t) * 3 * e->sr_cap, sizeof(int32_t) * 3 * w, N,
                                             cudaMemcpyDeviceToDevice, st));
    }
    e->last_launches = g_launch_count;
    return AIMNET_OK;
}

}  // namespace aimnet

// ---------------------------------------------------------------------------------------------------------------
extern "C" const char* aimnet2_last_error(void) { return aimnet::g_error.c_str(); }
extern "C" int aimnet2_abi_version(void) { return AIMNET2_ABI_VERSION; }
extern "C" int aimnet2_abi_struct_sizes(int* weights, int* options, int* system, int* result) {
    if (weights) *weights = (int)sizeof(aimnet2_weights_t);
    if (options) *options = (int)sizeof(aimnet2_options_t);
    if (system) *system = (int)sizeof(aimnet2_system_t);
    if (result) *result = (int)sizeof(aimnet2_result_t);
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_create(aimnet2_engine_t** out, const aimnet2_weights_t* w, int device) {
    AIM_REQUIRE(out && w, "engine_create: null argument");
    AIM_REQUIRE(w->num_charge_channels == 1 || w->num_charge_channels == 2, "engine_create: num_charge_channels must be 1 or 2");
    AIM_CUDA_CHECK(cudaSetDevice(device));
    aimnet2_engine* e = new aimnet2_engine();
    e->device = device;
    e->C = w->num_charge_channels;
    int C = e->C, rc;
    for (int g = 0; g < kG; ++g) e->aev.shifts[g] = w->shifts_s[g];
    e->aev.eta = w->eta_s;
    e->aev.rc = w->rc_s;
    // embedding rows of unimplemented species are NaN in exported models (train/export_model.py:74-80): kept as is
    if ((rc = upload(e, &e->afv, w->afv, (size_t)64 * kAG))) return rc;
    {   // gather layout of conv.cu: index(a, g) = ((a >> 2) * 16 + g) * 4 + (a & 3)
        e->n_impl_species = 0;
        for (int z = 0; z < 64; ++z) {
            bool ok = true;
            for (int k = 0; k < kAG && ok; ++k) ok = !std::isnan(w->afv[(size_t)z * kAG + k]);
            e->n_impl_species += ok ? 1 : 0;
        }
        std::vector<float> t((size_t)64 * kAG);
        for (int z = 0; z < 64; ++z)
            for (int a = 0; a < kA; ++a)
                for (int g = 0; g < kG; ++g) t[(size_t)z * kAG + ((a >> 2) * kG + g) * 4 + (a & 3)] = w->afv[(size_t)z * kAG + a * kG + g];
        if ((rc = upload(e, &e->afvT, t.data(), t.size()))) return rc;
    }
    if ((rc = upload(e, &e->agh_a, w->agh_a, (size_t)kA * kG * kH))) return rc;
    if ((rc = upload(e, &e->agh_q, w->agh_q, (size_t)C * kG * kH))) return rc;
    if ((rc = upload(e, &e->sae, w->sae, 64))) return rc;
    for (int p = 0; p < 3; ++p) {
        int nl = w->n_layers[p];
        AIM_REQUIRE(nl >= 2 && nl <= 4, "engine_create: 2..4 Linear layers per pass are supported");
        e->mlp[p].resize(nl);
        for (int l = 0; l < nl; ++l) {
            int in = w->layer_dims[p][l], outd = w->layer_dims[p][l + 1];
            AIM_REQUIRE(in >= 1 && outd >= 1 && in <= 768 && outd <= 512, "engine_create: layer size out of range");
            if ((rc = make_linear(e, e->mlp[p][l], w->mlp_w[p][l], w->mlp_b[p][l], in, outd))) return rc;
        }
        int in0 = 2 * kAG + kAH + (p > 0 ? C * (1 + kG + kH) : 0);
        int out_last = (p < 2) ? kAG + 2 * C : 256;
        AIM_REQUIRE(w->layer_dims[p][0] == in0, "engine_create: first layer width does not match the AEV/conv layout");
        AIM_REQUIRE(w->layer_dims[p][nl] == out_last, "engine_create: last layer width mismatch");
    }
    if ((rc = make_linear(e, e->head[0], w->head_w[0], w->head_b[0], 256, 128))) return rc;
    if ((rc = make_linear(e, e->head[1], w->head_w[1], w->head_b[1], 128, 128))) return rc;
    if ((rc = upload(e, &e->w3, w->head_w[2], 128))) return rc;
    e->b3 = w->head_b[2][0];
    e->sr_rc = w->sr_rc;
    e->sr_envelope = w->sr_envelope;
    if (w->d3_c6ref) {
        {   // rows padded to kC6Row floats so that the pair kernels read them with 16-byte loads
            std::vector<float> padded((size_t)95 * 95 * kC6Row, 0.f);
            for (size_t r = 0; r < (size_t)95 * 95; ++r)
                for (int k = 0; k < 25; ++k) padded[r * kC6Row + k] = w->d3_c6ref[r * 25 + k];
            if ((rc = upload(e, &e->d3_c6ref, padded.data(), padded.size()))) return rc;
        }
        if ((rc = upload(e, &e->d3_cnref, w->d3_cnref, (size_t)95 * 5))) return rc;
        if ((rc = upload(e, &e->d3_rcov, w->d3_rcov, 95))) return rc;
        if ((rc = upload(e, &e->d3_r4r2, w->d3_r4r2, 95))) return rc;
    }
    e->opt.coulomb_method = AIMNET_COULOMB_SIMPLE;
    e->opt.dsf_alpha = 0.2f;
    e->opt.dsf_rc = 15.0f;
    e->opt.ewald_accuracy = 1e-6f;
    e->opt.dispersion = 0;
    e->opt.d3_s6 = 1.0f;
    e->opt.d3_s8 = 0.3908f;
    e->opt.d3_a1 = 0.566f;
    e->opt.d3_a2 = 3.128f;
    e->opt.d3_cutoff = 15.0f;
    e->opt.d3_smoothing = 0.2f;
    e->opt.sr_cutoff = 5.0f;
    e->gemm_backend = gemm_tc_available() ? 2 : 0;
    AIM_CUDA_CHECK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    AIM_CUDA_CHECK(cudaMallocHost((void**)&e->pinned_int, 64));
    for (int k = 0; k < 6; ++k) AIM_CUDA_CHECK(cudaEventCreate(&e->ev[k]));
    // the uploads above are cudaMemcpy calls from pageable memory: they return once the data is staged, the DMA may
    // still be in flight, and own_stream / a caller's non-blocking stream is not ordered after the legacy stream
    AIM_CUDA_CHECK(cudaDeviceSynchronize());
    *out = e;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_destroy(aimnet2_engine_t* e) {
    if (!e) return AIMNET_OK;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (void* p : e->owned) cudaFree(p);
    if (e->ws) cudaFree(e->ws);
    if (e->stage) cudaFree(e->stage);
    if (e->graph.exec) cudaGraphExecDestroy(e->graph.exec);
    for (EwaldPlan& pl : e->ewald) ewald_release(pl);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->pinned_int) cudaFreeHost(e->pinned_int);
    for (int k = 0; k < 6; ++k)
        if (e->ev[k]) cudaEventDestroy(e->ev[k]);
    for (cudaEvent_t ev : e->gemm_ev) cudaEventDestroy(ev);
    delete e;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_set_options(aimnet2_engine_t* e, const aimnet2_options_t* opt) {
    AIM_REQUIRE(e && opt, "set_options: null argument");
    AIM_REQUIRE(opt->coulomb_method >= 0 && opt->coulomb_method <= 3, "set_options: bad coulomb_method");
    AIM_REQUIRE(opt->sr_cutoff > 0.f, "set_options: sr_cutoff must be positive");
    e->opt = *opt;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_set_gemm_backend(aimnet2_engine_t* e, int backend) {
    AIM_REQUIRE(e, "set_gemm_backend: null engine");
    AIM_REQUIRE(backend >= 0 && backend <= 5,
                "set_gemm_backend: backend must be 0 (SIMT), 1 (3xTF32), 2 (3xFP16), 3 (3xFP16 with the experimental pipelined "
                "epilogue), 4 (3xFP16, two tile streams per SM) or 5 (two tile streams on CTA pairs)");
    AIM_REQUIRE(backend == 0 || gemm_tc_available(), "set_gemm_backend: tcgen05 backends not available in this build");
    e->gemm_backend = backend;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_set_species_first_pass(aimnet2_engine_t* e, int on) {
    AIM_REQUIRE(e, "set_species_first_pass: null engine");
    e->species_pass0 = on ? 1 : 0;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_set_conv_impl(aimnet2_engine_t* e, int impl) {
    AIM_REQUIRE(e, "set_conv_impl: null engine");
    AIM_REQUIRE(impl >= 0 && impl <= 2, "set_conv_impl: 0 = list kernels always, 1 = dense shared-memory forward for batches of small molecules (default), 2 = dense forward and backward");
    e->conv_impl = impl;
    e->skin.valid = false;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_set_dense_min_molecules(aimnet2_engine_t* e, int n_mol) {
    AIM_REQUIRE(e, "set_dense_min_molecules: null engine");
    AIM_REQUIRE(n_mol >= 1, "set_dense_min_molecules: need at least one molecule");
    e->dense_min_mol = n_mol;
    e->skin.valid = false;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_set_small_m_rows(aimnet2_engine_t* e, int rows) {
    AIM_REQUIRE(e, "set_small_m_rows: null engine");
    AIM_REQUIRE(rows >= 0 && rows <= kSmallM, "set_small_m_rows: rows must be in [0, 512]");
    e->small_m_rows = rows;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_debug_poison(aimnet2_engine_t* e, int byte) {
    AIM_REQUIRE(e, "debug_poison: null engine");
    AIM_REQUIRE(byte >= -1 && byte <= 255, "debug_poison: byte must be -1 (off) or 0..255");
    e->poison = byte;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_debug_layout(const aimnet2_engine_t* e, char* buf, int buf_bytes) {
    AIM_REQUIRE(e && buf && buf_bytes > 0, "debug_layout: null argument");
    std::string s;
    for (const Region& r : e->layout) s += std::string(r.name) + " " + std::to_string(r.off) + " " + std::to_string(r.bytes) + "\n";
    AIM_REQUIRE((int)s.size() < buf_bytes, "debug_layout: buffer too small");
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_debug_read_workspace(aimnet2_engine_t* e, void* host_dst, int64_t offset, int64_t bytes) {
    AIM_REQUIRE(e && host_dst && offset >= 0 && bytes >= 0, "debug_read_workspace: bad argument");
    AIM_REQUIRE((size_t)(offset + bytes) <= e->ws_bytes, "debug_read_workspace: range outside the workspace");
    AIM_CUDA_CHECK(cudaSetDevice(e->device));
    AIM_CUDA_CHECK(cudaDeviceSynchronize());
    AIM_CUDA_CHECK(cudaMemcpy(host_dst, e->ws + offset, (size_t)bytes, cudaMemcpyDeviceToHost));
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_set_deterministic(aimnet2_engine_t* e, int on) {
    AIM_REQUIRE(e, "set_deterministic: null engine");
    e->deterministic = on != 0;   // per engine; every kernel uses fixed chunking and no atomics, so both settings agree
    return AIMNET_OK;
}

namespace aimnet {

// ---- CUDA-graph replay ------------------------------------------------------------------------------------------
// What a captured graph depends on besides the kernels' code: sizes, request flags, options, kernel selection, the
// neighbor capacities and every device pointer baked into the kernel nodes (the staging buffers and the workspace).
static std::vector<unsigned char> graph_signature(const aimnet2_engine* e, const aimnet2_system_t* ds, const aimnet2_result_t* dr,
                                                  int flags, cudaStream_t st) {
    std::vector<unsigned char> sig;
    auto put = [&sig](const void* p, size_t n) { sig.insert(sig.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
    const int ints[] = {ds->n_atoms, ds->n_mol, ds->n_cells, flags, e->sr_cap, e->lr_cap, e->gemm_backend, e->conv_impl,
                        e->dense_min_mol, e->small_m_rows, e->dense_now ? 1 : 0, e->last_max_seg};
    put(ints, sizeof(ints));
    const void* ptrs[] = {ds->coord, ds->numbers, ds->mol_idx, ds->charge, ds->mult, ds->cell, dr->energy, dr->charges,
                          dr->spin_charges, dr->forces, dr->stress, e->ws, (const void*)st};
    put(ptrs, sizeof(ptrs));
    put(&e->opt, sizeof(e->opt));
    return sig;
}

static bool graph_eligible(const aimnet2_engine* e, const aimnet2_system_t* ds, const aimnet2_result_t* dr) {
    const aimnet2_options_t& o = e->opt;
    return e->graph.enabled && e->timing == 0 && e->poison < 0 && ds->n_atoms > 0 && ds->nbmat == nullptr &&
           dr->nbmat_out == nullptr && o.coulomb_method != AIMNET_COULOMB_EWALD && !(o.neighbor_skin > 0.f) &&
           ds->pbc_host == nullptr && !(ds->cell != nullptr && ds->n_atoms >= 512);   // the cell-list builder allocates and sorts: not captured
}

// ds / dr point into the engine's staging buffers (stable addresses from call to call)
static int eval_staged(aimnet2_engine* e, const aimnet2_system_t* ds, const aimnet2_result_t* dr, int flags, cudaStream_t st) {
    if (!graph_eligible(e, ds, dr)) return eval_impl(e, ds, dr, flags, st);
    auto& g = e->graph;
    if (g.valid && g.sig == graph_signature(e, ds, dr, flags, st)) {
        AIM_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
        // the replayed neighbor build cannot grow its buffers: check the widest rows afterwards
        Buffers b;
        Bump bp{e->ws};
        const int ldx = pad32(2 * kAG + kAH + e->C * (1 + kG + kH));
        const bool need_lr = ds->cell != nullptr && (e->opt.coulomb_method != AIMNET_COULOMB_NONE || e->opt.dispersion);
        carve(e, bp, b, ds->n_atoms, ds->n_mol, e->sr_cap, e->lr_cap, ds->cell != nullptr, need_lr, ldx);
        AIM_CUDA_CHECK(cudaMemcpyAsync(e->pinned_int + 4, b.graph_counts, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        AIM_CUDA_CHECK(cudaStreamSynchronize(st));
        if (e->pinned_int[4] <= e->sr_cap && e->pinned_int[5] <= e->lr_cap) {
            g.launches++;
            e->last_launches = 1;
            return AIMNET_OK;
        }
        g.valid = false;   // a row outgrew its buffer: redo this step eagerly (grows the buffers), capture again later
        g.fallbacks++;
    }
    int rc = eval_impl(e, ds, dr, flags, st);
    if (rc != AIMNET_OK) return rc;
    std::vector<unsigned char> sig = graph_signature(e, ds, dr, flags, st);
    if (!(g.pending && g.pending_sig == sig)) {   // first sighting of this shape: remember it, capture when it comes back
        g.pending = true;
        g.pending_sig = sig;
        return AIMNET_OK;
    }
    // second evaluation with the same signature: record the step (nothing executes during capture; this call's results are
    // those of the eager evaluation above)
    cudaGraph_t graph = nullptr;
    AIM_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    g.capturing = true;
    const int saved_launches = e->last_launches;
    rc = eval_impl(e, ds, dr, flags, st);
    g.capturing = false;
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    e->last_launches = saved_launches;
    if (rc != AIMNET_OK || ce != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        g.pending = false;
        return AIMNET_OK;   // the eager results stand; stay on the eager path
    }
    if (g.exec) {
        cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
    }
    ce = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        g.exec = nullptr;
        g.valid = g.pending = false;
        return AIMNET_OK;
    }
    g.sig = sig;
    g.valid = true;
    g.pending = false;
    g.captures++;
    return AIMNET_OK;
}

struct Staged {
    aimnet2_system_t ds;
    aimnet2_result_t dr;
};

// engine-owned device copies of every input / output array (grow-only), carved in a fixed order
static int stage_buffers(aimnet2_engine* e, const aimnet2_system_t* sys, const aimnet2_result_t* res, Staged& s) {
    const int N = sys->n_atoms, B = sys->n_mol, nc = sys->n_cells;
    const size_t n = (size_t)std::max(N, 1);
    auto carve_stage = [&](Bump& bp) {
        s.ds = *sys;
        s.dr = *res;
        s.ds.coord = bp.take<float>(n * 3);
        s.ds.numbers = bp.take<int32_t>(n);
        s.ds.mol_idx = sys->mol_idx ? bp.take<int32_t>(n) : nullptr;
        s.ds.charge = bp.take<float>(B);
        s.ds.mult = sys->mult ? bp.take<float>(B) : nullptr;
        s.ds.cell = sys->cell ? bp.take<float>((size_t)9 * nc) : nullptr;
        s.dr.energy = bp.take<double>(B);
        s.dr.charges = bp.take<float>(n);
        s.dr.spin_charges = res->spin_charges ? bp.take<float>(n) : nullptr;
        s.dr.forces = res->forces ? bp.take<float>(n * 3) : nullptr;
        s.dr.stress = res->stress ? bp.take<float>((size_t)9 * std::max(nc, 1)) : nullptr;
        s.dr.nbmat_out = nullptr;
        s.dr.shifts_out = nullptr;
    };
    Bump probe{nullptr};
    carve_stage(probe);
    if (probe.off + 1024 > e->stage_bytes) {
        AIM_CUDA_CHECK(cudaDeviceSynchronize());
        if (e->stage) AIM_CUDA_CHECK(cudaFree(e->stage));
        e->stage = nullptr;
        AIM_CUDA_CHECK(cudaMalloc((void**)&e->stage, probe.off + 1024));
        e->stage_bytes = probe.off + 1024;
        e->graph.valid = e->graph.pending = false;
    }
    Bump bp{e->stage};
    carve_stage(bp);
    return AIMNET_OK;
}

static int stage_copies(const aimnet2_system_t* sys, const aimnet2_result_t* res, const Staged& s, int flags, bool in,
                        cudaMemcpyKind kind, cudaStream_t st) {
    const int N = sys->n_atoms, B = sys->n_mol, nc = sys->n_cells;
#define CP(dst, src, bytes) AIM_CUDA_CHECK(cudaMemcpyAsync((void*)(dst), (src), (bytes), kind, st))
    if (in) {
        if (N > 0) {
            CP(s.ds.coord, sys->coord, sizeof(float) * 3 * N);
            CP(s.ds.numbers, sys->numbers, sizeof(int32_t) * N);
            if (sys->mol_idx) CP(s.ds.mol_idx, sys->mol_idx, sizeof(int32_t) * N);
        }
        CP(s.ds.charge, sys->charge, sizeof(float) * B);
        if (sys->mult) CP(s.ds.mult, sys->mult, sizeof(float) * B);
        if (sys->cell) CP(s.ds.cell, sys->cell, sizeof(float) * 9 * nc);
    } else {
        CP(res->energy, s.dr.energy, sizeof(double) * B);
        if (N > 0) {
            CP(res->charges, s.dr.charges, sizeof(float) * N);
            if (res->spin_charges) CP(res->spin_charges, s.dr.spin_charges, sizeof(float) * N);
            if (res->forces && (flags & AIMNET_WANT_FORCES)) CP(res->forces, s.dr.forces, sizeof(float) * 3 * N);
        }
        if (res->stress && (flags & AIMNET_WANT_STRESS)) CP(res->stress, s.dr.stress, sizeof(float) * 9 * nc);
    }
#undef CP
    return AIMNET_OK;
}

}  // namespace aimnet

extern "C" int aimnet2_engine_eval(aimnet2_engine_t* e, const aimnet2_system_t* sys, const aimnet2_result_t* res, int flags,
                                   void* stream) {
    AIM_REQUIRE(e && sys && res, "engine_eval: null argument");
    AIM_CUDA_CHECK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (e->graph.enabled && graph_eligible(e, sys, res) && sys->n_mol >= 1) {
        // Graph replay needs stable addresses and a capturable stream: inputs / outputs go through the engine's staging
        // buffers (device-to-device) and the step runs on the engine's own stream (the caller's may be the legacy default
        // stream, which cannot be captured), ordered after / before the caller's stream with two events.
        cudaStream_t own = e->own_stream;
        AIM_CUDA_CHECK(cudaEventRecord(e->ev[5], st));
        AIM_CUDA_CHECK(cudaStreamWaitEvent(own, e->ev[5], 0));
        Staged s;
        AIM_TRY(stage_buffers(e, sys, res, s));
        s.ds.host_cell = sys->host_cell;
        AIM_TRY(stage_copies(sys, res, s, flags, true, cudaMemcpyDeviceToDevice, own));
        rc = eval_staged(e, &s.ds, &s.dr, flags, own);
        if (rc == AIMNET_OK) rc = stage_copies(sys, res, s, flags, false, cudaMemcpyDeviceToDevice, own);
        AIM_CUDA_CHECK(cudaEventRecord(e->ev[5], own));
        AIM_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev[5], 0));
    } else {
        rc = eval_impl(e, sys, res, flags, st);
    }
    if (rc == AIMNET_OK && e->timing) {
        AIM_CUDA_CHECK(cudaStreamSynchronize(st));
        collect_timing(e);
    }
    return rc;
}

extern "C" int aimnet2_engine_eval_host(aimnet2_engine_t* e, const aimnet2_system_t* sys, const aimnet2_result_t* res,
                                        int flags) {
    AIM_REQUIRE(e && sys && res, "engine_eval_host: null argument");
    AIM_CUDA_CHECK(cudaSetDevice(e->device));
    cudaStream_t st = e->own_stream;
    AIM_REQUIRE(sys->n_atoms >= 0 && sys->n_mol >= 1, "engine_eval_host: bad sizes");
    AIM_REQUIRE(sys->nbmat == nullptr, "engine_eval_host: caller-supplied nbmat is only supported by engine_eval");
    Staged s;
    AIM_TRY(stage_buffers(e, sys, res, s));
    s.ds.host_cell = sys->cell;   // the caller's cell IS host memory here
    AIM_TRY(stage_copies(sys, res, s, flags, true, cudaMemcpyHostToDevice, st));
    int rc = eval_staged(e, &s.ds, &s.dr, flags, st);
    if (rc != AIMNET_OK) return rc;
    AIM_TRY(stage_copies(sys, res, s, flags, false, cudaMemcpyDeviceToHost, st));
    AIM_CUDA_CHECK(cudaStreamSynchronize(st));
    if (e->timing) collect_timing(e);
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_enable_cuda_graph(aimnet2_engine_t* e, int on) {
    AIM_REQUIRE(e, "enable_cuda_graph: null engine");
    e->graph.enabled = on != 0;
    e->graph.valid = e->graph.pending = false;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_graph_stats(const aimnet2_engine_t* e, int* captures, int* launches, int* fallbacks) {
    AIM_REQUIRE(e, "graph_stats: null engine");
    if (captures) *captures = e->graph.captures;
    if (launches) *launches = e->graph.launches;
    if (fallbacks) *fallbacks = e->graph.fallbacks;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_last_launches(const aimnet2_engine_t* e) { return e ? e->last_launches : 0; }

extern "C" int aimnet2_engine_info(const aimnet2_engine_t* e, int* sr_width, int* lr_width, int64_t* workspace_bytes) {
    AIM_REQUIRE(e, "engine_info: null engine");
    if (sr_width) *sr_width = e->last_sr_width;
    if (lr_width) *lr_width = e->last_lr_width;
    if (workspace_bytes) *workspace_bytes = (int64_t)e->ws_bytes;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_neighbor_caps(const aimnet2_engine_t* e, int* sr_cap, int* lr_cap) {
    AIM_REQUIRE(e, "neighbor_caps: null engine");
    if (sr_cap) *sr_cap = e->sr_cap;
    if (lr_cap) *lr_cap = e->lr_cap;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_conv_mode(const aimnet2_engine_t* e, int* impl, int* dense_last, int* max_molecule_last) {
    AIM_REQUIRE(e, "conv_mode: null engine");
    if (impl) *impl = e->conv_impl;
    if (dense_last) *dense_last = e->dense_now ? 1 : 0;
    if (max_molecule_last) *max_molecule_last = e->last_max_seg;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_skin_stats(const aimnet2_engine_t* e, int* builds, int* reuses) {
    AIM_REQUIRE(e, "skin_stats: null engine");
    if (builds) *builds = e->skin.builds;
    if (reuses) *reuses = e->skin.reuses;
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_enable_timing(aimnet2_engine_t* e, int on) {
    AIM_REQUIRE(e, "enable_timing: null engine");
    e->timing = on < 0 ? 0 : (on > 2 ? 2 : on);   // 1: phase events, 2: + one event pair per GEMM launch
    return AIMNET_OK;
}

extern "C" int aimnet2_engine_last_timing(const aimnet2_engine_t* e, float* ms, int n) {
    AIM_REQUIRE(e && ms, "last_timing: null argument");
    int k = 0;
    for (; k < n && k < 5; ++k) ms[k] = e->last_ms[k];
    if (k < n) ms[k++] = e->last_gemm_ms;                    // summed GEMM device time (timing level 2)
    if (k < n) ms[k++] = (float)e->last_gemm_launches;
    if (k < n) ms[k++] = e->last_conv_ms;                    // summed AEV / conv_sv device time (timing level 2)
    if (k < n) ms[k++] = (float)e->last_conv_launches;       // event pairs (one per conv_fwd / conv_bwd call)
    return k;
}
