// Operator seams for the external pair terms (SURVEY.md §8b B3 iii): the kernels of lr.cu / ewald.cu behind C entry
// points shaped like the third-party calls the reference makes (aimnet/modules/lr.py:526-540 dsf_coulomb,
// :1204-1228 dftd3; aimnet/calculators/calculator.py:1566-1587 estimate_ewald_parameters).  Thin host code: scratch from
// the stream-ordered allocator, the engine's own launchers, per-system reductions.
#include <vector>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

// out[s] = sign * sum_{i in system s} virial_atom[i]   (9 components, fp64)
__global__ void __launch_bounds__(256) virial_reduce_kernel(const int32_t* __restrict__ mol_ptr,
                                                            const double* __restrict__ virial_atom, double sign,
                                                            double* __restrict__ out) {
    __shared__ double red[8];
    const int s = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p0 = mol_ptr[s], p1 = mol_ptr[s + 1];
    for (int k = 0; k < 9; ++k) {
        double acc = 0.0;
        for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) acc += virial_atom[(size_t)i * 9 + k];
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[w] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int j = 0; j < 8; ++j) t += red[j];
            out[9 * s + k] = sign * t;
        }
    }
}

namespace {

size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

// scratch shared by the pair-term seams: per-atom energies, charge / CN gradients, per-atom virials, segment pointers
struct SeamScratch {
    char* base = nullptr;
    double* e_atom = nullptr;
    float* f0 = nullptr;       // gq (Coulomb) / cn (D3)
    float* f1 = nullptr;       // D3: dEdCN
    float* wtab = nullptr;     // D3: (n, 16) weight table
    double* virial_atom = nullptr;
    int32_t* mol_ptr = nullptr;
    float* forces = nullptr;   // used when the caller wants a virial but no forces
    cudaStream_t st = nullptr;
    ~SeamScratch() {           // stream-ordered: runs after everything enqueued on st before the seam returns
        if (base) cudaFreeAsync(base, st);
    }
};

int seam_alloc(SeamScratch& s, int n, int n_systems, bool d3, bool virial, bool own_forces, cudaStream_t st) {
    const size_t nn = (size_t)(n > 0 ? n : 1);
    size_t off = 0;
    const size_t o_e = off;      off += up256(nn * sizeof(double));
    const size_t o_f0 = off;     off += up256(nn * sizeof(float));
    const size_t o_f1 = off;     off += d3 ? up256(nn * sizeof(float)) : 0;
    const size_t o_w = off;      off += d3 ? up256(nn * 16 * sizeof(float)) : 0;
    const size_t o_v = off;      off += virial ? up256(nn * 9 * sizeof(double)) : 0;
    const size_t o_p = off;      off += up256(((size_t)n_systems + 2) * sizeof(int32_t));
    const size_t o_fr = off;     off += own_forces ? up256(nn * 3 * sizeof(float)) : 0;
    AIM_CUDA_CHECK(cudaMallocAsync((void**)&s.base, off, st));
    s.st = st;
    s.e_atom = reinterpret_cast<double*>(s.base + o_e);
    s.f0 = reinterpret_cast<float*>(s.base + o_f0);
    s.f1 = d3 ? reinterpret_cast<float*>(s.base + o_f1) : nullptr;
    s.wtab = d3 ? reinterpret_cast<float*>(s.base + o_w) : nullptr;
    s.virial_atom = virial ? reinterpret_cast<double*>(s.base + o_v) : nullptr;
    s.mol_ptr = reinterpret_cast<int32_t*>(s.base + o_p);
    s.forces = own_forces ? reinterpret_cast<float*>(s.base + o_fr) : nullptr;
    return AIMNET_OK;
}

int check_lists(const char* who, int n_atoms, const float* cell, int n_cells, int n_systems, const int32_t* nbmat,
                const int32_t* shifts, int nb_width) {
    (void)who;
    AIM_REQUIRE(n_atoms >= 0 && n_systems >= 1, "pair-term seam: bad sizes");
    AIM_REQUIRE(nbmat != nullptr && nb_width >= 1, "pair-term seam: a neighbor matrix is required");
    AIM_REQUIRE((cell == nullptr) == (n_cells == 0), "pair-term seam: cell / n_cells mismatch");
    AIM_REQUIRE(n_cells == 0 || n_cells == 1 || n_cells == n_systems, "pair-term seam: n_cells must be 0, 1 or n_systems");
    AIM_REQUIRE(cell == nullptr || shifts != nullptr, "pair-term seam: shifts are required with a cell");
    return AIMNET_OK;
}

// everything after the pair kernels: per-system energy, optional per-system virial (W = -dE/d eps, the convention of
// aimnet/calculators/derivatives.py:128-131)
int seam_finish(SeamScratch& s, int n_systems, double* energy, double* virial, cudaStream_t st) {
    AIM_TRY(launch_energy_reduce(n_systems, s.mol_ptr, s.e_atom, nullptr, nullptr, nullptr, energy, st));
    if (virial) {
        virial_reduce_kernel<<<n_systems, 256, 0, st>>>(s.mol_ptr, s.virial_atom, -1.0, virial);
        AIM_LAUNCH_CHECK();
    }
    return AIMNET_OK;
}

}  // namespace
}  // namespace aimnet

using namespace aimnet;

extern "C" int aimnet2_dsf_coulomb(const float* positions, const float* charges, int n_atoms, float cutoff, float alpha,
                                   const float* cell, int n_cells, const int32_t* batch_idx, int n_systems,
                                   const int32_t* nbmat, const int32_t* shifts, int nb_width, int fill_value,
                                   double* energy, float* forces, float* charge_grad, double* virial, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AIM_TRY(check_lists("dsf_coulomb", n_atoms, cell, n_cells, n_systems, nbmat, shifts, nb_width));
    AIM_REQUIRE(positions && charges && energy, "dsf_coulomb: null argument");
    AIM_REQUIRE(cutoff > 0.f && alpha >= 0.f, "dsf_coulomb: cutoff must be positive and alpha non-negative");
    SeamScratch s;
    const bool own_forces = virial != nullptr && forces == nullptr;   // the pair kernel writes virials next to forces
    AIM_TRY(seam_alloc(s, n_atoms, n_systems, false, virial != nullptr, own_forces, st));
    AIM_TRY(launch_mol_ptr(batch_idx, n_atoms, n_systems, s.mol_ptr, nullptr, st));
    float* F = forces ? forces : s.forces;
    float* gq = charge_grad ? charge_grad : s.f0;
    if (n_atoms > 0) {
        AIM_CUDA_CHECK(cudaMemsetAsync(gq, 0, sizeof(float) * n_atoms, st));
        if (F) AIM_CUDA_CHECK(cudaMemsetAsync(F, 0, sizeof(float) * 3 * n_atoms, st));
        if (virial) AIM_CUDA_CHECK(cudaMemsetAsync(s.virial_atom, 0, sizeof(double) * 9 * n_atoms, st));
    }
    // DSF constants exactly as the engine forms them (engine.cu, aimnet/modules/lr.py:594-606); factor 1/2: energies
    // in e^2/A, the caller multiplies by Hartree*Bohr (lr.py:542-543)
    const double a = alpha, R = cutoff;
    const double erfc_rc = std::erfc(a * R);
    CoulombParams cp;
    cp.rc = cutoff;
    cp.alpha = alpha;
    cp.shift_val = (float)(erfc_rc / R);
    cp.shift_slope = (float)(erfc_rc / (R * R) + 2.0 * a / std::sqrt(M_PI) * std::exp(-a * a * R * R) / R);
    cp.self_coeff = (float)(-(erfc_rc / R / 2.0 + a / std::sqrt(M_PI)));
    cp.factor = 0.5;
    PairSource ps{NbView{nbmat, shifts, nullptr, nb_width, fill_value}, batch_idx, s.mol_ptr, 0.f};
    CellView cv{cell, n_cells};
    AIM_TRY(launch_coulomb(PAIR_DSF, n_atoms, ps, positions, cv, charges, cp, s.e_atom, gq, F, s.virial_atom, 0, st));
    return seam_finish(s, n_systems, energy, virial, st);
}

extern "C" int aimnet2_dftd3(const float* positions, const int32_t* numbers, int n_atoms, float s6, float s8, float a1,
                             float a2, float r_on_bohr, float r_off_bohr, const float* c6ref, const float* cnref,
                             const float* rcov, const float* r4r2, const float* cell, int n_cells,
                             const int32_t* batch_idx, int n_systems, const int32_t* nbmat, const int32_t* shifts,
                             int nb_width, int fill_value, double* energy, float* forces, float* coord_num,
                             double* virial, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AIM_TRY(check_lists("dftd3", n_atoms, cell, n_cells, n_systems, nbmat, shifts, nb_width));
    AIM_REQUIRE(positions && numbers && energy && c6ref && cnref && rcov && r4r2, "dftd3: null argument");
    AIM_REQUIRE(r_off_bohr > r_on_bohr && r_on_bohr >= 0.f, "dftd3: need 0 <= r_on < r_off");
    AIM_REQUIRE(((uintptr_t)c6ref & 15) == 0, "dftd3: c6ref must be 16-byte aligned ((95,95,28) padded rows)");
    SeamScratch s;
    const bool own_forces = virial != nullptr && forces == nullptr;
    AIM_TRY(seam_alloc(s, n_atoms, n_systems, true, virial != nullptr, own_forces, st));
    AIM_TRY(launch_mol_ptr(batch_idx, n_atoms, n_systems, s.mol_ptr, nullptr, st));
    float* F = forces ? forces : s.forces;
    float* cn = coord_num ? coord_num : s.f0;
    if (n_atoms > 0) {
        if (F) AIM_CUDA_CHECK(cudaMemsetAsync(F, 0, sizeof(float) * 3 * n_atoms, st));
        if (virial) AIM_CUDA_CHECK(cudaMemsetAsync(s.virial_atom, 0, sizeof(double) * 9 * n_atoms, st));
    }
    D3Params dp{c6ref, cnref, rcov, r4r2, s6, s8, a1, a2, r_on_bohr, r_off_bohr};
    PairSource ps{NbView{nbmat, shifts, nullptr, nb_width, fill_value}, batch_idx, s.mol_ptr, 0.f};
    CellView cv{cell, n_cells};
    AIM_TRY(launch_d3(n_atoms, ps, positions, cv, numbers, dp, cn, s.wtab, s.f1, s.e_atom, F, s.virial_atom, st));
    return seam_finish(s, n_systems, energy, virial, st);
}

// Ewald summation; replaces nvalchemiops...ewald_summation as called at aimnet/modules/lr.py:687-696 (energy-only call
// there: forces / stress / charge response come from autograd through the returned energies; here they are explicit
// optional outputs like in the DSF seam).  Per-system splitting parameters from `accuracy` (calculator.py:663-666), real
// space over the caller's neighbor matrix (which must reach every system's real-space cutoff), reciprocal space, self and
// neutralising-background terms per system.  Units e^2/A: the caller multiplies by Hartree*Bohr (lr.py:697).
//   host_cell            the cells in host memory (n_systems x 9 floats): k vectors are enumerated on the host
//   host_system_offsets  n_systems + 1 ints in HOST memory: atoms [off[s], off[s+1]) form system s (NULL when n_systems == 1)
//   energies_per_atom    (n_atoms) f64 device out.  The real-space and self terms are per atom; the reciprocal-space
//                        energy of a system is booked on its first atom (only per-system sums are defined:
//                        lr.py:698-703 scatter-adds them)
// Reciprocal-space plans (k vectors, structure factors) are cached per system slot in process-wide storage: one caller
// thread at a time, the reference's own threading contract.
static std::vector<EwaldPlan> g_seam_plans;

extern "C" int aimnet2_ewald_summation(const float* positions, const float* charges, int n_atoms, const float* cell,
                                       const float* host_cell, const int32_t* batch_idx, const int32_t* host_system_offsets,
                                       int n_systems, const int32_t* nbmat, const int32_t* shifts, int nb_width, int fill_value,
                                       double accuracy, double* energies_per_atom, float* forces, float* charge_grad,
                                       double* virial, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    AIM_TRY(check_lists("ewald_summation", n_atoms, cell, n_systems, n_systems, nbmat, shifts, nb_width));
    AIM_REQUIRE(positions && charges && energies_per_atom && cell && host_cell, "ewald_summation: null argument (a cell per system is required)");
    AIM_REQUIRE(n_systems == 1 || (host_system_offsets != nullptr && batch_idx != nullptr), "ewald_summation: batches need batch_idx and host_system_offsets");
    AIM_REQUIRE(accuracy > 0.0 && accuracy < 1.0, "ewald_summation: accuracy must be in (0, 1)");
    SeamScratch s;
    const bool own_forces = virial != nullptr && forces == nullptr;
    AIM_TRY(seam_alloc(s, n_atoms, n_systems, false, virial != nullptr, own_forces, st));
    AIM_TRY(launch_mol_ptr(batch_idx, n_atoms, n_systems, s.mol_ptr, nullptr, st));
    float* F = forces ? forces : s.forces;
    float* gq = charge_grad ? charge_grad : s.f0;
    if (n_atoms > 0) {
        AIM_CUDA_CHECK(cudaMemsetAsync(gq, 0, sizeof(float) * n_atoms, st));
        if (F) AIM_CUDA_CHECK(cudaMemsetAsync(F, 0, sizeof(float) * 3 * n_atoms, st));
        if (virial) AIM_CUDA_CHECK(cudaMemsetAsync(s.virial_atom, 0, sizeof(double) * 9 * n_atoms, st));
    }
    if ((int)g_seam_plans.size() < n_systems) g_seam_plans.resize(n_systems);
    PairSource ps{NbView{nbmat, shifts, nullptr, nb_width, fill_value}, batch_idx, s.mol_ptr, 0.f};
    CellView cv{cell, n_systems};
    for (int k = 0; k < n_systems; ++k) {
        const int lo = host_system_offsets ? host_system_offsets[k] : 0;
        const int ns = (host_system_offsets ? host_system_offsets[k + 1] : n_atoms) - lo;
        AIM_REQUIRE(lo >= 0 && ns >= 0 && lo + ns <= n_atoms, "ewald_summation: bad system offsets");
        if (ns == 0) continue;
        EwaldPlan& pl = g_seam_plans[k];
        AIM_TRY(ewald_prepare(pl, host_cell + 9 * k, ns, accuracy, 0.0, st));
        CoulombParams cp{(float)pl.rc, (float)pl.alpha, 0.f, 0.f, 0.f, 0.5};
        AIM_TRY(launch_coulomb(PAIR_EWALD, ns, ps, positions, cv, charges, cp, s.e_atom, gq, F, s.virial_atom, 0, st, lo));
        AIM_TRY(launch_ewald_recip(pl, ns, positions + 3 * (size_t)lo, charges + lo, s.e_atom + lo, gq + lo,
                                   F ? F + 3 * (size_t)lo : nullptr, s.virial_atom ? s.virial_atom + 9 * (size_t)lo : nullptr, st, 1.0));
    }
    if (n_atoms > 0)
        AIM_CUDA_CHECK(cudaMemcpyAsync(energies_per_atom, s.e_atom, sizeof(double) * n_atoms, cudaMemcpyDeviceToDevice, st));
    if (virial) {
        virial_reduce_kernel<<<n_systems, 256, 0, st>>>(s.mol_ptr, s.virial_atom, -1.0, virial);
        AIM_LAUNCH_CHECK();
    }
    return AIMNET_OK;
}

extern "C" int aimnet2_estimate_ewald_parameters(const float* host_cell, int n_atoms, double accuracy, double* alpha,
                                                 double* real_space_cutoff, double* reciprocal_space_cutoff) {
    AIM_REQUIRE(host_cell && n_atoms >= 1, "estimate_ewald_parameters: need a cell and at least one atom");
    AIM_REQUIRE(accuracy > 0.0 && accuracy < 1.0, "estimate_ewald_parameters: accuracy must be in (0, 1)");
    double a = 0, rc = 0, kc = 0, vol = 0;
    ewald_parameters(host_cell, n_atoms, accuracy, 0.0, a, rc, kc, vol);
    AIM_REQUIRE(vol > 0.0, "estimate_ewald_parameters: singular cell");
    if (alpha) *alpha = a;
    if (real_space_cutoff) *real_space_cutoff = rc;
    if (reciprocal_space_cutoff) *reciprocal_space_cutoff = kc;
    return AIMNET_OK;
}
