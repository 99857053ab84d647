/*
 * aimnet2_b200.h — C ABI of the B200-native AIMNet2 hot path (libaimnet2_b200.so).
 *
 * Plain pointers and sizes only; no torch types.  Every entry returns an int status:
 *     0  AIMNET_OK
 *     1  AIMNET_NEIGHBOR_OVERFLOW   a row needs more slots than `max_neighbors`; the host retries with a wider buffer,
 *                                   the contract of NeighborOverflowError in the reference
 *                                   (aimnet/calculators/neighbors.py:127-130)
 *    <0  invalid argument / CUDA failure (aimnet2_last_error() has the text); the Python host raises
 *        ValueError / RuntimeError like aimnet/kernels/conv_sv_2d_sp_wp.py:649-658
 *
 * All device pointers are plain CUDA device memory owned by the caller (torch on the Python side); work is
 * enqueued on `stream` (a cudaStream_t passed as void*), like the reference's Warp ops launch on
 * torch.cuda.current_stream (aimnet/kernels/conv_sv_2d_sp_wp.py:78-82).
 *
 * Which reference interface each entry replaces (paths relative to the reference tree):
 *   aimnet2_neighbor_matrix      nvalchemiops.torch.neighbors.neighbor_list as called at
 *                                aimnet/calculators/neighbors.py:106-125 and aimnet/modules/lr.py:388-396
 *   aimnet2_conv_sv_2d_sp_fwd    torch.ops.aimnet.conv_sv_2d_sp_fwd   aimnet/kernels/conv_sv_2d_sp_wp.py:252-276
 *   aimnet2_conv_sv_2d_sp_bwd    torch.ops.aimnet.conv_sv_2d_sp_bwd   aimnet/kernels/conv_sv_2d_sp_wp.py:285-330
 *   aimnet2_engine_*             AIMNet2Calculator.eval hot path       aimnet/calculators/calculator.py:879-947
 *                                = AIMNet2.forward (aimnet/models/aimnet2.py:141-187) + SRCoulomb/LRCoulomb/DFTD3
 *                                (aimnet/modules/lr.py:311-331, 559-615, 986-1032, 1580-1657) + the autograd
 *                                force/stress pass (aimnet/calculators/derivatives.py:96-146), as analytic kernels
 */
#ifndef AIMNET2_B200_H
#define AIMNET2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIMNET_OK 0
#define AIMNET_NEIGHBOR_OVERFLOW 1
#define AIMNET_EINVAL (-1)
#define AIMNET_ECUDA (-2)

/* Coulomb method of the external LRCoulomb module (aimnet/modules/lr.py:212-331) */
#define AIMNET_COULOMB_NONE 0
#define AIMNET_COULOMB_SIMPLE 1
#define AIMNET_COULOMB_DSF 2
#define AIMNET_COULOMB_EWALD 3

/* output request flags */
#define AIMNET_WANT_FORCES 1
#define AIMNET_WANT_STRESS 2

const char* aimnet2_last_error(void);
/* Bumped whenever a struct below changes (2: aimnet2_options_t.neighbor_skin).  A binding checks it, and the sizes of the
 * four structs as this library was compiled, against its own declarations before the first call. */
#define AIMNET2_ABI_VERSION 2
int aimnet2_abi_version(void);
int aimnet2_abi_struct_sizes(int* weights, int* options, int* system, int* result);

/* ---------------------------------------------------------------------------------------------------------------
 * Neighbor matrix (full list, both directions).  Canonical form: `d2 < cutoff^2` in float32, rows sorted by
 * (j, sx, sy, sz), unused slots = fill_value, zero shifts.
 *   positions   (n_atoms,3) f32 device        batch_idx  (n_atoms) i32 device, sorted, or NULL (one system)
 *   cell        (n_cells,3,3) f32 device or NULL ; host_cell = the same values in host memory (grid sizing)
 *   pbc         3*n_cells bytes on host (0/1), or NULL = all periodic when cell != NULL
 *   nbmat       (n_atoms, max_neighbors) i32 device out
 *   shifts      (n_atoms, max_neighbors, 3) i32 device out, or NULL when cell == NULL
 *   num_neighbors (n_atoms) i32 device out (true counts, may exceed max_neighbors on overflow)
 *   max_count_host   optional host int*, receives max(num_neighbors) (forces a stream sync)
 *   sorted      1 = canonical row order (always for the naive builder; per-row sort for the cell-list builder)
 * ------------------------------------------------------------------------------------------------------------- */
int aimnet2_neighbor_matrix(const float* positions, int n_atoms, float cutoff, const float* cell,
                            const float* host_cell, const uint8_t* pbc, int n_cells, const int32_t* batch_idx,
                            int n_systems, int max_neighbors, int fill_value, int sorted, int32_t* nbmat,
                            int32_t* shifts, int32_t* num_neighbors, int* max_count_host, void* stream);

/* wrap positions into the cell on periodic axes: ((x @ inv(cell)) mod 1) @ cell  (aimnet/calculators/neighbors.py:331) */
int aimnet2_wrap_positions(const float* positions, float* wrapped, int n_atoms, const float* cell, int n_cells,
                           const uint8_t* pbc_host, const int32_t* batch_idx, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * conv_sv_2d_sp operator seam:  a (B,A,G) f32, idx (B,M) i32 with padding value B-1, g (B,M,G,4) f32.
 * out (B,A,G,4); row B-1 is the padding row and is written as zero.
 * ------------------------------------------------------------------------------------------------------------- */
int aimnet2_conv_sv_2d_sp_fwd(const float* a, const int32_t* idx, const float* g, float* out, int B, int A, int G,
                              int M, void* stream);
int aimnet2_conv_sv_2d_sp_bwd(const float* grad_out, const float* a, const int32_t* idx, const float* g,
                              float* grad_a, float* grad_g, int B, int A, int G, int M, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Engine: weights resident in HBM, one call = one E(+F, +stress) evaluation.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct aimnet2_engine aimnet2_engine_t;

/* Weights in the reference's state_dict layout, host pointers, float32 unless noted.
 * Linear weights are (out,in) row-major as torch stores them. */
typedef struct {
    int num_charge_channels;          /* 1 closed-shell, 2 NSE */
    const float* afv;                 /* (64, 256) */
    const float* agh_a;               /* (16,16,12) conv_a.agh */
    const float* agh_q;               /* (C,16,12)  conv_q.agh */
    const float* shifts_s;            /* (16) aev.shifts_s */
    float eta_s;                      /* aev.eta_s */
    float rc_s;                       /* aev.rc_s */
    int n_layers[3];                  /* Linear layers per pass MLP (3,3,4) */
    const int* layer_dims[3];         /* per pass: n_layers+1 ints: in, hidden..., out */
    const float* const* mlp_w[3];     /* per pass: n_layers pointers */
    const float* const* mlp_b[3];
    const float* head_w[3];           /* 256->128->128->1 */
    const float* head_b[3];
    const double* sae;                /* (64) outputs.atomic_shift.shifts.weight, float64 */
    float sr_rc;                      /* outputs.srcoulomb.rc (4.6) */
    int sr_envelope;                  /* 0 exp, 1 cosine */
    /* DFT-D3 tables (aimnet/dftd3_data.pt), may be NULL when dispersion is never requested */
    const float* d3_c6ref;            /* (95,95,5,5) */
    const float* d3_cnref;            /* (95,5) */
    const float* d3_rcov;             /* (95) */
    const float* d3_r4r2;             /* (95) */
} aimnet2_weights_t;

typedef struct {
    int coulomb_method;               /* AIMNET_COULOMB_* ; external LRCoulomb(subtract_sr=False) */
    float dsf_alpha, dsf_rc;          /* 0.2, 15.0 */
    float ewald_accuracy;             /* 1e-6 */
    int dispersion;                   /* 1 = external DFTD3 */
    float d3_s6, d3_s8, d3_a1, d3_a2;
    float d3_cutoff, d3_smoothing;    /* 15.0, 0.2 */
    float sr_cutoff;                  /* 5.0 */
    /* Verlet skin in Angstrom (0 = rebuild the neighbor lists on every evaluation, the reference's behaviour).  With a
     * skin the engine builds its lists at cutoff + skin and keeps them until an atom has moved by more than skin / 2
     * since the build (same atoms, cell and options); every pair kernel applies its own cutoff, so results do not
     * depend on the skin.  Ignored when the caller supplies nbmat. */
    float neighbor_skin;
} aimnet2_options_t;

/* One evaluation.  Flat (mode-1) layout, real atoms only (the engine adds no padding atom; sentinel = n_atoms).
 * Device pointers unless noted. */
typedef struct {
    int n_atoms, n_mol;
    const float* coord;               /* (n_atoms,3) */
    const int32_t* numbers;           /* (n_atoms) */
    const int32_t* mol_idx;           /* (n_atoms) sorted; NULL = single molecule */
    const float* charge;              /* (n_mol) */
    const float* mult;                /* (n_mol) or NULL (NSE only) */
    const float* cell;                /* (n_cells,3,3) or NULL */
    const float* host_cell;           /* host copy of cell (grid sizing), NULL iff cell NULL */
    int n_cells;                      /* 0, 1 or n_mol */
    const uint8_t* pbc_host;          /* 3*n_cells host bytes or NULL */
    /* optional caller-supplied short-range neighbor matrix in the reference layout (n_atoms+1 or n_atoms rows) */
    const int32_t* nbmat;             /* (rows, nb_width) sentinel = n_atoms, or NULL = build it */
    const int32_t* shifts;            /* (rows, nb_width, 3) or NULL */
    int nb_width;
} aimnet2_system_t;

typedef struct {
    double* energy;                   /* (n_mol) f64 */
    float* charges;                   /* (n_atoms) */
    float* spin_charges;              /* (n_atoms) or NULL */
    float* forces;                    /* (n_atoms,3) or NULL */
    float* stress;                    /* (n_cells,3,3) or NULL */
    /* optional: the short-range neighbor matrix the engine built, reference layout with the padding row */
    int32_t* nbmat_out;               /* (n_atoms+1, nbmat_out_width) or NULL */
    int32_t* shifts_out;              /* (n_atoms+1, nbmat_out_width, 3) or NULL */
    int nbmat_out_width;
} aimnet2_result_t;

int aimnet2_engine_create(aimnet2_engine_t** out, const aimnet2_weights_t* w, int device);
int aimnet2_engine_destroy(aimnet2_engine_t* e);
int aimnet2_engine_set_options(aimnet2_engine_t* e, const aimnet2_options_t* opt);
/* GEMM backend for the per-atom MLPs: 0 = fp32 SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05 3xFP16 with row-chunk scaling
 * (default when available), 3 = backend 2's arithmetic with the experimental pipelined tile epilogue (gemm_tc16p.cu; not
 * validated on a GPU yet) */
int aimnet2_engine_set_gemm_backend(aimnet2_engine_t* e, int backend);
/* Backward of the first convolution (whose input features are the species embedding, aimnet/models/aimnet2.py:144-147) through
 * per-species contraction tables instead of per-pair contractions (csrc/conv.cu: species_scan / conv0_table / conv0_force).
 * 1 (default) = on; 0 = the generic pair kernel for every pass.  With more than 16 species in one evaluation the generic
 * kernel runs regardless (decided on the device). */
int aimnet2_engine_set_species_first_pass(aimnet2_engine_t* e, int on);

/* Row capacities of the engine's neighbor matrices.  They grow when a build overflows (the build is retried) and shrink
 * with the reference's hysteresis (aimnet/calculators/neighbors.py:127-140: to widest / 0.75 once the widest row is below
 * half of the capacity, never below the capacities a fresh engine starts with: 64 / 256); the device workspace is re-allocated smaller after 32 evaluations in a row that needed less than
 * half of it. */
int aimnet2_engine_neighbor_caps(const aimnet2_engine_t* e, int* sr_cap, int* lr_cap);

/* AEV / conv_sv kernels: 0 = the list kernels (csrc/conv.cu: one centre per warp walks its matrix row, neighbour rows gathered
 * through L1 / L2) for every input.  For batches of small non-periodic molecules (at most ~100 atoms each): 1 (default) = the
 * FORWARD convolutions take the dense walk of csrc/conv_dense.cu (the molecule's feature table staged into shared memory with
 * TMA, every centre walks all atoms of its molecule; measured 13 % faster than the list forward on 1024 x 50 atoms), the
 * backward pass keeps the list kernel; 2 = dense forward and dense backward (the dense backward executes 50 % more
 * instructions than the list backward and is slower: kept as a measured alternative, profiles/r2k_convd_*).  Pairs beyond the
 * cutoff contribute exactly zero, so the results do not depend on the choice beyond fp32 round-off.  Periodic systems, large
 * molecules and caller-supplied matrices always take the list kernels.  aimnet2_engine_conv_mode reports the setting, whether
 * the last evaluation took the dense walk, and the largest molecule of that batch. */
int aimnet2_engine_set_conv_impl(aimnet2_engine_t* e, int impl);
int aimnet2_engine_conv_mode(const aimnet2_engine_t* e, int* impl, int* dense_last, int* max_molecule_last);
/* The dense walk is taken for batches of at least this many molecules (default 64: one CTA works on one molecule at a time,
 * fewer molecules leave SMs idle and the list kernels win); tests lower it to run small fixtures through the dense kernels. */
int aimnet2_engine_set_dense_min_molecules(aimnet2_engine_t* e, int n_mol);
/* Evaluations with at most `rows` atoms run the MLPs on the small-M fp32 SIMT kernel whatever the backend (the
 * tensor-core pipelines are latency-bound for a single molecule); default 512, 0 = never. */
int aimnet2_engine_set_small_m_rows(aimnet2_engine_t* e, int rows);
/* Test seam: fill the engine's device workspace with `byte` (0..255) before every evaluation, so that a kernel reading
 * scratch memory it has not written shows up in the results (0xFF = NaN patterns); -1 switches it off (default). */
int aimnet2_engine_debug_poison(aimnet2_engine_t* e, int byte);
/* Debug seams for fault isolation (tools/first_touch.py): the names, byte offsets and sizes of the workspace buffers of the
 * last evaluation as text lines "name offset bytes\n" (buf_bytes must hold them all), and a synchronous copy of a
 * workspace range to host memory.  The layout is only meaningful until the next evaluation. */
int aimnet2_engine_debug_layout(const aimnet2_engine_t* e, char* buf, int buf_bytes);
int aimnet2_engine_debug_read_workspace(aimnet2_engine_t* e, void* host_dst, int64_t offset, int64_t bytes);
/* 1 = bitwise run-to-run reproducible results (every kernel of the engine already uses fixed chunking and no atomics, so this
 * only records the request; the flag is per engine) — the counterpart of AIMNet2Calculator(deterministic=True), aimnet/calculators/calculator.py:76-84 */
int aimnet2_engine_set_deterministic(aimnet2_engine_t* e, int on);

/* device-resident inputs/outputs; flags = AIMNET_WANT_* */
int aimnet2_engine_eval(aimnet2_engine_t* e, const aimnet2_system_t* sys, const aimnet2_result_t* res, int flags,
                        void* stream);
/* same with every pointer of sys/res in HOST memory: H2D copies, evaluation, D2H copies, stream sync inside */
int aimnet2_engine_eval_host(aimnet2_engine_t* e, const aimnet2_system_t* sys, const aimnet2_result_t* res, int flags);

/* CUDA-graph replay of the fixed-shape step (the MD / optimizer loop of SURVEY.md section 8f f1; the reference's
 * counterpart is the static-shape contract of compile_model / cache_static, aimnet/calculators/calculator.py:1091-1238).
 * With on = 1, an evaluation whose sizes, flags and options repeat is captured once (on its second occurrence) and from
 * then on replayed as one graph launch; inputs / outputs pass through engine-owned staging buffers so that the recorded
 * addresses stay valid (engine_eval: device-to-device copies, engine_eval_host: the H2D / D2H copies it makes anyway).  The
 * replayed neighbor build cannot grow its buffers: the widest rows are read back after the replay and, on overflow, the step
 * is redone eagerly.  Not captured (such calls silently take the eager path): Ewald, Verlet skin, caller-supplied
 * matrices, partial pbc, periodic systems of >= 512 atoms (cell-list builder), timing / poison modes.
 * graph_stats: graphs captured, replays, replays redone eagerly after an overflow. */
int aimnet2_engine_enable_cuda_graph(aimnet2_engine_t* e, int on);
int aimnet2_engine_graph_stats(const aimnet2_engine_t* e, int* captures, int* launches, int* fallbacks);

/* introspection: kernels launched by the last eval, last short-range / long-range list widths, workspace bytes */
int aimnet2_engine_last_launches(const aimnet2_engine_t* e);
int aimnet2_engine_info(const aimnet2_engine_t* e, int* sr_width, int* lr_width, int64_t* workspace_bytes);
/* Verlet skin bookkeeping: evaluations that built the lists / reused them (options.neighbor_skin > 0) */
int aimnet2_engine_skin_stats(const aimnet2_engine_t* e, int* builds, int* reuses);
/* per-phase device times (ms) of the last eval when timing was enabled (level 1: phase events, level 2: also one
 * event pair around every GEMM launch); slots: 0 neighbors, 1 forward,
 * 2 long-range, 3 backward, 4 total, then (level 2) 5 summed GEMM ms, 6 GEMM launches, 7 summed AEV / conv_sv ms,
 * 8 conv calls; returns the number of slots written */
int aimnet2_engine_enable_timing(aimnet2_engine_t* e, int on);
int aimnet2_engine_last_timing(const aimnet2_engine_t* e, float* ms, int n);

/* ---------------------------------------------------------------------------------------------------------------
 * Pair-term operator seams: the external long-range kernels behind entry points shaped like the third-party calls of
 * the reference (SURVEY.md section 8b, B3 iii).  Common arguments:
 *   positions (n_atoms,3) f32 device, Angstrom      cell (n_cells,3,3) f32 device or NULL, n_cells in {0, 1, n_systems}
 *   batch_idx (n_atoms) i32 device, sorted, or NULL (one system)
 *   nbmat (n_atoms, nb_width) i32 device, full list (both directions), unused slots = fill_value
 *   shifts (n_atoms, nb_width, 3) i32 device integer lattice vectors (required with a cell)
 *   energy (n_systems) f64 device out              forces (n_atoms,3) f32 device out or NULL
 *   virial (n_systems,3,3) f64 device out or NULL: W = -dE/d(strain), the convention of
 *   aimnet/calculators/derivatives.py:128-131
 * ------------------------------------------------------------------------------------------------------------- */

/* Damped-shifted-force Coulomb; replaces nvalchemiops...dsf_coulomb as called at aimnet/modules/lr.py:526-540 (closed
 * form: lr.py:594-611).  Energies in e^2/Angstrom, forces in e^2/Angstrom^2: the caller multiplies by Hartree*Bohr
 * (lr.py:542-547).  charge_grad (n_atoms) f32 out or NULL: dE/dq_i at fixed geometry (what autograd supplies through the
 * graph-attached charges in the reference). */
int aimnet2_dsf_coulomb(const float* positions, const float* charges, int n_atoms, float cutoff, float alpha,
                        const float* cell, int n_cells, const int32_t* batch_idx, int n_systems, const int32_t* nbmat,
                        const int32_t* shifts, int nb_width, int fill_value, double* energy, float* forces,
                        float* charge_grad, double* virial, void* stream);

/* DFT-D3(BJ) two-body dispersion with the quintic 5th-order switch; replaces nvalchemiops...dftd3 as called at
 * aimnet/modules/lr.py:1204-1228 (closed form: lr.py:1580-1657).  Positions in Angstrom, energies in eV, forces in
 * eV/Angstrom (the Python wrapper converts from / to the Bohr / Hartree units of that call site); r_on / r_off in Bohr.
 * Tables in the packed form of this library: c6ref (95,95,28) f32 = the (95,95,5,5) reference C6 with every 25-value
 * row padded to 28 (16-byte aligned), cnref (95,5) f32 = reference CN of element z's a-th reference system, -1 where
 * that reference does not exist; rcov, r4r2 (95) f32.  coord_num (n_atoms) f32 out or NULL. */
int aimnet2_dftd3(const float* positions, const int32_t* numbers, int n_atoms, float s6, float s8, float a1, float a2,
                  float r_on_bohr, float r_off_bohr, const float* c6ref, const float* cnref, const float* rcov,
                  const float* r4r2, const float* cell, int n_cells, const int32_t* batch_idx, int n_systems,
                  const int32_t* nbmat, const int32_t* shifts, int nb_width, int fill_value, double* energy,
                  float* forces, float* coord_num, double* virial, void* stream);

/* Ewald summation; replaces nvalchemiops...ewald_summation as called at aimnet/modules/lr.py:687-696 (an energy-only call
 * there: forces, stress and the charge response come from autograd through the returned energies; here they are explicit
 * optional outputs, like in the DSF seam).  Splitting parameters per system from `accuracy`
 * (aimnet/calculators/calculator.py:663-666); real space over the caller's neighbor matrix, which must reach every system's
 * real-space cutoff (aimnet2_estimate_ewald_parameters); reciprocal space, self and neutralising-background terms per system.
 * Units e^2/Angstrom (the caller multiplies by Hartree*Bohr, lr.py:697).
 *   cell                (n_systems,3,3) f32 device, one cell per system; host_cell: the same values in host memory
 *   host_system_offsets n_systems + 1 ints in HOST memory, atoms [off[s], off[s+1]) = system s; NULL when n_systems == 1
 *   energies_per_atom   (n_atoms) f64 device out; real-space and self terms per atom, the reciprocal-space energy of a
 *                       system booked on its first atom (only per-system sums are defined: lr.py:698-703)
 *   forces (n_atoms,3) f32 / charge_grad (n_atoms) f32 / virial (n_systems,3,3) f64 device out, each may be NULL
 * Reciprocal-space plans are cached per system slot in process-wide storage: one caller thread at a time. */
int aimnet2_ewald_summation(const float* positions, const float* charges, int n_atoms, const float* cell,
                            const float* host_cell, const int32_t* batch_idx, const int32_t* host_system_offsets,
                            int n_systems, const int32_t* nbmat, const int32_t* shifts, int nb_width, int fill_value,
                            double accuracy, double* energies_per_atom, float* forces, float* charge_grad, double* virial,
                            void* stream);

/* Ewald splitting parameters for a target accuracy (host arithmetic only); replaces
 * nvalchemiops...estimate_ewald_parameters as used at aimnet/calculators/calculator.py:1566-1587, formulas
 * calculator.py:663-666: eta = (V^2/N)^(1/6)/sqrt(2 pi), r_c = sqrt(-2 ln eps) eta, k_c = sqrt(-2 ln eps)/eta,
 * alpha = 1/(sqrt(2) eta).  host_cell: 9 floats in host memory; outputs may be NULL. */
int aimnet2_estimate_ewald_parameters(const float* host_cell, int n_atoms, double accuracy, double* alpha,
                                      double* real_space_cutoff, double* reciprocal_space_cutoff);

/* standalone NT GEMM with fused epilogue (test seam for the MLP kernels):
 * Y[M,N] = act(A[M,K] @ W[N,K]^T + bias); mode 0 none, 1 bias, 2 bias+GELU (also writes gelu'(z) to aux when non-NULL),
 * 3 multiply by aux[M,N].  lda/ldw/ldy in elements.  backend: 0 fp32 SIMT, 1 tcgen05 3xTF32, 2 tcgen05 3xFP16 (what the
 * engine runs), 3 = experimental 3xFP16 kernel with the tile epilogue pipelined under the next tile's MMAs (not used by
 * the engine, not yet validated on a GPU).  mode | 16 (backends 2, 3): the kernel writes its output pre-split. */
int aimnet2_gemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float* Y, int ldy,
                    float* aux, int ldaux, int M, int N, int K, int mode, int backend, void* stream);

/* debug aid for the 3xFP16 GEMM (backend 2): when device_buf is non-NULL (16384 uint64 of device memory), CTA 0 of
 * every following launch records SM-clock stamps per pipeline stage (8 events x 2048 stages); NULL switches it off. */
int aimnet2_gemm_set_trace(void* device_buf);

#ifdef __cplusplus
}
#endif
#endif
