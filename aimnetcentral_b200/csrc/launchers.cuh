// Host-side entry points of the translation units in this directory (every launcher enqueues on `st` and returns an
// AIMNET_* status); one declaration each, shared by the definition and its callers.
#pragma once
#include "common.cuh"

namespace aimnet {

// ---- nblist.cu: neighbor matrices and position wrapping
int neighbor_matrix_impl(const float* positions, int n_atoms, float cutoff, const float* cell, const float* host_cell,
                         const uint8_t* pbc_host, int n_cells, const int32_t* batch_idx, int n_systems, int max_nb,
                         int fill_value, int sorted, int32_t* nbmat, int32_t* shifts, int32_t* nnb,
                         int* max_count_host, cudaStream_t st, bool prefer_cells, int32_t* scratch = nullptr,
                         int32_t* pinned_host = nullptr);
int wrap_positions_impl(const float* positions, float* wrapped, int n_atoms, const float* cell, int n_cells,
                        const uint8_t* pbc_host, const int32_t* batch_idx, cudaStream_t st);

// ---- conv.cu: AEV + conv_sv message passing (forward, backward)
int launch_conv_fwd(int C, int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                    const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q, const float* agh_a,
                    const float* agh_q, float* x, int ldx, float* T_a, float* T_q, int with_q, cudaStream_t st);
int launch_conv_bwd(int C, int n_atoms, const NbView& nb, const float* coord, const CellView& cv,
                    const int32_t* mol_idx, const AevParams& aev, const float* aT, const float* q, const float* dx,
                    int ldx, const float* T_a, const float* T_q, const float* agh_a, const float* agh_q, float* dS_a,
                    float* dS_q, float* grad_a, float* grad_q, float* forces, double* virial_atom, int with_q,
                    int want_grad_a, cudaStream_t st, const int* skip_if = nullptr, bool prep = true);
// first convolution's backward by species (conv.cu): slots of the present species, then table + pair kernels
int launch_species_scan(int n_atoms, const int32_t* numbers, int* info, uint8_t* atom_slot, cudaStream_t st);
int launch_conv0_bwd_species(int n_atoms, const NbView& nb, const float* coord, const CellView& cv, const int32_t* mol_idx,
                             const AevParams& aev, const int* info, const uint8_t* atom_slot, const float* afvT,
                             const float* dS_a, float* P, float* forces, double* virial_atom, cudaStream_t st);
int conv0_species_bytes_per_atom();
int conv0_species_info_ints();
int conv0_species_max_slots();

int launch_conv_bwd_prep(int C, int n_atoms, const float* dx, int ldx, const float* T_a, const float* T_q, const float* agh_a,
                         const float* agh_q, float* dS_a, float* dS_q, int with_q, int permute, cudaStream_t st);

// ---- conv_dense.cu: the same convolutions for batches of small molecules: feature tables of one molecule staged into
// shared memory with TMA, every pair of the molecule walked from there
int conv_dense_max_atoms(int C);
int launch_conv_dense_fwd(int C, int n_atoms, int n_mol, int max_seg, const int32_t* mol_ptr, const float* coord,
                          const AevParams& aev, const float* aT, const float* q, const float* agh_a, const float* agh_q,
                          float* x, int ldx, float* T_a, float* T_q, int with_q, cudaStream_t st);
int launch_conv_dense_bwd_gather(int C, int n_atoms, int n_mol, int max_seg, const int32_t* mol_ptr, const float* coord,
                                 const AevParams& aev, const float* aT, const float* q, const float* dS_a, const float* dS_q,
                                 float* grad_a, float* grad_q, float* forces, float* f_part, float* gq_part, int with_q,
                                 int want_grad_a, cudaStream_t st);

// ---- gemm.cu: per-atom MLP GEMMs: backend dispatch
int gemm_nt(const float* A, int lda, const WeightView& w, const float* bias, float* Y, int ldy, float* aux, int ldaux,
            int M, int N, int K, int mode, int backend, cudaStream_t st);
int gemm_nt_split(const SplitMat& A, const WeightView& w, const float* bias, float* Y, int ldy,
                  const SplitMat* Ysplit, float* aux, int ldaux, int M, int N, int K, int mode, int variant,
                  cudaStream_t st);

// ---- gemm_tc.cu: tcgen05 3xTF32 backend
bool gemm_tc_available();
int split_tf32(const float* w, float* hi, float* lo, size_t n, cudaStream_t st);
int gemm_nt_tc(const float* A, int lda, const float* Whi, const float* Wlo, int ldw, const float* bias, float* Y,
               int ldy, float* aux, int ldaux, int M, int N, int K, int mode, cudaStream_t st);

// ---- gemm_tc16.cu: tcgen05 3xFP16 backend, pre-split activations
void gemm_tc16_set_trace(unsigned long long* buf);
unsigned long long* gemm_tc16_get_trace();
int split_fp16_device(const float* w, void* hi, void* lo, float* inv_scale, unsigned int* scratch, size_t n,
                      cudaStream_t st);
int presplit_f32(const float* X, int ldx, int M, int K, const SplitMat& out, cudaStream_t st);
int unsplit_f32(const SplitMat& in, int M, int N, float* Y, int ldy, cudaStream_t st);
int gemm_nt_tc16(const SplitMat& A, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw,
                 const float* bias, float* Y, int ldy, const SplitMat* Ysplit, float* aux, int ldaux, int M, int N,
                 int K, int mode, cudaStream_t st);

// ---- gemm_tc16p.cu: experimental backend 3 (software-pipelined tile epilogue); same contract as gemm_nt_tc16
int gemm_nt_tc16p(const SplitMat& A, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw,
                  const float* bias, float* Y, int ldy, const SplitMat* Ysplit, float* aux, int ldaux, int M, int N,
                  int K, int mode, cudaStream_t st);

// ---- gemm_tc16d.cu: backend 4 (two tile streams per SM: epilogue of one under the MMAs of the other); same contract
int gemm_nt_tc16d(const SplitMat& A, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw,
                  const float* bias, float* Y, int ldy, const SplitMat* Ysplit, float* aux, int ldaux, int M, int N,
                  int K, int mode, cudaStream_t st);

// ---- gemm_tc16c.cu: backend 5 (the two-stream kernel on CTA pairs, tcgen05 cta_group::2); same contract
int gemm_nt_tc16c(const SplitMat& A, const void* Whi, const void* Wlo, const float* w_inv_scale, int ldw,
                  const float* bias, float* Y, int ldy, const SplitMat* Ysplit, float* aux, int ldaux, int M, int N,
                  int K, int mode, cudaStream_t st);

// ---- pointwise.cu: embedding, NSE charge equilibration, reductions, Verlet-skin bookkeeping
int launch_embed(int n, const int32_t* numbers, const float* afv, float* a0, cudaStream_t st);
int launch_mol_ptr(const int32_t* mol_idx, int n, int n_mol, int32_t* ptr, int32_t* max_segment, cudaStream_t st);
int launch_nse_fwd(int C, int n, int n_mol, const int32_t* mol_idx, const int32_t* mol_ptr, const float* charge,
                   const float* mult, const float* y, int ldy, const float* q_prev, float* sumq, float* sumf,
                   const float* a_old, float* a_new, float* q_new, cudaStream_t st);
int launch_nse_bwd(int C, int n, int n_mol, const int32_t* mol_idx, const int32_t* mol_ptr, const float* charge,
                   const float* mult, const float* y, int ldy, const float* gq, const float* sumq, const float* sumf,
                   float* s1, const float* da_tot, const float* gp_last, int ldgp, float* dz, int lddz,
                   float* dq_prev, cudaStream_t st);
int launch_accum_grads(int C, int n, const float* dx, int ldx, const float* grad_a, const float* grad_q,
                       const float* base_q, int base_q_stride, float* da_tot, int accumulate, float* dq,
                       cudaStream_t st);
int launch_head_tail(int n, const float* h2, int ldh, const float* gp2, const float* w3, float b3,
                     const int32_t* numbers, const double* sae, double* e_atom, float* dz2, cudaStream_t st);
int launch_energy_reduce(int n_mol, const int32_t* mol_ptr, const double* e0, const double* e1, const double* e2,
                         const double* e3, double* energy, cudaStream_t st);
int launch_stress_reduce(const int32_t* mol_ptr, int n_cells, int n, const double* virial_atom, const float* cell,
                         float* stress, cudaStream_t st);
int launch_skin_check(int n, const float* x, const float* ref, float thr2, const int32_t* mol, const int32_t* mol_ref,
                      int32_t* flag, cudaStream_t st);
int launch_skin_save(int n, const float* x, const float* wrapped, float* ref, float* off, const int32_t* mol,
                     int32_t* mol_ref, cudaStream_t st);
int launch_skin_apply(int n, const float* x, const float* off, float* out, cudaStream_t st);
int launch_charges_out(int C, int n, const float* q, float* charges, float* spin, cudaStream_t st);

// ---- lr.cu: pair terms: Coulomb (simple / DSF / Ewald real space), DFT-D3
int launch_coulomb(int mode, int n, const PairSource& ps, const float* coord, const CellView& cv, const float* q,
                   const CoulombParams& p, double* e_atom, float* gq, float* forces, double* virial_atom,
                   int accumulate_e, cudaStream_t st, int atom_lo = 0);
int launch_d3(int n, const PairSource& ps, const float* coord, const CellView& cv, const int32_t* numbers,
              const D3Params& p, float* cn, float* wtab, float* dEdCN, double* e_atom, float* forces,
              double* virial_atom, cudaStream_t st);

// ---- ewald.cu: Ewald reciprocal space
// ke: Coulomb constant of the output units (eV A by default; 1 for the operator seam, whose caller applies Hartree*Bohr)
int launch_ewald_recip(const EwaldPlan& pl, int n, const float* coord, const float* q, double* e_atom, float* gq,
                       float* forces, double* virial_atom, cudaStream_t st, double ke = kHartree * kBohr);

}  // namespace aimnet
