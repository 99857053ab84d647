// tcgen05 3xTF32 backend of the per-atom MLP GEMMs (placeholder until the TMEM kernel lands).
#include "common.cuh"

namespace aimnet {

bool gemm_tc_available() { return false; }

int gemm_nt_tc(const float*, int, const float*, int, const float*, float*, int, float*, int, int, int, int, int,
               cudaStream_t) {
    set_error("gemm: tcgen05 backend not built");
    return AIMNET_EINVAL;
}

}  // namespace aimnet
