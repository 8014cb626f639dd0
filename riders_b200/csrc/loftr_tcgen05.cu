// A3 fused LoFTR encoder layer on the 5th-generation tensor cores (placeholder until the
// tcgen05 kernel lands in this file; the entry points report "unavailable" so that callers fail
// loudly instead of silently computing something else).
#include "common.cuh"

namespace riders {

bool loftr_tc_built() { return false; }

size_t loftr_tc_workspace_bytes(int N, int tokens, int C) {
    return (size_t)N * tokens * C * 4 * sizeof(float) + 256;
}

int loftr_layer_tc(const float*, const float*, const riders_loftr_weights&, int, int, int, int, int, int,
                   float, float, float*, void*, cudaStream_t) {
    set_error("loftr: tcgen05 layer not built into this library");
    return RIDERS_EINVAL;
}

}  // namespace riders
