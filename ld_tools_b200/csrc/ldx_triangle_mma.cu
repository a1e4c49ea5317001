// ldx_triangle_mma.cu -- K5: all-pairs (1,1) counts as an exact int8 Gram matrix on tcgen05.
// (placeholder until the tensor-core engine lands; the popcount engine serves every request)
#include "ldx_internal.h"

namespace ldx {
bool triangle_mma_available() { return false; }
int launch_triangle_mma(ldx_store *, const int64_t *, int64_t, int, int, int, uint32_t *, int32_t *) {
    return set_error(LDX_ERR_ARG, "the tcgen05 engine is not available in this build");
}
}  // namespace ldx
