// ldx_fixup.cuh -- device side of the near-tie list (see LDX_R2_NEARTIE in ldx_common.cuh).
#pragma once
#include "ldx_common.cuh"

namespace ldx {

struct FixupSink {
    FixupRec *recs;
    uint32_t *count;
    uint32_t capacity;
};

__device__ __forceinline__ void fixup_append(const FixupSink &s, uint64_t out_index, int32_t n11,
                                             int32_t n1a, int32_t n1b, uint32_t packed) {
    const uint32_t slot = atomicAdd(s.count, 1u);
    if (slot < s.capacity) {
        FixupRec r;
        r.out_index = out_index; r.n11 = n11; r.n1a = n1a; r.n1b = n1b; r.packed = packed;
        s.recs[slot] = r;
    }
}

}  // namespace ldx
