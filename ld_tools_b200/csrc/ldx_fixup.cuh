// ldx_fixup.cuh -- device side of the near-tie list (see LDX_R2_NEARTIE in ldx_common.cuh).
#pragma once
#include "ldx_common.cuh"

namespace ldx {

struct FixupSink {
    FixupRec *recs;
    uint32_t *count;
    uint32_t capacity;
    uint64_t tag;          // ORed into out_index: which enqueued call the record belongs to (bits 48..63)
};
constexpr int FIX_TAG_SHIFT = 48;
constexpr uint64_t FIX_INDEX_MASK = (1ull << FIX_TAG_SHIFT) - 1;

__device__ __forceinline__ void fixup_append(const FixupSink &s, uint64_t out_index, int32_t n11,
                                             int32_t n1a, int32_t n1b, uint32_t packed) {
    const uint32_t slot = atomicAdd(s.count, 1u);
    if (slot < s.capacity) {
        FixupRec r;
        r.out_index = out_index | s.tag; r.n11 = n11; r.n1a = n1a; r.n1b = n1b; r.packed = packed;
        s.recs[slot] = r;
    }
}

}  // namespace ldx
