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
        r.n_pair = 0; r.n0a = 0; r.n0b = 0; r.pad = 0;
        s.recs[slot] = r;
    }
}
// The same for a pair of the general route: the host needs all six counts (n_pair != 0 marks the record).
__device__ __forceinline__ void fixup_append_general(const FixupSink &s, uint64_t out_index, const GenCounts &c, uint32_t packed) {
    const uint32_t slot = atomicAdd(s.count, 1u);
    if (slot < s.capacity) {
        FixupRec r;
        r.out_index = out_index | s.tag; r.n11 = c.n11; r.n1a = c.n1a; r.n1b = c.n1b; r.packed = packed;
        r.n_pair = c.n_pair; r.n0a = c.n0a; r.n0b = c.n0b; r.pad = 0;
        s.recs[slot] = r;
    }
}

// The completion record of an enqueued call goes to pinned host memory as ONE aligned 64-bit store (single-copy atomic,
// one PCIe write): sequence number in bits 0-31, near-tie count (saturated) in bits 32-55, error flag in bits 56-63.  The
// host polls that word: it can never see a new sequence number with stale counts, and no system-wide fence has to sit
// between payload and flag on the kernel's critical path (it cost the last CTA of the single-wave kernel ~2.4 us).
__device__ __forceinline__ void publish_record(volatile uint32_t *mailbox, uint32_t seq, uint32_t n_fix, uint32_t err) {
    const unsigned long long w = (unsigned long long)seq | ((unsigned long long)(n_fix < 0xffffffu ? n_fix : 0xffffffu) << 32) |
                                 ((unsigned long long)(err & 0xffu) << 56);
    asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(mailbox), "l"(w) : "memory");
}

}  // namespace ldx
