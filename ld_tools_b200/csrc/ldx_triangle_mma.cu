// ldx_triangle_mma.cu -- K5: all-pairs (alt, alt) haplotype counts as an EXACT int8 Gram matrix
// on the 5th-generation tensor cores (tcgen05.mma kind::i8, int32 accumulators in TMEM), with the
// D / D' / r2 finalisation fused into the epilogue.
//
// Replaces the double loop at ld_triangle.py:133-230 (var_1 = row variant, var_2 = column
// variant, ld_triangle.py:193); the counting step is calc_ld.py:30-32 for 128 x N pairs at once:
//     n11[r][c] = sum_h  a[r][h] * a[c][h],   a[v][h] = (plane[v] & mask) bit h
// Operand bytes are 0 or a power of two chosen so that every (alt, alt) product is 2^7 (see
// widen4): the int32 accumulator holds the count * 128 exactly; the epilogue shifts it back.
//
// MINOR-ALLELE ENCODING.  The operands are not the alt bits themselves: a variant whose alt allele is the
// major one under the mask (2 * n1 > N) enters with its bits complemented (under the mask), so that every
// operand row has at most N/2 bits set.  r2 and |D'| of calc_ld.py:50-90 do not change when an allele is
// relabelled (D changes sign), and with counts <= N/2 every integer the screening arithmetic of the epilogue
// needs -- n11' * N, n1a' * n1b', Dn, the D' bound -- is below 2^24 (N <= 5792; one rounding up to N = 8192)
// and therefore EXACT in single precision: the epilogue runs on FFMA2 / FMUL2 alone, no integer multiply-add
// chain and no int -> float conversion of Dn (see fast_pair2).  The exact path (finalise_pair, the reference's
// own operation sequence in fp64) gets the true counts back: true_n11().  Results equal the popcount engine's
// bit for bit (tests/test_parity_gpu.py).
//
// Data movement is designed around the L2: the operands stay BIT-PACKED in global memory/L2
// (16 B per variant per 128-haplotype chunk, gathered once per call by gather_bits_kernel, O(V))
// and are widened to bytes inside the SM.  Streaming pre-widened int8 panels from L2 was measured
// first (profiles/r01_ncu_full_mma_v1_int8_from_l2_tile64.txt): 7.6 TB/s of L2->SM traffic with
// both the tensor and the fp64 pipe under 14% busy.  Bit-packed operands cut that traffic 8x and
// keep a 100k-variant operand set (64 MB) resident in the 126 MB L2.
//
// Persistent, warp-specialised CTAs (one per SM, 28 warps), each looping over 128 x N tiles of
// the lower triangle(s) -- one launch takes up to MMA_MAX_SETS independent variant sets (ldx_triangle_batch_dev):
//   warp 0      producer: cp.async.bulk (UBLKCP, the TMA engine's linear mode) brings the tiles'
//               bit blocks (2 KB per 128 variants per chunk) into an 8-deep shared-memory ring;
//               completion is counted on mbarriers (complete_tx).  Runs ahead across tiles.
//               DIRECT mode (one-wave calls on contiguous store rows): no gather kernel at all -- the
//               producer reads 128-row x 32-byte boxes of the store's planes through a TMA tensor map
//               (cp.async.bulk.tensor.2d, UTMALDG) and the wideners apply mask, allele flip and bit reversal.
//   warps 4-19  wideners: each thread turns one variant-row of a stage (256 bits) into operand bytes: the row
//               operand straight into TENSOR MEMORY (tcgen05.st), the column operand into the 128-byte-swizzled
//               shared-memory tile tcgen05 reads, then fence.proxy.async + mbarrier arrive.  Four teams, one
//               operand stage each.
//   warp 1      MMA issuer: one elected lane issues 8 x tcgen05.mma (K = 32 each) per stage,
//               tcgen05.commit's the operand stage back to the wideners and, per tile, the TMEM
//               accumulator to the epilogue; owns the TMEM allocation (2 accumulators of N columns).
//   warps 20-27 epilogue: tcgen05.ld the int32 counts (16 lanes x 256 bits: a lane owns two row variants and
//               eight column variants per load), screen them in single precision, store the packed words
//               straight from registers.  Works on accumulator t while the tensor pipe fills t+1.
//
// Roofline: int8 tensor pipe, 2 * n_hap int8 ops per pair; co-bounds are the epilogue's instruction issue
// and the widening ALU work ((128 + N) rows per 128 * N pairs).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ldx_internal.h"
#include "ldx_fixup.cuh"

#define LDX_TRY(expr) do { int rc__ = (expr); if (rc__ != LDX_OK) return rc__; } while (0)

namespace ldx {

constexpr int MMA_M = 128;            // rows (variants) per tile = TMEM lanes
constexpr int KCHUNK = 128;           // haplotypes (= bytes) per shared-memory row: one swizzle atom
constexpr int MMA_K = 32;             // int8 K of one tcgen05.mma


// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// Bounded wait: a pipeline bug must end in an error code, never in a hung GPU.  The common case
// (phase already complete) is one try_wait; the slow path polls, optionally sleeping between
// polls so that long waits (the epilogue waiting out a whole K loop) do not steal issue slots
// from the warps doing the work, and checks the abort flag / a 2 s deadline now and then.
__device__ __forceinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, volatile int *abort_s, int32_t *err, uint32_t sleep_ns) {
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try_wait(bar, parity)) return true;
        if (sleep_ns) __nanosleep(sleep_ns);
        if ((spins & 0x3f) == 0) {
            if (*abort_s) return false;
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (!t0) t0 = t;
            else if (t - t0 > 2000000000ull) { *abort_s = 1; atomicExch(err, 1); return false; }
        }
    }
}
template <bool CLUSTER = false>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int *abort_s, int32_t *err, uint32_t sleep_ns = 0) {
    if (CLUSTER) {      // the barrier also takes arrives from the peer CTA of a pair: acquire at cluster scope
        unsigned long long t0 = 0;
        for (uint32_t spins = 1;; ++spins) {
            if (mbar_try_wait_cluster(bar, parity)) return true;
            if ((spins & 0x3f) == 0) {
                if (*abort_s) return false;
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                if (!t0) t0 = t;
                else if (t - t0 > 2000000000ull) { *abort_s = 1; atomicExch(err, 1); return false; }
            }
        }
    }
    if (mbar_try_wait(bar, parity)) return true;
    return mbar_wait_slow(bar, parity, abort_s, err, sleep_ns);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// One box of a 2-D tensor map (direct mode: 128 store rows x 32 bytes of their planes) -> dense shared-memory tile.
// Rows beyond the tensor are zero-filled and still count towards complete_tx.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int32_t x, int32_t y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(x), "r"(y) : "memory");
}
// One lane of a fully converged warp; the surrounding control flow stays warp-uniform so that the
// compiler keeps descriptors/addresses in uniform registers (issuing tcgen05/bulk-copy instructions
// from a lane-divergent region costs an R2UR waterfall loop per instruction).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Programmatic dependent launch: the three kernels of a call (gather -> all-pairs -> deferred pairs) are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, so that a kernel's launch latency and
// prologue overlap the tail of its predecessor.  pdl_wait() blocks until the predecessor grid has completed
// and its writes are visible (a no-op without the attribute); pdl_launch_dependents() lets the successor
// start being scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): two CTAs of a cluster share one 256 x N tile; rank 0 issues the MMAs
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> the same variable in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrive on a barrier of the pair's other CTA.  What the arrive announces -- operand bytes in THIS CTA's shared and tensor
// memory -- has been completed by the caller's tcgen05.wait::st / fence.proxy.async / __syncwarp before this point, and it is
// consumed in place (the pair's MMA reads each CTA's operands from that CTA's own memories).  A cluster-scope RELEASE here
// made the arriving thread wait ~0.6 us for a fence it does not need, once per stage and warp, on every team's critical
// path: the pair kernel ran slower than the single-CTA one (2.39 vs 2.10 ms at 32,768 variants).  A CTA-scope fence plus a
// relaxed cluster-scope arrive: 1.89 ms.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_i8_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// The same with the A operand (row variants) read from TENSOR MEMORY: lane = row, 32-bit column c holds
// K elements 4c..4c+3.  Shared-memory bandwidth is what bounds this kernel (the widened operands are
// written once and read once per chunk), and this variant takes A's write AND read off that port.
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane (lane 32*(warp%4) + laneid).
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                    "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                    "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16 lanes x 256 bits, four times along the columns (32 columns): see the epilogue for the register map.
__device__ __forceinline__ void tmem_ld16x256(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor):
// [0,14) start >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) | [32,46) SBO >> 4 = 1024 B
// between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 @ [4,6); a/b format
// UINT8 = 0 @ [7,10)/[10,13); K-major A and B (bits 15, 16 = 0); N >> 3 @ [17,23); M >> 4 @ [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ gather
// Store rows (gathered through rows[], masked, minor-allele encoded) -> chunk-blocked bit panels:
//   bits[panel][chunk][row 0..127][16 B],  panel = matrix_row / 128,  chunk = haplotype / 128
// so that the 128 rows of one pipeline stage are ONE contiguous 2 KB block (one bulk copy).
// A second copy with each byte bit-reversed feeds the column operand.  The variant sets of a launch sit one after
// the other in this row space, each padded to whole 256-row panels: matrix row r of set k is global row base_row + r.
struct GatherSet {
    const uint64_t *planes, *mask; const VarFreq *freq; const int64_t *rows;
    int64_t v, v_pad, base_row; int32_t stride_words, n_sel;
};
struct GatherArgs {
    GatherSet set[MMA_MAX_SETS];
    int32_t kc_count; uint4 *bits, *bits_rev; VarFreq *freq_rows;
};

__device__ __forceinline__ uint32_t brev_bytes(uint32_t x) { return __byte_perm(__brev(x), 0, 0x0123); }   // every BYTE bit-reversed

__global__ void __launch_bounds__(256)
gather_bits_kernel(const __grid_constant__ GatherArgs G) {
    pdl_launch_dependents();                                            // the all-pairs kernel may start its prologue
    pdl_wait();                                                         // the scratch is still read by the previous call's kernels
    const GatherSet &S = G.set[blockIdx.z];
    const int kc = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;          // matrix row of this set
    if (r >= S.v_pad) return;
    const int64_t gr = S.base_row + r;                                  // row in the launch's row space
    uint4 out = make_uint4(0, 0, 0, 0);
    if (r < S.v) {
        const int64_t srow = S.rows[r];
        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(S.planes + srow * S.stride_words) + kc);
        const uint4 m = __ldg(reinterpret_cast<const uint4 *>(S.mask) + kc);
        const VarFreq f = S.freq[srow];
        const uint32_t flip = 2 * f.n1 > S.n_sel ? 0xffffffffu : 0u;    // alt is the major allele: operand = ref bits
        out = make_uint4((x.x ^ flip) & m.x, (x.y ^ flip) & m.y, (x.z ^ flip) & m.z, (x.w ^ flip) & m.w);
        if (kc == 0) G.freq_rows[gr] = f;                               // the TRUE counts: true_n11() undoes the flip
    } else if (kc == 0) {
        VarFreq z; z.p = 0.0; z.q = 0.0; z.pq = 0.0; z.n1 = 0; z.p_e4 = 0;
        G.freq_rows[gr] = z;
    }
    const int64_t o = ((gr >> 7) * G.kc_count + kc) * 128 + (gr & 127);
    G.bits[o] = out;
    // the same bits with every BYTE bit-reversed: the column operand's source (see widen_row)
    G.bits_rev[o] = make_uint4(brev_bytes(out.x), brev_bytes(out.y), brev_bytes(out.z), brev_bytes(out.w));
}

// ------------------------------------------------------------------------------------------ widen
// 32 haplotype bits -> 32 operand bytes with ONE logic instruction per 4 bytes and no shifts:
//     row operand     byte(j, p) = w  & (0x01 << p)        -> value 2^p      for haplotype 8j+p
//     column operand  byte(j, p) = w' & (0x80 >> p)        -> value 2^(7-p)  (w' = w with every
//                                                              byte bit-reversed, so bit 7-p of
//                                                              byte j is haplotype 8j+p again)
// The two operands scale the same haplotype by 2^p and 2^(7-p): every (alt, alt) product is 2^7,
// the int32 accumulator holds n11 * 128 exactly, and the epilogue shifts it back.  The haplotype
// ORDER inside a chunk is permuted by this, identically for both operands -- a dot product does
// not care.  (All LOPs: the ALU pipe issues them at 64 lanes/clk/SM.)
template <bool COLUMN>
__device__ __forceinline__ uint4 widen4(uint32_t w, int e) {   // e = 0: p = 0..3, e = 1: p = 4..7
    uint4 o;
    if (!COLUMN) {
        o.x = w & (0x01010101u << (4 * e));     o.y = w & (0x01010101u << (4 * e + 1));
        o.z = w & (0x01010101u << (4 * e + 2)); o.w = w & (0x01010101u << (4 * e + 3));
    } else {
        o.x = w & (0x80808080u >> (4 * e));     o.y = w & (0x80808080u >> (4 * e + 1));
        o.z = w & (0x80808080u >> (4 * e + 2)); o.w = w & (0x80808080u >> (4 * e + 3));
    }
    return o;
}
// One variant-row of a chunk: 128 bits -> 8 x 16-byte units at the 128B-swizzled positions
// (unit j of row r lives at j ^ (r & 7)); a quarter-warp's STS.128 then covers all 32 banks.
template <bool COLUMN>
__device__ __forceinline__ void expand_row(uint8_t *row_base, int r7, const uint4 &b) {
    const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<uint4 *>(row_base + (((2 * q) ^ r7) << 4)) = widen4<COLUMN>(w[q], 0);
        *reinterpret_cast<uint4 *>(row_base + (((2 * q + 1) ^ r7) << 4)) = widen4<COLUMN>(w[q], 1);
    }
}
constexpr int ACC_SHIFT = 7;   // 2^p * 2^(7-p) = 2^7 per (alt, alt) haplotype

// ------------------------------------------------------------------------------------------ screening arithmetic
// The epilogue's fast path.  With complete biallelic data every quantity of calc_ld.py:33-90 is a
// ratio of small integers.  In the minor-allele encoding (a' = min(n1a, N - n1a) <= N/2, likewise c',
// n11' = the accumulator's count; relabelling an allele leaves r2 and |D'| unchanged):
//     Dn = n11'*N - a'*c'                     |d| = |Dn| / N^2                    (calc_ld.py:50)
//     m  = Dn > 0 ? N*min(a', c') - a'*c'     D'  = |Dn| / m                      (calc_ld.py:63-76)
//                 : a'*c'                          [min(a'c', a0'c0') = a'c' because a0' >= a', c0' >= c']
//     den = (a'*a0') * (c'*c0')               r2  = Dn^2 / den                    (calc_ld.py:86-88)
// a'*c' <= N^2/4, n11'*N <= N^2/2, a'*a0' <= N^2/4: for N <= 5792 every one of these integers is below 2^24 and the
// single-precision FMAs that form them are EXACT; up to N = 8192 Dn and the positive bound may be rounded once
// (an FMA rounds the exact value), which the error budget below already carries.  x = value * 10^4:
//     fD = fma(n11', N, -a'c')                      exact, or 1 rounding (sign and zero test always exact)
//     x_dp = (fD * 1e4) * rcp(m)                    : fD 1 + mul 1 + m 1 + rcp.approx 2 + mul 1 = 6 u, u = 2^-24
//     x_r2 = ((fD * 1e4) * ra) * (fD * rc)          : (2 + 2 + 1) + (1 + 2 + 1) + 1 = 10 u, ra = rcp(a'a0'), rc = rcp(c'c0')
// The reference rounds ITS OWN fp64 chain, whose distance from the exact ratio is bounded by the
// cancellation in f11 - p1*p2: |x_ref - x_exact| <= 10^4 * 16 * 2^-53 * N^2 for r2 (half of that for
// D').  Whenever x is farther than both bounds together (+1e-5) from every k + 1/2, Python's
// round(x_ref, 4) and the rounding of x agree, and the packed word is final.  Otherwise -- and when
// Dn == 0 for two polymorphic variants, where only the reference's own rounding errors decide
// between int 0 and 0.0 -- `slow` is set and the pair is redone with the reference's operation
// sequence in fp64 on the TRUE counts (finalise_pair, slow_pairs_kernel): about 1% of the pairs.
// Monomorphic variants (den == 0) need no arithmetic: d is exactly 0 and the bound is exactly 0 in
// the reference as well (calc_ld.py:68-69, :89-90): both int-0 flags.
constexpr float SCREEN_U = 5.9604644775390625e-08f;      // 2^-24
constexpr float SCREEN_C_DP = 8.0f * SCREEN_U;            // 6 u proven + 2 u slack
constexpr float SCREEN_C_R2 = 12.0f * SCREEN_U;           // 10 u proven + 2 u slack
constexpr float ROUND_MAGIC = 12582912.0f;                // 1.5 * 2^23: x + MAGIC has ulp 1 for 0 <= x < 2^22
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// The (alt, alt) count of the TRUE alleles from the count n11p of the minor-allele operands (a_t, b_t: true alt counts):
// a flipped row contributes its ref haplotypes, so  |~a & b| = b_t - n11,  |a & ~b| = a_t - n11,  |~a & ~b| = N - a_t - b_t + n11.
__device__ __forceinline__ int32_t true_n11(int32_t n11p, int32_t a_t, int32_t b_t, int32_t N) {
    const bool fa = 2 * a_t > N, fb = 2 * b_t > N;
    return fa ? (fb ? n11p - N + a_t + b_t : b_t - n11p) : (fb ? a_t - n11p : n11p);
}
// ------------------------------------------------------------------------------------------ GEMM + epilogue
// One variant set of a launch: its place in the launch's row space and its outputs.
struct SetRec {
    int64_t base_row;            // first row of the set in bits / freq_rows (a multiple of 256)
    int64_t v;                   // variants
    int64_t out_off;             // packed index of the first pair of the call's row range: outputs are relative to it
    uint32_t *packed; int32_t *n11;
    uint64_t fix_tag;            // ORed into the out_index of the set's near-tie records (FIX_TAG_SHIFT)
    // the general route (variants with missing calls / haploid samples / other codes, ldx_common.cuh): the set's store and the
    // store rows of its matrix rows (rows == nullptr: store row = row0 + matrix row); gen == nullptr: the store has none
    const GenStore *gen; const int64_t *rows; int64_t row0;
};
struct MmaArgs {
    const uint4 *bits, *bits_rev; int32_t kc_count;
    const VarFreq *freq_rows; FinalCtx fc;
    const int4 *tiles; int32_t n_tiles;      // {row panel, column block (both in the launch's row space), set, -}
    int measure, has_thres, thres_e4;
    int32_t n_sel;               // N = selected haplotypes
    float lim_dp, lim_r2;        // 0.5 - guard band of the screening arithmetic at x = 0 (see fast_pair2)
    uint4 *slow; uint32_t *slow_count; uint32_t slow_cap;   // deferred pairs {row, col, n11', set} for slow_pairs_kernel
    // single-wave calls (at most one tile per CTA): no follow-up kernel -- every epilogue warp settles its own deferred
    // pairs and the last CTA publishes the completion record (counters = d_fix_count, see slow_pairs_kernel)
    int inline_settle; uint32_t *counters; volatile uint32_t *mailbox; uint32_t seq;
    uint32_t pool_cap;           // single-wave calls: capacity of the CTA-wide list (tests shrink it to force the overflow paths)
    int help;                    // single-wave calls: the wideners, idle once the K loop is done, take half of the epilogue
    FixupSink fix;
    int32_t *error_flag;
    unsigned long long *trace;   // optional [512]: globaltimer stamps of CTA 0 (diagnostics)
    int dbg;                     // diagnostics build only (LDX_DEBUG_MMA; results invalid): 1 skip row-operand widening,
                                 // 2 skip column-operand widening, 4 skip the whole epilogue, 8 skip the MMAs, 16 no result stores,
                                 // 32 nobody is deferred, 64 no screening arithmetic (16 + 32 + 64: what is left is TMEM loads and set-up)
    // direct mode: the planes of the store itself through a TMA tensor map; matrix row r = store row row0 + r
    const uint4 *mask; int32_t row0;
    alignas(64) CUtensorMap tmap;
    SetRec set[MMA_MAX_SETS];    // single-wave kernels work on set[0]
};

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory"); return t; }
// diagnostics: per-CTA life-cycle stamps at trace[512 + 8 * cta + k] (k: 0 entry, 1 prologue done, 2 first accumulator ready,
// 3 first epilogue done, 4 all roles done, 5 SM id, 6/7 settlement barrier reached / passed; trace[2048 + 2 * cta]: settled)
#define LDX_CTA_STAMP(k) do { if (TRACE && A.trace && blockIdx.x < 192) A.trace[512 + 8 * blockIdx.x + (k)] = gtime(); } while (0)

// Measured alternatives to this split (A/B runs of library builds on one box, tools/gpu_call_ab.sh; 32,768 variants, this
// build 1.82-1.85 ms): three widener teams + twelve epilogue warps at 96 registers 2.23 ms (three teams no longer keep the
// tensor pipe fed, although ncu shows the sixteen wideners waiting for more than half of their time and the eight epilogue
// warps busy for 95% of theirs, profiles/r02_ncu_source_pair_v16384.txt); deferred-pair entries that carry the true counts
// (no frequency loads in the settlement) 1.98 ms; the epilogue as a loop over 16 x 32 work units with the row counts
// prefetched 1.89-2.01 ms.  The epilogue loop is latency-bound and at the edge of the instruction cache: small changes in
// code layout move it by 5-10%.
constexpr int N_WIDEN_WARPS = 16, N_EPI_WARPS = 8;  // wideners: four teams of 4 warps, team k takes pipeline stages g = k mod 4
constexpr int WIDEN_TEAMS = 4, TEAM_WARPS = N_WIDEN_WARPS / WIDEN_TEAMS;
// Warp roles, aligned to warpgroups (4 warps) so that setmaxnreg can move registers between them:
//   warps 0-3   producer (0), MMA issuer (1), two spare warps      -> 40 registers
//   warps 4-19  wideners (latency-bound: many warps, few registers) -> 56 registers (64 in the single-wave kernel, where
//               they also run a share of the epilogue)
//   warps 20-27 epilogue (sixteen pairs in flight per lane)         -> 104 registers
// The pool is the CTA's own 896 x 72 registers: setmaxnreg.inc BLOCKS until enough have been released,
// so the budget must close: 128*(72-40) + 512*(72-64) = 8192 freed >= 256*(104-72) = 8192 needed (single-wave kernel).
constexpr int REGS_LAUNCH = 72, REGS_CTRL = 40, REGS_WIDEN_SINGLE = 64 /* 56 in the multi-wave kernel */, REGS_EPI = 104;
static_assert(128 * (REGS_LAUNCH - REGS_CTRL) + 32 * N_WIDEN_WARPS * (REGS_LAUNCH - REGS_WIDEN_SINGLE) >= 32 * N_EPI_WARPS * (REGS_EPI - REGS_LAUNCH), "setmaxnreg budget");
constexpr int FIRST_WIDEN_WARP = 4, FIRST_EPI_WARP = FIRST_WIDEN_WARP + N_WIDEN_WARPS;
constexpr int MMA_THREADS = 32 * (FIRST_EPI_WARP + N_EPI_WARPS);   // 896
static_assert(MMA_THREADS * REGS_LAUNCH <= 65536, "register file");
constexpr int SLOW_BUF = 64;                                           // deferred pairs buffered per epilogue warp
// Result words leave with streaming stores (st.global.cs: first to be evicted from L2): a 100,000-variant call writes 20 GB of
// them through the cache its 64 MB of operand bits live in.  Measured: 1.84 ms instead of 1.87 ms at 32,768 variants.
#define LDX_ST(p, v) __stcs((p), (v))
constexpr int EPI_PITCH = 16;                                          // words per parked row (rare path: bank conflicts do not matter)

// Per PAIR of neighbouring column variants (2j, 2j + 1), what the screening arithmetic of the epilogue needs, laid out as
// the packed f32x2 operands fast_pair2 uses: one 16-byte load per two pairs and no register shuffling.
struct __align__(16) ColPair {
    float n1f[2];    // c' = min(n1, N - n1): the minor-allele counts, exact
    float rc[2];     // rcp.approx(c' * (N - c')); +inf for a monomorphic variant
};
__device__ __forceinline__ float minor_f(int32_t n1t, int32_t Nn) { return __int2float_rn(min(n1t, Nn - n1t)); }
__device__ __forceinline__ float rcp_nn(int32_t n1t, int32_t Nn) {         // 1 / (c' * c0'): the product is exact (<= 2^24)
    const int32_t c = min(n1t, Nn - n1t);
    // a variant of the general route (coded, negative count): NaN, so that the screen of every pair with it fails and the pair
    // is deferred to settle_slow_pair, which takes the general route
    return n1t < 0 ? __int_as_float(0x7fc00000) : rcp_approx(__int2float_rn(c * (Nn - c)));
}
// The same per row variant of an epilogue lane, packed once per pass.
struct RowP { uint64_t na2, ra2; float n1f; };    // {-a', -a'}, {ra, ra}, a'
__device__ __forceinline__ RowP make_rowp(int32_t n1t, int32_t Nn) {
    RowP p;
    p.n1f = minor_f(n1t, Nn);
    const float ra = rcp_nn(n1t, Nn);
    asm("mov.b64 %0, {%1, %1};" : "=l"(p.na2) : "f"(-p.n1f));
    asm("mov.b64 %0, {%1, %1};" : "=l"(p.ra2) : "f"(ra));
    return p;
}

// PAIR: two CTAs of a cluster work on one 256 x N tile with tcgen05.mma.cta_group::2; each widens its own 128
// row variants (TMEM) and HALF of the column variants (N/2 rows of the shared-memory operand, which the
// hardware reads from both CTAs): 128 + N/2 instead of 128 + N variant-rows per 128 x N results.
template <int N, bool PAIR = false> struct MmaCfg {
    static constexpr int NB = PAIR ? N / 2 : N;               // column variants widened by this CTA
    static constexpr int ROWS = MMA_M + NB;                   // variant-rows widened per chunk
    static constexpr int CH = 2;                              // 128-haplotype chunks per pipeline stage
    static constexpr int B_ROWS = NB < 128 ? NB : 128;        // column variants per 128-row bit block
    static constexpr int B_PARTS = (NB + 127) / 128;
    static constexpr int OP_STAGES = WIDEN_TEAMS;             // widened operand stages (A in TMEM, B in smem): one per team
    static constexpr int OP_BYTES = NB * KCHUNK * CH;         // B tile of one stage: CH swizzle atoms side by side
    static constexpr int BIT_STAGES = N <= 128 ? 8 : 4;       // bit blocks in flight from L2 (latency: deep)
    static constexpr int BIT_BYTES = ROWS * 16 * CH;
    static constexpr int A_COLS = CH * KCHUNK / 4;            // TMEM columns of one A stage (64)
    static constexpr int ACC_BUFS = N <= 128 ? 2 : 1;         // accumulators (512 TMEM columns in total)
    static constexpr int TMEM_A0 = ACC_BUFS * N;              // first A column
    static constexpr int TMEM_COLS = 512;
    static_assert(TMEM_A0 + OP_STAGES * A_COLS <= 512, "TMEM budget");
    static constexpr int N_BARS = 2 * OP_STAGES + 2 * BIT_STAGES + 4;
    static constexpr int EPI_BYTES = N_EPI_WARPS * (32 * EPI_PITCH * 4 + SLOW_BUF * 16);   // staging + deferred-pair buffer per warp
    static constexpr size_t SMEM = 1024 /*align slack*/ + (size_t)OP_STAGES * OP_BYTES + (size_t)BIT_STAGES * BIT_BYTES +
                                   N_EPI_WARPS * (N / 4) * sizeof(ColPair) + EPI_BYTES + N_BARS * 8 + 64;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert(!PAIR || (N == 128 && CH == 2), "the pair kernel splits the column widening as 64 rows x 2 chunks");
};

// ---- two pairs at a time with the packed single-precision instructions of sm_100 (FMUL2 / FADD2 / FFMA2:
// one issue slot for two results; IEEE round-to-nearest per half, so the error analysis above is unchanged).
// The epilogue is bound by instruction issue, which it shares with the wideners.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// What fast_pair2 needs about the call, packed for the f32x2 instructions once per warp.
struct ScreenK {
    uint64_t Ns2;        // {N / 128, N / 128}: the accumulator holds n11' * 128 (exactly representable: a power-of-two scale)
    uint64_t N2;         // {N, N}
    float lim_dp, lim_r2;
    uint32_t m_shift, thres;
};
__device__ __forceinline__ ScreenK make_screen(const MmaArgs &A) {
    ScreenK K;
    K.m_shift = A.measure == LDX_MEASURE_R2 ? 0u : (uint32_t)LDX_DP_SHIFT;
    K.thres = A.has_thres ? (uint32_t)A.thres_e4 : 0u;      // rounded values are >= 0: 0 flags nothing
    const float nf = __int2float_rn(A.n_sel);
    K.Ns2 = pk2(nf * 0.0078125f, nf * 0.0078125f); K.N2 = pk2(nf, nf);
    K.lim_dp = A.lim_dp; K.lim_r2 = A.lim_r2;
    return K;
}

// Pairs (row, column 2j) and (row, column 2j + 1): the screening arithmetic described above.  The bound m carries the
// sign of Dn (m = -a'c' for Dn <= 0), so that Dn / m is |D'| without an absolute value -- the packed instructions have no
// operand modifiers.  A monomorphic variant shows as ra * rc = +inf (rcp.approx(0) = +inf; the product of two finite
// reciprocals is at most 1).
template <bool THRES>
__device__ __forceinline__ void fast_pair2(uint32_t acc0, uint32_t acc1, const ScreenK &K, const RowP &R, const ColPair &C,
                                           uint32_t &w0, uint32_t &w1, bool &slow0, bool &slow1) {
    const uint64_t cn = pk2(C.n1f[0], C.n1f[1]);
    const uint64_t accf = pk2(__int2float_rn((int32_t)acc0), __int2float_rn((int32_t)acc1));   // n11' * 128 <= 2^19: exact
    const uint64_t nP = mul2(R.na2, cn);                                            // -a'c': exact (<= 2^24)
    const uint64_t fD = fma2(accf, K.Ns2, nP);                                     // Dn = n11'*N - a'c'
    const uint64_t mp = fma2(K.N2, pk2(fminf(R.n1f, C.n1f[0]), fminf(R.n1f, C.n1f[1])), nP);   // N*min(a', c') - a'c' > 0
    float d0, d1, mp0, mp1, np0, np1;
    upk2(fD, d0, d1); upk2(mp, mp0, mp1); upk2(nP, np0, np1);
    const uint64_t R1 = pk2(rcp_approx(d0 > 0.0f ? mp0 : np0), rcp_approx(d1 > 0.0f ? mp1 : np1));
    const uint64_t rinv = mul2(R.ra2, pk2(C.rc[0], C.rc[1]));                     // 1 / den; +inf <=> a monomorphic variant
    const uint64_t D4 = mul2(fD, pk2(1.0e4f, 1.0e4f));
    const uint64_t x_dp = mul2(D4, R1);
    const uint64_t x_r2 = mul2(mul2(D4, fD), rinv);
    const uint64_t magic = pk2(ROUND_MAGIC, ROUND_MAGIC), nmagic = pk2(-ROUND_MAGIC, -ROUND_MAGIC), neg1 = pk2(-1.0f, -1.0f);
    const uint64_t t_dp = add2(x_dp, magic), t_r2 = add2(x_r2, magic);
    const uint64_t f_dp = fma2(add2(t_dp, nmagic), neg1, x_dp);          // exact: x - nearest integer
    const uint64_t f_r2 = fma2(add2(t_r2, nmagic), neg1, x_r2);
    const uint64_t l_dp = fma2(x_dp, pk2(-SCREEN_C_DP, -SCREEN_C_DP), pk2(K.lim_dp, K.lim_dp));   // guard band grows with x
    const uint64_t l_r2 = fma2(x_r2, pk2(-SCREEN_C_R2, -SCREEN_C_R2), pk2(K.lim_r2, K.lim_r2));
    float td0, td1, tr0, tr1, fd0, fd1, fr0, fr1, ld0, ld1, lr0, lr1, ri0, ri1;
    upk2(t_dp, td0, td1); upk2(t_r2, tr0, tr1); upk2(f_dp, fd0, fd1); upk2(f_r2, fr0, fr1); upk2(l_dp, ld0, ld1); upk2(l_r2, lr0, lr1);
    upk2(rinv, ri0, ri1);
    const float inf = __int_as_float(0x7f800000);
    const bool mono0 = ri0 == inf, mono1 = ri1 == inf;
    // (<=) with everything else falling through: a NaN from an unforeseen input must take the exact path, never pass
    const bool fine0 = (fabsf(fd0) <= ld0) && (fabsf(fr0) <= lr0) && (d0 != 0.0f);
    const bool fine1 = (fabsf(fd1) <= ld1) && (fabsf(fr1) <= lr1) && (d1 != 0.0f);
    slow0 = !(fine0 || mono0);
    slow1 = !(fine1 || mono1);
    // x + MAGIC has the rounded integer k (< 2^14) in its low mantissa bits: word = k_dp << 16 | k_r2 is one byte permute
    uint32_t a = __byte_perm((uint32_t)__float_as_int(tr0), (uint32_t)__float_as_int(td0), 0x5410);
    uint32_t b = __byte_perm((uint32_t)__float_as_int(tr1), (uint32_t)__float_as_int(td1), 0x5410);
    a = mono0 ? (LDX_DP_INT0 | LDX_R2_INT0) : a;
    b = mono1 ? (LDX_DP_INT0 | LDX_R2_INT0) : b;
    if (THRES) {
        a |= (((a >> K.m_shift) & LDX_R2_MASK) < K.thres) ? LDX_BELOW_THRES : 0u;
        b |= (((b >> K.m_shift) & LDX_R2_MASK) < K.thres) ? LDX_BELOW_THRES : 0u;
    }
    w0 = a; w1 = b;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// One deferred pair {row, col, n11', set} (row / col relative to the set) redone with the reference's own operation
// sequence (finalise_pair) on the true counts.
__device__ __forceinline__ void settle_slow_pair(const uint4 e, const MmaArgs &A, uint32_t m_shift, uint32_t thres) {
    const SetRec &S = A.set[e.w];
    const int64_t r = e.x, col = e.y;
    const VarFreq fa = A.freq_rows[S.base_row + r], fb = A.freq_rows[S.base_row + col];
    const uint64_t idx = (uint64_t)(r * (r - 1) / 2 + col - S.out_off);
    if ((fa.n1 | fb.n1) < 0 && S.gen) {                                // a variant of the general route: explicit counts from the store's planes
        const int64_t sa = S.rows ? S.rows[r] : S.row0 + r, sb = S.rows ? S.rows[col] : S.row0 + col;
        const GenCounts c = general_pair_counts(*S.gen, sa, fa.n1, sb, fb.n1);
        uint32_t w = finalise_general(c).packed;
        w |= (((w >> m_shift) & LDX_R2_MASK) < thres) ? LDX_BELOW_THRES : 0u;
        if (w & LDX_R2_NEARTIE) {
            FixupSink fx = A.fix;
            fx.tag = S.fix_tag;
            fixup_append_general(fx, idx, c, w);
        }
        S.packed[idx] = w;
        if (S.n11) S.n11[idx] = c.n11;
        return;
    }
    const int32_t n11 = true_n11((int32_t)e.z, fa.n1, fb.n1, A.n_sel);
    const PairFinal f = finalise_pair(n11, fa, fb, A.fc);              // var_1 = row, var_2 = column
    uint32_t w = f.packed;
    w |= (((w >> m_shift) & LDX_R2_MASK) < thres) ? LDX_BELOW_THRES : 0u;
    if (w & LDX_R2_NEARTIE) {
        FixupSink fx = A.fix;
        fx.tag = S.fix_tag;
        fixup_append(fx, idx, n11, fa.n1, fb.n1, w);
    }
    S.packed[idx] = w;
}

// Warp-collective: move a warp's buffered deferred pairs to the global list.  Returns the new count (0).
// Entries that do not fit the list any more (degenerate inputs with far more near-boundary pairs than
// the ~1% the list is sized for) are settled right here -- slower, never wrong.  Every buffered entry
// belongs to a chunk whose provisional words this warp has already stored (and __syncwarp orders the
// warp's stores), so the settled word is the one that stays.
template <bool SINGLE>
__device__ __forceinline__ uint32_t flush_slow(const MmaArgs &A, const uint4 *sbuf, uint32_t cnt, int lane, uint32_t m_shift, uint32_t thres) {
    if (cnt == 0) return 0;
    __syncwarp();
    if (SINGLE) {          // warp-uniform: one deferred pair per lane, right here
        for (uint32_t i = lane; i < cnt; i += 32) settle_slow_pair(sbuf[i], A, m_shift, thres);
        __syncwarp();
        return 0;
    }
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(A.slow_count, cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t i = lane; i < cnt; i += 32) {
        if (base + i < A.slow_cap) A.slow[base + i] = sbuf[i];
        else settle_slow_pair(sbuf[i], A, m_shift, thres);
    }
    __syncwarp();
    return 0;
}

// The pairs the epilogue's screening could not settle: one thread per pair, fully parallel.  The last
// block to finish publishes the call's completion record (near-tie count, error flag, sequence number)
// to the host mailbox.
__global__ void __launch_bounds__(256)
slow_pairs_kernel(const __grid_constant__ MmaArgs A) {
    pdl_launch_dependents();
    pdl_wait();                                                 // everything below depends on the all-pairs kernel
    uint32_t *counters = A.counters;
    const uint32_t total = counters[2];
    const uint32_t n = total < A.slow_cap ? total : A.slow_cap;    // the rest was settled in the epilogue (flush_slow)
    const uint32_t m_shift = A.measure == LDX_MEASURE_R2 ? 0u : (uint32_t)LDX_DP_SHIFT;
    const uint32_t thres = A.has_thres ? (uint32_t)A.thres_e4 : 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        settle_slow_pair(A.slow[i], A, m_shift, thres);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(&counters[3], 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            counters[3] = 0;
            counters[2] = 0;
            if (A.mailbox) publish_record(A.mailbox, A.seq, *(volatile uint32_t *)&counters[0], *(volatile uint32_t *)&counters[1]);
        }
    }
}

// SINGLE: the call is one wave of tiles (every CTA has exactly one): deferred pairs are settled in this kernel, the last CTA
// publishes the completion record, and the wideners share the epilogue.  The multi-wave instantiation carries none of that
// code: its hot loops are sensitive to every change in register allocation and code layout.
// DIRECT (single-wave, one CTA per 128 x 128 tile only): no gather kernel ran; the producer reads the store's planes through
// the tensor map in A.tmap and the wideners apply mask, minor-allele flip and the column operand's bit reversal.
template <int N, bool THRES, bool TRACE, bool PAIR, bool SINGLE, bool DIRECT>
__global__ void __launch_bounds__(MMA_THREADS, 1)
triangle_mma_kernel(const __grid_constant__ MmaArgs A) {
    using Cfg = MmaCfg<N, PAIR>;
    static_assert(!DIRECT || (SINGLE && !PAIR && N == 128), "direct mode: one CTA per 128 x 128 tile of a one-wave call");
    // work units: a CTA (tile = 128 x N), or a CTA pair (tile = 256 x N, this CTA owns rows 128 * rank ...)
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t *smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);      // swizzle-128B needs 1 KB alignment
    uint8_t *op_s = smem;                                                        // [OP_STAGES][N][128 B] column operand
    uint8_t *bit_s = smem + Cfg::OP_STAGES * Cfg::OP_BYTES;                      // [BIT_STAGES][ROWS][16 B]
    ColPair *col_s = reinterpret_cast<ColPair *>(bit_s + Cfg::BIT_STAGES * Cfg::BIT_BYTES);   // [N_EPI_WARPS][N / 4]: private to each epilogue warp
    uint32_t *epi_s = reinterpret_cast<uint32_t *>(col_s + N_EPI_WARPS * (N / 4));  // [N_EPI_WARPS][32][EPI_PITCH]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(epi_s) + Cfg::EPI_BYTES);
    const uint32_t op_full = smem_u32(bars), op_empty = op_full + 8 * Cfg::OP_STAGES;
    const uint32_t bit_full = op_empty + 8 * Cfg::OP_STAGES, bit_empty = bit_full + 8 * Cfg::BIT_STAGES;
    const uint32_t tmem_full = bit_empty + 8 * Cfg::BIT_STAGES, tmem_empty = tmem_full + 16;
    uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(bars + Cfg::N_BARS);
    volatile int *abort_s = reinterpret_cast<volatile int *>(tmem_ptr_s + 1);
    // single-wave kernel: CTA-wide list of deferred pairs.  It lives in the operand stages, which are free once the
    // tile's accumulator is complete (every writer has waited for tmem_full first).
    uint32_t *pool_cnt = tmem_ptr_s + 2;
    uint4 *pool = reinterpret_cast<uint4 *>(op_s);
    const uint32_t POOL_CAP = min((uint32_t)(Cfg::OP_STAGES * Cfg::OP_BYTES / 16), A.pool_cap);
    constexpr int POOL_THREADS = 32 * (N_WIDEN_WARPS + N_EPI_WARPS);      // wideners + epilogue warps settle the list together

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kc_count = A.kc_count;
    if (TRACE && A.trace && blockIdx.x == 0 && threadIdx.x == 0) A.trace[7] = gtime();    // kernel entry
    if (threadIdx.x == 0) {
        LDX_CTA_STAMP(0);
        if (TRACE && A.trace && blockIdx.x < 192) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); A.trace[512 + 8 * blockIdx.x + 5] = smid; }
    }

    if (threadIdx.x == 0) {
        // pair: the leader's op_full / tmem_empty also count the peer's wideners / epilogue warps (remote arrives)
        for (int s = 0; s < Cfg::OP_STAGES; ++s) { mbar_init(op_full + 8 * s, (PAIR ? 2 : 1) * TEAM_WARPS); mbar_init(op_empty + 8 * s, 1); }
        for (int s = 0; s < Cfg::BIT_STAGES; ++s) { mbar_init(bit_full + 8 * s, 1); mbar_init(bit_empty + 8 * s, TEAM_WARPS); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + 8 * s, 1); mbar_init(tmem_empty + 8 * s, (PAIR ? 2 : 1) * N_EPI_WARPS); }
        *abort_s = 0;
        if (SINGLE) *pool_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (DIRECT) asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&A.tmap)) : "memory");
    }
    if (warp == 1) {   // one warp allocates TMEM (power-of-two columns >= 32) and later frees it
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(tmem_ptr_s)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(tmem_ptr_s)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer's barriers are initialised before anyone arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    // the tile list is uploaded by the host (and cached), not written by the preceding kernel: the first tile's
    // coordinates can be fetched while that kernel is still running
    const int4 first_tile = SINGLE && unit < A.n_tiles ? A.tiles[unit] : make_int4(0, 0, 0, 0);
    if (SINGLE && threadIdx.x == 32) {
        // the near-tie list is touched by a handful of pairs per call: without this the first of them pays a DRAM
        // round trip for the counter and one for the record on the kernel's critical path
        asm volatile("prefetch.global.L2 [%0];" :: "l"(A.fix.count));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(A.fix.recs));
    }
    pdl_launch_dependents();                                            // deferred-pairs kernel: launch latency hidden behind this grid
    pdl_wait();                                                         // bit panels / frequencies come from the gather kernel
    if (TRACE && A.trace && blockIdx.x == 0 && threadIdx.x == 0) A.trace[0] = gtime();   // prologue done
    if (threadIdx.x == 0) LDX_CTA_STAMP(1);

    const int ks_count = kc_count / Cfg::CH;                  // pipeline stages per tile (kc_count is even)
    if (warp == 0) {
        // ===== producer: bit blocks L2 -> shared memory ring.  One stage = CH chunks of the row panel
        // (CH x 2 KB, contiguous in global memory) + the same for the tile's column variants.
        // The whole warp runs the loop; one elected lane issues the copies.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        uint32_t g = 0;
        for (int t = unit; t < A.n_tiles; t += n_units) {
            const int4 tile = SINGLE ? first_tile : A.tiles[t];
            const int64_t c0 = (int64_t)tile.y * N + (PAIR ? rank * Cfg::NB : 0);      // pair: this CTA's half of the columns
            const uint4 *a_src = A.bits + (int64_t)(PAIR ? tile.x * 2 + rank : tile.x) * kc_count * 128;
            for (int ks = 0; ks < ks_count; ++ks, ++g) {
                const uint32_t s = g % Cfg::BIT_STAGES, it = g / Cfg::BIT_STAGES;
                if (!mbar_wait(bit_empty + 8 * s, (it & 1) ^ 1, abort_s, A.error_flag)) goto done;
                const uint32_t bar = bit_full + 8 * s;
                const uint32_t dst = smem_u32(bit_s + s * Cfg::BIT_BYTES);
                const int kc = ks * Cfg::CH;
                if (elect_one()) {
                    if (TRACE && A.trace && blockIdx.x == 0 && g < 48) A.trace[128 + g] = gtime();
                    mbar_arrive_expect_tx(bar, Cfg::BIT_BYTES);
                    if (DIRECT) {
                        // [128 rows][32 B] of the row variants, then of the column variants, straight from the store
                        tma_load_2d(dst, &A.tmap, kc * 4, A.row0 + tile.x * MMA_M, bar);
                        tma_load_2d(dst + MMA_M * 16 * Cfg::CH, &A.tmap, kc * 4, A.row0 + (int32_t)c0, bar);
                    } else {
                    bulk_g2s(dst, a_src + (int64_t)kc * 128, Cfg::CH * MMA_M * 16, bar);       // [CH][128] rows
#pragma unroll
                    for (int part = 0; part < Cfg::B_PARTS; ++part) {
                        const int64_t crow = c0 + part * 128;             // first column variant of this part
                        const uint4 *b_src = A.bits_rev + ((crow >> 7) * kc_count + kc) * 128 + (crow & 127);
                        const uint32_t b_dst = dst + (Cfg::CH * MMA_M + part * Cfg::CH * Cfg::B_ROWS) * 16;   // [part][CH][B_ROWS]
                        if (Cfg::B_ROWS == 128) {
                            bulk_g2s(b_dst, b_src, Cfg::CH * 128 * 16, bar);              // whole blocks: contiguous
                        } else {
#pragma unroll
                            for (int c = 0; c < Cfg::CH; ++c) bulk_g2s(b_dst + c * Cfg::B_ROWS * 16, b_src + c * 128, Cfg::B_ROWS * 16, bar);
                        }
                    }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        constexpr uint32_t idesc = make_idesc(PAIR ? 2 * MMA_M : MMA_M, N);
        uint32_t g = 0, tl = 0;
        if (PAIR && rank != 0) goto done;      // only the leader of a pair issues (and waits for both CTAs' operands)
        for (int t = unit; t < A.n_tiles; t += n_units, ++tl) {
            const uint32_t buf = tl % Cfg::ACC_BUFS;
            if (!mbar_wait<PAIR>(tmem_empty + 8 * buf, ((tl / Cfg::ACC_BUFS) & 1) ^ 1, abort_s, A.error_flag)) goto done;   // epilogue(s) drained it
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + buf * N;
            for (int ks = 0; ks < ks_count; ++ks, ++g) {
                const uint32_t s = g % Cfg::OP_STAGES, it = g / Cfg::OP_STAGES;
                if (!mbar_wait<PAIR>(op_full + 8 * s, it & 1, abort_s, A.error_flag)) goto done;
                tc_fence_after();
                const uint64_t db = make_smem_desc(smem_u32(op_s + s * Cfg::OP_BYTES));
                const uint32_t ta = tmem_base + Cfg::TMEM_A0 + s * Cfg::A_COLS;
                if (elect_one()) {
                    if (TRACE && A.trace && blockIdx.x == 0 && g < 48) A.trace[64 + g] = gtime();
                    if (!(TRACE && (A.dbg & 8)))
#pragma unroll
                    for (int k = 0; k < Cfg::CH * KCHUNK / MMA_K; ++k) {   // K = 32: 8 TMEM columns of A; B: atom k/4 (NB*128 B apart), +32 B per step inside
                        const uint64_t dbk = db + (uint64_t)((k >> 2) * (Cfg::NB * KCHUNK >> 4) + (k & 3) * 2);
                        if (PAIR) umma_i8_ts_pair(tmem_acc, ta + 8 * k, dbk, idesc, (uint32_t)((ks | k) != 0));
                        else umma_i8_ts(tmem_acc, ta + 8 * k, dbk, idesc, (uint32_t)((ks | k) != 0));
                    }
                    if (PAIR) {
                        umma_commit_pair(op_empty + 8 * s);       // both CTAs' wideners get their stage back
                        if (ks == ks_count - 1) umma_commit_pair(tmem_full + 8 * buf);
                    } else {
                        umma_commit(op_empty + 8 * s);            // stage reusable once these MMAs have read it
                        if (ks == ks_count - 1) umma_commit(tmem_full + 8 * buf);   // accumulator complete
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp < FIRST_WIDEN_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // spare warps: hand their registers over and wait
    } else if (warp < FIRST_EPI_WARP) {
        if (SINGLE) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ===== wideners: bit rows -> operand bytes.  Four teams of four warps take pipeline stages in turn,
        // so that the fixed latencies of a stage (two mbarrier waits, the TMEM store, the proxy fence)
        // of one team overlap the work of the others.  Thread wt of a team widens row variant wt of the
        // stage into TENSOR MEMORY (lane = row; warp%4 is the TMEM quadrant the warp may write) and
        // column variant(s) wt (+128) into the 128B-swizzled shared-memory tile.
        const int team = (warp - FIRST_WIDEN_WARP) / TEAM_WARPS;
        const int wt = ((warp - FIRST_WIDEN_WARP) & 3) * 32 + lane;     // 0..127
        const uint32_t tmem_lane = (uint32_t)(((warp - FIRST_WIDEN_WARP) & 3) * 32) << 16;
        // the wideners do not care which tile a stage belongs to: team k takes stages g = k, k + TEAMS, ...
        // of this CTA's stream, bit stage g % BIT_STAGES -> operand stage g % OP_STAGES (= its own: k)
        const uint32_t my_tiles = (uint32_t)(A.n_tiles - unit + n_units - 1) / (uint32_t)n_units;
        const uint32_t op_full_leader = PAIR ? mapa_u32(op_full, 0) : 0u;     // pair: everybody reports to rank 0's barrier
        const uint32_t g_end = my_tiles * (uint32_t)ks_count;
        static_assert(Cfg::OP_STAGES == WIDEN_TEAMS, "one operand stage per team");
        bool dead = false;            // single-wave kernel: an aborted warp still meets the others at the settlement barrier
        // direct mode: this thread's row and column variant enter complemented when their alt allele is the major one
        uint32_t flip_a = 0, flip_b = 0;
        if (DIRECT && my_tiles) {
            flip_a = 2 * A.freq_rows[(int64_t)first_tile.x * MMA_M + wt].n1 > A.n_sel ? 0xffffffffu : 0u;
            flip_b = 2 * A.freq_rows[(int64_t)first_tile.y * N + wt].n1 > A.n_sel ? 0xffffffffu : 0u;
        }
        {
            for (uint32_t g = (uint32_t)team; g < g_end; g += WIDEN_TEAMS) {
                const uint32_t sb = g % Cfg::BIT_STAGES, itb = g / Cfg::BIT_STAGES;
                const uint32_t so = (uint32_t)team, ito = g / Cfg::OP_STAGES;
                uint4 mk[Cfg::CH];
                if (DIRECT) {                   // the mask words of this stage's haplotypes (one tile: stage g is chunk pair g)
#pragma unroll
                    for (int c = 0; c < Cfg::CH; ++c) mk[c] = __ldg(A.mask + g * Cfg::CH + c);
                }
                if (!mbar_wait(bit_full + 8 * sb, itb & 1, abort_s, A.error_flag)) { if (SINGLE) { dead = true; break; } goto done; }
                if (TRACE && A.trace && blockIdx.x == 0 && g < 48 && wt == 0) A.trace[192 + g] = gtime();
                const uint4 *bsrc = reinterpret_cast<const uint4 *>(bit_s + sb * Cfg::BIT_BYTES);
                uint4 ba[Cfg::CH], bb[Cfg::CH];
#pragma unroll
                for (int c = 0; c < Cfg::CH; ++c) {
                    if (DIRECT) {               // [128 rows][CH x 16 B] per operand, as the tensor map's box lands
                        const uint4 x = bsrc[wt * Cfg::CH + c], y = bsrc[(MMA_M + wt) * Cfg::CH + c];
                        ba[c] = make_uint4((x.x ^ flip_a) & mk[c].x, (x.y ^ flip_a) & mk[c].y, (x.z ^ flip_a) & mk[c].z, (x.w ^ flip_a) & mk[c].w);
                        bb[c] = make_uint4(brev_bytes((y.x ^ flip_b) & mk[c].x), brev_bytes((y.y ^ flip_b) & mk[c].y),
                                           brev_bytes((y.z ^ flip_b) & mk[c].z), brev_bytes((y.w ^ flip_b) & mk[c].w));
                    } else ba[c] = bsrc[c * MMA_M + wt];
                }
                if (!mbar_wait(op_empty + 8 * so, (ito & 1) ^ 1, abort_s, A.error_flag)) { if (SINGLE) { dead = true; break; } goto done; }
                tc_fence_after();
                if (TRACE && A.trace && blockIdx.x == 0 && g < 48 && wt == 0) A.trace[256 + g] = gtime();
                if (!(TRACE && (A.dbg & 1)))
#pragma unroll
                for (int c = 0; c < Cfg::CH; ++c) {
                    const uint32_t w[4] = {ba[c].x, ba[c].y, ba[c].z, ba[c].w};
                    uint32_t v[32];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int p = 0; p < 8; ++p) v[q * 8 + p] = w[q] & (0x01010101u << p);   // same byte order as expand_row
                    tmem_st32(tmem_base + Cfg::TMEM_A0 + so * Cfg::A_COLS + c * (KCHUNK / 4) + tmem_lane, v);
                }
                uint8_t *ops = op_s + so * Cfg::OP_BYTES;
                if (DIRECT) {
                    if (!(TRACE && (A.dbg & 2))) {
#pragma unroll
                        for (int c = 0; c < Cfg::CH; ++c) expand_row<true>(ops + c * (N * KCHUNK) + wt * KCHUNK, wt & 7, bb[c]);
                    }
                } else if (PAIR) {
                    // 64 column variants x 2 chunks over the team's 128 threads: one row-chunk each
                    const int brow = wt & 63, c = wt >> 6;
                    if (!(TRACE && (A.dbg & 2)))
                        expand_row<true>(ops + c * (Cfg::NB * KCHUNK) + brow * KCHUNK, brow & 7, bsrc[Cfg::CH * MMA_M + c * Cfg::B_ROWS + brow]);
                } else
                if (wt < Cfg::B_ROWS && !(TRACE && (A.dbg & 2))) {
#pragma unroll
                    for (int c = 0; c < Cfg::CH; ++c)
#pragma unroll
                        for (int i = 0; i < Cfg::B_PARTS; ++i)
                            expand_row<true>(ops + c * (N * KCHUNK) + (wt + i * 128) * KCHUNK, wt & 7,
                                             bsrc[Cfg::CH * MMA_M + (i * Cfg::CH + c) * Cfg::B_ROWS + wt]);
                }
                if (TRACE && A.trace && blockIdx.x == 0 && g < 48 && wt == 0) A.trace[320 + g] = gtime();
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to tcgen05
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_remote(op_full_leader + 8 * so); else mbar_arrive(op_full + 8 * so);
                    mbar_arrive(bit_empty + 8 * sb);
                }
                if (TRACE && A.trace && blockIdx.x == 0 && g < 48 && wt == 0) A.trace[8 + g] = gtime();
            }
        }
        if (SINGLE && N == 128 && !PAIR && A.help && my_tiles && !dead) {
            // ===== single wave: this CTA's only tile has no successor to widen for.  Widener warp (quadrant, team)
            // takes the epilogue of rows 16..31 of its TMEM quadrant x columns 32*team..+31 -- the same arithmetic
            // as the epilogue warps below (which then do rows 0..15 only), with the column records formed on the
            // fly.  A pair the screen cannot settle leaves its count in the result word and is redone by the lane
            // that found it (finalise_pair); a lane's own store is visible to its own later load.
            const SetRec &S = A.set[0];
            const int quad = warp & 3, lq = lane >> 2, lr = lane & 3;
            const ScreenK K = make_screen(A);
            const int32_t Nn = A.n_sel;
            const int64_t r0 = (int64_t)first_tile.x * MMA_M, c0 = (int64_t)first_tile.y * N;
            const int64_t rmin = r0 + quad * 32 + 16;
            const int cb = team * 32;
            const int64_t cg0 = c0 + cb;
            const bool mine = N == 128 && rmin < S.v && cg0 < rmin + 15 && cg0 < S.v && !(TRACE && (A.dbg & 4));   // warp-uniform
            const int64_t ra = rmin + lq, rb = ra + 8;
            int32_t n1a = 0, n1b = 0, cn1[8];
            if (mine) {                             // the counts this lane needs, fetched while the tensor pipe finishes
                n1a = A.freq_rows[ra].n1; n1b = A.freq_rows[rb].n1;
#pragma unroll
                for (int k = 0; k < 4; ++k) { cn1[2 * k] = A.freq_rows[cg0 + 8 * k + 2 * lr].n1; cn1[2 * k + 1] = A.freq_rows[cg0 + 8 * k + 2 * lr + 1].n1; }
            }
            dead = !mbar_wait(tmem_full, 0, abort_s, A.error_flag, 64);
            tc_fence_after();
            if (mine && !dead) {
                const RowP Ra = make_rowp(n1a, Nn), Rb = make_rowp(n1b, Nn);
                uint32_t *pa = S.packed + (ra * (ra - 1) / 2 - S.out_off + cg0 + 2 * lr);
                uint32_t *pb = S.packed + (rb * (rb - 1) / 2 - S.out_off + cg0 + 2 * lr);
                int32_t *qa = S.n11 ? S.n11 + (ra * (ra - 1) / 2 - S.out_off + cg0 + 2 * lr) : nullptr;
                int32_t *qb = S.n11 ? S.n11 + (rb * (rb - 1) / 2 - S.out_off + cg0 + 2 * lr) : nullptr;
                uint32_t acc[16];
                tmem_ld16x256(tmem_base + ((uint32_t)(quad * 32 + 16) << 16) + (uint32_t)cb, acc);
                uint32_t slow = 0;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const int k = i >> 2, g = (i >> 1) & 1;
                    const int64_t col = cg0 + 8 * k + 2 * lr;
                    ColPair cp;
                    cp.n1f[0] = minor_f(cn1[2 * k], Nn); cp.n1f[1] = minor_f(cn1[2 * k + 1], Nn);
                    cp.rc[0] = rcp_nn(cn1[2 * k], Nn); cp.rc[1] = rcp_nn(cn1[2 * k + 1], Nn);
                    uint32_t w0, w1;
                    bool s0, s1;
                    if (TRACE && (A.dbg & 64)) { w0 = acc[i] + cp.rc[0]; w1 = acc[i + 1]; s0 = s1 = false; }
                    else
                    fast_pair2<THRES>(acc[i], acc[i + 1], K, g ? Rb : Ra, cp, w0, w1, s0, s1);
                    if (TRACE && (A.dbg & 32)) s0 = s1 = false;
                    const int64_t row = g ? rb : ra;
                    const bool v0 = row < S.v && col < row && !(TRACE && (A.dbg & 16) && w0 != 0x12345678u), v1 = row < S.v && col + 1 < row && !(TRACE && (A.dbg & 16) && w1 != 0x12345678u);
                    if (v0) {
                        LDX_ST((g ? pb : pa) + 8 * k, s0 ? (acc[i] >> ACC_SHIFT) : w0);
                        if (qa) (g ? qb : qa)[8 * k] = true_n11((int32_t)(acc[i] >> ACC_SHIFT), g ? n1b : n1a, cn1[2 * k], Nn);
                        if (s0) {                  // deferred: onto the CTA's list (a full list: redone by this lane below)
                            const uint32_t slot = atomicAdd(pool_cnt, 1u);
                            if (slot < POOL_CAP) pool[slot] = make_uint4((uint32_t)row, (uint32_t)col, acc[i] >> ACC_SHIFT, 0u);
                            else slow |= 1u << i;
                        }
                    }
                    if (v1) {
                        LDX_ST((g ? pb : pa) + 8 * k + 1, s1 ? (acc[i + 1] >> ACC_SHIFT) : w1);
                        if (qa) (g ? qb : qa)[8 * k + 1] = true_n11((int32_t)(acc[i + 1] >> ACC_SHIFT), g ? n1b : n1a, cn1[2 * k + 1], Nn);
                        if (s1) {
                            const uint32_t slot = atomicAdd(pool_cnt, 1u);
                            if (slot < POOL_CAP) pool[slot] = make_uint4((uint32_t)row, (uint32_t)(col + 1), acc[i + 1] >> ACC_SHIFT, 0u);
                            else slow |= 1u << (i + 1);
                        }
                    }
                }
                while (slow) {
                    const int i = __ffs(slow) - 1;
                    slow &= slow - 1;
                    const int k = i >> 2, e = i & 1;
                    const int64_t row = (i & 2) ? rb : ra, col = cg0 + 8 * k + 2 * lr + e;
                    const uint32_t n11p = ((i & 2) ? pb : pa)[8 * k + e];
                    settle_slow_pair(make_uint4((uint32_t)row, (uint32_t)col, n11p, 0u), A, K.m_shift, K.thres);
                }
            }
            tc_fence_before();
        }
        if (SINGLE) {
            // the CTA's deferred pairs (about 1% of the tile), one per thread over the wideners and the epilogue warps:
            // one pass of the fp64 chain.  bar.sync orders every warp's provisional stores before the settled words.
            if (warp == FIRST_WIDEN_WARP && lane == 0) LDX_CTA_STAMP(6);         // first widener warp at the settlement barrier
            asm volatile("bar.sync 1, %0;" :: "n"(POOL_THREADS) : "memory");
            if (warp == FIRST_WIDEN_WARP && lane == 0) LDX_CTA_STAMP(7);         // ... past it (all 24 warps arrived)
            if (warp == FIRST_WIDEN_WARP && lane == 0 && TRACE && A.trace && blockIdx.x < 192) A.trace[2048 + 2 * 192 + blockIdx.x] = *pool_cnt;
            if (!dead) {
                const uint32_t m_shift = A.measure == LDX_MEASURE_R2 ? 0u : (uint32_t)LDX_DP_SHIFT;
                const uint32_t thres = A.has_thres ? (uint32_t)A.thres_e4 : 0u;
                const uint32_t total = min(*pool_cnt, POOL_CAP);
                for (uint32_t j = threadIdx.x - 32 * FIRST_WIDEN_WARP; j < total; j += POOL_THREADS)
                    settle_slow_pair(pool[j], A, m_shift, thres);
            }
            if (warp == FIRST_WIDEN_WARP && lane == 0 && TRACE && A.trace && blockIdx.x < 192) A.trace[512 + 8 * 192 + 2 * blockIdx.x] = gtime();       // settled
            // no fence here: thread 0's fence after the CTA-wide barrier below is cumulative over everything the
            // barrier made it observe (the pattern of a grid-wide sync)
        }
    } else {
        // ===== epilogue.  Warp w may only touch TMEM lanes 32*(w%4) .. +31.  Accumulators are read with
        // the 16-lane x 256-bit shape: register i = 4k + 2g + e of lane l holds
        //     row  rmin + l/4 + 8g,   column  cb + 8k + 2(l%4) + e            (tools/probe/tmem_layout.cu)
        // so a lane works on TWO row variants and eight column variants per load, and the four lanes of
        // a row own 8 consecutive result words (one 32-byte sector): results go straight from registers
        // to global memory, no transposition through shared memory.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int ew = warp - FIRST_EPI_WARP;                     // 0..7
        const int quad = warp & 3, half = ew >> 2;
        const int lq = lane >> 2, lr = lane & 3;
        uint32_t *stage = epi_s + ew * (32 * EPI_PITCH + SLOW_BUF * 4);
        uint4 *sbuf = reinterpret_cast<uint4 *>(stage + 32 * EPI_PITCH);
        uint32_t slow_cnt = 0;                                    // warp-uniform: entries buffered in sbuf
        const uint32_t lanemask_lt = (1u << lane) - 1u;
        const int32_t Nn = A.n_sel;
        const ScreenK K = make_screen(A);
        const uint32_t m_shift = K.m_shift, thres = K.thres;
        uint32_t tl = 0;
        const uint32_t tmem_empty_leader = PAIR ? mapa_u32(tmem_empty, 0) : 0u;
        for (int t = unit; t < A.n_tiles; t += n_units, ++tl) {
            const uint32_t buf = tl % Cfg::ACC_BUFS;
            const int4 tile = SINGLE ? first_tile : A.tiles[t];
            const SetRec &S = A.set[SINGLE ? 0 : tile.z];
            // rows / columns in the launch's row space (operands, frequencies) and relative to the set (outputs)
            const int64_t r0g = (int64_t)(PAIR ? tile.x * 2 + rank : tile.x) * MMA_M, c0g = (int64_t)tile.y * N;
            const int64_t base = SINGLE ? 0 : S.base_row, v = S.v;
            const int64_t r0 = r0g - base, c0 = c0g - base;
            uint32_t *const packed = S.packed; int32_t *const n11o = S.n11; const int64_t out_off = S.out_off;
            // the column variants this warp works on (its half of the tile's columns): a private copy per warp costs
            // a few redundant loads and saves a barrier across the eight epilogue warps on every tile
            ColPair *cols = col_s + ew * (N / 4) - half * (N / 4);        // indexed by HALF the column's offset in the tile
            __syncwarp();                                                 // the previous tile's readers are done
            {
                float *cf = reinterpret_cast<float *>(cols);              // column i: n1f at [4 * (i / 2) + (i & 1)], rc two floats on
                for (int i = half * (N / 2) + lane; i < (half + 1) * (N / 2); i += 32) {
                    const int32_t n1t = A.freq_rows[c0g + i].n1;
                    cf[4 * (i >> 1) + (i & 1)] = minor_f(n1t, Nn);
                    cf[4 * (i >> 1) + (i & 1) + 2] = rcp_nn(n1t, Nn);
                }
            }
            __syncwarp();
            // the first pass's row counts, fetched before the wait (freq_rows is padded to whole tiles)
            const int32_t n1a_first = SINGLE ? A.freq_rows[r0g + quad * 32 + lq].n1 : 0, n1b_first = SINGLE ? A.freq_rows[r0g + quad * 32 + lq + 8].n1 : 0;
            if (!mbar_wait(tmem_full + 8 * buf, (tl / Cfg::ACC_BUFS) & 1, abort_s, A.error_flag, 128)) {
                if (SINGLE) break;            // aborted: still meet the other epilogue warps at the barrier below
                goto done;
            }
            tc_fence_after();
            if (TRACE && A.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && tl < 3) A.trace[1 + 2 * tl] = gtime();   // accumulator ready
            if (ew == 0 && lane == 0 && tl == 0) LDX_CTA_STAMP(2);
            if (TRACE && A.trace && blockIdx.x == 1 && ew == 0 && lane == 0) A.trace[39] = gtime();
            const int h_end = SINGLE && N == 128 && !PAIR && A.help ? 1 : 2;     // single wave: rows 16..31 of the quadrant belong to the wideners
#pragma unroll 1
            for (int h = 0; h < h_end; ++h) {
                const int64_t rmin = r0 + quad * 32 + 16 * h;    // this pass: rows rmin .. rmin + 15 of the set
                if (rmin >= v || (TRACE && (A.dbg & 4))) break;  // warp-uniform
                // the lane's two row variants
                const int64_t ra = rmin + lq, rb = ra + 8;
                const int32_t n1a = SINGLE && h == 0 ? n1a_first : A.freq_rows[base + ra].n1, n1b = SINGLE && h == 0 ? n1b_first : A.freq_rows[base + rb].n1;
                const RowP Ra = make_rowp(n1a, Nn), Rb = make_rowp(n1b, Nn);
                // result words of (row, column c0 + 2 lr + j) live at p?[j]
                uint32_t *pa = packed + (ra * (ra - 1) / 2 - out_off + c0 + 2 * lr);
                uint32_t *pb = packed + (rb * (rb - 1) / 2 - out_off + c0 + 2 * lr);
                int32_t *qa = n11o ? n11o + (ra * (ra - 1) / 2 - out_off + c0 + 2 * lr) : nullptr;
                int32_t *qb = n11o ? n11o + (rb * (rb - 1) / 2 - out_off + c0 + 2 * lr) : nullptr;
                const uint32_t tmem_acc = tmem_base + buf * N + ((uint32_t)(quad * 32 + 16 * h) << 16);
#pragma unroll 1
                for (int cb = half * (N / 2); cb < (half + 1) * (N / 2); cb += 32) {
                    const int64_t cg0 = c0 + cb;                  // first column of this load
                    if (cg0 >= rmin + 15 || cg0 >= v) break;      // warp-uniform: nothing below the diagonal
                    uint32_t acc[16];
                    if (TRACE && A.trace && blockIdx.x == 1 && ew == 0 && lane == 0) A.trace[40 + (cb >> 5) * 3] = gtime();     // before the load
                    tmem_ld16x256(tmem_acc + (uint32_t)cb, acc);
                    if (TRACE && A.trace && blockIdx.x == 1 && ew == 0 && lane == 0) A.trace[41 + (cb >> 5) * 3] = gtime();     // accumulators in registers
                    uint32_t word[16];
                    uint32_t slow = 0;                            // bit i: pair i must be redone exactly
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {                 // registers i, i + 1: same row, adjacent columns
                        const int k = i >> 2, g = (i >> 1) & 1;
                        const ColPair cp = cols[(cb >> 1) + 4 * k + lr];
                        bool s0, s1;
                        if (TRACE && (A.dbg & 64)) { word[i] = acc[i] + cp.rc[0]; word[i + 1] = acc[i + 1]; s0 = s1 = false; }   // diagnostics: no screen arithmetic
                        else
                        fast_pair2<THRES>(acc[i], acc[i + 1], K, g ? Rb : Ra, cp, word[i], word[i + 1], s0, s1);
                        if (s0) slow |= 1u << i;
                        if (s1) slow |= 2u << i;
                    }
                    if (TRACE && (A.dbg & 32)) slow = 0;              // diagnostics: nobody is deferred
                    if (TRACE && (A.dbg & 16)) {                      // diagnostics: no result stores (keep the values alive)
                        uint32_t x = 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) x ^= word[i];
                        if (x == 0x12345678u) pa[0] = x;
                    } else
                    {
                    const bool interior = cg0 + 32 <= rmin && rmin + 15 < v;   // warp-uniform: every pair is below the diagonal
                    if (interior) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int k = i >> 2, g = (i >> 1) & 1, e = i & 1;
                            LDX_ST((g ? pb : pa) + cb + 8 * k + e, word[i]);
                        }
                    } else {
                        uint32_t valid = 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int k = i >> 2, g = (i >> 1) & 1, e = i & 1;
                            const int64_t row = g ? rb : ra, col = cg0 + 8 * k + 2 * lr + e;
                            if (row < v && col < row) {           // only pairs of the triangle exist
                                valid |= 1u << i;
                                LDX_ST((g ? pb : pa) + cb + 8 * k + e, word[i]);
                            }
                        }
                        slow &= valid;
                    }
                    }
                    if (n11o) {                                   // warp-uniform; the counts are a diagnostic / test output
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int k = i >> 2, g = (i >> 1) & 1, e = i & 1;
                            const int64_t row = g ? rb : ra, col = cg0 + 8 * k + 2 * lr + e;
                            if (row < v && col < row)
                                (g ? qb : qa)[cb + 8 * k + e] = true_n11((int32_t)(acc[i] >> ACC_SHIFT), g ? n1b : n1a, A.freq_rows[base + col].n1, Nn);
                        }
                    }
                    // Rare (about 1% of the pairs): a screened value sits next to a rounding boundary, or D
                    // is exactly 0 for two polymorphic variants.  Those pairs are not redone here -- one lane
                    // running the reference's operation sequence would stall the warp for a microsecond --
                    // but queued for slow_pairs_kernel, which runs after this kernel and overwrites their
                    // words.  The queue is per warp in shared memory (ballot compaction, no atomics) and is
                    // flushed to the global list when it fills up and at the end.
                    if (TRACE && A.trace && blockIdx.x == 1 && ew == 0 && lane == 0) A.trace[42 + (cb >> 5) * 3] = gtime();     // words stored
                    if (__any_sync(0xffffffffu, slow != 0)) {
                        // a lane with flagged pairs parks its counts in shared memory and fetches them by index
                        if (slow) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                *reinterpret_cast<uint4 *>(stage + lane * EPI_PITCH + i) = make_uint4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                        }
                        __syncwarp();
                        while (true) {
                            const uint32_t bal = __ballot_sync(0xffffffffu, slow != 0);
                            if (bal == 0) break;
                            if (slow_cnt > SLOW_BUF - 32) slow_cnt = flush_slow<SINGLE>(A, sbuf, slow_cnt, lane, m_shift, thres);
                            if (slow) {
                                const int i = __ffs(slow) - 1;
                                slow &= slow - 1;
                                const uint32_t a = stage[lane * EPI_PITCH + i];
                                const int64_t row = (i & 2) ? rb : ra, col = cg0 + 8 * (i >> 2) + 2 * lr + (i & 1);
                                sbuf[slow_cnt + __popc(bal & lanemask_lt)] = make_uint4((uint32_t)row, (uint32_t)col, a >> ACC_SHIFT, SINGLE ? 0u : (uint32_t)tile.z);
                            }
                            slow_cnt += __popc(bal);
                            __syncwarp();
                        }
                    }
                }
            }
            // this warp's TMEM reads of the tile are complete: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_remote(tmem_empty_leader + 8 * buf); else mbar_arrive(tmem_empty + 8 * buf); }
            if (TRACE && A.trace && blockIdx.x == 0 && ew == 0 && lane == 0 && tl < 3) A.trace[2 + 2 * tl] = gtime();       // epilogue of this tile done
            if (ew == 0 && lane == 0 && tl == 0) LDX_CTA_STAMP(3);
            if (TRACE && A.trace && blockIdx.x == 1 && ew == 0 && lane == 0) A.trace[46] = gtime();
        }
        if (SINGLE) {
            // Single wave: no follow-up kernel.  This warp's buffered pairs go onto the CTA's list (what does not
            // fit is settled right here), then everybody settles the list: see the wideners' side above.
            __syncwarp();
            uint32_t base = 0;
            if (lane == 0 && slow_cnt) base = atomicAdd(pool_cnt, slow_cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t i = lane; i < slow_cnt; i += 32) {
                if (base + i < POOL_CAP) pool[base + i] = sbuf[i];
                else settle_slow_pair(sbuf[i], A, m_shift, thres);
            }
            asm volatile("bar.sync 1, %0;" :: "n"(POOL_THREADS) : "memory");
            const uint32_t total = min(*pool_cnt, POOL_CAP);
            for (uint32_t j = threadIdx.x - 32 * FIRST_WIDEN_WARP; j < total; j += POOL_THREADS)
                settle_slow_pair(pool[j], A, m_shift, thres);
        } else {
            flush_slow<false>(A, sbuf, slow_cnt, lane, m_shift, thres);
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (SINGLE && threadIdx.x == 0) {
        __threadfence();
        const uint32_t ticket = atomicAdd(&A.counters[3], 1u);
        if (ticket == gridDim.x - 1) {            // last CTA: what slow_pairs_kernel's last block does otherwise
            __threadfence();
            A.counters[3] = 0;
            if (A.mailbox) publish_record(A.mailbox, A.seq, *(volatile uint32_t *)&A.counters[0], *(volatile uint32_t *)&A.counters[1]);
        }
    }
    if (TRACE && A.trace && blockIdx.x == 0 && threadIdx.x == 0) A.trace[56] = gtime();   // all roles done
    if (threadIdx.x == 0) LDX_CTA_STAMP(4);
    if (PAIR) cluster_sync_all();      // the peer may still read this CTA's operands / arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

bool triangle_mma_available() { return true; }
// The screening arithmetic's guard band grows with N^2 (see fast_pair2): up to this many selected
// haplotypes fewer than 1% of the pairs are deferred; beyond it ENGINE_AUTO uses the popcount engine.
int triangle_mma_max_haplotypes() { return 8192; }

template <int N, bool THRES, bool TRACE, bool PAIR, bool SINGLE, bool DIRECT>
static int launch_tiles_t(ldx_ctx *ctx, const MmaArgs &A) {
    static bool attr_set[64] = {};            // a function attribute is per device: a process may hold contexts on several
    const int dv = ctx->device & 63;
    if (!attr_set[dv]) {
        LDX_CUDA(cudaFuncSetAttribute(triangle_mma_kernel<N, THRES, TRACE, PAIR, SINGLE, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)MmaCfg<N, PAIR>::SMEM));
        attr_set[dv] = true;
    }
    // persistent: one CTA per SM; a pair kernel runs sm_count / 2 clusters of two
    const int units = PAIR ? ctx->sm_count / 2 : ctx->sm_count;
    const int grid = (A.n_tiles < units ? A.n_tiles : units) * (PAIR ? 2 : 1);
    timing_begin(ctx);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(MMA_THREADS); cfg.dynamicSmemBytes = MmaCfg<N, PAIR>::SMEM; cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = PAIR ? 2 : 1;
        LDX_CUDA(cudaLaunchKernelEx(&cfg, triangle_mma_kernel<N, THRES, TRACE, PAIR, SINGLE, DIRECT>, A));
    }
    timing_end(ctx);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "triangle_mma_kernel");
    return LDX_OK;
}
template <int N, bool PAIR, bool SINGLE, bool DIRECT>
static int launch_tiles_s(ldx_ctx *ctx, const MmaArgs &A) {
    if (A.trace && !A.has_thres && N == 128) return launch_tiles_t<N, false, N == 128, PAIR, SINGLE, DIRECT>(ctx, A);   // diagnostics build of the kernel
    return A.has_thres ? launch_tiles_t<N, true, false, PAIR, SINGLE, DIRECT>(ctx, A) : launch_tiles_t<N, false, false, PAIR, SINGLE, DIRECT>(ctx, A);
}

// The planes of a store as a 2-D tensor ([n_variants][stride_words * 2] uint32), box = 8 words (32 B: one pipeline stage's two
// haplotype chunks) x 128 rows.  Encoded once per store through the driver entry point (no link-time libcuda dependency).
static int ensure_tensor_map(ldx_store *s) {
    if (s->tmap_ready) return LDX_OK;
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return set_error(LDX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    static_assert(sizeof(CUtensorMap) == sizeof(s->tmap), "tensor map size");
    const cuuint64_t dims[2] = {(cuuint64_t)s->stride_words * 2, (cuuint64_t)std::max<int64_t>(s->n_variants, 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)s->stride_words * 8};
    const cuuint32_t box[2] = {8, (cuuint32_t)MMA_M}, estr[2] = {1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap *>(s->tmap), CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, s->d_planes, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(LDX_ERR_CUDA, "cuTensorMapEncodeTiled failed for the store's planes");
    s->tmap_ready = true;
    return LDX_OK;
}

int launch_triangle_mma(ldx_ctx *ctx, const MmaSetDesc *sets, int n_sets, const int64_t *d_rows, int measure, int has_thres,
                        int thres_e4, uint32_t publish_seq) {
    if (n_sets < 1 || n_sets > MMA_MAX_SETS) return set_error(LDX_ERR_ARG, "tcgen05 engine: 1..32 variant sets per launch");
    ldx_store *s0 = sets[0].s;
    int64_t v_max = 0;
    for (int k = 0; k < n_sets; ++k) {
        const MmaSetDesc &d = sets[k];
        if (d.s->ctx != ctx) return set_error(LDX_ERR_ARG, "tcgen05 engine: every store of a launch must belong to the calling context");
        if (d.s->n_hap != s0->n_hap || d.s->n_sel != s0->n_sel || d.s->stride_words != s0->stride_words)
            return set_error(LDX_ERR_ARG, "tcgen05 engine: the variant sets of one launch must share the haplotype count and the number of selected haplotypes");
        if (d.row_begin % MMA_M) return set_error(LDX_ERR_ARG, "tcgen05 engine: row_begin must be a multiple of 128");
        if (n_sets > 1 && d.row_begin != 0) return set_error(LDX_ERR_ARG, "tcgen05 engine: row ranges are for single-set calls");
        if (!d.d_packed) return set_error(LDX_ERR_ARG, "tcgen05 engine: the packed output is required");
        v_max = std::max(v_max, d.v);
    }
    if (v_max < 2) return LDX_OK;
    if (s0->n_sel > triangle_mma_max_haplotypes())
        return set_error(LDX_ERR_ARG, "tcgen05 engine: more than 8192 selected haplotypes (use LDX_ENGINE_POPC or LDX_ENGINE_AUTO)");
    if (s0->n_hap > (1 << 23)) return set_error(LDX_ERR_ARG, "tcgen05 engine: more than 2^23 haplotypes would overflow the int32 accumulator");
    // 128-haplotype chunks, rounded up to whole pipeline stages of two (rows are 128 B multiples,
    // i.e. a multiple of 8 chunks, so the padding chunk is inside the row and zero)
    const int kc_count = ((s0->n_hap + KCHUNK - 1) / KCHUNK + 1) / 2 * 2;
    // the launch's row space: the sets one after the other, each padded to whole 256-row panels
    int64_t base_row[MMA_MAX_SETS], total_rows = 0;
    for (int k = 0; k < n_sets; ++k) { base_row[k] = total_rows; total_rows += (sets[k].v + 255) / 256 * 256; }
    const int64_t panels = total_rows / MMA_M;
    // tile width: narrow tiles give every SM several tiles to overlap for small matrices, wide
    // tiles amortise the widening work for large ones
    int n_tile = ctx->mma_tile_n;
    if (n_tile == 0) n_tile = (n_sets == 1 && v_max <= 1024) ? 64 : 128;
    // CTA pairs (256 x 128 tiles, cta_group::2): a quarter less widening work per result
    // mma_pair: 0 off, 1 on, -1 (default) on for multi-wave calls -- a one-wave call has nothing to overlap the pair's longer
    // prologue and signalling with
    bool pair = ctx->mma_pair != 0 && n_tile == 128 && sets[0].row_begin % (2 * MMA_M) == 0 && ctx->sm_count % 2 == 0;
    auto count_tiles = [&](int64_t ph) {
        size_t n = 0;
        for (int k = 0; k < n_sets; ++k)
            for (int64_t bi = sets[k].row_begin / ph; bi < (sets[k].v + ph - 1) / ph; ++bi)
                n += (size_t)((std::min<int64_t>(bi * ph + ph - 1, sets[k].v - 1) + n_tile - 1) / n_tile);
        return n;
    };
    if (pair && ctx->mma_pair < 0) pair = count_tiles(MMA_M) > (size_t)ctx->sm_count;
    const int64_t ph = pair ? 2 * MMA_M : MMA_M;             // tile height
    const size_t n_tiles = count_tiles(ph);
    if (n_tiles == 0) return LDX_OK;
    if (n_tiles > 0x7fffffffull) return set_error(LDX_ERR_ARG, "tcgen05 engine: too many tiles");
    // one wave of tiles (e.g. the 2,000-variant workload: 136 tiles on 148 SMs): a follow-up kernel for ~1% of the
    // pairs costs more (launch boundary + its own tail) than settling them in the epilogue warps
    static const int inline_env = getenv("LDX_INLINE_SETTLE") ? atoi(getenv("LDX_INLINE_SETTLE")) : -1;
    const bool one_wave = n_tiles <= (size_t)(pair ? ctx->sm_count / 2 : ctx->sm_count);
    const bool single = n_sets == 1 && (inline_env >= 0 ? inline_env != 0 && one_wave : one_wave);
    // direct mode: a one-wave call on contiguous store rows needs no gathered copy of the operands at all
    static const int direct_env = getenv("LDX_MMA_DIRECT") ? atoi(getenv("LDX_MMA_DIRECT")) : -1;
    const bool direct = single && !pair && n_tile == 128 && sets[0].contiguous && (direct_env >= 0 ? direct_env != 0 : ctx->mma_direct != 0) &&
                        sets[0].row0 + total_rows < (1ll << 31);
    // ---- scratch: bit panels (both operand forms), frequencies, tile list, deferred-pair list
    const size_t bits_bytes = direct ? 0 : (size_t)panels * kc_count * 128 * sizeof(uint4);
    const size_t freq_bytes = direct ? 0 : (size_t)total_rows * sizeof(VarFreq);
    const size_t tile_bytes = (n_tiles * sizeof(int4) + 15) / 16 * 16;
    // deferred-pair list: the guard band defers < 1% of the pairs (triangle_mma_max_haplotypes)
    uint64_t n_pairs = 0;
    for (int k = 0; k < n_sets; ++k) {
        const uint64_t v = (uint64_t)sets[k].v, rb = (uint64_t)sets[k].row_begin;
        if (v >= 2) n_pairs += v * (v - 1) / 2 - rb * (rb > 0 ? rb - 1 : 0) / 2;
    }
    const uint64_t slow_cap64 = n_pairs / 64 + 65536;
    if (slow_cap64 > 0xffffffffull) return set_error(LDX_ERR_ARG, "tcgen05 engine: too many pairs for one call");
    const size_t slow_bytes = single ? 0 : (size_t)slow_cap64 * sizeof(uint4);
    const size_t need = 2 * bits_bytes + freq_bytes + tile_bytes + slow_bytes + 1024;
    if (ctx->mma_ops_bytes < need) {
        if (ctx->d_mma_ops) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_mma_ops); ctx->d_mma_ops = nullptr; ctx->mma_ops_bytes = 0; }
        if (cudaMalloc(&ctx->d_mma_ops, need) != cudaSuccess) { cudaGetLastError(); return set_error(LDX_ERR_NOMEM, "tcgen05 operand scratch allocation failed"); }
        ctx->mma_ops_bytes = need;
        ctx->mma_tiles_key.clear();
    }
    uint8_t *base = reinterpret_cast<uint8_t *>(ctx->d_mma_ops);
    uint4 *d_bits = reinterpret_cast<uint4 *>(base), *d_bits_rev = reinterpret_cast<uint4 *>(base + bits_bytes);
    VarFreq *d_freq_rows = reinterpret_cast<VarFreq *>(base + 2 * bits_bytes);
    int4 *d_tiles = reinterpret_cast<int4 *>(base + 2 * bits_bytes + freq_bytes);
    uint4 *d_slow = reinterpret_cast<uint4 *>(base + 2 * bits_bytes + freq_bytes + tile_bytes);
    // ---- tile list (cached): every tile holding at least one pair with row > col, set by set, row-panel major, in the launch's
    // row space.  All tiles cost the same K loop; the order keeps the tiles a wave of CTAs works on in neighbouring panels in L2.
    std::vector<int64_t> key;
    key.reserve(2 + 2 * (size_t)n_sets);
    key.push_back(n_tile); key.push_back(pair ? 1 : 0);
    for (int k = 0; k < n_sets; ++k) { key.push_back(sets[k].v); key.push_back(sets[k].row_begin); }
    if (ctx->mma_tiles_ptr != d_tiles || ctx->mma_tiles_key != key) {
        std::vector<int4> tiles;
        tiles.reserve(n_tiles);
        for (int k = 0; k < n_sets; ++k) {
            const int64_t pb = base_row[k] / ph, cb = base_row[k] / n_tile;
            for (int64_t bi = sets[k].row_begin / ph; bi < (sets[k].v + ph - 1) / ph; ++bi) {
                const int64_t rmax = std::min<int64_t>(bi * ph + ph - 1, sets[k].v - 1);
                for (int64_t bj = 0; bj * n_tile < rmax; ++bj) tiles.push_back(make_int4((int)(pb + bi), (int)(cb + bj), k, 0));
            }
        }
        LDX_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), n_tiles * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));   // `tiles` is a local
        ctx->mma_tiles_key = key; ctx->mma_tiles_ptr = d_tiles;
    }

    if (!direct) {
        GatherArgs G;
        memset(&G, 0, sizeof G);
        for (int k = 0; k < n_sets; ++k) {
            GatherSet &g = G.set[k];
            g.planes = sets[k].s->d_planes; g.mask = sets[k].s->d_mask; g.freq = sets[k].s->d_freq; g.rows = d_rows + sets[k].rows_off;
            g.v = sets[k].v; g.v_pad = (sets[k].v + 255) / 256 * 256; g.base_row = base_row[k];
            g.stride_words = sets[k].s->stride_words; g.n_sel = sets[k].s->n_sel;
        }
        G.kc_count = kc_count; G.bits = d_bits; G.bits_rev = d_bits_rev; G.freq_rows = d_freq_rows;
        const int64_t v_pad_max = (v_max + 255) / 256 * 256;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(v_pad_max / 256), (unsigned)kc_count, (unsigned)n_sets); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        LDX_CUDA(cudaLaunchKernelEx(&cfg, gather_bits_kernel, G));
        ctx->launches++;
        LDX_LAUNCHED(ctx, "gather_bits_kernel");
    }

    MmaArgs A;
    memset(&A, 0, sizeof A);
    A.bits = d_bits; A.bits_rev = d_bits_rev; A.kc_count = kc_count; A.freq_rows = d_freq_rows; A.fc = s0->fc; A.tiles = d_tiles;
    A.n_tiles = (int32_t)n_tiles;
    A.measure = measure; A.has_thres = has_thres; A.thres_e4 = thres_e4;
    for (int k = 0; k < n_sets; ++k) {
        SetRec &r = A.set[k];
        r.base_row = base_row[k]; r.v = sets[k].v; r.out_off = sets[k].row_begin * (sets[k].row_begin - 1) / 2;
        r.packed = sets[k].d_packed; r.n11 = sets[k].d_n11; r.fix_tag = sets[k].fix_tag;
        r.gen = sets[k].s->n_nonsimple > 0 ? sets[k].s->d_gen : nullptr;
        r.rows = d_rows + sets[k].rows_off; r.row0 = sets[k].row0;
    }
    A.n_sel = s0->n_sel;
    {   // guard band of the screening arithmetic (see fast_pair2).  Its fp32 half needs the minor-allele products exact
        // (<= 2^24, i.e. N <= 8192); should a caller ever get past the check above with more, every pair
        // takes the exact path.
        const double n = (double)s0->n_sel, g = 1.0e4 * 16.0 * 1.1102230246251565e-16 * n * n;
        const bool ok = s0->n_sel <= 8192;
        A.lim_r2 = ok ? (float)(0.5 - (g + 1.0e-5)) : -1.0f;
        A.lim_dp = ok ? (float)(0.5 - (0.5 * g + 1.0e-5)) : -1.0f;
    }
    A.fix = FixupSink{ctx->d_fix, ctx->d_fix_count, ctx->fix_capacity, sets[0].fix_tag};
    A.error_flag = reinterpret_cast<int32_t *>(ctx->d_fix_count + 1);
    A.slow = d_slow; A.slow_count = ctx->d_fix_count + 2;
    A.slow_cap = ctx->defer_cap ? (uint32_t)std::min<uint64_t>(slow_cap64, (uint64_t)ctx->defer_cap) : (uint32_t)slow_cap64;
    A.pool_cap = ctx->defer_cap ? (uint32_t)ctx->defer_cap : 0xffffffffu;
    A.trace = ctx->d_trace;
    A.dbg = getenv("LDX_DEBUG_MMA") ? atoi(getenv("LDX_DEBUG_MMA")) : 0;
    A.inline_settle = single ? 1 : 0;
    A.counters = ctx->d_fix_count; A.mailbox = publish_seq ? ctx->d_mailbox : nullptr; A.seq = publish_seq;
    A.help = single && n_tile == 128 && !pair && !getenv("LDX_NO_HELP");
    if (direct) {
        LDX_TRY(ensure_tensor_map(s0));
        memcpy(&A.tmap, s0->tmap, sizeof A.tmap);
        A.mask = reinterpret_cast<const uint4 *>(s0->d_mask);
        A.row0 = (int32_t)sets[0].row0;
        A.freq_rows = s0->d_freq + sets[0].row0;          // padded by STORE_FREQ_PAD zeroed entries: whole tiles may be read
        A.bits = A.bits_rev = nullptr;
        A.set[0].rows = nullptr;                          // store row = row0 + matrix row
    }
    int rc;
    if (n_tile == 64) rc = single ? launch_tiles_s<64, false, true, false>(ctx, A) : launch_tiles_s<64, false, false, false>(ctx, A);
    else if (n_tile == 128) {
        if (direct) rc = launch_tiles_s<128, false, true, true>(ctx, A);
        else if (pair) rc = single ? launch_tiles_s<128, true, true, false>(ctx, A) : launch_tiles_s<128, true, false, false>(ctx, A);
        else rc = single ? launch_tiles_s<128, false, true, false>(ctx, A) : launch_tiles_s<128, false, false, false>(ctx, A);
    } else return set_error(LDX_ERR_ARG, "tcgen05 tile width must be 64 or 128");
    if (rc != LDX_OK) return rc;
    if (single) return LDX_OK;
    // deferred pairs + completion record (d_fix_count: [0] near-ties, [1] error flag, [2] deferred pairs, [3] ticket)
    // ~1% of the pairs are deferred and each costs a long, serial fp64 chain: one pair per thread at twice that
    // rate (idle blocks are cheap, a thread looping over several pairs is not)
    const int sgrid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * 8, std::max<uint64_t>(1, n_pairs / (50u * 256u) + 1));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)sgrid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        LDX_CUDA(cudaLaunchKernelEx(&cfg, slow_pairs_kernel, A));
    }
    ctx->launches++;
    LDX_LAUNCHED(ctx, "slow_pairs_kernel");
    return LDX_OK;
}

}  // namespace ldx
