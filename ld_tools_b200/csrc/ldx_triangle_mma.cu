// ldx_triangle_mma.cu -- K5: all-pairs (alt, alt) haplotype counts as an EXACT int8 Gram matrix
// on the 5th-generation tensor cores (tcgen05.mma kind::i8, int32 accumulators in TMEM), with the
// fp64 D / D' / r2 finalisation fused into the epilogue.
//
// Replaces the double loop at ld_triangle.py:133-230 (var_1 = row variant, var_2 = column
// variant, ld_triangle.py:193); the counting step is calc_ld.py:30-32 for 128 x N pairs at once:
//     n11[r][c] = sum_h  A[r][h] * A[c][h],   A[v][h] = (plane[v] & mask) bit h  in {0, 1}
// Products of 0/1 bytes accumulated in int32 are exact, so the counts equal the popcount
// engine's bit for bit (tests/test_parity_gpu.py compares them).
//
// Pipeline of one CTA (= one 128 x N tile of the lower triangle, 10 warps):
//   expand_kernel (separate launch, O(V)): bit planes -> 0/1 bytes, written to global memory
//       ALREADY in the shared-memory image tcgen05 wants: [panel of 128 variants][128-haplotype
//       chunk][row][128 B] with the 128-byte swizzle (16-byte unit j of row r stored at j ^ (r&7)).
//       Every pipeline stage is then ONE contiguous 16 KB block per operand panel.
//   warp 0   producer: cp.async.bulk (UBLKCP, the TMA engine's linear mode) global -> shared,
//            completion counted on an mbarrier (complete_tx).  No tensor map needed.
//   warp 1   MMA issuer: one elected lane issues 4 x tcgen05.mma (K = 32 each) per stage and
//            tcgen05.commit's the stage back to the producer; owns the TMEM allocation.
//   warps 2-9 epilogue: tcgen05.ld the int32 counts (lane = row, column = column variant),
//            run calc_ld.py:33-97 in fp64 (ldx_common.cuh) and store packed results.
// Several CTAs are resident per SM (N <= 128), so one tile's tensor work overlaps another tile's
// fp64 epilogue without an intra-CTA software pipeline.
//
// Roofline: int8 tensor pipe, 2 * n_hap int8 ops per pair; co-bounds are the fp64 epilogue
// (~45 fp64 instructions per pair) and L2 -> SM operand traffic ((128 + N) * 128 B per stage).
#include <vector>

#include "ldx_internal.h"
#include "ldx_fixup.cuh"

namespace ldx {

constexpr int MMA_M = 128;            // rows (variants) per tile = TMEM lanes
constexpr int KCHUNK = 128;           // haplotypes (= bytes) per shared-memory row: one swizzle atom
constexpr int MMA_K = 32;             // int8 K of one tcgen05.mma
constexpr int PANEL_BYTES = MMA_M * KCHUNK;   // 16 KB: one (panel, chunk) block
constexpr int MMA_THREADS = 320;      // 1 producer + 1 MMA + 8 epilogue warps

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
// Bounded wait: a pipeline bug must end in an error code, never in a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int *abort_s, int32_t *err) {
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try_wait(bar, parity)) return true;
        if (*abort_s) return false;
        if ((spins & 0x3ff) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (!t0) t0 = t;
            else if (t - t0 > 2000000000ull) { *abort_s = 1; atomicExch(err, 1); return false; }
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
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

// ------------------------------------------------------------------------------------------ expand
// bit planes -> swizzled 0/1 byte panels.  One thread = one 16-byte unit (16 haplotypes of one
// variant).  Four bits at a time: x * 0x00204081 puts bit i of the nibble at bit 8*i, & 0x01010101.
__global__ void __launch_bounds__(256)
expand_kernel(const uint64_t *__restrict__ planes, const uint64_t *__restrict__ mask, int32_t stride_words,
              const int64_t *__restrict__ rows, int64_t v, int64_t v_pad, int32_t kc_count,
              const VarFreq *__restrict__ freq, uint8_t *__restrict__ ops, VarFreq *__restrict__ freq_rows) {
    const int kc = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);   // matrix row
    const int j = threadIdx.x & 7;                                      // 16-byte unit in the 128 B row
    if (r >= v_pad) return;
    uint32_t bits = 0;
    if (r < v) {
        const int64_t srow = rows[r];
        const int h0 = kc * KCHUNK + j * 16;                            // first haplotype of this unit
        const uint64_t w = planes[srow * stride_words + (h0 >> 6)] & mask[h0 >> 6];
        bits = (uint32_t)(w >> (h0 & 63)) & 0xffffu;
        if (kc == 0 && j == 0) freq_rows[r] = freq[srow];
    } else if (kc == 0 && j == 0) {
        VarFreq z; z.p = 0.0; z.q = 0.0; z.pq = 0.0; z.n1 = 0; z.p_e4 = 0;
        freq_rows[r] = z;
    }
    uint4 out;
    out.x = ((bits & 0xf) * 0x00204081u) & 0x01010101u;
    out.y = (((bits >> 4) & 0xf) * 0x00204081u) & 0x01010101u;
    out.z = (((bits >> 8) & 0xf) * 0x00204081u) & 0x01010101u;
    out.w = (((bits >> 12) & 0xf) * 0x00204081u) & 0x01010101u;
    const int64_t panel = r >> 7;
    const int rl = (int)(r & 127);
    uint8_t *dst = ops + (panel * kc_count + kc) * (int64_t)PANEL_BYTES + rl * KCHUNK + ((j ^ (rl & 7)) << 4);
    *reinterpret_cast<uint4 *>(dst) = out;
}

// ------------------------------------------------------------------------------------------ GEMM + epilogue
struct MmaArgs {
    const uint8_t *ops; int32_t kc_count;
    const VarFreq *freq_rows; FinalCtx fc;
    const int2 *tiles;
    int64_t v; int measure, has_thres, thres_e4;
    uint32_t *packed; int32_t *n11;
    FixupSink fix;
    int32_t *error_flag;
};

template <int N> struct MmaCfg {
    static constexpr int STAGES = N == 64 ? 4 : (N == 128 ? 3 : 4);
    static constexpr int B_BYTES = N * KCHUNK;
    static constexpr int STAGE_BYTES = PANEL_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = N < 32 ? 32 : N;
    static constexpr int CTAS_PER_SM = N <= 128 ? 2 : 1;
    static constexpr size_t SMEM = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + N * sizeof(VarFreq) + 256;
};

template <int N>
__global__ void __launch_bounds__(MMA_THREADS, MmaCfg<N>::CTAS_PER_SM)
triangle_mma_kernel(const MmaArgs A) {
    using Cfg = MmaCfg<N>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t *smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);      // swizzle-128B needs 1 KB alignment
    VarFreq *fb_s = reinterpret_cast<VarFreq *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(fb_s + N);
    const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + Cfg::STAGES);
    const uint32_t tmem_full_bar = smem_u32(bars + 2 * Cfg::STAGES);
    uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(bars + 2 * Cfg::STAGES + 1);
    volatile int *abort_s = reinterpret_cast<volatile int *>(tmem_ptr_s + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int2 tile = A.tiles[blockIdx.x];
    const int64_t r0 = (int64_t)tile.x * MMA_M, c0 = (int64_t)tile.y * N;
    const int kc_count = A.kc_count;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        mbar_init(tmem_full_bar, 1);
        *abort_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // one warp allocates TMEM (power-of-two columns >= 32) and later frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_ptr_s)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < N; i += MMA_THREADS) fb_s[i] = A.freq_rows[c0 + i];   // column variants
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_ptr_s;

    if (warp == 0) {
        // ===== producer: two linear bulk copies per stage (A panel chunk, B rows chunk)
        if (lane == 0) {
            const uint8_t *a_src = A.ops + (int64_t)tile.x * kc_count * PANEL_BYTES;
            for (int kc = 0; kc < kc_count; ++kc) {
                const int s = kc % Cfg::STAGES, it = kc / Cfg::STAGES;
                if (!mbar_wait(empty_bar + 8 * s, (it & 1) ^ 1, abort_s, A.error_flag)) break;
                const uint32_t bar = full_bar + 8 * s;
                mbar_arrive_expect_tx(bar, Cfg::STAGE_BYTES);
                const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                bulk_g2s(sa, a_src + (int64_t)kc * PANEL_BYTES, PANEL_BYTES, bar);
#pragma unroll
                for (int part = 0; part < (N + 127) / 128; ++part) {
                    const int64_t crow = c0 + part * 128;                 // first column variant of this part
                    const int rows_here = N < 128 ? N : 128;
                    const uint8_t *b_src = A.ops + ((crow >> 7) * kc_count + kc) * (int64_t)PANEL_BYTES + (crow & 127) * KCHUNK;
                    bulk_g2s(sa + PANEL_BYTES + part * PANEL_BYTES, b_src, rows_here * KCHUNK, bar);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(MMA_M, N);
            bool ok = true;
            for (int kc = 0; kc < kc_count && ok; ++kc) {
                const int s = kc % Cfg::STAGES, it = kc / Cfg::STAGES;
                ok = mbar_wait(full_bar + 8 * s, it & 1, abort_s, A.error_flag);
                if (!ok) break;
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + PANEL_BYTES);
#pragma unroll
                for (int k = 0; k < KCHUNK / MMA_K; ++k)     // +32 B along K inside the swizzle atom = +2 encoded
                    umma_i8(tmem_acc, da + 2 * k, db + 2 * k, idesc, (uint32_t)((kc | k) != 0));
                umma_commit(empty_bar + 8 * s);               // stage reusable once these MMAs have read it
            }
            if (ok) umma_commit(tmem_full_bar);               // accumulator complete
        }
    } else {
        // ===== epilogue: warp w may only touch TMEM lanes 32*(w%4) .. +31
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int64_t r = r0 + quad * 32 + lane;
        const bool ok = mbar_wait(tmem_full_bar, 0, abort_s, A.error_flag);
        tc_fence_after();
        if (ok) {
            const VarFreq fa = A.freq_rows[r];
            const int64_t rbase = r * (r - 1) / 2;
            const int64_t warp_rmax = r0 + quad * 32 + 31;
#pragma unroll 1
            for (int c = half * (N / 2); c < (half + 1) * (N / 2); c += 16) {
                if (c0 + c >= warp_rmax || c0 + c >= A.v) break;        // warp-uniform: nothing below the diagonal
                uint32_t acc[16];
                tmem_ld16(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, acc);
                if (r < A.v) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int64_t col = c0 + c + j;
                        if (col < r) {
                            const VarFreq fb = fb_s[c + j];
                            const int32_t cnt = (int32_t)acc[j];
                            const PairFinal f = finalise_pair(cnt, fa, fb, A.fc);   // var_1 = row, var_2 = column
                            uint32_t word = f.packed;
                            if (A.has_thres && measure_e4(word, A.measure) < A.thres_e4) word |= LDX_BELOW_THRES;
                            const int64_t o = rbase + col;
                            if (A.packed) {
                                A.packed[o] = word;
                                if (word & LDX_R2_NEARTIE) fixup_append(A.fix, (uint64_t)o, cnt, fa.n1, fb.n1, word);
                            }
                            if (A.n11) A.n11[o] = cnt;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_acc), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

bool triangle_mma_available() { return true; }

template <int N>
static int launch_tiles(ldx_ctx *ctx, const MmaArgs &A, int n_tiles) {
    static bool attr_set = false;
    if (!attr_set) {
        LDX_CUDA(cudaFuncSetAttribute(triangle_mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmaCfg<N>::SMEM));
        attr_set = true;
    }
    triangle_mma_kernel<N><<<n_tiles, MMA_THREADS, MmaCfg<N>::SMEM, ctx->stream>>>(A);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    return LDX_OK;
}

int launch_triangle_mma(ldx_store *s, const int64_t *d_rows, int64_t v, int measure, int has_thres,
                        int thres_e4, uint32_t *d_packed, int32_t *d_n11) {
    if (v < 2) return LDX_OK;
    ldx_ctx *ctx = s->ctx;
    const int kc_count = (s->n_hap + KCHUNK - 1) / KCHUNK;
    const int64_t v_pad = (v + 255) / 256 * 256;
    const int64_t panels = v_pad / MMA_M;
    // tile width: narrow tiles fill the SMs for small matrices, wide tiles cut L2 traffic for large ones
    int n_tile = ctx->mma_tile_n;
    if (n_tile == 0) n_tile = v <= 4096 ? 64 : 128;
    // ---- tile list: every 128 x N tile that holds at least one pair with row > col, row-panel major
    std::vector<int2> tiles;
    for (int64_t bi = 0; bi < (v + MMA_M - 1) / MMA_M; ++bi) {
        const int64_t rmax = std::min<int64_t>(bi * MMA_M + MMA_M - 1, v - 1);
        for (int64_t bj = 0; bj * n_tile < rmax; ++bj) tiles.push_back(make_int2((int)bi, (int)bj));
    }
    if (tiles.empty()) return LDX_OK;
    // ---- scratch: operand panels | gathered VarFreq | tile list
    const size_t ops_bytes = (size_t)panels * kc_count * PANEL_BYTES;
    const size_t freq_bytes = (size_t)v_pad * sizeof(VarFreq);
    const size_t tile_bytes = tiles.size() * sizeof(int2);
    const size_t need = ops_bytes + freq_bytes + tile_bytes + 1024;
    if (ctx->mma_ops_bytes < need) {
        if (ctx->d_mma_ops) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_mma_ops); ctx->d_mma_ops = nullptr; ctx->mma_ops_bytes = 0; }
        if (cudaMalloc(&ctx->d_mma_ops, need) != cudaSuccess) { cudaGetLastError(); return set_error(LDX_ERR_NOMEM, "tcgen05 operand scratch allocation failed"); }
        ctx->mma_ops_bytes = need;
    }
    uint8_t *d_ops = reinterpret_cast<uint8_t *>(ctx->d_mma_ops);
    VarFreq *d_freq_rows = reinterpret_cast<VarFreq *>(d_ops + ops_bytes);
    int2 *d_tiles = reinterpret_cast<int2 *>(d_ops + ops_bytes + freq_bytes);
    LDX_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tile_bytes, cudaMemcpyHostToDevice, ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));   // `tiles` is a local (pageable copy is staged, but be explicit)

    dim3 egrid((unsigned)((v_pad + 31) / 32), (unsigned)kc_count);
    expand_kernel<<<egrid, 256, 0, ctx->stream>>>(s->d_planes, s->d_mask, s->stride_words, d_rows, v, v_pad, kc_count,
                                                  s->d_freq, d_ops, d_freq_rows);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());

    MmaArgs A;
    A.ops = d_ops; A.kc_count = kc_count; A.freq_rows = d_freq_rows; A.fc = s->fc; A.tiles = d_tiles;
    A.v = v; A.measure = measure; A.has_thres = has_thres; A.thres_e4 = thres_e4;
    A.packed = d_packed; A.n11 = d_n11;
    A.fix = FixupSink{ctx->d_fix, ctx->d_fix_count, ctx->fix_capacity};
    A.error_flag = reinterpret_cast<int32_t *>(ctx->d_fix_count + 1);
    switch (n_tile) {
        case 64: return launch_tiles<64>(ctx, A, (int)tiles.size());
        case 128: return launch_tiles<128>(ctx, A, (int)tiles.size());
        case 256: return launch_tiles<256>(ctx, A, (int)tiles.size());
        default: return set_error(LDX_ERR_ARG, "tcgen05 tile width must be 64, 128 or 256");
    }
}

}  // namespace ldx
