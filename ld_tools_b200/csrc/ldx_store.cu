// ldx_store.cu -- kernels that build and summarise the bit-plane store.
//
//   K1 pack_gt_kernel      VCF GT text -> one bit per haplotype (replaces the per-sample
//                          `rec.samples[name]['GT']` loops, ld_area.py:182-187 / :230-235,
//                          ld_triangle.py:158-186, ld_lite.py:109-137)
//   K2 variant_freq_kernel n1 = popcount(mask & plane), p, q, p*q, round(p,4)
//                          (calc_ld.py:37-44, :96-97; ld_area.py:188-189)
//   subset_kernel          gather selected haplotype columns into a narrower store
#include "ldx_internal.h"
#include "ldx_fixup.cuh"

namespace ldx {

// ------------------------------------------------------------------------------------------ K1
// One CTA per variant row (grid-stride).  The row's text (4 bytes per sample, arbitrary byte alignment inside the VCF line) is
// staged into shared memory word by word.  A thread then takes FOUR samples: five words of the staged text, four
// funnel shifts undo the row's byte skew, and one XOR against "0|0" per sample gives both allele bits and the validity test
// (every other bit of the three bytes must be zero); its 8 haplotype bits go to a byte of the row's image in shared memory --
// haplotype 2*s+a lives at bit 2*s+a, so a thread's byte IS byte s/4 of the row -- and the image leaves with coalesced 8-byte
// stores.  ~9 instructions per sample (the first version: one sample per lane, three byte loads, two ballots and a bit
// interleave per 32 samples, ~40).
constexpr int PACK_THREADS = 256;

__host__ __device__ inline int pack_text_granules(int32_t n_samples) { return (int)((4ll * n_samples + 15 + 16) / 16 + 1); }

__global__ void __launch_bounds__(PACK_THREADS)
pack_gt_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ row_off, int64_t row_pitch,
               int64_t n_rows, int32_t n_samples, uint64_t *__restrict__ planes, int32_t stride_words,
               uint8_t *__restrict__ status) {
    extern __shared__ uint4 stage[];                 // row text, 16-byte granules; then the row's packed image
    __shared__ int s_bad;
    const int64_t row_bytes = 4ll * n_samples;
    uint8_t *outb = reinterpret_cast<uint8_t *>(stage + pack_text_granules(n_samples));      // [stride_words * 8]
    const int out_bytes = stride_words * 8, n_quads = (n_samples + 3) / 4;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t off = row_off ? row_off[r] : r * row_pitch;
        if (off < 0) {                               // a row without genotype text (ldx_store_ingest_vcf: malformed line): all reference
            for (int w = threadIdx.x; w < stride_words; w += PACK_THREADS) planes[r * (int64_t)stride_words + w] = 0;
            if (threadIdx.x == 0 && status) status[r] = 0;
            continue;                                // block-uniform
        }
        // staged word by word from the row's 4-byte-aligned start (coalesced 128-byte requests), so that sample 4b begins in word
        // 4b of the staging area: the consumer's 16-byte shared-memory loads are aligned and free of bank conflicts
        const int skew = (int)(off & 3);
        const int n_words = (int)((skew + row_bytes + 3) >> 2) + 1;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(text + (off - skew));
        uint32_t *stage32 = reinterpret_cast<uint32_t *>(stage);
        if (threadIdx.x == 0) s_bad = 0;
        for (int g = threadIdx.x; g < n_words; g += PACK_THREADS) stage32[g] = __ldcs(src + g);
        __syncthreads();
        const uint32_t *w32 = stage32;
        const int q8 = skew * 8;
        uint32_t bad = 0;
        for (int b = threadIdx.x; b < out_bytes; b += PACK_THREADS) {
            uint32_t byte = 0;
            if (b < n_quads) {
                const uint32_t *p = w32 + 4 * b;     // (the words past the row's end are slack of the staging area: masked below)
                const uint4 x03 = *reinterpret_cast<const uint4 *>(p);
                const uint32_t x0 = x03.x, x1 = x03.y, x2 = x03.z, x3 = x03.w, x4 = p[4];
                const uint32_t v[4] = {__funnelshift_r(x0, x1, q8), __funnelshift_r(x1, x2, q8), __funnelshift_r(x2, x3, q8), __funnelshift_r(x3, x4, q8)};
                // "0|0" -> 0: byte 0 / byte 2 = allele ^ '0', byte 1 = separator ^ '|'; a sample is plain iff nothing but bit 0 of bytes
                // 0 and 2 is left.  (A row with any other sample is flagged and re-packed whole by the general kernel, so the bits
                // taken from such a sample here do not matter.)
                const int live = n_samples - 4 * b;          // >= 4 except in the row's last byte
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t x = (v[j] ^ 0x00307c30u) & 0x00ffffffu;
                    if (j >= live) x = 0;
                    bad |= x & 0x00fefffeu;
                    byte |= ((x & 1u) | ((x >> 15) & 2u)) << (2 * j);
                }
            }
            outb[b] = (uint8_t)byte;                 // pad bytes of the row are written as zero
        }
        if (__any_sync(0xffffffffu, bad != 0) && (threadIdx.x & 31) == 0) atomicOr(&s_bad, 1);
        __syncthreads();
        uint64_t *out = planes + r * (int64_t)stride_words;
        for (int w = threadIdx.x; w < stride_words; w += PACK_THREADS) out[w] = reinterpret_cast<const uint64_t *>(outb)[w];
        if (threadIdx.x == 0 && status) status[r] = (uint8_t)(s_bad ? 1 : 0);
        __syncthreads();
    }
}

int launch_pack_gt(ldx_ctx *ctx, const uint8_t *d_text, const int64_t *d_row_off, int64_t row_pitch,
                   int64_t n_rows, int32_t n_samples, uint64_t *d_planes_first, int32_t stride_words,
                   uint8_t *d_status) {
    if (n_rows <= 0) return LDX_OK;
    const size_t smem = (size_t)pack_text_granules(n_samples) * 16 + (size_t)stride_words * 8;
    if (smem > 200 * 1024) return set_error(LDX_ERR_ARG, "pack_gt: row too long for shared memory");
    if (smem > 48 * 1024)
        LDX_CUDA(cudaFuncSetAttribute(pack_gt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (int64_t)ctx->sm_count * 8;
    const int grid = (int)(n_rows < want ? n_rows : want);
    timing_begin(ctx);
    pack_gt_kernel<<<grid, PACK_THREADS, smem, ctx->stream>>>(d_text, d_row_off, row_pitch, n_rows,
                                                               n_samples, d_planes_first, stride_words,
                                                               d_status);
    timing_end(ctx);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ K2
// Eight lanes per variant: lane l of a group loads 16-byte granules l, l+8, ... of the 128-byte
// aligned row (each group load is one full 128-byte line), popcounts under the mask, and a
// 3-step shuffle tree sums the group.  HBM-bound: one pass over the store.
constexpr int FREQ_THREADS = 256;

__global__ void __launch_bounds__(FREQ_THREADS)
variant_freq_kernel(const uint4 *__restrict__ planes, const uint4 *__restrict__ mask, int32_t stride_u4,
                    int64_t n_variants, FinalCtx fc, VarFreq *__restrict__ freq) {
    const int sub = threadIdx.x & 7;
    const int64_t group0 = ((int64_t)blockIdx.x * FREQ_THREADS + threadIdx.x) >> 3;
    const int64_t n_groups = ((int64_t)gridDim.x * FREQ_THREADS) >> 3;   // multiple of 4
    // the loop bound is warp-uniform (first group of the warp) so the shuffles stay converged
    for (int64_t vb = group0 & ~3ll; vb < n_variants; vb += n_groups) {
        const int64_t v = vb + (group0 & 3);
        const bool valid = v < n_variants;
        const uint4 *row = planes + (valid ? v : 0) * stride_u4;
        int cnt = 0;
        if (valid)
            for (int g = sub; g < stride_u4; g += 8) cnt += popc_and_u4(ldg_u4_stream(row + g), __ldg(mask + g));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
        if (sub == 0 && valid) {
            VarFreq f;
            f.n1 = cnt;
            f.p = __ddiv_rn((double)cnt, fc.n_hap);                        // calc_ld.py:41
            f.q = __ddiv_rn((double)((int)fc.n_hap - cnt), fc.n_hap);      // calc_ld.py:42
            f.pq = __dmul_rn(f.p, f.q);
            bool tie;
            f.p_e4 = (int32_t)round4_e4(f.p, tie);                         // calc_ld.py:96
            freq[v] = f;
        }
    }
}

// One thread per variant, after variant_freq_kernel: rows of the general route get their coded record.
__global__ void __launch_bounds__(256)
general_freq_kernel(const GenStore *__restrict__ G, const int32_t *__restrict__ kind, int64_t n_variants, int32_t n_sel, VarFreq *__restrict__ freq,
                    int32_t *__restrict__ row_len, int32_t *__restrict__ row_n1) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_variants) return;
    const int32_t k = kind[v];
    if (k == -1) { row_len[v] = n_sel; row_n1[v] = freq[v].n1; return; }
    const int32_t code = k >= 0 ? -1 - k : GEN_FULL;
    int32_t len = 0, n1 = 0;
    for (int w = 0; w < G->words; ++w) {
        uint64_t present, alt, ref;
        gen_row_word(*G, v, code, w, present, alt, ref);
        len += __popcll(present); n1 += __popcll(alt);
    }
    VarFreq f;
    f.n1 = code; f.p = 0.0; f.q = 0.0; f.pq = 0.0; f.p_e4 = 0;
    if (len > 0) {
        bool tie;
        f.p_e4 = (int32_t)round4_e4_wide(__ddiv_rn((double)n1, (double)len), tie);
    }
    freq[v] = f;
    row_len[v] = len; row_n1[v] = n1;
}

int launch_variant_freq(ldx_store *s) {
    if (s->n_variants <= 0) return LDX_OK;
    ldx_ctx *ctx = s->ctx;
    // whole warps must stay converged for the shuffles: every group of a warp iterates the same
    // number of times because the loop bound is rounded up per warp below.
    const int64_t groups_needed = (s->n_variants + 3) / 4 * 4;
    int64_t blocks = (groups_needed * 8 + FREQ_THREADS - 1) / FREQ_THREADS;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    variant_freq_kernel<<<(int)blocks, FREQ_THREADS, 0, ctx->stream>>>(
        reinterpret_cast<const uint4 *>(s->d_planes), reinterpret_cast<const uint4 *>(s->d_mask),
        s->stride_words / 2, s->n_variants, s->fc, s->d_freq);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    if (s->n_nonsimple > 0) {
        // rows of the general route: their own list length and allele counts under the selection (calc_ld.py:31, :37-40 for
        // the variant alone), the coded VarFreq.n1 that sends every pair with them down the general route, and
        // round(alt freq, 4) of their OWN list (ld_area.py:188-189)
        const size_t nv = (size_t)s->n_variants;
        if (!s->d_row_len) LDX_CUDA(cudaMalloc(&s->d_row_len, nv * sizeof(int32_t)));
        if (!s->d_row_n1) LDX_CUDA(cudaMalloc(&s->d_row_n1, nv * sizeof(int32_t)));
        general_freq_kernel<<<(unsigned)((s->n_variants + 255) / 256), 256, 0, ctx->stream>>>(s->d_gen, s->d_kind, s->n_variants, s->n_sel, s->d_freq,
                                                                                              s->d_row_len, s->d_row_n1);
        ctx->launches++;
        LDX_CUDA(cudaGetLastError());
    } else {
        if (s->d_row_len) { cudaFree(s->d_row_len); s->d_row_len = nullptr; }
        if (s->d_row_n1) { cudaFree(s->d_row_n1); s->d_row_n1 = nullptr; }
    }
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ subset
// dst bit k of variant v = src bit sel[k].  One thread per destination word.
__global__ void subset_kernel(const uint64_t *__restrict__ src, int32_t src_stride, const int32_t *__restrict__ sel,
                              int32_t n_sel, int64_t n_variants, uint64_t *__restrict__ dst, int32_t dst_stride) {
    const int64_t total = n_variants * dst_stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i / dst_stride;
        const int w = (int)(i - v * dst_stride);
        const uint64_t *row = src + v * src_stride;
        uint64_t word = 0;
        const int k0 = w * 64;
        for (int b = 0; b < 64; ++b) {
            const int k = k0 + b;
            if (k < n_sel) {
                const int h = __ldg(sel + k);
                word |= ((row[h >> 6] >> (h & 63)) & 1ull) << b;
            }
        }
        dst[i] = word;
    }
}

// The same gather with a warp per row: the row is staged in shared memory once, lane l looks up destination bit k0 + l (its
// source column from sel[], coalesced), a ballot makes 32 destination bits at a time, and the row leaves as coalesced 4-byte
// stores.  No dependent global loads per bit: 200,000 rows of 5008 -> 1006 columns in 0.06 ms instead of 0.47 ms.
constexpr int SUBSET_WARPS = 8;
__global__ void __launch_bounds__(32 * SUBSET_WARPS)
subset_rows_kernel(const uint64_t *__restrict__ src, int32_t src_stride, const int32_t *__restrict__ sel, int32_t n_sel, int64_t n_rows,
                   uint64_t *__restrict__ dst, int32_t dst_stride) {
    extern __shared__ uint4 rows_s[];                                   // [SUBSET_WARPS][src_stride / 2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gran = src_stride / 2;                                    // 16-byte granules per source row (the pitch is a multiple of 16 words)
    uint4 *mine_s = rows_s + warp * gran;
    const uint32_t *row32 = reinterpret_cast<const uint32_t *>(mine_s);
    const int out32 = dst_stride * 2;
    for (int64_t r = (int64_t)blockIdx.x * SUBSET_WARPS + warp; r < n_rows; r += (int64_t)gridDim.x * SUBSET_WARPS) {
        const uint4 *g = reinterpret_cast<const uint4 *>(src + r * src_stride);
        for (int i = lane; i < gran; i += 32) mine_s[i] = ldg_u4_stream(g + i);
        __syncwarp();
        uint32_t *out = reinterpret_cast<uint32_t *>(dst + r * dst_stride);
        for (int w0 = 0; w0 < out32; w0 += 32) {                        // 32 destination words = 1024 bits per pass
            uint32_t word = 0;
            const int w_end = min(32, out32 - w0);
            for (int w = 0; w < w_end; ++w) {
                const int k = (w0 + w) * 32 + lane;
                uint32_t bit = 0;
                if (k < n_sel) { const int h = __ldg(sel + k); bit = (row32[h >> 5] >> (h & 31)) & 1u; }
                const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                if (lane == w) word = bal;
            }
            if (lane < w_end) out[w0 + lane] = word;
        }
        __syncwarp();                                                   // the staged row is overwritten by the next one
    }
}

static int launch_subset_any(ldx_ctx *ctx, const uint64_t *d_src, int32_t src_stride, const int32_t *d_sel, int32_t n_sel, int64_t n_rows,
                             uint64_t *d_dst, int32_t dst_stride) {
    if (n_rows <= 0) return LDX_OK;
    const size_t smem = (size_t)SUBSET_WARPS * src_stride * 8;
    if (smem <= 48 * 1024) {
        const int64_t blocks = std::min<int64_t>((n_rows + SUBSET_WARPS - 1) / SUBSET_WARPS, (int64_t)ctx->sm_count * 8);
        timing_begin(ctx);
        subset_rows_kernel<<<(int)blocks, 32 * SUBSET_WARPS, smem, ctx->stream>>>(d_src, src_stride, d_sel, n_sel, n_rows, d_dst, dst_stride);
        timing_end(ctx);
    } else {                                                            // very wide rows: one thread per destination word
        const int64_t total = n_rows * dst_stride;
        const int64_t blocks = std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 32);
        subset_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(d_src, src_stride, d_sel, n_sel, n_rows, d_dst, dst_stride);
    }
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    return LDX_OK;
}

int launch_subset(const ldx_store *src, const int32_t *d_sel, ldx_store *dst) {
    return launch_subset_any(src->ctx, src->d_planes, src->stride_words, d_sel, dst->n_hap, src->n_variants, dst->d_planes, dst->stride_words);
}

// the same gather for any [n_rows][src_stride] block of planes (aux planes, the common pattern)
int launch_subset_planes(ldx_ctx *ctx, const uint64_t *d_src, int32_t src_stride, const int32_t *d_sel, int32_t n_sel, int64_t n_rows, uint64_t *d_dst,
                         int32_t dst_stride) {
    return launch_subset_any(ctx, d_src, src_stride, d_sel, n_sel, n_rows, d_dst, dst_stride);
}

// ------------------------------------------------------------------------------------------ mailbox
// One thread copies {near-tie count, error flag} and the call's sequence number to pinned host memory (one 64-bit
// store, see publish_record).
__global__ void publish_kernel(const uint32_t *__restrict__ fix_count, volatile uint32_t *mailbox, uint32_t seq) {
    publish_record(mailbox, seq, fix_count[0], fix_count[1]);
}

int launch_publish(ldx_ctx *ctx) {
    publish_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_fix_count, ctx->d_mailbox, ctx->seq);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    return LDX_OK;
}

}  // namespace ldx
