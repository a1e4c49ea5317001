// ldx_triangle_popc.cu -- K5': all-pairs LD of a variant list by AND + POPC on bit planes.
//
// Replaces the double loop at ld_triangle.py:133-230: for row > col, var_1 = rows[row],
// var_2 = rows[col] (ld_triangle.py:193); output = lower triangle packed by rows.
//
// Roofline: the integer POPC pipe.  A pair costs ceil(n_hap/32) 32-bit POPCs (157 at 5008
// haplotypes); HBM traffic is negligible because a 64 x 64 tile of pairs re-uses its 128 rows
// from shared memory.  This is the reference engine for the tensor-core path
// (ldx_triangle_mma.cu), which must reproduce its counts bit for bit, and the engine of choice
// when the tcgen05 path is unavailable.
//
// Tiling: one CTA per 64 x 64 tile of the lower triangle (tiles with bi >= bj), 256 threads as a
// 16 x 16 grid, 4 x 4 pairs per thread.  Rows are staged in shared memory as [variant][word]
// with a pitch of W+2 words, which spreads the 16 column groups of a warp over distinct banks.
// The mask is folded into the row tile while staging.  Words are processed in slabs of at most
// 80 so that any haplotype count fits the same shared-memory footprint.
#include "ldx_internal.h"
#include "ldx_fixup.cuh"

namespace ldx {

constexpr int TRI_TILE = 64;
constexpr int TRI_THREADS = 256;
constexpr int TRI_SLAB = 80;               // words per slab (5008 haplotypes = one slab)
constexpr int TRI_PITCH = TRI_SLAB + 2;    // u64 units

struct TriArgs {
    const uint64_t *planes; const uint64_t *mask; int32_t stride_words;
    const VarFreq *freq; FinalCtx fc;
    const int64_t *rows; int64_t v; int64_t n_tiles_side;
    int64_t tile_begin, out_off;   // first tile (triangular order) and first packed index of the call's row range
    int measure, has_thres, thres_e4;
    uint32_t *packed; int32_t *n11;
    FixupSink fix;
    const GenStore *gen;           // the general route, or nullptr
};

__device__ __forceinline__ void tile_coords(int64_t t, int64_t &bi, int64_t &bj) {
    // t = bi*(bi+1)/2 + bj with 0 <= bj <= bi
    bi = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (bi * (bi + 1) / 2 > t) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    bj = t - bi * (bi + 1) / 2;
}

__global__ void __launch_bounds__(TRI_THREADS)
triangle_popc_kernel(const TriArgs A) {
    extern __shared__ __align__(16) uint64_t smem[];
    uint64_t *sa = smem;                                  // [64][PITCH] row variants (mask folded in)
    uint64_t *sb = smem + TRI_TILE * TRI_PITCH;           // [64][PITCH] column variants
    __shared__ VarFreq fa_s[TRI_TILE], fb_s[TRI_TILE];
    __shared__ int64_t ra_s[TRI_TILE], rb_s[TRI_TILE];

    int64_t bi, bj;
    tile_coords(A.tile_begin + blockIdx.x, bi, bj);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t r0 = bi * TRI_TILE, c0 = bj * TRI_TILE;

    if (tid < TRI_TILE) {
        const int64_t r = r0 + tid;
        const int64_t sr = A.rows[r < A.v ? r : A.v - 1];
        ra_s[tid] = sr; fa_s[tid] = A.freq[sr];
    } else if (tid < 2 * TRI_TILE) {
        const int64_t c = c0 + tid - TRI_TILE;
        const int64_t sr = A.rows[c < A.v ? c : A.v - 1];
        rb_s[tid - TRI_TILE] = sr; fb_s[tid - TRI_TILE] = A.freq[sr];
    }
    __syncthreads();

    int cnt[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cnt[i][j] = 0;

    for (int w0 = 0; w0 < A.stride_words; w0 += TRI_SLAB) {
        const int wn = min(TRI_SLAB, A.stride_words - w0);   // multiple of 16
        const int gran = wn / 2;                              // 16-byte granules per row
        // stage both tiles: thread -> (variant, granule), granule fastest => coalesced rows
        for (int idx = tid; idx < TRI_TILE * gran; idx += TRI_THREADS) {
            const int vloc = idx / gran, g = idx - vloc * gran;
            const uint4 m = __ldg(reinterpret_cast<const uint4 *>(A.mask + w0) + g);
            uint4 a = __ldg(reinterpret_cast<const uint4 *>(A.planes + ra_s[vloc] * A.stride_words + w0) + g);
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(A.planes + rb_s[vloc] * A.stride_words + w0) + g);
            a.x &= m.x; a.y &= m.y; a.z &= m.z; a.w &= m.w;
            reinterpret_cast<uint4 *>(sa + vloc * TRI_PITCH)[g] = a;
            reinterpret_cast<uint4 *>(sb + vloc * TRI_PITCH)[g] = b;
        }
        __syncthreads();
        for (int g = 0; g < gran; ++g) {
            uint4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = reinterpret_cast<const uint4 *>(sa + (ty * 4 + i) * TRI_PITCH)[g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = reinterpret_cast<const uint4 *>(sb + (tx * 4 + j) * TRI_PITCH)[g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) cnt[i][j] += popc_and_u4(a[i], b[j]);
        }
        __syncthreads();
    }

    // epilogue: 16 pairs per thread; pair (r, c) is written iff r > c (strict lower triangle)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty * 4 + i;
        if (r >= A.v) continue;
        const VarFreq fa = fa_s[ty * 4 + i];
        const int64_t rbase = r * (r - 1) / 2 - A.out_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = c0 + tx * 4 + j;
            if (c >= r) continue;
            const VarFreq fb = fb_s[tx * 4 + j];
            const int64_t o = rbase + c;
            if (A.gen && (fa.n1 | fb.n1) < 0) {               // a variant of the general route
                const GenCounts gc = general_pair_counts(*A.gen, ra_s[ty * 4 + i], fa.n1, rb_s[tx * 4 + j], fb.n1);
                uint32_t word = finalise_general(gc).packed;
                if (A.has_thres && measure_e4(word, A.measure) < A.thres_e4) word |= LDX_BELOW_THRES;
                if (A.packed) {
                    A.packed[o] = word;
                    if (word & LDX_R2_NEARTIE) fixup_append_general(A.fix, (uint64_t)o, gc, word);
                }
                if (A.n11) A.n11[o] = gc.n11;
                continue;
            }
            const PairFinal f = finalise_pair(cnt[i][j], fa, fb, A.fc);   // var_1 = row, var_2 = col
            uint32_t word = f.packed;
            if (A.has_thres && measure_e4(word, A.measure) < A.thres_e4) word |= LDX_BELOW_THRES;
            if (A.packed) {
                A.packed[o] = word;
                if (word & LDX_R2_NEARTIE) fixup_append(A.fix, (uint64_t)o, cnt[i][j], fa.n1, fb.n1, word);
            }
            if (A.n11) A.n11[o] = cnt[i][j];
        }
    }
}

int launch_triangle_popc(ldx_store *s, const int64_t *d_rows, int64_t v, int64_t row_begin, int measure, int has_thres,
                         int thres_e4, uint32_t *d_packed, int32_t *d_n11) {
    if (v < 2) return LDX_OK;
    if (row_begin % TRI_TILE) return set_error(LDX_ERR_ARG, "popcount engine: row_begin must be a multiple of 64");
    ldx_ctx *ctx = s->ctx;
    TriArgs A;
    A.planes = s->d_planes; A.mask = s->d_mask; A.stride_words = s->stride_words;
    A.freq = s->d_freq; A.fc = s->fc; A.rows = d_rows; A.v = v;
    A.n_tiles_side = (v + TRI_TILE - 1) / TRI_TILE;
    A.measure = measure; A.has_thres = has_thres; A.thres_e4 = thres_e4;
    A.packed = d_packed; A.n11 = d_n11;
    A.fix = FixupSink{ctx->d_fix, ctx->d_fix_count, ctx->fix_capacity, ctx->fix_tag};
    A.gen = s->n_nonsimple > 0 ? s->d_gen : nullptr;
    const int64_t bi_begin = row_begin / TRI_TILE;
    A.tile_begin = bi_begin * (bi_begin + 1) / 2;
    A.out_off = row_begin * (row_begin - 1) / 2;
    const int64_t n_tiles = A.n_tiles_side * (A.n_tiles_side + 1) / 2 - A.tile_begin;
    if (n_tiles > 0x7fffffffll) return set_error(LDX_ERR_ARG, "triangle: too many tiles for one launch");
    const size_t smem = (size_t)2 * TRI_TILE * TRI_PITCH * sizeof(uint64_t);
    static bool attr_set[64] = {};            // a function attribute is per device: a process may hold contexts on several
    const int dv = ctx->device & 63;
    if (!attr_set[dv]) {
        LDX_CUDA(cudaFuncSetAttribute(triangle_popc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dv] = true;
    }
    timing_begin(ctx);
    triangle_popc_kernel<<<(int)n_tiles, TRI_THREADS, smem, ctx->stream>>>(A);
    timing_end(ctx);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "triangle_popc_kernel");
    return LDX_OK;
}

}  // namespace ldx
