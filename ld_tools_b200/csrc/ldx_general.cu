// ldx_general.cu -- the store side of the general route: genotype rows that are not complete
// phased diploid 0/1 rows.
//
// The reference builds a variant's genotype list with `+= rec.samples[name]['GT']` (ld_area.py:182-187, :230-235;
// ld_triangle.py:158-186; ld_lite.py:118-123) -- pysam hands it a tuple per sample: (0, 1) for "0|1" or "0/1", (1,) for a
// haploid "1", (None, 1) for ".|1", (2, 0) for "2|0" -- and calc_ld counts over those lists (calc_ld.py:30-40).  K1's fast
// kernel (pack_gt_kernel) flags every row that has a field outside the plain alphabet; this file
//   * parses those rows in full (pack_gt_general_kernel): any ploidy up to 2, '.', allele codes other than 0 / 1, '/' or '|',
//     sub-fields after ':' -- into the alt plane plus two aux planes (present, ref);
//   * chooses the store's common presence pattern (that of its middle row: all slots for an autosome, "males haploid" for the
//     non-pseudo-autosomal bulk of chrX) and classifies every row against it (classify_rows_kernel): rows that match it with
//     0/1 alleles only stay on the fast paths, under the selection mask ANDed with the pattern; the others carry a negative
//     VarFreq.n1 and every pair with one of them takes general_pair_counts / finalise_general (ldx_common.cuh).
#include <algorithm>
#include <vector>

#include "ldx_internal.h"

#define LDX_TRY(expr) do { int rc__ = (expr); if (rc__ != LDX_OK) return rc__; } while (0)

namespace ldx {

constexpr int GEN_THREADS = 256;

// allele token at text[i ..): returns the code (0 ref, 1 alt, 2 other) and advances i past it; an empty token is "other"
__device__ __forceinline__ int parse_allele(const uint8_t *t, int &i, int end) {
    if (i >= end) return 2;
    const uint8_t c = t[i];
    if (c == '.') { ++i; return 2; }
    if (c < '0' || c > '9') return 2;                       // not a token at all: caller sees no progress and flags the row
    int code = c == '0' ? 0 : c == '1' ? 1 : 2;
    ++i;
    while (i < end && t[i] >= '0' && t[i] <= '9') { code = 2; ++i; }     // "10", "01": pysam's int() of it is neither 0 nor 1 ("01" does not occur)
    return code;
}

__global__ void __launch_bounds__(GEN_THREADS)
pack_gt_general_kernel(const uint8_t *__restrict__ text, int64_t text_bytes, const int64_t *__restrict__ row_off, int64_t row_pitch,
                       const int64_t *__restrict__ rows_idx, int64_t n, int32_t n_samples, uint64_t *__restrict__ planes, int32_t stride_words,
                       uint64_t *__restrict__ aux_out, uint8_t *__restrict__ status_out, int32_t max_len) {
    extern __shared__ __align__(16) uint8_t smem_g[];
    unsigned long long *p_alt = reinterpret_cast<unsigned long long *>(smem_g);
    unsigned long long *p_ref = p_alt + stride_words, *p_pre = p_ref + stride_words;
    uint8_t *row = reinterpret_cast<uint8_t *>(p_pre + stride_words);
    __shared__ int s_end, s_bad, s_scan[GEN_THREADS];
    const int tid = threadIdx.x;
    for (int64_t k = blockIdx.x; k < n; k += gridDim.x) {
        const int64_t r = rows_idx[k];
        const int64_t off = row_off ? row_off[r] : r * row_pitch;
        const int avail = (int)min((int64_t)max_len, text_bytes - off);
        if (tid == 0) { s_end = avail; s_bad = 0; }
        for (int w = tid; w < 3 * stride_words; w += GEN_THREADS) p_alt[w] = 0;
        __syncthreads();
        // the row's text -> shared memory; its end = the first newline
        for (int i = tid; i < avail; i += GEN_THREADS) {
            const uint8_t c = text[off + i];
            row[i] = c;
            if (c == '\n' || c == '\r') atomicMin(&s_end, i);
        }
        __syncthreads();
        const int end = s_end;
        // field j starts after the j-th tab: tabs per thread segment, scanned
        const int seg = (end + GEN_THREADS - 1) / GEN_THREADS, a = min(tid * seg, end), b = min(a + seg, end);
        int tabs = 0;
        for (int i = a; i < b; ++i) tabs += row[i] == '\t';
        s_scan[tid] = tabs;
        __syncthreads();
        for (int d = 1; d < GEN_THREADS; d <<= 1) {
            const int v = tid >= d ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        int j = s_scan[tid] - tabs;                          // fields that start before this segment's first tab: index of the next field = j + 1
        const int total_fields = s_scan[GEN_THREADS - 1] + 1;
        int bad = 0;
        for (int i = tid == 0 ? -1 : a; i < b; ++i) {        // a thread takes the fields that start after the tabs of its segment; i = -1: field 0
            if (i >= 0 && row[i] != '\t') continue;
            const int f = i == -1 ? 0 : ++j;
            if (f >= n_samples) break;
            int p = i + 1;
            const int p0 = p;
            const int a0 = parse_allele(row, p, end);
            if (p == p0) { bad = 1; continue; }                                    // empty / non-numeric field
            unsigned long long bit = 1ull << ((2 * f) & 63);
            int w = (2 * f) >> 6;
            atomicOr(&p_pre[w], bit);
            if (a0 == 1) atomicOr(&p_alt[w], bit); else if (a0 == 0) atomicOr(&p_ref[w], bit);
            if (p < end && (row[p] == '|' || row[p] == '/')) {                    // a second allele: diploid
                ++p;
                const int p1 = p;
                const int a1 = parse_allele(row, p, end);
                if (p == p1) bad = 1;
                bit = 1ull << ((2 * f + 1) & 63); w = (2 * f + 1) >> 6;
                atomicOr(&p_pre[w], bit);
                if (a1 == 1) atomicOr(&p_alt[w], bit); else if (a1 == 0) atomicOr(&p_ref[w], bit);
                if (p < end && (row[p] == '|' || row[p] == '/')) bad = 1;          // ploidy > 2: outside the store's two slots per sample
            }
            if (p < end && row[p] != ':' && row[p] != '\t') bad = 1;               // GT is the first sub-field; anything else ends it
        }
        if (bad || (tid == 0 && total_fields < n_samples)) atomicOr(&s_bad, 1);
        __syncthreads();
        const bool rejected = s_bad != 0;                        // such a row is left empty (no slot present): every pair with it is int 0 / int 0
        for (int w = tid; w < stride_words; w += GEN_THREADS) {
            planes[r * (int64_t)stride_words + w] = rejected ? 0ull : p_alt[w];
            aux_out[(k * 2 + 0) * (int64_t)stride_words + w] = rejected ? 0ull : p_pre[w];
            aux_out[(k * 2 + 1) * (int64_t)stride_words + w] = rejected ? 0ull : p_ref[w];
        }
        if (tid == 0) status_out[k] = s_bad ? 8 : 0;
        __syncthreads();
    }
}

int launch_pack_gt_general(ldx_ctx *ctx, const uint8_t *d_text, int64_t text_bytes, const int64_t *d_row_off, int64_t row_pitch, const int64_t *d_rows_idx,
                           int64_t n, int32_t n_samples, uint64_t *d_planes_first, int32_t stride_words, uint64_t *d_aux_out, uint8_t *d_status_out) {
    if (n <= 0) return LDX_OK;
    // a GT field is at most "aa|bb" plus sub-fields; rows longer than this (very long sub-fields) are cut and flagged
    const int64_t max_len = std::min<int64_t>((int64_t)16 * n_samples + 64, 160 * 1024);
    const size_t smem = (size_t)3 * stride_words * 8 + (size_t)max_len + 16;
    if (smem > 200 * 1024) return set_error(LDX_ERR_ARG, "general genotype parser: row too long for shared memory");
    static bool attr_set[64] = {};
    if (smem > 48 * 1024 && !attr_set[ctx->device & 63]) {
        LDX_CUDA(cudaFuncSetAttribute(pack_gt_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[ctx->device & 63] = true;
    }
    const int grid = (int)std::min<int64_t>(n, (int64_t)ctx->sm_count * 4);
    pack_gt_general_kernel<<<grid, GEN_THREADS, smem, ctx->stream>>>(d_text, text_bytes, d_row_off, row_pitch, d_rows_idx, n, n_samples, d_planes_first,
                                                                     stride_words, d_aux_out, d_status_out, (int32_t)max_len);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "pack_gt_general_kernel");
    return LDX_OK;
}

// Rows [first_row, first_row + n_rows) have just been packed by the fast kernel; h_status[r] bit 0 = row r has a field outside
// its alphabet.  Those rows get aux slots and are parsed again in full.
int store_pack_general(ldx_store *s, int64_t first_row, int64_t n_rows, const uint8_t *d_text, int64_t text_bytes, const int64_t *d_row_off,
                       int64_t row_pitch, int32_t n_samples, uint8_t *h_status) {
    ldx_ctx *ctx = s->ctx;
    // rows packed again give their old aux slots up
    for (int64_t &r : s->aux_rows)
        if (r >= first_row && r < first_row + n_rows) r = -1;
    std::vector<int64_t> flagged;
    for (int64_t r = 0; r < n_rows; ++r)
        if (h_status[r] & 1) flagged.push_back(r);
    s->classify_dirty = true;
    if (flagged.empty()) return LDX_OK;
    const int64_t n = (int64_t)flagged.size(), old = (int64_t)s->aux_rows.size();
    if (old + n > s->aux_capacity) {                                  // grow the aux planes (amortised doubling)
        const int64_t cap = std::max<int64_t>(old + n, 2 * s->aux_capacity);
        uint64_t *bigger = nullptr;
        if (cudaMalloc(&bigger, (size_t)cap * 2 * s->stride_words * sizeof(uint64_t)) != cudaSuccess) { cudaGetLastError(); return set_error(LDX_ERR_NOMEM, "store: aux planes"); }
        if (old) LDX_CUDA(cudaMemcpyAsync(bigger, s->d_aux, (size_t)old * 2 * s->stride_words * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(s->d_aux);
        s->d_aux = bigger; s->aux_capacity = cap;
    }
    int64_t *d_idx = nullptr; uint8_t *d_st = nullptr;
    LDX_CUDA(cudaMalloc(&d_idx, (size_t)n * sizeof(int64_t)));
    cudaError_t e = cudaMalloc(&d_st, (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, flagged.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    int rc = e == cudaSuccess ? (int)LDX_OK : cuda_fail(e, "general genotype parser: staging");
    if (rc == LDX_OK)
        rc = launch_pack_gt_general(ctx, d_text, text_bytes, d_row_off, row_pitch, d_idx, n, n_samples, s->d_planes + first_row * s->stride_words, s->stride_words,
                                    s->d_aux + old * 2 * s->stride_words, d_st);
    std::vector<uint8_t> st((size_t)n, 0);
    if (rc == LDX_OK && cudaMemcpyAsync(st.data(), d_st, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "general genotype parser: status");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == LDX_OK) rc = cuda_fail(cudaGetLastError(), "pack_gt_general_kernel");
    cudaFree(d_idx); cudaFree(d_st);
    if (rc != LDX_OK) return rc;
    for (int64_t k = 0; k < n; ++k) {
        s->aux_rows.push_back(first_row + flagged[(size_t)k]);
        h_status[flagged[(size_t)k]] |= st[(size_t)k];                 // bit 3: not even the general parser takes the row
    }
    return LDX_OK;
}

// kind[v] on entry: aux slot of a flagged row, -1 otherwise.  On exit: -1 simple, slot g >= 0, -2 = GEN_FULL.
// counters: [0] rows with kind != -1, [1] rows that keep their aux planes.
__global__ void __launch_bounds__(256)
classify_rows_kernel(const uint64_t *__restrict__ planes, const uint64_t *__restrict__ aux, const uint64_t *__restrict__ common,
                     const uint64_t *__restrict__ all_slots, int32_t stride_words, int32_t words, int64_t n_variants, int32_t *__restrict__ kind,
                     unsigned long long *__restrict__ counters) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_variants) return;
    const int32_t g = kind[v];
    bool simple = true;
    if (g >= 0) {
        const uint64_t *pre = aux + (int64_t)g * 2 * stride_words, *ref = pre + stride_words, *alt = planes + v * stride_words;
        for (int w = 0; w < words && simple; ++w) simple = pre[w] == common[w] && ref[w] == (~alt[w] & common[w]);
        kind[v] = simple ? -1 : g;
    } else {
        for (int w = 0; w < words && simple; ++w) simple = common[w] == all_slots[w];
        kind[v] = simple ? -1 : -2;
    }
    if (!simple) {
        atomicAdd(&counters[0], 1ull);
        if (g >= 0) atomicAdd(&counters[1], 1ull);
    }
}

__global__ void copy_plane_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, int32_t n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

int store_publish_gen(ldx_store *s) {
    if (!s->d_gen) {
        if (cudaMalloc(&s->d_gen, sizeof(GenStore)) != cudaSuccess) { cudaGetLastError(); return set_error(LDX_ERR_NOMEM, "store: general-route descriptor"); }
    }
    GenStore g;
    g.planes = s->d_planes; g.aux = s->d_aux; g.mask_user = s->d_mask_user; g.common = s->d_common; g.all_slots = s->d_all_slots;
    g.stride_words = s->stride_words; g.words = s->words;
    LDX_CUDA(cudaMemcpyAsync(s->d_gen, &g, sizeof g, cudaMemcpyHostToDevice, s->ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(s->ctx->stream));             // `g` is a local
    return LDX_OK;
}

// The rows parsed by the general kernel sit in s->aux_rows (aux slot k belongs to row aux_rows[k]).  Chooses the common
// pattern (the middle row's presence) and classifies every row.
int store_classify_rows(ldx_store *s) {
    ldx_ctx *ctx = s->ctx;
    s->classify_dirty = false;
    s->n_general = 0; s->n_nonsimple = 0;
    const int64_t n_aux = (int64_t)s->aux_rows.size();
    // without flagged rows every row is a plain diploid row: common = all slots, everything simple (unless a pattern was loaded)
    if (n_aux == 0 && !s->common_loaded) {
        if (s->d_kind) { cudaFree(s->d_kind); s->d_kind = nullptr; }
        LDX_CUDA(cudaMemcpyAsync(s->d_common, s->d_all_slots, sizeof(uint64_t) * s->stride_words, cudaMemcpyDeviceToDevice, ctx->stream));
        s->h_common.resize((size_t)s->stride_words);
        LDX_CUDA(cudaMemcpyAsync(s->h_common.data(), s->d_all_slots, sizeof(uint64_t) * s->stride_words, cudaMemcpyDeviceToHost, ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));
        return LDX_OK;
    }
    const size_t nv = (size_t)std::max<int64_t>(s->n_variants, 1);
    if (!s->d_kind) LDX_CUDA(cudaMalloc(&s->d_kind, nv * sizeof(int32_t)));
    std::vector<int32_t> kind(nv, -1);
    int64_t mid_slot = -1;
    const int64_t mid = s->n_variants / 2;
    for (int64_t k = 0; k < n_aux; ++k) {
        if (s->aux_rows[(size_t)k] < 0) continue;                      // a slot given up by a re-packed row
        kind[(size_t)s->aux_rows[(size_t)k]] = (int32_t)k;
        if (s->aux_rows[(size_t)k] == mid) mid_slot = k;
    }
    LDX_CUDA(cudaMemcpyAsync(s->d_kind, kind.data(), nv * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (!s->common_loaded) {
        const uint64_t *src = mid_slot >= 0 ? s->d_aux + mid_slot * 2 * s->stride_words : s->d_all_slots;
        LDX_CUDA(cudaMemcpyAsync(s->d_common, src, sizeof(uint64_t) * s->stride_words, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    unsigned long long *d_cnt;
    LDX_CUDA(cudaMalloc(&d_cnt, 2 * sizeof(unsigned long long)));
    LDX_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), ctx->stream));
    classify_rows_kernel<<<(unsigned)((s->n_variants + 255) / 256), 256, 0, ctx->stream>>>(s->d_planes, s->d_aux, s->d_common, s->d_all_slots, s->stride_words,
                                                                                          s->words, s->n_variants, s->d_kind, d_cnt);
    ctx->launches++;
    unsigned long long h_cnt[2] = {0, 0};
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);        // `kind` is a local, too
    cudaFree(d_cnt);
    if (e != cudaSuccess) return cuda_fail(e, "classify_rows_kernel");
    s->n_nonsimple = (int64_t)h_cnt[0];
    s->n_general = (int64_t)h_cnt[1];
    if (s->n_nonsimple == 0) { cudaFree(s->d_kind); s->d_kind = nullptr; }
    s->h_common.resize((size_t)s->stride_words);
    LDX_CUDA(cudaMemcpy(s->h_common.data(), s->d_common, sizeof(uint64_t) * s->stride_words, cudaMemcpyDeviceToHost));
    return store_publish_gen(s);
}

}  // namespace ldx
