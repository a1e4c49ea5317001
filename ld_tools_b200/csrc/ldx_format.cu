// ldx_format.cu -- the table writer of ld_triangle.py:351-360 at scale.
//
// The reference prints the V x V matrix with '\t'.join(map(str, ld_two_dim[row])): V^2 str() calls and one Python
// object per cell.  Here the packed lower triangle that the all-pairs kernels left in HBM is turned into that text
// on the GPU, byte for byte: a cell is
//     "0"                      above or on the diagonal (the matrix starts as int zeros, ld_triangle.py:114; only row > col is filled, :150),
//     "0"                      below the caller's threshold (:223-225) or where calc_ld returned its int 0 (:68-90),
//     str(round(x, 4))         otherwise -- and because the result word carries k = round(x, 4) * 10^4 as an integer,
//                              str() is digit arithmetic: k / 10^4 is the double nearest to the decimal k * 10^-4, whose
//                              shortest round-tripping representation (float_repr_style 'short') is that decimal with
//                              trailing zeros dropped and at least one fractional digit: "0.0", "0.5", "0.8125", "1.0".
// Line r = prefix[r] (rsID '\t' position '\t', supplied by the caller) + the cells joined by '\t' + '\n'.
//
// Byte work: 4 B read per lower-triangle cell, 2..7 B written per cell, two passes over the words (line lengths, then
// the text) with a device-wide scan of the line lengths in between.  HBM is the roofline by bytes (0.37 of the measured
// peak at 20,000 x 20,000 cells); what the text kernel is bound by is the byte stores into shared memory.
#include <algorithm>
#include <cstring>

#include <cub/cub.cuh>

#include "ldx_internal.h"

#define LDX_TRY(expr) do { int rc__ = (expr); if (rc__ != LDX_OK) return rc__; } while (0)
#define LDX_REQUIRE(cond, msg) do { if (!(cond)) return ldx::set_error(LDX_ERR_ARG, msg); } while (0)

namespace ldx {

int scratch_get(ldx_ctx *ctx, int which, size_t bytes, void **out);      // ldx_api.cu
void scratch_trim(ldx_ctx *ctx, size_t keep_bytes);

constexpr int FMT_THREADS = 256, FMT_CELLS_PER_THREAD = 8, FMT_CHUNK = FMT_THREADS * FMT_CELLS_PER_THREAD;
constexpr int FMT_CELL_MAX = 7;   // "1.6383" + separator

// str(k / 10000.0) for 0 <= k < 20000: characters in the low bytes (first character lowest), length in *len.
__host__ __device__ __forceinline__ uint64_t text_of_e4(uint32_t k, int *len) {
    const uint32_t ip = k / 10000u, fr = k - ip * 10000u;
    const uint32_t d1 = fr / 1000u, r1 = fr - d1 * 1000u, d2 = r1 / 100u, r2 = r1 - d2 * 100u, d3 = r2 / 10u, d4 = r2 - d3 * 10u;
    *len = 2 + (d4 ? 4 : d3 ? 3 : d2 ? 2 : 1);
    return (uint64_t)('0' + ip) | ((uint64_t)'.' << 8) | ((uint64_t)('0' + d1) << 16) | ((uint64_t)('0' + d2) << 24) |
           ((uint64_t)('0' + d3) << 32) | ((uint64_t)('0' + d4) << 40);
}

// A matrix cell given its result word: "0" for the reference's int 0 (threshold, monomorphic, d' == 0), else the float.
// Branch-free (the digits are computed for every word and dropped for an int 0): the lanes of a warp never diverge on
// the data and the eight words of a thread can be loaded ahead of their use.  dp_shift = 0 / 16 and int0_mask select
// the measure.
__device__ __forceinline__ void text_of_word(uint32_t w, int dp_shift, uint32_t int0_mask, uint32_t *lo, uint32_t *hi, int *len) {
    int n;
    const uint64_t t = text_of_e4((w >> dp_shift) & LDX_R2_MASK, &n);
    const bool int0 = (w & int0_mask) != 0;
    *lo = int0 ? (uint32_t)'0' : (uint32_t)t;
    *hi = int0 ? 0u : (uint32_t)(t >> 32);
    *len = int0 ? 1 : n;
}

__device__ __forceinline__ int64_t tri64(int64_t r) { return r * (r - 1) / 2; }

// Pass 1: bytes of every line.  One CTA per line (grid-stride); cells at or above the diagonal are two bytes each.
__global__ void __launch_bounds__(FMT_THREADS)
matrix_line_bytes_kernel(const uint32_t *__restrict__ packed, int64_t v, int64_t row_begin, int64_t n_lines, int measure,
                         const int64_t *__restrict__ prefix_off, int64_t *__restrict__ line_bytes) {
    using Reduce = cub::BlockReduce<int, FMT_THREADS>;
    __shared__ typename Reduce::TempStorage tmp;
    const int dp_shift = measure == LDX_MEASURE_R2 ? 0 : LDX_DP_SHIFT;
    const uint32_t int0_mask = LDX_BELOW_THRES | (measure == LDX_MEASURE_R2 ? LDX_R2_INT0 : LDX_DP_INT0);
    for (int64_t l = blockIdx.x; l < n_lines; l += gridDim.x) {
        const int64_t r = row_begin + l;
        const uint32_t *__restrict__ words = packed + (tri64(r) - tri64(row_begin));
        int sum = 0, c = threadIdx.x;
        for (; c + 3 * FMT_THREADS < (int)r; c += 4 * FMT_THREADS) {           // four loads in flight per thread
            uint32_t w[4], lo, hi;
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = __ldg(words + c + k * FMT_THREADS);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int len;
                text_of_word(w[k], dp_shift, int0_mask, &lo, &hi, &len);
                sum += len + 1;
            }
        }
        for (; c < (int)r; c += FMT_THREADS) {
            uint32_t lo, hi;
            int len;
            text_of_word(__ldg(words + c), dp_shift, int0_mask, &lo, &hi, &len);
            sum += len + 1;
        }
        const int total = Reduce(tmp).Sum(sum);
        if (threadIdx.x == 0) line_bytes[l] = (prefix_off[r + 1] - prefix_off[r]) + (int64_t)total + 2 * (v - r);
        __syncthreads();
    }
}

// One cell into the chunk buffer: up to six text bytes, then the tab at dst[n].  With `exact` false the six byte stores are
// unconditional -- what they write beyond the cell's length lands in LATER cells of the same thread (the caller checks
// that dst + 6 stays inside the thread's own bytes) and is overwritten by them in program order; cells at the end of a
// thread's run must not touch the neighbour's bytes and take the predicated form.
__device__ __forceinline__ unsigned char *put_cell(unsigned char *dst, uint32_t lo, uint32_t hi, int n, bool exact) {
    if (!exact) {
        dst[0] = (unsigned char)lo; dst[1] = (unsigned char)(lo >> 8); dst[2] = (unsigned char)(lo >> 16);
        dst[3] = (unsigned char)(lo >> 24); dst[4] = (unsigned char)hi; dst[5] = (unsigned char)(hi >> 8);
    } else {
        dst[0] = (unsigned char)lo;
        if (n > 1) { dst[1] = (unsigned char)(lo >> 8); dst[2] = (unsigned char)(lo >> 16); }       // a float has at least "0.0"
        if (n > 3) dst[3] = (unsigned char)(lo >> 24);
        if (n > 4) dst[4] = (unsigned char)hi;
        if (n > 5) dst[5] = (unsigned char)(hi >> 8);
    }
    dst[n] = '\t';
    return dst + n + 1;
}

// Pass 2: the text.  A CTA formats one line.  Cells left of the diagonal, in chunks of 2,048: every thread builds its
// eight cells in registers, a block scan places them, the chunk is assembled in shared memory at the alignment it will
// have in global memory and copied out with 16-byte stores.  The diagonal and everything right of it is the constant
// pattern "0\t0\t...0\n", stored directly.
__global__ void __launch_bounds__(FMT_THREADS)
matrix_text_kernel(const uint32_t *__restrict__ packed, int64_t v, int64_t row_begin, int64_t n_lines, int measure,
                   const char *__restrict__ prefixes, const int64_t *__restrict__ prefix_off,
                   const int64_t *__restrict__ line_off, char *__restrict__ text) {
    using Scan = cub::BlockScan<int, FMT_THREADS>;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ __align__(16) unsigned char buf[FMT_CHUNK * FMT_CELL_MAX + 32];
    const int tid = threadIdx.x;
    const int dp_shift = measure == LDX_MEASURE_R2 ? 0 : LDX_DP_SHIFT;
    const uint32_t int0_mask = LDX_BELOW_THRES | (measure == LDX_MEASURE_R2 ? LDX_R2_INT0 : LDX_DP_INT0);
    for (int64_t l = blockIdx.x; l < n_lines; l += gridDim.x) {
        const int64_t r = row_begin + l;
        char *out = text + line_off[l];
        const int64_t p0 = prefix_off[r], plen = prefix_off[r + 1] - p0;
        for (int64_t i = tid; i < plen; i += FMT_THREADS) out[i] = prefixes[p0 + i];
        int64_t pos = plen;
        const uint32_t *__restrict__ words = packed + (tri64(r) - tri64(row_begin));
        for (int cb = 0; cb < (int)r; cb += FMT_CHUNK) {
            uint32_t w[FMT_CELLS_PER_THREAD], lo[FMT_CELLS_PER_THREAD], hi[FMT_CELLS_PER_THREAD];
            int len[FMT_CELLS_PER_THREAD], mine = 0;
            const int c0 = cb + tid * FMT_CELLS_PER_THREAD;
            const int n_valid = min(max((int)r - c0, 0), FMT_CELLS_PER_THREAD);       // this thread's cells left of the diagonal
#pragma unroll
            for (int k = 0; k < FMT_CELLS_PER_THREAD; ++k) w[k] = k < n_valid ? __ldg(words + c0 + k) : 0u;
#pragma unroll
            for (int k = 0; k < FMT_CELLS_PER_THREAD; ++k) {
                text_of_word(w[k], dp_shift, int0_mask, &lo[k], &hi[k], &len[k]);
                if (k >= n_valid) len[k] = 0;
                mine += k < n_valid ? len[k] + 1 : 0;
            }
            int off, total;
            Scan(tmp).ExclusiveSum(mine, off, total);
            const int shift = (int)(reinterpret_cast<uintptr_t>(out + pos) & 15);
            unsigned char *dst = buf + shift + off;
            const unsigned char *const end = dst + mine;             // where this thread's bytes stop and its neighbour's start
#pragma unroll
            for (int k = 0; k < FMT_CELLS_PER_THREAD; ++k)
                if (len[k]) dst = put_cell(dst, lo[k], hi[k], len[k], k + 1 == FMT_CELLS_PER_THREAD || dst + 6 > end);
            __syncthreads();
            char *g = out + pos;
            const int head = shift ? min(16 - shift, total) : 0;
            if (tid < head) g[tid] = (char)buf[shift + tid];
            const int nvec = (total - head) >> 4;
            const uint4 *src = reinterpret_cast<const uint4 *>(buf + shift + head);     // shift + head is 0 or 16 when nvec > 0
            uint4 *gv = reinterpret_cast<uint4 *>(g + head);
            for (int i = tid; i < nvec; i += FMT_THREADS) gv[i] = src[i];
            const int done = head + (nvec << 4);
            if (tid < total - done) g[done + tid] = (char)buf[shift + done + tid];
            pos += total;
            __syncthreads();
        }
        // cells r .. v-1: "0\t" each, the last one "0\n"
        {
            char *g = out + pos;
            const int64_t total = 2 * (v - r);
            const int shift = (int)(reinterpret_cast<uintptr_t>(g) & 15);
            const int64_t head = shift ? min((int64_t)(16 - shift), total - 1) : 0;
            if (tid < head) g[tid] = (tid & 1) ? '\t' : '0';
            const int64_t nvec = (total - 1 - head) >> 4;                                // the final byte is never part of a vector
            const uint32_t w = (head & 1) ? 0x30093009u : 0x09300930u;
            uint4 *gv = reinterpret_cast<uint4 *>(g + head);
            for (int64_t i = tid; i < nvec; i += FMT_THREADS) gv[i] = make_uint4(w, w, w, w);
            const int64_t done = head + (nvec << 4);
            if (tid < total - done) g[done + tid] = done + tid == total - 1 ? '\n' : (((done + tid) & 1) ? '\t' : '0');
        }
    }
}

}  // namespace ldx

using namespace ldx;

extern "C" int32_t ldx_format_e4(int32_t value_e4, char *out8) {
    LDX_REQUIRE(out8 && value_e4 >= 0 && value_e4 < 20000, "bad argument");
    int len;
    const uint64_t s = text_of_e4((uint32_t)value_e4, &len);
    for (int b = 0; b < 8; ++b) out8[b] = b < len ? (char)(s >> (8 * b)) : '\0';
    return LDX_OK;
}

extern "C" int32_t ldx_triangle_text(ldx_ctx *ctx, const uint32_t *packed, int64_t v, int64_t row_begin, int64_t row_end,
                                     int32_t measure, const char *prefixes, const int64_t *prefix_off, int32_t flags,
                                     char *text, int64_t cap, int64_t *n_bytes) {
    LDX_REQUIRE(ctx && n_bytes, "NULL argument");
    *n_bytes = 0;
    LDX_REQUIRE(measure == LDX_MEASURE_R2 || measure == LDX_MEASURE_DPRIME, "bad measure");
    LDX_REQUIRE(v >= 0 && v < (1ll << 31) && row_begin >= 0 && row_begin <= row_end && row_end <= v, "bad row range");
    LDX_REQUIRE(prefix_off && (flags & ~3) == 0 && cap >= 0 && (text || cap == 0), "bad argument");
    const int64_t n_lines = row_end - row_begin;
    if (n_lines == 0) return LDX_OK;
    LDX_REQUIRE(prefix_off[0] == 0, "prefix_off[0] must be 0");
    for (int64_t r = 0; r < v; ++r) LDX_REQUIRE(prefix_off[r + 1] >= prefix_off[r], "prefix_off must not decrease");
    const int64_t prefix_bytes = prefix_off[v];
    LDX_REQUIRE(prefixes || prefix_bytes == 0, "prefixes is NULL");
    const int64_t n_words = (row_end > 1 ? row_end * (row_end - 1) / 2 : 0) - (row_begin > 1 ? row_begin * (row_begin - 1) / 2 : 0);
    LDX_REQUIRE(packed || n_words == 0, "packed is NULL");
    const bool packed_on_device = flags & LDX_TEXT_PACKED_ON_DEVICE, text_on_device = flags & LDX_TEXT_OUT_ON_DEVICE;
    LDX_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->pending_q.empty()) LDX_TRY(ldx_resolve(ctx, nullptr));     // device results are final only after the settlement
    cudaStream_t st = ctx->stream;

    // ---- scratch: line lengths / offsets, the prefixes, the scan's workspace; the words when they come from the host
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (int64_t *)nullptr, (int64_t *)nullptr, (int)(n_lines + 1), st);
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    uint8_t *base = nullptr;
    const size_t b_lines = pad(((size_t)n_lines + 1) * 8), b_poff = pad(((size_t)v + 1) * 8), b_pfx = pad((size_t)prefix_bytes + 16);
    LDX_TRY(scratch_get(ctx, 1, 2 * b_lines + b_poff + b_pfx + pad(cub_bytes) + 256, (void **)&base));
    int64_t *d_line_bytes = reinterpret_cast<int64_t *>(base), *d_line_off = reinterpret_cast<int64_t *>(base + b_lines);
    int64_t *d_poff = reinterpret_cast<int64_t *>(base + 2 * b_lines);
    char *d_pfx = reinterpret_cast<char *>(base + 2 * b_lines + b_poff);
    void *d_cub = base + 2 * b_lines + b_poff + b_pfx;
    const uint32_t *d_packed = packed;
    if (!packed_on_device && n_words) {
        uint32_t *d_up = nullptr;
        LDX_TRY(scratch_get(ctx, 2, (size_t)n_words * 4, (void **)&d_up));
        LDX_CUDA(cudaMemcpyAsync(d_up, packed, (size_t)n_words * 4, cudaMemcpyHostToDevice, st));
        d_packed = d_up;
    }
    LDX_CUDA(cudaMemcpyAsync(d_poff, prefix_off, ((size_t)v + 1) * 8, cudaMemcpyHostToDevice, st));
    if (prefix_bytes) LDX_CUDA(cudaMemcpyAsync(d_pfx, prefixes, (size_t)prefix_bytes, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemsetAsync(d_line_bytes + n_lines, 0, 8, st));

    // ---- pass 1 + scan: where every line starts, and the size of the text
    const unsigned grid = (unsigned)std::min<int64_t>(n_lines, (int64_t)ctx->sm_count * 8);
    matrix_line_bytes_kernel<<<grid, FMT_THREADS, 0, st>>>(d_packed, v, row_begin, n_lines, measure, d_poff, d_line_bytes);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "matrix_line_bytes_kernel");
    LDX_CUDA(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_line_bytes, d_line_off, (int)(n_lines + 1), st));
    int64_t total = 0;
    LDX_CUDA(cudaMemcpyAsync(&total, d_line_off + n_lines, 8, cudaMemcpyDeviceToHost, st));
    LDX_CUDA(cudaStreamSynchronize(st));
    *n_bytes = total;
    if (total > cap) return set_error(LDX_ERR_CAPACITY, "matrix text: buffer too small (*n_bytes holds the size needed)");

    // ---- pass 2: the text, straight into the caller's device buffer or through the arena to the host
    char *d_text = text;
    if (!text_on_device) LDX_TRY(scratch_get(ctx, 0, (size_t)total + 16, (void **)&d_text));
    timing_begin(ctx);
    matrix_text_kernel<<<grid, FMT_THREADS, 0, st>>>(d_packed, v, row_begin, n_lines, measure, d_pfx, d_poff, d_line_off, d_text);
    timing_end(ctx);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "matrix_text_kernel");
    if (!text_on_device) LDX_CUDA(cudaMemcpyAsync(text, d_text, (size_t)total, cudaMemcpyDeviceToHost, st));
    LDX_CUDA(cudaStreamSynchronize(st));
    scratch_trim(ctx, (size_t)256 << 20);
    return LDX_OK;
}
