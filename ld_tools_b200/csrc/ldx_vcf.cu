// ldx_vcf.cu -- GPU-side indexing of decompressed VCF text (the ingest path).
//
// The reference reaches a record's fields through pysam, one record at a time: rec.pos / rec.id / rec.ref /
// rec.info / rec.samples[name]['GT'] (ld_area.py:215-235, ld_triangle.py:128-186, prep_intgen_data.py:163-177).
// Here the whole decompressed <chrom>.vcf is uploaded once and four small kernels do what a per-line host loop
// would: (1) newline index, (2) one thread per line finds the nine fixed columns and parses what the fused
// ld_area filters need -- POS, len(REF), the rs number, the MULTI_ALLELIC key -- (3) the records are compacted
// and their annotations written straight into the store, (4) K1 (pack_gt_kernel) bit-packs the genotype
// columns from the same device copy of the text.  The host gets back one fixed-size record per variant (field
// offsets included, so that the text columns the writers print can be sliced lazily).
//
// All of it is byte work bound by the one pass over the text (10 KB per 2504-sample variant): the upload over
// PCIe is the limit, not the kernels.
#include <algorithm>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>

#include "ldx_internal.h"

#define LDX_TRY(expr) do { int rc__ = (expr); if (rc__ != LDX_OK) return rc__; } while (0)
#define LDX_REQUIRE(cond, msg) do { if (!(cond)) return ldx::set_error(LDX_ERR_ARG, msg); } while (0)

namespace ldx {

constexpr int NL_THREADS = 256, NL_BYTES_PER_THREAD = 64, NL_BLOCK_BYTES = NL_THREADS * NL_BYTES_PER_THREAD;

__device__ __forceinline__ int count_nl_64(const uint8_t *__restrict__ text, int64_t begin, int64_t n) {
    int c = 0;
    if (begin + NL_BYTES_PER_THREAD <= n) {                 // whole 64-byte run: four 16-byte loads (begin is 64-aligned)
        const uint4 *p = reinterpret_cast<const uint4 *>(text + begin);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 v = __ldg(p + g);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = w[k] ^ 0x0a0a0a0au;        // zero byte <=> newline
                c += __popc(((x - 0x01010101u) & ~x & 0x80808080u));
            }
        }
        // the borrow trick can flag the byte above a true zero byte (0x01 after 0x00): recount exactly when anything was found
        if (c) {
            c = 0;
            for (int i = 0; i < NL_BYTES_PER_THREAD; ++i) c += text[begin + i] == '\n';
        }
    } else {
        for (int64_t i = begin; i < n; ++i) c += text[i] == '\n';
    }
    return c;
}

__global__ void __launch_bounds__(NL_THREADS)
count_newlines_kernel(const uint8_t *__restrict__ text, int64_t n, uint32_t *__restrict__ block_counts) {
    using BlockReduce = cub::BlockReduce<int, NL_THREADS>;
    __shared__ typename BlockReduce::TempStorage tmp;
    const int64_t begin = (int64_t)blockIdx.x * NL_BLOCK_BYTES + (int64_t)threadIdx.x * NL_BYTES_PER_THREAD;
    const int c = begin < n ? count_nl_64(text, begin, n) : 0;
    const int total = BlockReduce(tmp).Sum(c);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)total;
}

__global__ void __launch_bounds__(NL_THREADS)
write_newlines_kernel(const uint8_t *__restrict__ text, int64_t n, const uint32_t *__restrict__ block_base, int64_t *__restrict__ nl) {
    using BlockScan = cub::BlockScan<int, NL_THREADS>;
    __shared__ typename BlockScan::TempStorage tmp;
    const int64_t begin = (int64_t)blockIdx.x * NL_BLOCK_BYTES + (int64_t)threadIdx.x * NL_BYTES_PER_THREAD;
    const int c = begin < n ? count_nl_64(text, begin, n) : 0;
    int base;
    BlockScan(tmp).ExclusiveSum(c, base);
    if (c) {
        int64_t *out = nl + block_base[blockIdx.x] + base;
        const int64_t end = begin + NL_BYTES_PER_THREAD < n ? begin + NL_BYTES_PER_THREAD : n;
        for (int64_t i = begin; i < end; ++i)
            if (text[i] == '\n') *out++ = i;
    }
}

// One thread per line.  Line k is text[start, end) with start = nl[k-1] + 1, end = nl[k] (the newline).
__global__ void __launch_bounds__(256)
parse_lines_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ nl, int64_t n_lines, int32_t n_samples,
                   ldx_vcf_row *__restrict__ tmp_rows, uint32_t *__restrict__ is_rec) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    const int64_t start = k ? nl[k - 1] + 1 : 0;
    int64_t end = nl[k];
    if (end > start && text[end - 1] == '\r') --end;
    if (end <= start || text[start] == '#') { is_rec[k] = 0; return; }
    ldx_vcf_row r;
    r.line_off = start;
    r.idnum = -1;
    r.status = 0; r.eligible = 0; r.multi = 0;
    r.pad[0] = r.pad[1] = r.pad[2] = r.pad[3] = r.pad[4] = 0;
    int32_t field_off[10];
    int nf = 1;
    field_off[0] = 0;
    for (int64_t i = start; i < end && nf < 10; ++i)
        if (text[i] == '\t') field_off[nf++] = (int32_t)(i + 1 - start);
    // a record line has the nine fixed columns and n_samples genotype fields of at least one character each (plain diploid
    // fields take 4 * n_samples - 1 bytes; haploid ones are shorter: K1 sorts that out)
    const int64_t need = nf == 10 ? (int64_t)field_off[9] + 2ll * n_samples - 1 : 0;
    if (nf < 10 || need > end - start) {
        // not a full record line: kept as a row (the drivers' row numbering follows the file) and flagged
        r.status = 2;
        for (int f = nf; f < 10; ++f) field_off[f] = (int32_t)(end - start);
    }
    r.id_off = field_off[2]; r.ref_off = field_off[3]; r.alt_off = field_off[4]; r.info_off = field_off[7];
    r.fmt_off = field_off[8]; r.gt_off = field_off[9];
    // POS
    int64_t pos = 0;
    for (int32_t i = field_off[1]; i < field_off[2] - 1; ++i) {
        const uint8_t c = text[start + i];
        if (c < '0' || c > '9' || pos > 214748363) { pos = 0; r.status |= 4; break; }
        pos = pos * 10 + (c - '0');
    }
    r.pos = (int32_t)pos;
    r.ref_len = max(field_off[4] - 1 - field_off[3], 0);
    // ID: rs\d+$ (ld_area.py:223, prep_intgen_data.py:166)
    {
        const int32_t a = field_off[2], b = field_off[3] - 1;
        bool rs = b - a >= 3 && b - a <= 20 && text[start + a] == 'r' && text[start + a + 1] == 's';
        int64_t num = 0;
        for (int32_t i = a + 2; rs && i < b; ++i) {
            const uint8_t c = text[start + i];
            if (c < '0' || c > '9') rs = false; else num = num * 10 + (c - '0');
        }
        if (rs) r.idnum = num;
        r.eligible = rs ? 1 : 0;
    }
    // INFO: the MULTI_ALLELIC key (ld_area.py:224 `'MULTI_ALLELIC' in rec.info`): a ';'-delimited key, with or without a value
    {
        const char key[] = "MULTI_ALLELIC";
        const int32_t a = field_off[7], b = field_off[8] - 1;
        bool at_key = true;
        for (int32_t i = a; i < b; ++i) {
            if (at_key && b - i >= 13) {
                bool m = true;
                for (int j = 0; j < 13; ++j) m &= text[start + i + j] == (uint8_t)key[j];
                if (m && (i + 13 == b || text[start + i + 13] == ';' || text[start + i + 13] == '=')) { r.multi = 1; break; }
            }
            at_key = text[start + i] == ';';
        }
        if (r.multi || r.status) r.eligible = 0;
    }
    // INFO: END=<n> extends the record's interval for the window fetch (ld_area.py:215-217 goes through the tabix index, and
    // htslib's tbx.c takes a VCF record's end from INFO/END when it has one -- structural variants -- else from len(REF);
    // an END at or before POS-1 is ignored there too).  ref_len carries the interval length either way.
    {
        const int32_t a = field_off[7], b = field_off[8] - 1;
        bool at_key = true;
        for (int32_t i = a; i + 4 < b + 1; ++i) {
            if (at_key && text[start + i] == 'E' && text[start + i + 1] == 'N' && text[start + i + 2] == 'D' && text[start + i + 3] == '=') {
                int64_t v = 0;
                int32_t j = i + 4;
                for (; j < b && text[start + j] >= '0' && text[start + j] <= '9' && v < 4000000000ll; ++j) v = v * 10 + (text[start + j] - '0');
                if (j > i + 4 && v > pos - 1 && v - (pos - 1) < 2147483647ll && !(r.status & 4)) r.ref_len = (int32_t)(v - (pos - 1));
                break;
            }
            at_key = text[start + i] == ';';
        }
    }
    tmp_rows[k] = r;
    is_rec[k] = 1;
}

// Record lines -> dense rows; the window annotations go straight into the store's device arrays.
__global__ void __launch_bounds__(256)
compact_rows_kernel(const ldx_vcf_row *__restrict__ tmp_rows, const uint32_t *__restrict__ is_rec, const uint32_t *__restrict__ rec_index,
                    int64_t n_lines, ldx_vcf_row *__restrict__ rows, int64_t *__restrict__ gt_abs, int32_t *__restrict__ pos0,
                    int32_t *__restrict__ end0, int64_t *__restrict__ idnum, uint8_t *__restrict__ eligible, int64_t row_base) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines || !is_rec[k]) return;
    const int64_t row = rec_index[k];                          // relative to this text's first record (store row = row_base + row)
    ldx_vcf_row r = tmp_rows[k];
    if (r.idnum < 0) r.idnum = -1 - (row_base + row);          // unique per row: never equal to a query's id (ld_area.py:222)
    rows[row] = r;
    gt_abs[row] = (r.status & 2) ? -1 : r.line_off + r.gt_off;
    pos0[row] = r.pos - 1;
    end0[row] = r.pos - 1 + r.ref_len;
    idnum[row] = r.idnum;
    eligible[row] = r.eligible;
}

// pack status (bit 0) and zeroed planes of malformed rows
__global__ void __launch_bounds__(256)
merge_status_kernel(ldx_vcf_row *__restrict__ rows, const uint8_t *__restrict__ pack_status, int64_t n_rows, uint8_t *__restrict__ eligible) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows || (rows[r].status & 2)) return;
    rows[r].status |= pack_status[r] & 9;      // bit 0: not plain "a|b" fields (general route), bit 3: no parser takes the genotypes
    if (pack_status[r] & 8) { rows[r].eligible = 0; eligible[r] = 0; }      // kept as a row, never paired by a window scan
}

int launch_pack_gt(ldx_ctx *ctx, const uint8_t *d_text, const int64_t *d_row_off, int64_t row_pitch, int64_t n_rows, int32_t n_samples,
                   uint64_t *d_planes_first, int32_t stride_words, uint8_t *d_status);
int store_alloc_annotations(ldx_store *s);                               // ldx_api.cu: pos0 / end0 / idnum / eligible of every row
int store_shrink(ldx_store *s, int64_t n_variants);                      // ldx_api.cu: keep the first n_variants rows
int scratch_get(ldx_ctx *ctx, int which, size_t bytes, void **out);      // ldx_api.cu: block `which` (0..2) of the context's arena
void scratch_trim(ldx_ctx *ctx, size_t keep_bytes);                      // ... and: free those blocks if larger than keep_bytes

// Scratch comes from the context's arena (three blocks, one per phase: their sizes depend on counts only known after the
// previous phase) -- cudaMalloc / cudaFree of a dozen buffers per call cost more than the kernels.
struct Carve {
    uint8_t *base = nullptr;
    size_t off = 0;
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    template <typename T> T *take(size_t n) {
        T *p = reinterpret_cast<T *>(base + off);
        off += pad(n * sizeof(T));
        return p;
    }
};

}  // namespace ldx

using namespace ldx;

// VCF text (whole lines) -> rows of a store.  *store_io == nullptr: a new store of exactly the text's records is created;
// else the records land at rows row_base ... of the given (annotated, large enough) store.  rows_out[rows_cap] receives the
// records (line_off relative to `text`).
static int ingest_text(ldx_ctx *ctx, const uint8_t *text, int64_t text_bytes, int32_t n_samples, ldx_store **store_io, int64_t row_base,
                       ldx_vcf_row *rows_out, int64_t rows_cap, int64_t *n_rows_out, bool trim_scratch) {
    *n_rows_out = 0;
    LDX_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // ---- phase 1: the text, once, with a final newline and slack for K1's aligned 16-byte loads; newline counts per block
    const bool add_nl = text[text_bytes - 1] != '\n';
    const int64_t n = text_bytes + (add_nl ? 1 : 0);
    const int64_t n_blocks = (n + NL_BLOCK_BYTES - 1) / NL_BLOCK_BYTES;
    LDX_REQUIRE(n_blocks < (1ll << 31), "vcf ingest: text too large for one call");
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(n_blocks + 1), st);
    Carve c1;
    LDX_TRY(scratch_get(ctx, 0, Carve::pad((size_t)n + 4 * (size_t)n_samples + 64) + 2 * Carve::pad(((size_t)n_blocks + 1) * 4) + Carve::pad(cub_bytes) + 256, (void **)&c1.base));
    const size_t text_slack = (size_t)4 * n_samples + 64;      // K1's fast kernel reads a whole plain row from every genotype offset
    uint8_t *d_text = c1.take<uint8_t>((size_t)n + text_slack);
    uint32_t *d_counts = c1.take<uint32_t>((size_t)n_blocks + 1), *d_base = c1.take<uint32_t>((size_t)n_blocks + 1);
    void *d_cub = c1.take<uint8_t>(cub_bytes);
    LDX_CUDA(cudaMemcpyAsync(d_text, text, (size_t)text_bytes, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemsetAsync(d_text + text_bytes, '\n', (size_t)(n - text_bytes) + text_slack, st));
    LDX_CUDA(cudaMemsetAsync(d_counts + n_blocks, 0, sizeof(uint32_t), st));
    count_newlines_kernel<<<(unsigned)n_blocks, NL_THREADS, 0, st>>>(d_text, n, d_counts);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    LDX_CUDA(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_counts, d_base, (int)(n_blocks + 1), st));
    uint32_t n_lines32 = 0;
    LDX_CUDA(cudaMemcpyAsync(&n_lines32, d_base + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    LDX_CUDA(cudaStreamSynchronize(st));
    const int64_t n_lines = n_lines32;
    LDX_REQUIRE(n_lines < (1ll << 31), "vcf ingest: too many lines for one call");
    if (n_lines == 0) return LDX_OK;
    // ---- phase 2: newline positions, per-line parse, record numbering
    size_t cub_bytes2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes2, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(n_lines + 1), st);
    Carve c2;
    LDX_TRY(scratch_get(ctx, 1, Carve::pad((size_t)n_lines * 8) + Carve::pad((size_t)n_lines * sizeof(ldx_vcf_row)) +
                                    2 * Carve::pad(((size_t)n_lines + 1) * 4) + Carve::pad(cub_bytes2) + 256, (void **)&c2.base));
    int64_t *d_nl = c2.take<int64_t>((size_t)n_lines);
    ldx_vcf_row *d_tmp = c2.take<ldx_vcf_row>((size_t)n_lines);
    uint32_t *d_isrec = c2.take<uint32_t>((size_t)n_lines + 1), *d_recidx = c2.take<uint32_t>((size_t)n_lines + 1);
    void *d_cub2 = c2.take<uint8_t>(cub_bytes2);
    write_newlines_kernel<<<(unsigned)n_blocks, NL_THREADS, 0, st>>>(d_text, n, d_base, d_nl);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    LDX_CUDA(cudaMemsetAsync(d_isrec + n_lines, 0, sizeof(uint32_t), st));
    const unsigned lgrid = (unsigned)((n_lines + 255) / 256);
    parse_lines_kernel<<<lgrid, 256, 0, st>>>(d_text, d_nl, n_lines, n_samples, d_tmp, d_isrec);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    LDX_CUDA(cub::DeviceScan::ExclusiveSum(d_cub2, cub_bytes2, d_isrec, d_recidx, (int)(n_lines + 1), st));
    uint32_t n_rec32 = 0;
    LDX_CUDA(cudaMemcpyAsync(&n_rec32, d_recidx + n_lines, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    LDX_CUDA(cudaStreamSynchronize(st));
    const int64_t n_rec = n_rec32;
    *n_rows_out = n_rec;
    if (n_rec > rows_cap) return set_error(LDX_ERR_CAPACITY, "vcf ingest: rows buffer too small");
    // ---- phase 3: dense rows
    Carve c3;
    LDX_TRY(scratch_get(ctx, 2, Carve::pad((size_t)n_rec * sizeof(ldx_vcf_row)) + Carve::pad((size_t)n_rec * 8) + Carve::pad((size_t)n_rec) + 256,
                        (void **)&c3.base));
    ldx_vcf_row *d_rows = c3.take<ldx_vcf_row>((size_t)n_rec);
    int64_t *d_gt = c3.take<int64_t>((size_t)n_rec);
    uint8_t *d_status = c3.take<uint8_t>((size_t)n_rec);
    // ---- the store, its annotations, the genotype planes
    ldx_store *s = *store_io;
    const bool own_store = s == nullptr;
    int rc = LDX_OK;
    if (own_store) {
        LDX_TRY(ldx_store_create(ctx, n_rec, 2 * n_samples, &s));
        rc = store_alloc_annotations(s);
    } else if (row_base + n_rec > s->n_variants) return set_error(LDX_ERR_CAPACITY, "vcf ingest: more records than the store has rows");
    if (rc == LDX_OK && n_rec > 0) {
        compact_rows_kernel<<<lgrid, 256, 0, st>>>(d_tmp, d_isrec, d_recidx, n_lines, d_rows, d_gt, s->d_pos0 + row_base, s->d_end0 + row_base,
                                                   s->d_idnum + row_base, s->d_eligible + row_base, row_base);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = set_error(LDX_ERR_CUDA, "vcf ingest: compact launch failed");
    }
    if (rc == LDX_OK && n_rec > 0)
        rc = launch_pack_gt(ctx, d_text, d_gt, 0, n_rec, n_samples, s->d_planes + row_base * s->stride_words, s->stride_words, d_status);   // rows with offset -1: zeros
    if (rc == LDX_OK && n_rec > 0) {
        // rows with a field outside the plain "a|b" alphabet: parsed again in full, aux planes, the general route (ldx_general.cu)
        std::vector<uint8_t> h_status((size_t)n_rec);
        if (cudaMemcpyAsync(h_status.data(), d_status, (size_t)n_rec, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "vcf ingest: genotype status");
        if (rc == LDX_OK) rc = store_pack_general(s, row_base, n_rec, d_text, n, d_gt, 0, n_samples, h_status.data());
        if (rc == LDX_OK && cudaMemcpyAsync(d_status, h_status.data(), (size_t)n_rec, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "vcf ingest: status");
        if (rc == LDX_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "vcf ingest: status");      // h_status is a local
    }
    if (rc == LDX_OK && n_rec > 0) {
        merge_status_kernel<<<(unsigned)((n_rec + 255) / 256), 256, 0, st>>>(d_rows, d_status, n_rec, s->d_eligible + row_base);
        ctx->launches++;
        if (cudaMemcpyAsync(rows_out, d_rows, sizeof(ldx_vcf_row) * (size_t)n_rec, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "vcf ingest: rows download");
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == LDX_OK) rc = cuda_fail(cudaGetLastError(), "vcf ingest");
    if (trim_scratch) scratch_trim(ctx, (size_t)256 << 20);      // the text of a whole chromosome does not stay behind in the arena
    if (rc != LDX_OK) { if (own_store) ldx_store_destroy(s); return rc; }
    s->annotated = true;
    s->mask_set = false;
    *store_io = s;
    return LDX_OK;
}

extern "C" int32_t ldx_store_ingest_vcf(ldx_ctx *ctx, const uint8_t *text, int64_t text_bytes, int32_t n_samples,
                                        ldx_store **store_out, ldx_vcf_row *rows_out, int64_t rows_cap, int64_t *n_rows_out) {
    LDX_REQUIRE(ctx && text && store_out && n_rows_out, "NULL argument");
    LDX_REQUIRE(text_bytes > 0 && n_samples > 0 && n_samples <= (1 << 23), "bad text size or sample count");
    LDX_REQUIRE(rows_cap >= 0 && (rows_out || rows_cap == 0), "bad rows buffer");
    *store_out = nullptr;
    ldx_store *s = nullptr;
    LDX_TRY(ingest_text(ctx, text, text_bytes, n_samples, &s, 0, rows_out, rows_cap, n_rows_out, true));
    if (!s) {                                  // a text without a single line: an empty store
        LDX_TRY(ldx_store_create(ctx, 0, 2 * n_samples, &s));
        const int rc = store_alloc_annotations(s);
        if (rc != LDX_OK) { ldx_store_destroy(s); return rc; }
        s->annotated = true;
    }
    *store_out = s;
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ a whole <chrom>.vcf.gz, slab by slab
// a real 1000 Genomes chromosome is ~65 GB of text (6.4 M records x 10 KB) for a 4 GB store; inflating it into
// one host buffer and uploading it as one device buffer needs 65 GB of each.  Here the BGZF members are inflated in groups of
// `slab_bytes` into ONE pinned host buffer, each slab is cut at its last newline (the partial line is carried over), indexed,
// parsed and packed into the store at the next row, and its records' fixed columns are kept: host and device hold one slab of
// text at a time.  The store is allocated for an upper bound of the record count (a record line has at least 2 * n_samples + 18
// bytes) and trimmed to the rows found.
extern "C" int32_t ldx_store_ingest_vcf_file(ldx_ctx *ctx, const char *path, int32_t n_samples, int64_t slab_bytes, int32_t threads,
                                             ldx_store **store_out, ldx_vcf_row **rows_out, int64_t *n_rows_out, uint8_t **blob_out,
                                             int64_t **blob_off_out, int64_t *text_bytes_out) {
    LDX_REQUIRE(ctx && path && store_out && rows_out && n_rows_out && blob_out && blob_off_out, "NULL argument");
    LDX_REQUIRE(n_samples > 0 && n_samples <= (1 << 23), "bad sample count");
    *store_out = nullptr; *rows_out = nullptr; *n_rows_out = 0; *blob_out = nullptr; *blob_off_out = nullptr;
    if (text_bytes_out) *text_bytes_out = 0;
    if (slab_bytes <= 0) slab_bytes = (int64_t)256 << 20;
    slab_bytes = std::max<int64_t>(slab_bytes, 4096);       // a line (or a BGZF member) that does not fit a slab is an error, not a hang
    LDX_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint8_t> in;
    LDX_TRY(read_whole_file(path, in));
    std::vector<BgzfMember> members;
    size_t total = 0;
    const bool bgzf = !in.empty() && bgzf_scan(in.data(), in.size(), members, &total);
    std::vector<ldx_vcf_row> rows;
    std::vector<uint8_t> blob;
    std::vector<int64_t> off(1, 0);
    ldx_store *s = nullptr;
    int rc = LDX_OK;
    auto keep_columns = [&](const uint8_t *text, int64_t n_text, int64_t first, int64_t n_new, int64_t text_base) {
        // the nine fixed columns of the new records (what the writers print), and their line offsets made file-wide
        for (int64_t r = first; r < first + n_new; ++r) {
            ldx_vcf_row &x = rows[(size_t)r];
            if (x.line_off < 0 || x.line_off > n_text) return set_error(LDX_ERR_STATE, "vcf ingest: record outside its slab");
            const int64_t len = std::max<int64_t>(0, std::min<int64_t>(x.gt_off, n_text - x.line_off));   // a refused record (status 2): what there is
            blob.insert(blob.end(), text + x.line_off, text + x.line_off + len);
            off.push_back((int64_t)blob.size());
            x.line_off += text_base;
        }
        return (int)LDX_OK;
    };
    if (!bgzf) {
        // plain gzip (one stream): no block table to cut by -- the whole text at once, as ldx_inflate_gz_file + ldx_store_ingest_vcf
        uint8_t *text = nullptr; int64_t n_text = 0;
        LDX_TRY(ldx_inflate_gz_file(path, threads, &text, &n_text, nullptr));
        int64_t n_rec = 0;
        if (n_text > 0) {
            rows.resize((size_t)(n_text / std::max<int64_t>(2ll * n_samples + 18, 1) + 16));
            rc = ingest_text(ctx, text, n_text, n_samples, &s, 0, rows.data(), (int64_t)rows.size(), &n_rec, true);
            rows.resize((size_t)n_rec);
            if (rc == LDX_OK) rc = keep_columns(text, n_text, 0, n_rec, 0);
        }
        std::free(text);
        total = (size_t)n_text;
    } else {
        const int64_t min_line = 2ll * n_samples + 18;
        const int64_t n_max = (int64_t)total / min_line + 16;
        rc = ldx_store_create(ctx, n_max, 2 * n_samples, &s);
        if (rc == LDX_OK) rc = store_alloc_annotations(s);
        slab_bytes = std::min<int64_t>(slab_bytes, std::max<int64_t>((int64_t)total, 4096));       // a small file: a small pinned buffer
        uint8_t *h_slab = nullptr;
        size_t slab_cap = (size_t)slab_bytes + (1u << 16) + 64;                       // + one member + the final newline
        if (rc == LDX_OK && cudaMallocHost((void **)&h_slab, slab_cap) != cudaSuccess) { cudaGetLastError(); rc = set_error(LDX_ERR_NOMEM, "vcf ingest: pinned slab buffer"); }
        int64_t carry = 0, row_base = 0, text_base = 0;               // text_base: file-wide offset of h_slab[0]
        size_t m = 0;
        while (rc == LDX_OK && (m < members.size() || carry > 0)) {
            // the next group of members behind the carried-over partial line
            size_t e = m;
            int64_t fill = carry;
            while (e < members.size() && fill + (int64_t)members[e].out_len <= (int64_t)slab_bytes) fill += (int64_t)members[e++].out_len;
            if (e == m && m < members.size()) {
                // not even one more member fits behind the carried-over text (a slab smaller than a member, or a line longer than
                // the slab): the buffer grows to what this step needs, up to 1 GiB of text without a newline
                const int64_t need = carry + (int64_t)members[m].out_len;
                if (need > ((int64_t)1 << 30)) { rc = set_error(LDX_ERR_DATA, "vcf ingest: a line longer than 1 GiB"); break; }
                uint8_t *bigger = nullptr;
                if (cudaMallocHost((void **)&bigger, (size_t)need + (1u << 16) + 64) != cudaSuccess) { cudaGetLastError(); rc = set_error(LDX_ERR_NOMEM, "vcf ingest: pinned slab buffer"); break; }
                if (carry > 0) std::memcpy(bigger, h_slab, (size_t)carry);
                cudaFreeHost(h_slab);
                h_slab = bigger; slab_cap = (size_t)need + (1u << 16) + 64; slab_bytes = need;
                continue;
            }
            if (!bgzf_inflate_range(in.data(), members, m, e, h_slab + carry, threads)) { rc = set_error(LDX_ERR_ARG, "inflate: corrupt BGZF block (deflate error or CRC mismatch)"); break; }
            const bool last = e == members.size();
            int64_t cut = fill;
            if (!last) {
                while (cut > 0 && h_slab[cut - 1] != '\n') --cut;     // whole lines only; the rest is carried over (cut == 0: all of it)
            }
            if (cut > 0) {
                const size_t before = rows.size();
                rows.resize(before + (size_t)(cut / min_line + 16));
                int64_t n_rec = 0;
                rc = ingest_text(ctx, h_slab, cut, n_samples, &s, row_base, rows.data() + before, (int64_t)(rows.size() - before), &n_rec, false);
                rows.resize(before + (size_t)(rc == LDX_OK ? n_rec : 0));
                if (rc == LDX_OK) rc = keep_columns(h_slab, cut, (int64_t)before, n_rec, text_base);
                row_base += n_rec;
            }
            carry = fill - cut;
            if (carry > 0) std::memmove(h_slab, h_slab + cut, (size_t)carry);
            text_base += cut;
            m = e;
            if (last) { if (carry > 0) rc = set_error(LDX_ERR_STATE, "vcf ingest: text left over"); break; }
        }
        if (h_slab) cudaFreeHost(h_slab);
        scratch_trim(ctx, (size_t)256 << 20);
        if (rc == LDX_OK) rc = store_shrink(s, row_base);
    }
    if (rc == LDX_OK && !s) {
        rc = ldx_store_create(ctx, 0, 2 * n_samples, &s);
        if (rc == LDX_OK) rc = store_alloc_annotations(s);
        if (rc == LDX_OK) s->annotated = true;
    }
    ldx_vcf_row *r_out = nullptr; uint8_t *b_out = nullptr; int64_t *o_out = nullptr;
    if (rc == LDX_OK) {
        r_out = static_cast<ldx_vcf_row *>(std::malloc(std::max<size_t>(rows.size() * sizeof(ldx_vcf_row), 1)));
        b_out = static_cast<uint8_t *>(std::malloc(std::max<size_t>(blob.size(), 1)));
        o_out = static_cast<int64_t *>(std::malloc(off.size() * sizeof(int64_t)));
        if (!r_out || !b_out || !o_out) rc = set_error(LDX_ERR_NOMEM, "vcf ingest: host tables");
    }
    if (rc != LDX_OK) { std::free(r_out); std::free(b_out); std::free(o_out); if (s) ldx_store_destroy(s); return rc; }
    std::memcpy(r_out, rows.data(), rows.size() * sizeof(ldx_vcf_row));
    std::memcpy(b_out, blob.data(), blob.size());
    std::memcpy(o_out, off.data(), off.size() * sizeof(int64_t));
    *store_out = s; *rows_out = r_out; *n_rows_out = (int64_t)rows.size(); *blob_out = b_out; *blob_off_out = o_out;
    if (text_bytes_out) *text_bytes_out = (int64_t)total;
    return LDX_OK;
}


/* Host helper: the nine fixed columns of every record, back to back (record r = out[off[r], off[r+1])), so that the
 * caller can drop the multi-gigabyte text and still print ID / REF / ALT / INFO of the rows it reports. */
extern "C" int32_t ldx_vcf_copy_prefixes(const uint8_t *text, int64_t text_bytes, const ldx_vcf_row *rows, int64_t n_rows,
                                         uint8_t *out, int64_t out_cap, int64_t *off_out) {
    LDX_REQUIRE(text && (rows || n_rows == 0) && off_out && n_rows >= 0, "bad argument");
    int64_t total = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        off_out[r] = total;
        LDX_REQUIRE(rows[r].line_off >= 0 && rows[r].gt_off >= 0 && rows[r].line_off + rows[r].gt_off <= text_bytes, "row outside the text");
        total += rows[r].gt_off;
    }
    off_out[n_rows] = total;
    if (!out) return LDX_OK;                                     // size query
    if (total > out_cap) return set_error(LDX_ERR_CAPACITY, "prefix buffer too small");
    for (int64_t r = 0; r < n_rows; ++r) std::memcpy(out + off_out[r], text + rows[r].line_off, (size_t)rows[r].gt_off);
    return LDX_OK;
}
