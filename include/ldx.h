/* ldx.h -- C ABI of libldx.so, the B200 (sm_100a) engine behind ld-tools' LD hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no C++/torch types.  It is what
 * the reference's Python would bind with ctypes.  Every entry point names the
 * reference code it replaces; paths are relative to the ld-tools repository root.
 *
 * Conventions
 *   - Every function returns an int32 status: LDX_OK (0) or a negative LDX_ERR_* code;
 *     ldx_last_error() returns a thread-local message for the last failure.
 *   - No C++ exception crosses the boundary.  The caller allocates and owns every host buffer.
 *   - CUDA is initialised by ldx_init() only -- call it AFTER fork (the reference fans out with
 *     multiprocessing.Pool, ld_area.py:336 / ld_triangle.py:406); one ctx per process and device.
 *   - Calls on one ctx are not concurrent.  Distinct contexts/processes are independent.
 *   - There is no CPU fallback: without a CUDA device ldx_init() fails.
 *
 * Data model
 *   A *store* holds one chromosome's biallelic phased variants as bit planes in HBM: haplotype
 *   h (= 2*sample + allele slot, the order of the flat list the drivers build with
 *   `+= rec.samples[name]['GT']`, ld_area.py:182-187) is bit (h % 64) of 64-bit word (h / 64)
 *   of the variant's row.  Row pitch = ceil(n_hap/64) rounded up to 16 words (128 B); pad bits
 *   are zero.  5008 haplotypes -> 79 words -> 640-byte rows.  A *mask* plane selects the
 *   haplotypes of the chosen samples (get_sample_names.py:17-31); N = popcount(mask).
 *   Genotype rows that are not complete phased diploid 0/1 rows -- a missing call ('.', in N but in neither allele count),
 *   a haploid sample (one list element), an allele code other than 0 / 1, unphased '/' -- are kept with two more planes
 *   (slot present, allele == 0) and every pair with such a variant is computed from explicit counts with the pairing's own N,
 *   exactly as calc_ld.py:30-40 treats the lists the drivers build with `+= rec.samples[name]['GT']`, including zip()'s
 *   position-by-position pairing of lists of unequal ploidy.  A store's common ploidy pattern (e.g. "males haploid" on chrX)
 *   stays on the fast paths: it is folded into the mask.  Rounded values beyond 1.6383 (possible only with many non-0/1
 *   entries) saturate the 14-bit fields.
 *
 * Packed pair result (uint32) -- the reference's rounded outputs (calc_ld.py:94-95) as integers:
 *   bits  0..13  round(r2, 4) * 10^4          (0..10000)
 *   bit   14     internal (always 0 in results returned to the caller)
 *   bit   15     r2 is the reference's INT 0   (calc_ld.py:89-90; prints "0", not "0.0")
 *   bits 16..29  round(D', 4) * 10^4
 *   bit   30     pair is below the caller's threshold (triangle only; ld_triangle.py:223-225)
 *   bit   31     D' is the reference's INT 0   (ZeroDivisionError branch, calc_ld.py:68-69,75-76)
 *   value / 10000.0 in the caller's language reproduces Python's round(x, 4) bit for bit.
 */
#ifndef LDX_H
#define LDX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDX_ABI_VERSION 1

enum {
    LDX_OK = 0,
    LDX_ERR_ARG = -1,        /* bad argument */
    LDX_ERR_CUDA = -2,       /* CUDA runtime/driver failure (incl. no device) */
    LDX_ERR_NOMEM = -3,      /* host or device allocation failed */
    LDX_ERR_EMPTY = -4,      /* N == 0: the reference raises ZeroDivisionError (calc_ld.py:33) */
    LDX_ERR_CAPACITY = -5,   /* output buffer too small; *n_out holds the size needed */
    LDX_ERR_STATE = -6,      /* call order (e.g. compute before ldx_store_set_mask) */
    LDX_ERR_DATA = -7        /* input outside the supported domain */
};

enum { LDX_MEASURE_R2 = 0, LDX_MEASURE_DPRIME = 1 };      /* -l r_square | d_prime */
enum { LDX_ENGINE_AUTO = 0, LDX_ENGINE_POPC = 1, LDX_ENGINE_MMA = 2 };

#define LDX_R2_MASK      0x00003fffu
#define LDX_R2_INT0      0x00008000u
#define LDX_DP_SHIFT     16
#define LDX_DP_MASK      0x3fff0000u
#define LDX_BELOW_THRES  0x40000000u
#define LDX_DP_INT0      0x80000000u

typedef struct ldx_ctx ldx_ctx;
typedef struct ldx_store ldx_store;

/* One kept (query, opposing variant) pair of a window scan: the row ld_area.py:264-276 emits. */
typedef struct {
    int32_t query;     /* index into the call's query arrays */
    int32_t row;       /* store row of the opposing variant */
    int32_t n11;       /* (alt, alt) haplotype count, calc_ld.py:32 */
    uint32_t packed;   /* rounded r2 / D' + type flags, format above */
} ldx_hit;

/* Result of the list-level calculator (ldx_calc_ld_lists): every intermediate of calc_ld.py. */
typedef struct {
    int64_t n_hap, n_11, n_a1, n_a0, n_b1, n_b0;   /* calc_ld.py:31-32, :37-40 */
    double d, dprime, r2, p_a, p_b;                /* before rounding, calc_ld.py:33-90 */
    double r2_e4, dprime_e4, p_a_e4, p_b_e4;       /* round(x,4)*10^4 as exact integers in fp64 */
    int32_t dprime_is_int0, r2_is_int0;            /* the reference's int-0 sentinels */
} ldx_ld_result;

/* ---------------------------------------------------------------- lifecycle */
int32_t ldx_abi_version(void);
const char *ldx_last_error(void);
int32_t ldx_device_count(int32_t *n_out);
/* device < 0 selects the current device.  Creates the ctx's stream and scratch buffers. */
int32_t ldx_init(int32_t device, ldx_ctx **ctx_out);
int32_t ldx_destroy(ldx_ctx *ctx);
/* Launch on the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream) instead of
 * the ctx's own.  The handle is used as given: NULL is CUDA's legacy default stream.
 * ldx_use_own_stream() goes back to the ctx's private non-blocking stream. */
int32_t ldx_set_stream(ldx_ctx *ctx, void *cuda_stream);
int32_t ldx_use_own_stream(ldx_ctx *ctx);
int32_t ldx_synchronize(ldx_ctx *ctx);
/* Tuning knobs (benchmarks/tests; defaults are chosen per call):
 *   LDX_TUNE_MMA_TILE_N  column width of the tcgen05 all-pairs tile: 0 = heuristic, 64 or 128
 *   LDX_TUNE_MMA_MIN_V   LDX_ENGINE_AUTO uses the tcgen05 engine from this many variants (256)
 *   LDX_TUNE_MMA_PAIR    1 = 256 x 128 tiles on CTA pairs (tcgen05.mma.cta_group::2), 0 = one CTA per 128 x 128 tile,
 *                        -1 (default) = pairs for calls of more than one wave of tiles
 *   LDX_TUNE_DEFER_CAP   capacity of the tcgen05 engine's deferred-pair lists (0 = sized from the pair count); a
 *                        small value forces the overflow paths (pairs settled in place) -- for tests
 *   LDX_TUNE_WINDOW_MQ   1 (default) = window scans of several queries with monotone candidate ranges use the multi-query
 *                        kernel (a store row is loaded once per four queries); 0 = always one query per pass
 *   LDX_TUNE_MMA_DIRECT  -1 / 1 (default) = a one-wave all-pairs call whose rows[] are contiguous store rows reads the planes
 *                        through a TMA tensor map (no gather kernel, no operand scratch); 0 = always gather */
enum { LDX_TUNE_MMA_TILE_N = 1, LDX_TUNE_MMA_MIN_V = 2, LDX_TUNE_MMA_PAIR = 3, LDX_TUNE_DEFER_CAP = 4, LDX_TUNE_WINDOW_MQ = 5,
       LDX_TUNE_MMA_DIRECT = 6 };
int32_t ldx_set_tuning(ldx_ctx *ctx, int32_t key, int32_t value);
/* Diagnostics: with enable != 0 the tcgen05 all-pairs kernel's first CTA records %globaltimer
 * stamps (ns): [0] prologue done, [1]/[2] accumulator ready / epilogue done of its 1st tile,
 * [3]/[4] 2nd tile, [5]/[6] 3rd tile; [8+g] widener done / [64+g] MMA start / [128+g] producer issue /
 * [192+g] bits landed / [256+g] operand stage free (widener) / [320+g] widening stores issued, for its
 * first 48 pipeline stages g; [7] kernel entry, [56] all roles done; every CTA c < 192 also records its life
 * cycle at [512 + 8c + k], k = 0 entry, 1 prologue done, 2 first accumulator ready, 3 first epilogue done,
 * 4 all roles done, 5 SM id, 6/7 settlement barrier reached / passed (single-wave kernel), at [2048 + 2c] deferred
 * pairs settled, at [2432 + c] their number.  stamps8 (may be NULL, else 4096 entries) receives the last recording. */
int32_t ldx_debug_trace(ldx_ctx *ctx, int32_t enable, uint64_t *stamps8);
/* Dominant-kernel timing for roofline reports: while enabled, every all-pairs kernel
 * (triangle_mma_kernel / triangle_popc_kernel) and window kernel launch is bracketed by CUDA events
 * on the ctx stream.  Each call returns (and clears) the device time in ms and the number of launches
 * accumulated since the previous call, then sets the new state.  Reading synchronises with the last
 * recorded event. */
int32_t ldx_kernel_timing(ldx_ctx *ctx, int32_t enable, double *ms_out, int64_t *launches_out);
int32_t ldx_sm_count(ldx_ctx *ctx, int32_t *n_out);
/* Device memory for the outputs of the *_dev entry points, for callers that have no CUDA allocator of their own (the drivers
 * of ld_tools_b200/drivers.py: packed words of a batch of matrices that only ldx_triangle_text reads).  256-byte aligned. */
int32_t ldx_dev_alloc(ldx_ctx *ctx, int64_t bytes, void **dev_ptr_out);
int32_t ldx_dev_free(ldx_ctx *ctx, void *dev_ptr);
/* Kernels launched by this ctx since creation (bench.py's gpu_launches claim). */
int32_t ldx_launch_count(ldx_ctx *ctx, int64_t *n_out);

/* ---------------------------------------------------------------- the list-level calculator
 * Replaces backend/calc_ld.py:3-99 for one pair given as genotype codes: 0 = ref, 1 = alt,
 * any other byte = "other" (None, 2, ...: paired but in neither allele count, calc_ld.py:37-40).
 * Unequal lengths pair up to the shorter one (zip, calc_ld.py:30-31) while the allele counts run
 * over the full vectors.  len == 0 -> LDX_ERR_EMPTY. */
int32_t ldx_calc_ld_lists(ldx_ctx *ctx, const uint8_t *g_a, int64_t len_a, const uint8_t *g_b,
                          int64_t len_b, ldx_ld_result *out);

/* calc_ld.py:33-97 alone, for n pairs given as integer counts under n_hap haplotypes:
 * n11 = (alt, alt) haplotypes, n1a / n1b = alt alleles of var_1 / var_2 (ref = n_hap - alt).
 * Outputs are host arrays of n (each may be NULL): D, D', r2 before rounding and the packed word. */
int32_t ldx_finalise_counts(ldx_ctx *ctx, int32_t n_hap, const int32_t *n11, const int32_t *n1a,
                            const int32_t *n1b, int64_t n, double *d, double *dprime, double *r2,
                            uint32_t *packed);

/* ---------------------------------------------------------------- the bit-plane store
 * Replaces the per-pair pysam genotype extraction (ld_lite.py:109-137, ld_area.py:182-187 and
 * :230-235, ld_triangle.py:158-186) and the prep_intgen_data/create_src_dict cache as the
 * thing the calculator reads. */
int32_t ldx_store_create(ldx_ctx *ctx, int64_t n_variants, int32_t n_hap, ldx_store **store_out);
int32_t ldx_store_destroy(ldx_store *store);
int32_t ldx_store_shape(const ldx_store *store, int64_t *n_variants, int32_t *n_hap,
                        int32_t *stride_words);
/* Device address of the planes ([n_variants][stride_words] uint64), for zero-copy interop. */
int32_t ldx_store_planes_ptr(const ldx_store *store, void **dev_ptr_out);

/* GPU bit-packing of VCF genotype text (kernel K1).  `text` is host memory holding, for each of
 * n_rows variants, n_samples fields "a|b" each followed by one separator byte, starting at
 * byte row_off[i] (row_off == NULL: row i starts at i * row_pitch).  Rows land at store rows
 * first_row...  row_status[i] (may be NULL): bit 0 = a field outside the phased diploid biallelic alphabet {0|0, 0|1, 1|0,
 * 1|1}: the row was parsed again in full ("a", "a|b", "a/b" with a, b = '.' or a number; sub-fields after ':' ignored; the
 * row ends at its newline) and takes the general route; bit 3 = not even that parser takes it (more than two alleles in a
 * field, an empty field, fewer than n_samples fields): its planes are undefined and callers should refuse the file. */
int32_t ldx_store_pack_gt(ldx_store *store, int64_t first_row, int64_t n_rows, const uint8_t *text,
                          int64_t text_bytes, const int64_t *row_off, int64_t row_pitch,
                          int32_t n_samples, uint8_t *row_status);
/* ---------------------------------------------------------------- VCF text -> store, on the GPU
 * Replaces the per-record pysam access of the drivers (rec.pos / .id / .ref / .info / .samples[..]['GT'],
 * ld_area.py:215-235, ld_triangle.py:128-186) and the Python per-line loop an ingest would otherwise need.
 * `text` = a whole decompressed VCF (host memory; '##' / '#CHROM' lines are skipped, every other line is a
 * record: nine tab-separated fixed columns, then n_samples "a|b" genotype columns).  One upload; kernels
 * build the newline index, parse every line (POS, len(REF), the rs number, the MULTI_ALLELIC INFO key),
 * write the window annotations (pos0, end0, idnum, eligible = rs\d+$ and not MULTI_ALLELIC,
 * ld_area.py:222-225) into a NEW store and bit-pack the genotypes (K1).  rows_out[rows_cap] receives one
 * record per variant in file order (= store row); *n_rows_out = their number (LDX_ERR_CAPACITY if > rows_cap).
 * status: bit 0 = a genotype field outside {0|0,0|1,1|0,1|1} (the row was parsed in full and takes the general route, see
 * ldx_store_pack_gt), bit 1 = fewer than 9 + n_samples columns (row kept, all-reference, not eligible), bit 2 = POS is not a
 * number, bit 3 = a genotype field no parser takes (more than two alleles, empty). */
typedef struct ldx_vcf_row {
    int64_t line_off;                 /* byte offset of the line in the text */
    int64_t idnum;                    /* digits of an rs\d+ ID, else -1 - row */
    int32_t gt_off, id_off, ref_off, alt_off, info_off, fmt_off;   /* field starts, relative to line_off */
    int32_t pos;                      /* POS, 1-based */
    int32_t ref_len;                  /* the record covers [pos - 1, pos - 1 + ref_len) for the window fetch: len(REF), or INFO's END=<n> - (pos - 1) when it has one (htslib tbx.c) */
    uint8_t status, eligible, multi, pad[5];
} ldx_vcf_row;
int32_t ldx_store_ingest_vcf(ldx_ctx *ctx, const uint8_t *text, int64_t text_bytes, int32_t n_samples,
                             ldx_store **store_out, ldx_vcf_row *rows_out, int64_t rows_cap, int64_t *n_rows_out);
/* The whole ingest in one call, in bounded memory: <chrom>.vcf.gz (BGZF) is inflated slab by slab (`slab_bytes` of text at a
 * time, <= 0: 256 MiB; `threads` host threads, <= 0: all cores) into one pinned buffer, every slab is cut at its last newline,
 * indexed, parsed and packed on the GPU into the next rows of the store; the records and their fixed columns come back in
 * tables allocated by the library (*rows_out [*n_rows_out], *blob_out, *blob_off_out [*n_rows_out + 1]: as ldx_store_ingest_vcf /
 * ldx_vcf_copy_prefixes give them; line_off is file-wide; release each with ldx_free_host).  *text_bytes_out (may be NULL) = the
 * size of the decompressed text.  A plain gzip file (no block table) is read at once.  A slab smaller than a BGZF member or than
 * a line grows to what the step needs (a line of more than 1 GiB is an error). */
int32_t ldx_store_ingest_vcf_file(ldx_ctx *ctx, const char *path, int32_t n_samples, int64_t slab_bytes, int32_t threads,
                                  ldx_store **store_out, ldx_vcf_row **rows_out, int64_t *n_rows_out, uint8_t **blob_out,
                                  int64_t **blob_off_out, int64_t *text_bytes_out);
/* Host side of the ingest: a whole .gz file -> malloc'ed text (*text_out; release it with ldx_free_host).  A BGZF file
 * (what tabix indexes, prep_intgen_data.py:138: independent gzip members of <= 64 KiB announcing their sizes) is
 * inflated by `threads` host threads in parallel (<= 0: all cores), every block straight into its final place, CRCs
 * checked; any other gzip file (one stream or concatenated members) sequentially.  *was_bgzf_out may be NULL. */
int32_t ldx_inflate_gz_file(const char *path, int32_t threads, uint8_t **text_out, int64_t *text_bytes_out,
                            int32_t *was_bgzf_out);
int32_t ldx_free_host(void *p);
/* Host helper: the fixed columns of every record back to back (record r = out[off_out[r], off_out[r+1]); off_out has
 * n_rows + 1 entries), so that the caller can drop the text and still print ID / REF / ALT / INFO of the rows it
 * reports.  out == NULL: only off_out is filled (size query). */
int32_t ldx_vcf_copy_prefixes(const uint8_t *text, int64_t text_bytes, const ldx_vcf_row *rows, int64_t n_rows,
                              uint8_t *out, int64_t out_cap, int64_t *off_out);

/* Load / read back ready-made planes (host, [n_rows][stride_words] uint64). */
int32_t ldx_store_upload(ldx_store *store, int64_t first_row, int64_t n_rows, const uint64_t *planes);
int32_t ldx_store_download(const ldx_store *store, int64_t first_row, int64_t n_rows, uint64_t *planes);
/* ldx_store_upload without the wait: the copy is enqueued on the context's stream.  A pinned `planes` buffer must stay unchanged
 * until the next blocking call on this context (ldx_triangle*, ldx_window, ldx_synchronize ...). */
int32_t ldx_store_upload_async(ldx_store *store, int64_t first_row, int64_t n_rows, const uint64_t *planes);
/* The on-disk form of a store: what replaces the reference's per-run tabix/pysam access to <chrom>.vcf.gz
 * (prep_intgen_data.py:138-177 builds its cache once; so does this).  One file = 64-byte header
 * {"LDXSTOR1", n_variants, n_hap, stride_words, annotated} + the planes + (if annotated) pos0 | end0 | idnum |
 * eligible, all little-endian, streamed through pinned staging in 64 MiB pieces.  The sample mask is per run
 * and not saved.  ldx_store_load creates the store; a truncated or foreign file is LDX_ERR_ARG. */
int32_t ldx_store_save(const ldx_store *store, const char *path);
int32_t ldx_store_load(ldx_ctx *ctx, const char *path, ldx_store **store_out);

/* Select samples: mask[stride_words] (host).  Runs the per-variant count kernel (K2):
 * n1[v] = popcount(mask & plane[v]) (calc_ld.py:37,39), N = popcount(mask) (calc_ld.py:31), and
 * the per-variant frequencies.  N == 0 -> LDX_ERR_EMPTY. */
int32_t ldx_store_set_mask(ldx_store *store, const uint64_t *mask);
/* n1_out[n_variants], p_e4_out[n_variants] = round(n1/N, 4)*10^4 (ld_area.py:188-189); any may be NULL. */
int32_t ldx_store_counts(const ldx_store *store, int32_t *n1_out, int32_t *p_e4_out, int32_t *n_hap_sel_out);
/* Per-row facts of the general route under the current selection (any pointer may be NULL): n1_out[v] = the variant's alt
 * count, len_out[v] = the length of its own genotype list (calc_ld.py:31 for the variant alone: N for a complete row),
 * kind_out[v] = -1 for a row on the fast paths, >= 0 or -2 for a row of the general route; *n_general_out = how many of those
 * the store has.  The alt frequency calc_ld reports for var_2 of a pair (a, b) is round(n1[b] / min(len[a], len[b]), 4). */
int32_t ldx_store_row_counts(const ldx_store *store, int32_t *n1_out, int32_t *len_out, int32_t *kind_out, int64_t *n_general_out);
/* New store holding only the selected haplotype columns (sel[k] = source haplotype of new
 * haplotype k), mask = all ones: 5x less HBM traffic for a 503-sample subset. */
int32_t ldx_store_subset(const ldx_store *store, const int32_t *sel, int32_t n_sel, ldx_store **store_out);
/* Per-row annotations the fused ld_area filters need (host arrays of n_variants):
 * pos0 = POS-1, end0 = POS-1+len(REF) (pysam fetch overlap, ld_area.py:215-217),
 * idnum = digits of the rsID (ld_area.py:222 same-id test), eligible = id matches rs\d+$ and
 * the record is not MULTI_ALLELIC (ld_area.py:223-224). */
int32_t ldx_store_set_annotations(ldx_store *store, const int32_t *pos0, const int32_t *end0,
                                  const int64_t *idnum, const uint8_t *eligible);

/* ---------------------------------------------------------------- pair list (kernel K3)
 * Replaces calc_ld(g1, g2) at ld_lite.py:143 for n pairs of store rows (var_1 = ia[k],
 * var_2 = ib[k]).  Outputs are host arrays of n, each may be NULL. */
int32_t ldx_pairs(ldx_store *store, const int64_t *ia, const int64_t *ib, int64_t n, int32_t *n11,
                  double *d, double *dprime, double *r2, uint32_t *packed);

/* ---------------------------------------------------------------- window scan (kernel K4)
 * Replaces the ld_area.py:215-249 loop for nq queries at once.  Query k is store row q_row[k];
 * candidates are store rows [lo[k], hi[k]) (a superset chosen by the caller's position index);
 * the kernel keeps row j iff it overlaps the 0-based half-open window [win_start[k], win_end[k])
 * (pos0[j] < win_end && end0[j] > win_start), is eligible, has a different idnum than the query,
 * and its ROUNDED measure >= thres_e4 / 10^4 (ld_area.py:248).  var_1 = query, var_2 = row j.
 * hits[cap] is filled sorted by (query, row); *n_hits = number kept (LDX_ERR_CAPACITY if > cap,
 * with *n_hits = size needed).  *n_scanned (may be NULL) = candidate pairs evaluated. */
int32_t ldx_window(ldx_store *store, const int64_t *q_row, const int64_t *lo, const int64_t *hi,
                   const int32_t *win_start, const int32_t *win_end, int64_t nq, int32_t measure,
                   int32_t thres_e4, ldx_hit *hits, int64_t cap, int64_t *n_hits, int64_t *n_scanned);

/* ---------------------------------------------------------------- all pairs (kernels K5 / K5')
 * Replaces the ld_triangle.py:133-230 double loop.  rows[v] = store rows in matrix order (the
 * caller sorts by position, ld_triangle.py:88).  For row > col: var_1 = rows[row], var_2 =
 * rows[col] (ld_triangle.py:193).  Output is the lower triangle packed by rows:
 * index = row*(row-1)/2 + col, v*(v-1)/2 entries.  has_thres != 0 sets LDX_BELOW_THRES on pairs
 * whose rounded measure < thres_e4/10^4.  packed / n11 are host arrays and may be NULL. */
int32_t ldx_triangle(ldx_store *store, const int64_t *rows, int64_t v, int32_t measure,
                     int32_t has_thres, int32_t thres_e4, int32_t engine, uint32_t *packed,
                     int32_t *n11);

/* A slice of the same triangle: only matrix rows row_begin .. row_end-1 (row_begin a multiple of 128),
 * i.e. the pairs (row, col) with row_begin <= row < row_end, col < row.  The output holds
 * tri(row_end) - tri(row_begin) entries, tri(r) = r*(r-1)/2, and entry 0 is pair (row_begin, 0): it is
 * the contiguous range [tri(row_begin), tri(row_end)) of the full packed triangle.  This is the unit the
 * multi-GPU path shards by (ld_tools_b200/shard.py: row ranges balanced by tile count). */
int32_t ldx_triangle_rows(ldx_store *store, const int64_t *rows, int64_t v, int64_t row_begin,
                          int64_t row_end, int32_t measure, int32_t has_thres, int32_t thres_e4,
                          int32_t engine, uint32_t *packed, int32_t *n11);

/* Narrow outputs of the same call, for callers that want the numbers on the host: every byte crosses PCIe, and the drivers only
 * ever print ONE measure of a matrix (ld_triangle.py:230).
 * ldx_triangle_values: 2 bytes per pair for the measure asked for -- bits 0..13 = round(value, 4) * 10^4, LDX_V16_BELOW = below the
 * threshold (has_thres != 0), LDX_V16_INT0 = the reference prints the int 0 -- in the layout of ldx_triangle_rows.
 * ldx_triangle_hits: with a threshold (-z), only the pairs that pass it, in matrix order (row > col, indices into rows[]); hits[cap],
 * *n_hits = pairs found (LDX_ERR_CAPACITY if > cap, with *n_hits = the size needed). */
#define LDX_V16_VALUE 0x3fffu
#define LDX_V16_BELOW 0x4000u
#define LDX_V16_INT0  0x8000u
typedef struct ldx_pair_hit { int32_t row, col; uint32_t packed; } ldx_pair_hit;
int32_t ldx_triangle_values(ldx_store *store, const int64_t *rows, int64_t v, int64_t row_begin, int64_t row_end, int32_t measure,
                            int32_t has_thres, int32_t thres_e4, int32_t engine, uint16_t *values);
int32_t ldx_triangle_hits(ldx_store *store, const int64_t *rows, int64_t v, int32_t measure, int32_t thres_e4, int32_t engine,
                          ldx_pair_hit *hits, int64_t cap, int64_t *n_hits);

/* ---------------------------------------------------------------- device-resident variants
 * Same kernels, but outputs stay in HBM (caller-allocated device memory) and the call only
 * enqueues work on the ctx stream.  ldx_resolve() then finishes the rounding of the (very rare)
 * pairs whose r2 * 10^4 sits within 1e-9 of a rounding tie -- there the reference's libm pow decides
 * the last digit -- and must be called before the outputs are consumed (it synchronises).
 * Index arrays (rows, q_row, ...) are HOST arrays; pinned ones must stay valid until the next
 * ldx_synchronize()/ldx_resolve().
 * ldx_window_dev: dev_hits must be 16-byte aligned; dev_n_hits points at TWO int64 in device
 * memory: [0] = hits found (may exceed cap; only the first cap are stored), [1] = pairs scanned.
 * Hits are unordered; a hit that ldx_resolve() finds below the threshold after exact rounding
 * gets LDX_BELOW_THRES set in its packed word instead of being removed. */
int32_t ldx_triangle_dev(ldx_store *store, const int64_t *rows, int64_t v, int32_t measure,
                         int32_t has_thres, int32_t thres_e4, int32_t engine,
                         uint32_t *dev_packed, int32_t *dev_n11);
int32_t ldx_triangle_rows_dev(ldx_store *store, const int64_t *rows, int64_t v, int64_t row_begin,
                              int64_t row_end, int32_t measure, int32_t has_thres, int32_t thres_e4,
                              int32_t engine, uint32_t *dev_packed, int32_t *dev_n11);
int32_t ldx_window_dev(ldx_store *store, const int64_t *q_row, const int64_t *lo, const int64_t *hi,
                       const int32_t *win_start, const int32_t *win_end, int64_t nq,
                       int32_t measure, int32_t thres_e4, ldx_hit *dev_hits, int64_t cap,
                       int64_t *dev_n_hits);
int32_t ldx_resolve(ldx_ctx *ctx, int64_t *n_fixed_out);
/* Several independent variant sets in ONE launch of the all-pairs engine.  The reference's unit of work is one matrix per
 * (source file, chromosome) (ld_triangle.py:80-88; its own fan-out over them is the process pool of :406-408); a
 * 2,000-variant matrix alone is a single wave of tiles on 148 SMs and is bound by launch, pipeline-fill and drain
 * latencies, a batch of them is not.  Set k is the lower triangle of rows[0..v) of its store (as ldx_triangle_dev:
 * v*(v-1)/2 words into dev_packed, dev_n11 may be NULL).  Sets that share the haplotype count and the number of selected
 * haplotypes (e.g. the chromosomes of one sample selection) go into the same launch, up to 32 per launch; anything else
 * falls back to one ldx_triangle_dev call per set.  Enqueue only: ldx_resolve() as for the other *_dev calls.
 * rows[] arrays are HOST arrays. */
typedef struct {
    ldx_store *store;
    const int64_t *rows;
    int64_t v;
    uint32_t *dev_packed;
    int32_t *dev_n11;
} ldx_triangle_set;
int32_t ldx_triangle_batch_dev(ldx_ctx *ctx, const ldx_triangle_set *sets, int32_t n_sets, int32_t measure,
                               int32_t has_thres, int32_t thres_e4, int32_t engine);

/* ---------------------------------------------------------------- matrix text
 * Replaces the table writer's body loop, ld_triangle.py:356-360: `'\t'.join(map(str, ld_two_dim[row]))` for
 * matrix rows row_begin .. row_end-1 of a v x v matrix, formatted on the GPU from the packed words.
 * `packed` is what ldx_triangle_rows[_dev] produced for the same row range (entry 0 = pair (row_begin, 0));
 * it is a device pointer with LDX_TEXT_PACKED_ON_DEVICE in flags, else a host array.  Line r is
 *     prefixes[prefix_off[r] .. prefix_off[r+1])  +  cells 0..v-1 joined by '\t'  +  '\n'
 * (prefix_off has v+1 entries, indexed by absolute matrix row; the reference's prefix is rsID '\t' position '\t').
 * Cell (r, c): c >= r -> "0" (the matrix starts as int zeros, :114; only row > col is filled, :150); LDX_BELOW_THRES or the measure's INT0 flag
 * -> "0" (:223-225, calc_ld.py:68-90); otherwise str(k / 10000.0) exactly as Python prints it ("0.0", "0.5",
 * "0.8125", "1.0").  text[cap] (device memory with LDX_TEXT_OUT_ON_DEVICE) receives *n_bytes bytes; when
 * cap is too small the call returns LDX_ERR_CAPACITY with *n_bytes = the size needed (cap = 0 is a size query;
 * 7 * v * (row_end - row_begin) + the prefix bytes always suffices).  Pending *_dev calls are resolved first.
 * With ldx_kernel_timing enabled the text kernel's launches are counted like the all-pairs kernels'. */
enum { LDX_TEXT_PACKED_ON_DEVICE = 1, LDX_TEXT_OUT_ON_DEVICE = 2 };
int32_t ldx_triangle_text(ldx_ctx *ctx, const uint32_t *packed, int64_t v, int64_t row_begin, int64_t row_end,
                          int32_t measure, const char *prefixes, const int64_t *prefix_off, int32_t flags,
                          char *text, int64_t cap, int64_t *n_bytes);
/* ldx_triangle_rows_dev + ldx_resolve + ldx_triangle_text in one call: the pairs of matrix rows row_begin..row_end-1
 * (row_begin a multiple of 128) are computed into the ctx's scratch and only their text leaves the GPU -- what
 * `ld_triangle -o table` needs (ld_triangle.py:133-230 and :356-360 together).  flags: 0 or LDX_TEXT_OUT_ON_DEVICE. */
int32_t ldx_triangle_table(ldx_store *store, const int64_t *rows, int64_t v, int64_t row_begin, int64_t row_end,
                           int32_t measure, int32_t has_thres, int32_t thres_e4, int32_t engine,
                           const char *prefixes, const int64_t *prefix_off, int32_t flags,
                           char *text, int64_t cap, int64_t *n_bytes);
/* ---------------------------------------------------------------- ld_area rows
 * Replaces the per-hit writer of ld_area.py:252-283 for all queries of a window scan at once.  hits[n_hits] as ldx_window
 * returned them (sorted by (query, row)); q_row[nq] = the queries' store rows (the dist column, :272); blob / blob_off / rows =
 * the records' fixed columns and field offsets (ldx_vcf_copy_prefixes, ldx_store_ingest_vcf; n_rows of them); the alt_freq
 * column is p_e4[row] (ldx_store_counts: var_2_alt_freq of calc_ld.py:97 for complete genotypes).  overrides (may be NULL):
 * [n_hits][3] = {alt_freq, r2, D'} * 10^4 per hit, each < 0 = "as above": for the general route, where the alt frequency belongs
 * to the PAIR (n1 / len(zip(...))) and a pairing of lists of unequal ploidy can exceed the packed word's 1.6383.  Host code, `threads` host threads (<= 0: all cores).  *text_out is allocated by the library (release it
 * with ldx_free_host); query k's text is (*text_out)[query_off[k], query_off[k+1]) (query_off has nq + 1 entries; *n_bytes =
 * query_off[nq]):
 *   LDX_AREA_TSV    one line per hit: pos, rsID, ref, alt, type, alt_freq, r2, D', dist joined by tabs (str() of each, :264-274)
 *   LDX_AREA_JSON   per hit the element json.dump(trg_obj, fh, indent=4) writes (:275-283) preceded by its ",\n    " separator:
 *                   the caller writes json.dumps([meta, query], indent=4)[:-2], the chunk, then "\n]"
 *   LDX_AREA_RSIDS  one rsID per line (:258-260) */
enum { LDX_AREA_TSV = 0, LDX_AREA_JSON = 1, LDX_AREA_RSIDS = 2 };
int32_t ldx_area_format(const ldx_hit *hits, int64_t n_hits, const int64_t *q_row, int64_t nq, const uint8_t *blob,
                        const int64_t *blob_off, const ldx_vcf_row *rows, int64_t n_rows, const int32_t *p_e4,
                        const int32_t *overrides, int32_t format, int32_t threads, char **text_out,
                        int64_t *n_bytes, int64_t *query_off);
/* Host helper (no device needed): str(value_e4 / 10000.0) for 0 <= value_e4 < 20000 as the kernels print it,
 * NUL-padded to 8 bytes -- the same digit arithmetic, exposed so that it can be checked against Python's str(). */
int32_t ldx_format_e4(int32_t value_e4, char *out8);

#ifdef __cplusplus
}
#endif
#endif /* LDX_H */
