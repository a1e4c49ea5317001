// ldx_internal.h -- host-side structs and kernel launchers shared by the .cu files of libldx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ldx_common.cuh"

struct ldx_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;        // the stream kernels are launched on
    int64_t launches = 0;
    // near-tie fix-up list (device) + pinned host mirror
    ldx::FixupRec *d_fix = nullptr;
    uint32_t *d_fix_count = nullptr;      // [0] = records appended (may exceed capacity)
    uint32_t fix_capacity = 0;
    ldx::FixupRec *h_fix = nullptr;       // pinned
    uint32_t *h_fix_count = nullptr;      // pinned
    // What the fix-ups on the device list refer to: one entry per *_dev call enqueued since the last
    // ldx_resolve().  Call number k (1-based index into pending_q) tags its records with k << 48 in
    // out_index (fix_tag while its kernels are being launched); tag 0 = a host-buffer call, which
    // collects its own records before returning.
    struct Pending {
        int kind = 0;                     // 1 packed array, 2 hit array
        void *dev_out = nullptr;
        double n_hap = 0;
        int measure = 0, has_thres = 0, thres_e4 = 0;
    };
    std::vector<Pending> pending_q;
    uint64_t fix_tag = 0;
    // scratch for small per-call index arrays
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    // MMA-path scratch (expanded operand panels), grown on demand
    void *d_mma_ops = nullptr;
    size_t mma_ops_bytes = 0;
    int mma_tile_n = 0;                   // tcgen05 tile width override (0 = heuristic)
    // tile list cached in d_mma_ops: the key is everything the list depends on -- {tile width, pair mode, then per set
    // (v, row_begin)} -- plus the place in the scratch block it was uploaded to (depends on the haplotype count too)
    std::vector<int64_t> mma_tiles_key;
    const void *mma_tiles_ptr = nullptr;
    int mma_pair = -1;                    // LDX_TUNE_MMA_PAIR: 256 x 128 tiles on CTA pairs (tcgen05 cta_group::2): 0 off, 1 on, -1 auto
    int mma_direct = -1;                  // LDX_TUNE_MMA_DIRECT: one-wave calls on contiguous store rows read the planes through a
                                          // TMA tensor map (no gather kernel): 0 off, 1 / -1 (default) on
    // completion mailbox: pinned, device-mapped {seq, near-tie count, error flag}; a 1-thread kernel
    // publishes it after each *_dev call so that ldx_resolve() can poll host memory instead of
    // paying a stream synchronisation (tens of microseconds) per call
    volatile uint32_t *h_mailbox = nullptr;
    uint32_t *d_mailbox = nullptr;
    uint32_t seq = 0;
    // host copy of the variant list(s) last staged on the device, and what it was validated against: the same list may
    // arrive for another (smaller) store on the same ctx, and must then be bounds-checked again
    std::vector<int64_t> rows_cache;
    std::vector<int64_t> rows_cache_key;  // per set {store address, its n_variants, v}
    std::vector<char> rows_cache_contig;  // per set: rows[k] == rows[0] + k
    bool rows_cache_valid = false;
    unsigned long long *d_trace = nullptr;   // diagnostics: globaltimer stamps written by the tcgen05 kernel
    void *h_stage = nullptr;              // pinned bounce buffer of the store files (allocated on first use)
    uint8_t *h_lists = nullptr;           // pinned staging of ldx_calc_ld_lists: both genotype lists in, the result out
    size_t h_lists_bytes = 0;
    uint8_t *h_win = nullptr;             // pinned staging of a multi-query window scan's work lists (one copy per call)
    size_t h_win_bytes = 0;
    cudaEvent_t win_staged = nullptr;     // recorded behind the copy out of h_win
    int window_mq = 1;                    // LDX_TUNE_WINDOW_MQ: 0 = ld_area scans always use the one-query-per-pass kernel
    int defer_cap = 0;                    // LDX_TUNE_DEFER_CAP: capacity of the deferred-pair lists (0 = sized from the pair count)
    int mma_min_v = 256;                  // ENGINE_AUTO uses the tcgen05 engine from this many variants
    // dominant-kernel timing (ldx_kernel_timing): CUDA event pairs around the all-pairs / window kernel
    bool timing = false;
    std::vector<cudaEvent_t> timing_events;   // pool, used pairwise
    size_t timing_used = 0;               // events recorded since the last read
    double timing_ms = 0.0;               // accumulated by drains
    int64_t timing_launches = 0;
};

struct ldx_store {
    ldx_ctx *ctx = nullptr;
    int64_t n_variants = 0;
    int32_t n_hap = 0;          // haplotype columns in the planes
    int32_t words = 0;          // ceil(n_hap / 64)
    int32_t stride_words = 0;   // row pitch in uint64 (multiple of 16)
    uint64_t *d_planes = nullptr;
    uint64_t *d_mask = nullptr; // [stride_words]
    ldx::VarFreq *d_freq = nullptr;   // [n_variants], valid once mask_set
    int32_t n_sel = 0;          // N = popcount(mask)
    bool mask_set = false;
    ldx::FinalCtx fc{};
    // annotations for the fused window filters
    int32_t *d_pos0 = nullptr, *d_end0 = nullptr;
    int64_t *d_idnum = nullptr;
    uint8_t *d_eligible = nullptr;
    bool annotated = false;
    // Variants that are not complete phased diploid 0/1 rows with the store's common ploidy pattern (see
    // "the general route" in ldx_common.cuh).  kind[v]: -1 simple, g >= 0 = index of its aux planes, -2 = all slots present and
    // 0/1 without aux planes (only when the common pattern is not all-diploid).  All null / zero for a store without such rows.
    int32_t *d_kind = nullptr;            // [n_variants]
    uint64_t *d_aux = nullptr;            // [n_general][2][stride_words]: present, ref
    int64_t n_general = 0;                // rows with aux planes
    int64_t n_nonsimple = 0;              // rows with kind != -1
    uint64_t *d_common = nullptr;         // [stride_words] the common presence pattern (all 2 * n_samples slots unless told otherwise)
    uint64_t *d_all_slots = nullptr;      // [stride_words] n_hap ones
    uint64_t *d_mask_user = nullptr;      // [stride_words] the selection as given (d_mask = selection & common pattern)
    ldx::GenStore *d_gen = nullptr;       // device copy of the pointers above, for the kernels' general route
    std::vector<int64_t> aux_rows;        // host: aux slot k holds the planes of row aux_rows[k] (-1: slot abandoned)
    int64_t aux_capacity = 0;             // slots allocated in d_aux
    bool classify_dirty = false;          // rows were (re)packed since kind[] / the common pattern were last derived
    bool common_loaded = false;           // the common pattern came with a store file / a subset: do not re-derive it
    std::vector<uint64_t> h_common;       // host copy of d_common (ldx_store_set_mask folds it into the selection)
    uint64_t *h_mask_stage = nullptr;     // pinned [2][stride_words]: the masks on their way to the device (set_mask does not wait for them)
    cudaEvent_t mask_staged = nullptr;    // recorded behind the copies out of h_mask_stage
    int32_t *d_row_len = nullptr;         // [n_variants] K2: len of the variant's own genotype list under the selection (calc_ld.py:31)
    int32_t *d_row_n1 = nullptr;          // [n_variants] K2: its alt count under the selection (the true one, also for general rows)
    // TMA tensor map of the planes ([n_variants][stride_words * 2] uint32, box 8 x 128), encoded on first use by the
    // tcgen05 engine's direct mode (ldx_triangle_mma.cu); 128 bytes, 64-byte aligned as the driver requires
    alignas(64) unsigned char tmap[128] = {};
    bool tmap_ready = false;
};
constexpr int64_t STORE_FREQ_PAD = 512;   // zeroed VarFreq entries past the last variant: the all-pairs kernels read whole tiles

namespace ldx {

int set_error(int code, const std::string &msg);   // returns code
int cuda_fail(cudaError_t e, const char *what);    // records + returns LDX_ERR_CUDA

// After every kernel launch: the launch error, and -- with LDX_DEBUG_SYNC=1 in the environment -- a stream
// synchronisation, so that a faulting kernel is reported at its own launch site.
int after_launch(ldx_ctx *ctx, const char *kernel);
#define LDX_LAUNCHED(ctx, name) do { int rc__ = ldx::after_launch(ctx, name); if (rc__ != LDX_OK) return rc__; } while (0)

#define LDX_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return ldx::cuda_fail(e__, #call);       \
    } while (0)

// ---- kernel launchers (each enqueues on ctx->stream and bumps ctx->launches)
int launch_pack_gt(ldx_ctx *ctx, const uint8_t *d_text, const int64_t *d_row_off, int64_t row_pitch,
                   int64_t n_rows, int32_t n_samples, uint64_t *d_planes_first, int32_t stride_words,
                   uint8_t *d_status);
int launch_variant_freq(ldx_store *s);
// GT text of rows whose fields are not all plain "a|b" with a, b in {0, 1} (K1's status bit 0): any ploidy <= 2, '.' and other
// allele codes, '/' or '|', sub-fields after ':'.  d_rows_idx[n] = the rows (relative to d_planes_first), d_row_off their GT
// text offsets; writes the alt plane, and present / ref planes into d_aux_out[k] (k = position in d_rows_idx);
// d_status_out[k] bit 3 = a field with more than two alleles or a row that ends early.
int launch_pack_gt_general(ldx_ctx *ctx, const uint8_t *d_text, int64_t text_bytes, const int64_t *d_row_off, int64_t row_pitch, const int64_t *d_rows_idx,
                           int64_t n, int32_t n_samples, uint64_t *d_planes_first, int32_t stride_words, uint64_t *d_aux_out, uint8_t *d_status_out);
// After parsing: the common presence pattern and kind[] of every row (s->aux_rows says which rows have aux planes).
int store_classify_rows(ldx_store *s);
// Host side of K1 for rows the fast kernel flagged (status bit 0): aux slots, the general parser, bookkeeping.  d_status / h_status:
// the fast kernel's per-row status on the device / host (n_rows of them); h_status gets bit 3 for rows even the general parser rejects.
int store_pack_general(ldx_store *s, int64_t first_row, int64_t n_rows, const uint8_t *d_text, int64_t text_bytes, const int64_t *d_row_off,
                       int64_t row_pitch, int32_t n_samples, uint8_t *h_status);
int store_publish_gen(ldx_store *s);      // (re)build d_gen after the pointers changed
int launch_subset(const ldx_store *src, const int32_t *d_sel, ldx_store *dst);
int launch_subset_planes(ldx_ctx *ctx, const uint64_t *d_src, int32_t src_stride, const int32_t *d_sel, int32_t n_sel, int64_t n_rows, uint64_t *d_dst,
                         int32_t dst_stride);
int launch_pairs(ldx_store *s, const int64_t *d_ia, const int64_t *d_ib, int64_t n, int32_t *d_n11,
                 double *d_d, double *d_dp, double *d_r2, uint32_t *d_packed);
int launch_finalise_counts(ldx_ctx *ctx, const ldx::FinalCtx &fc, const int32_t *d_n11, const int32_t *d_n1a,
                           const int32_t *d_n1b, int64_t n, double *d_d, double *d_dp, double *d_r2, uint32_t *d_packed);
int launch_lists(ldx_ctx *ctx, const uint8_t *d_ga, int64_t len_a, const uint8_t *d_gb, int64_t len_b,
                 ldx_ld_result *d_out);
int launch_window(ldx_store *s, const int64_t *d_qrow, const int64_t *d_lo, const int64_t *d_hi,
                  const int32_t *d_ws, const int32_t *d_we, const int64_t *d_chunk_prefix, int64_t nq,
                  int64_t n_chunks, int measure, int thres_e4, ldx_hit *d_hits, int64_t cap,
                  unsigned long long *d_n_hits);
// Both all-pairs launchers work on the first v entries of d_rows and emit the pairs (row, col < row) of rows
// row_begin .. v-1 (row_begin a multiple of 128); output index 0 is pair (row_begin, 0).
int launch_triangle_popc(ldx_store *s, const int64_t *d_rows, int64_t v, int64_t row_begin, int measure, int has_thres,
                         int thres_e4, uint32_t *d_packed, int32_t *d_n11);
// The tcgen05 engine works on up to MMA_MAX_SETS variant sets per launch (ldx_triangle_batch_dev): set k is the first v
// entries of d_rows + rows_off and emits rows row_begin .. v-1 (row_begin = 0 unless n_sets == 1) into d_packed / d_n11;
// fix_tag is ORed into the out_index of its near-tie records.  All sets share the haplotype and selected-haplotype counts.
// contiguous: rows[k] == rows[0] + k (set 0 of a single-set call; host knowledge from stage_rows) -> direct mode may apply.
// publish_seq != 0: the engine's last kernel also publishes the completion record for that sequence number.
constexpr int MMA_MAX_SETS = 32;
struct MmaSetDesc {
    ldx_store *s; int64_t rows_off; int64_t v, row_begin; uint32_t *d_packed; int32_t *d_n11; uint64_t fix_tag; int64_t row0; bool contiguous;
};
int launch_triangle_mma(ldx_ctx *ctx, const MmaSetDesc *sets, int n_sets, const int64_t *d_rows, int measure, int has_thres,
                        int thres_e4, uint32_t publish_seq);
bool triangle_mma_available();
int triangle_mma_max_haplotypes();
int launch_publish(ldx_ctx *ctx);   // enqueue the mailbox update for ctx->seq
// Dominant-kernel timing: bracket a launch with events on ctx->stream when ctx->timing is on.
void timing_begin(ldx_ctx *ctx);
void timing_end(ldx_ctx *ctx);

// BGZF pieces of the ingest path (ldx_inflate.cu), shared with the slab-wise file ingest (ldx_vcf.cu)
struct BgzfMember { size_t in_off, in_len, out_off, out_len, data_off; };
bool bgzf_scan(const uint8_t *in, size_t n, std::vector<BgzfMember> &members, size_t *total_out);
bool bgzf_inflate_range(const uint8_t *in, const std::vector<BgzfMember> &members, size_t first, size_t last, uint8_t *out, int threads);
int read_whole_file(const char *path, std::vector<uint8_t> &in);

constexpr int WINDOW_CHUNK = 256;   // rows per work item of the window kernel
constexpr int WINDOW_MQ = 4;        // queries per work item of the multi-query window kernel (ldx_window.cu: MQ)
struct WindowMqBlock { int64_t base, first_item; int32_t a, b; };   // = ldx::MqBlock (ldx_window.cu)
// One record per query in sorted order (= ldx::MqQueryX): the host fills q .. we, a small kernel adds the query row's idnum, n1
// and its half of the r2 screen; the scan kernels read nothing else about a query.
struct WindowMqQueryX { int64_t idnum; int32_t q, qrow, lo, hi, ws, we, n1, pad[3]; };
bool window_mq_supported(const ldx_store *s);
int launch_window_mq(ldx_store *s, int64_t nq, const void *d_blocks, int64_t n_blocks, void *d_ext, int64_t n_sorted, unsigned int *d_next,
                     int measure, int thres_e4, ldx_hit *d_hits, int64_t cap, unsigned long long *d_counters);
constexpr size_t WINDOW_MQ_EXT_BYTES = 48;

}  // namespace ldx
