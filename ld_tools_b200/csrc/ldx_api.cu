// ldx_api.cu -- the C ABI declared in include/ldx.h: context, store, and the host-side halves of
// the compute entry points (staging, launch, copy-back, near-tie settlement, hit ordering).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cub/cub.cuh>

#include "ldx_internal.h"
#include "ldx_fixup.cuh"

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

namespace ldx {
int set_error(int code, const std::string &msg) { g_last_error = msg; return code; }
int cuda_fail(cudaError_t e, const char *what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();   // clear the sticky-less error state
    return LDX_ERR_CUDA;
}
int after_launch(ldx_ctx *ctx, const char *kernel) {
    static const bool debug_sync = getenv("LDX_DEBUG_SYNC") != nullptr;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug_sync) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, kernel);
    return LDX_OK;
}
}  // namespace ldx
using namespace ldx;

#define LDX_TRY(expr) do { int rc__ = (expr); if (rc__ != LDX_OK) return rc__; } while (0)
#define LDX_REQUIRE(cond, msg) do { if (!(cond)) return set_error(LDX_ERR_ARG, msg); } while (0)

// grow-only device scratch owned by the ctx (avoids cudaMalloc/cudaFree on every call)
struct Arena {
    enum { SLOTS = 13 };
    void *ptr[SLOTS] = {};
    size_t bytes[SLOTS] = {};
};
static Arena *arena_of(ldx_ctx *ctx) { return reinterpret_cast<Arena *>(ctx->d_scratch); }

static int arena_get(ldx_ctx *ctx, int slot, size_t bytes, void **out) {
    Arena *a = arena_of(ctx);
    if (bytes == 0) bytes = 16;
    if (a->bytes[slot] < bytes) {
        cudaSetDevice(ctx->device);                 // an allocation lands on the CURRENT device: a fresh host thread starts on device 0
        if (a->ptr[slot]) { cudaStreamSynchronize(ctx->stream); cudaFree(a->ptr[slot]); a->ptr[slot] = nullptr; a->bytes[slot] = 0; }
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&a->ptr[slot], want);
        if (e != cudaSuccess) { cudaGetLastError(); return set_error(LDX_ERR_NOMEM, "device scratch allocation failed"); }
        a->bytes[slot] = want;
    }
    *out = a->ptr[slot];
    return LDX_OK;
}
enum { S_IA = 0, S_IB, S_PACKED, S_N11, S_D, S_DP, S_R2, S_TEXT, S_ROWOFF, S_STATUS, S_HITS, S_MISC, S_ROWS };

namespace ldx {
// Arena blocks for code outside this file (ldx_vcf.cu): 0 = the text slot, 1 = row offsets, 2 = status.
int scratch_get(ldx_ctx *ctx, int which, size_t bytes, void **out) {
    static const int slot[3] = {S_TEXT, S_ROWOFF, S_STATUS};
    if (which < 0 || which > 2) return set_error(LDX_ERR_ARG, "bad scratch block");
    return arena_get(ctx, slot[which], bytes, out);
}
// A whole-chromosome ingest leaves blocks as large as its text (tens of GB) in the grow-only arena: give those back.
void scratch_trim(ldx_ctx *ctx, size_t keep_bytes) {
    Arena *a = arena_of(ctx);
    static const int slot[3] = {S_TEXT, S_ROWOFF, S_STATUS};
    for (int k = 0; k < 3; ++k)
        if (a->bytes[slot[k]] > keep_bytes) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(a->ptr[slot[k]]);
            a->ptr[slot[k]] = nullptr; a->bytes[slot[k]] = 0;
        }
}
}  // namespace ldx

// ------------------------------------------------------------------------------------------ lifecycle
extern "C" int32_t ldx_abi_version(void) { return LDX_ABI_VERSION; }
extern "C" const char *ldx_last_error(void) { return g_last_error.c_str(); }

extern "C" int32_t ldx_device_count(int32_t *n_out) {
    LDX_REQUIRE(n_out, "n_out is NULL");
    int n = 0;
    LDX_CUDA(cudaGetDeviceCount(&n));
    *n_out = n;
    return LDX_OK;
}

extern "C" int32_t ldx_init(int32_t device, ldx_ctx **ctx_out) {
    LDX_REQUIRE(ctx_out, "ctx_out is NULL");
    *ctx_out = nullptr;
    int n = 0;
    LDX_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) return set_error(LDX_ERR_CUDA, "no CUDA device: libldx has no CPU fallback");
    if (device < 0) LDX_CUDA(cudaGetDevice(&device));
    LDX_REQUIRE(device < n, "device index out of range");
    LDX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LDX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(LDX_ERR_CUDA, std::string("libldx is built for sm_100a only; device is ") + prop.name);
    ldx_ctx *ctx = new (std::nothrow) ldx_ctx();
    if (!ctx) return set_error(LDX_ERR_NOMEM, "ctx allocation failed");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->d_scratch = new (std::nothrow) Arena();
    ctx->fix_capacity = 1u << 20;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fix, sizeof(FixupRec) * (size_t)ctx->fix_capacity);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fix_count, 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_fix_count, 0, 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_fix_count, 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&ctx->h_mailbox, 4 * sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) { for (int i = 0; i < 4; ++i) ctx->h_mailbox[i] = 0; e = cudaHostGetDevicePointer((void **)&ctx->d_mailbox, (void *)ctx->h_mailbox, 0); }
    if (e != cudaSuccess) { ldx_destroy(ctx); return cuda_fail(e, "ldx_init"); }
    ctx->stream = ctx->own_stream;
    if (const char *e = getenv("LDX_MMA_PAIR")) ctx->mma_pair = atoi(e) < 0 ? -1 : atoi(e) != 0;    // default of LDX_TUNE_MMA_PAIR
    *ctx_out = ctx;
    return LDX_OK;
}

extern "C" int32_t ldx_destroy(ldx_ctx *ctx) {
    if (!ctx) return LDX_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (Arena *a = arena_of(ctx)) {
        for (int i = 0; i < Arena::SLOTS; ++i) if (a->ptr[i]) cudaFree(a->ptr[i]);
        delete a;
    }
    if (ctx->d_fix) cudaFree(ctx->d_fix);
    if (ctx->d_fix_count) cudaFree(ctx->d_fix_count);
    if (ctx->h_fix_count) cudaFreeHost(ctx->h_fix_count);
    if (ctx->h_mailbox) cudaFreeHost((void *)ctx->h_mailbox);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_lists) cudaFreeHost(ctx->h_lists);
    if (ctx->h_win) cudaFreeHost(ctx->h_win);
    if (ctx->win_staged) cudaEventDestroy(ctx->win_staged);
    if (ctx->d_mma_ops) cudaFree(ctx->d_mma_ops);
    if (ctx->d_trace) cudaFree(ctx->d_trace);
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return LDX_OK;
}

extern "C" int32_t ldx_set_stream(ldx_ctx *ctx, void *cuda_stream) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return LDX_OK;
}

extern "C" int32_t ldx_use_own_stream(ldx_ctx *ctx) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    ctx->stream = ctx->own_stream;
    return LDX_OK;
}

extern "C" int32_t ldx_set_tuning(ldx_ctx *ctx, int32_t key, int32_t value) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    switch (key) {
        case LDX_TUNE_MMA_TILE_N:
            LDX_REQUIRE(value == 0 || value == 64 || value == 128, "tile width must be 0, 64 or 128");
            ctx->mma_tile_n = value;
            return LDX_OK;
        case LDX_TUNE_MMA_PAIR:
            LDX_REQUIRE(value >= -1 && value <= 1, "pair mode must be -1 (auto), 0 or 1");
            ctx->mma_pair = value;
            return LDX_OK;
        case LDX_TUNE_MMA_MIN_V:
            LDX_REQUIRE(value >= 2, "minimum variant count must be >= 2");
            ctx->mma_min_v = value;
            return LDX_OK;
        case LDX_TUNE_WINDOW_MQ:
            LDX_REQUIRE(value == 0 || value == 1, "window multi-query mode must be 0 or 1");
            ctx->window_mq = value;
            return LDX_OK;
        case LDX_TUNE_MMA_DIRECT:
            LDX_REQUIRE(value >= -1 && value <= 1, "direct mode must be -1 (auto), 0 or 1");
            ctx->mma_direct = value;
            return LDX_OK;
        case LDX_TUNE_DEFER_CAP:
            LDX_REQUIRE(value >= 0, "deferred-pair list capacity must be >= 0");
            ctx->defer_cap = value;
            return LDX_OK;
        default:
            return set_error(LDX_ERR_ARG, "unknown tuning key");
    }
}

extern "C" int32_t ldx_debug_trace(ldx_ctx *ctx, int32_t enable, uint64_t *stamps8) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    LDX_CUDA(cudaSetDevice(ctx->device));
    if (enable && !ctx->d_trace) {
        LDX_CUDA(cudaMalloc(&ctx->d_trace, 4096 * sizeof(unsigned long long)));
        LDX_CUDA(cudaMemset(ctx->d_trace, 0, 4096 * sizeof(unsigned long long)));
    }
    if (stamps8 && ctx->d_trace) {
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));
        LDX_CUDA(cudaMemcpy(stamps8, ctx->d_trace, 4096 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    if (!enable && ctx->d_trace) { cudaFree(ctx->d_trace); ctx->d_trace = nullptr; }
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ kernel timing
namespace ldx {
static void timing_drain(ldx_ctx *ctx) {
    if (ctx->timing_used == 0) return;
    cudaEventSynchronize(ctx->timing_events[ctx->timing_used - 1]);
    for (size_t i = 0; i + 1 < ctx->timing_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->timing_events[i], ctx->timing_events[i + 1]) == cudaSuccess) ctx->timing_ms += ms;
        ctx->timing_launches++;
    }
    ctx->timing_used = 0;
}
void timing_begin(ldx_ctx *ctx) {
    if (!ctx->timing) return;
    if (ctx->timing_used + 2 > 8192) timing_drain(ctx);
    while (ctx->timing_events.size() < ctx->timing_used + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); ctx->timing = false; return; }
        ctx->timing_events.push_back(e);
    }
    cudaEventRecord(ctx->timing_events[ctx->timing_used], ctx->stream);
}
void timing_end(ldx_ctx *ctx) {
    if (!ctx->timing) return;
    cudaEventRecord(ctx->timing_events[ctx->timing_used + 1], ctx->stream);
    ctx->timing_used += 2;
}
}  // namespace ldx

extern "C" int32_t ldx_kernel_timing(ldx_ctx *ctx, int32_t enable, double *ms_out, int64_t *launches_out) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    LDX_CUDA(cudaSetDevice(ctx->device));
    timing_drain(ctx);
    if (ms_out) *ms_out = ctx->timing_ms;
    if (launches_out) *launches_out = ctx->timing_launches;
    ctx->timing_ms = 0.0; ctx->timing_launches = 0;
    ctx->timing = enable != 0;
    return LDX_OK;
}

extern "C" int32_t ldx_synchronize(ldx_ctx *ctx) {
    if (ctx) cudaSetDevice(ctx->device);
    LDX_REQUIRE(ctx, "ctx is NULL");
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));
    return LDX_OK;
}

extern "C" int32_t ldx_sm_count(ldx_ctx *ctx, int32_t *n_out) {
    LDX_REQUIRE(ctx && n_out, "NULL argument");
    *n_out = ctx->sm_count;
    return LDX_OK;
}

extern "C" int32_t ldx_dev_alloc(ldx_ctx *ctx, int64_t bytes, void **dev_ptr_out) {
    LDX_REQUIRE(ctx && dev_ptr_out && bytes >= 0, "bad argument");
    *dev_ptr_out = nullptr;
    LDX_CUDA(cudaSetDevice(ctx->device));
    if (cudaMalloc(dev_ptr_out, (size_t)std::max<int64_t>(bytes, 256)) != cudaSuccess) {
        cudaGetLastError();
        return set_error(LDX_ERR_NOMEM, "device allocation failed");
    }
    return LDX_OK;
}

extern "C" int32_t ldx_dev_free(ldx_ctx *ctx, void *dev_ptr) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    if (!dev_ptr) return LDX_OK;
    LDX_CUDA(cudaSetDevice(ctx->device));
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));          // enqueued kernels may still write it
    LDX_CUDA(cudaFree(dev_ptr));
    return LDX_OK;
}

extern "C" int32_t ldx_launch_count(ldx_ctx *ctx, int64_t *n_out) {
    LDX_REQUIRE(ctx && n_out, "NULL argument");
    *n_out = ctx->launches;
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------
// Near-tie settlement.  The kernels compute d*d; CPython computes pow(d, 2.0) (calc_ld.py:87),
// which glibc rounds differently for ~0.1% of inputs (1 ulp).  That can only change the printed
// value when r2*10^4 is within an ulp of k + 0.5; the kernels flag a far wider band (1e-9) and
// the few flagged pairs are re-evaluated here from their integer counts with the real libm pow.
// gcc folds pow(x, 2.0) into x*x, hence the volatile pointer.
static double (*volatile libm_pow)(double, double) = pow;

static double host_round4_e4(double x) {
    const double hi = x * 1.0e4;
    const double lo = std::fma(x, 1.0e4, -hi);
    const double k = std::floor(hi);
    const double frac = hi - k;
    if (frac > 0.5 || (frac == 0.5 && lo > 0.0)) return k + 1.0;
    if (frac == 0.5 && lo == 0.0) return k + (double)(((long long)k) & 1);
    return k;
}

// calc_ld.py:33-97 on the host for biallelic complete data; returns the packed word.
static uint32_t host_finalise_packed(double N, int32_t n11, int32_t n1a, int32_t n1b, double *r2_raw) {
    const double f11 = (double)n11 / N;
    const double pa = (double)n1a / N, qa = (double)((int32_t)N - n1a) / N;
    const double pb = (double)n1b / N, qb = (double)((int32_t)N - n1b) / N;
    const double t = pa * pb;
    const double d = f11 - t;
    double bound;
    if (d >= 0.0) { const double x = pa * qb, y = qa * pb; bound = (y < x) ? y : x; }
    else { const double x = -t, y = -(qa * qb); bound = (y > x) ? y : x; }
    if (r2_raw) *r2_raw = 0.0;
    if (bound == 0.0) return LDX_DP_INT0 | LDX_R2_INT0;
    const double dp = d / bound;
    uint32_t word = ((uint32_t)host_round4_e4(dp)) << LDX_DP_SHIFT;
    if (dp != 0.0) {
        const double r2 = libm_pow(d, 2.0) / (((pa * qa) * pb) * qb);
        if (r2_raw) *r2_raw = r2;
        word |= (uint32_t)host_round4_e4(r2);
    } else word |= LDX_R2_INT0;
    return word;
}

static inline int32_t word_measure(uint32_t w, int measure) {
    return measure == LDX_MEASURE_R2 ? (int32_t)(w & LDX_R2_MASK) : (int32_t)((w & LDX_DP_MASK) >> LDX_DP_SHIFT);
}

// Fetch (and clear) the device fix-up list.  Waits for the stream's work: through the mailbox
// (a host-memory poll) when the last enqueued call published one, else by synchronising.
static int collect_fixups(ldx_ctx *ctx, std::vector<FixupRec> &recs, bool use_mailbox = false) {
    recs.clear();
    uint32_t n, err;
    bool polled = false;
    uint64_t record = 0;                       // {sequence number : 32, near-tie count : 24, error flag : 8}, see publish_record
    if (use_mailbox && ctx->seq != 0) {
        const uint32_t want = ctx->seq;
        const volatile uint64_t *box = reinterpret_cast<const volatile uint64_t *>(ctx->h_mailbox);
        for (long spins = 0; spins < 200000000L; ++spins) {
            record = *box;
            if ((uint32_t)record == want) { polled = true; break; }
        }
    }
    if (polled) {
        std::atomic_thread_fence(std::memory_order_acquire);
        n = (uint32_t)(record >> 32) & 0xffffffu; err = (uint32_t)(record >> 56);
        if (n == 0xffffffu) n = ctx->fix_capacity + 1;      // saturated: reported as the overflow it is
    } else {
        // [0] = near-tie records appended, [1] = tcgen05 pipeline error flag
        LDX_CUDA(cudaMemcpyAsync(ctx->h_fix_count, ctx->d_fix_count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));
        n = ctx->h_fix_count[0]; err = ctx->h_fix_count[1];
    }
    if (err) {
        cudaMemsetAsync(ctx->d_fix_count, 0, 4 * sizeof(uint32_t), ctx->stream);
        if (err == 2) return set_error(LDX_ERR_CAPACITY, "tcgen05 engine: deferred-pair list overflow (degenerate input); use LDX_ENGINE_POPC for this variant set");
        return set_error(LDX_ERR_CUDA, "tcgen05 pipeline timed out (mbarrier wait exceeded 2 s); results are invalid");
    }
    if (n == 0) return LDX_OK;
    if (n > ctx->fix_capacity) {
        cudaMemsetAsync(ctx->d_fix_count, 0, sizeof(uint32_t), ctx->stream);
        return set_error(LDX_ERR_CAPACITY, "near-tie list overflow (more than 2^20 flagged pairs in one call)");
    }
    recs.resize(n);
    LDX_CUDA(cudaMemcpyAsync(recs.data(), ctx->d_fix, sizeof(FixupRec) * n, cudaMemcpyDeviceToHost, ctx->stream));
    LDX_CUDA(cudaMemsetAsync(ctx->d_fix_count, 0, sizeof(uint32_t), ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));
    return LDX_OK;
}

// A *_dev call: reserve the next slot of the pending queue and tag the fix-up records its kernels append.
static int resolve_pending(ldx_ctx *ctx, int64_t *n_fixed_out);
static int begin_dev_call(ldx_ctx *ctx) {
    if (ctx->pending_q.size() >= 4096) LDX_TRY(resolve_pending(ctx, nullptr));   // keeps the tag within 16 bits
    ctx->fix_tag = (uint64_t)(ctx->pending_q.size() + 1) << FIX_TAG_SHIFT;
    return LDX_OK;
}
static void end_dev_call(ldx_ctx *ctx, int kind, void *dev_out, double n_hap, int measure, int has_thres, int thres_e4) {
    ldx_ctx::Pending p;
    p.kind = kind; p.dev_out = dev_out; p.n_hap = n_hap; p.measure = measure; p.has_thres = has_thres; p.thres_e4 = thres_e4;
    ctx->pending_q.push_back(p);
    ctx->fix_tag = 0;
}
// Host-buffer entry points collect every record on the device list as their own: settle what enqueued
// device-resident calls left there first.
static int settle_before_host_call(ldx_ctx *ctx) {
    ctx->fix_tag = 0;
    return ctx->pending_q.empty() ? (int)LDX_OK : resolve_pending(ctx, nullptr);
}

// The general route's finalisation (ldx_common.cuh::finalise_general) on the host, with the real libm pow: calc_ld.py:33-97 from
// explicit counts.
static uint32_t host_finalise_general(const FixupRec &r, double *r2_raw) {
    if (r2_raw) *r2_raw = 0.0;
    if (r.n_pair <= 0) return LDX_DP_INT0 | LDX_R2_INT0;
    const double N = (double)r.n_pair;
    const double f11 = (double)r.n11 / N;
    const double pa = (double)r.n1a / N, qa = (double)r.n0a / N, pb = (double)r.n1b / N, qb = (double)r.n0b / N;
    const double t = pa * pb;
    const double d = f11 - t;
    double bound;
    if (d >= 0.0) { const double x = pa * qb, y = qa * pb; bound = (y < x) ? y : x; }
    else { const double x = -t, y = -(qa * qb); bound = (y > x) ? y : x; }
    if (bound == 0.0) return LDX_DP_INT0 | LDX_R2_INT0;
    const double dp = d / bound;
    uint32_t word = ((uint32_t)std::fmin(host_round4_e4(dp), 16383.0)) << LDX_DP_SHIFT;      // the 14-bit fields saturate (finalise_general)
    if (dp != 0.0) {
        const double r2 = libm_pow(d, 2.0) / (((pa * qa) * pb) * qb);
        if (r2_raw) *r2_raw = r2;
        word |= (uint32_t)std::fmin(host_round4_e4(r2), 16383.0);
    } else word |= LDX_R2_INT0;
    return word;
}
static uint32_t host_finalise_rec(const FixupRec &r, double N, double *r2_raw) {
    return r.n_pair != 0 ? host_finalise_general(r, r2_raw) : host_finalise_packed(N, r.n11, r.n1a, r.n1b, r2_raw);
}

static uint32_t settle_word(const FixupRec &r, double N, int measure, int has_thres, int thres_e4) {
    uint32_t w = host_finalise_rec(r, N, nullptr);
    if (has_thres && word_measure(w, measure) < thres_e4) w |= LDX_BELOW_THRES;
    return w;
}

// ------------------------------------------------------------------------------------------ lists
extern "C" int32_t ldx_calc_ld_lists(ldx_ctx *ctx, const uint8_t *g_a, int64_t len_a, const uint8_t *g_b,
                                     int64_t len_b, ldx_ld_result *out) {
    LDX_REQUIRE(ctx && out, "NULL argument");
    LDX_REQUIRE(len_a >= 0 && len_b >= 0 && (g_a || !len_a) && (g_b || !len_b), "bad genotype vectors");
    if (len_a == 0 || len_b == 0)
        return set_error(LDX_ERR_EMPTY, "division by zero");   // calc_ld.py:33 with an empty pairing
    LDX_CUDA(cudaSetDevice(ctx->device));
    uint8_t *d_buf; ldx_ld_result *d_out;
    const size_t off_b = ((size_t)len_a + 15) / 16 * 16;
    LDX_TRY(arena_get(ctx, S_TEXT, off_b + (size_t)len_b + 16, (void **)&d_buf));
    LDX_TRY(arena_get(ctx, S_MISC, sizeof(ldx_ld_result), (void **)&d_out));
    // both lists travel as ONE copy from the context's pinned staging (two pageable copies cost a runtime staging pass each), the
    // result comes back into it: a scalar call is two copies, one kernel and one wait
    const size_t res_off = (off_b + (size_t)len_b + 63) / 64 * 64, need = res_off + sizeof(ldx_ld_result);
    if (ctx->h_lists_bytes < need) {
        if (ctx->h_lists) cudaFreeHost(ctx->h_lists);
        ctx->h_lists = nullptr; ctx->h_lists_bytes = 0;
        const size_t cap = std::max<size_t>(need * 2, (size_t)1 << 16);
        LDX_CUDA(cudaMallocHost((void **)&ctx->h_lists, cap));
        ctx->h_lists_bytes = cap;
    }
    std::memcpy(ctx->h_lists, g_a, (size_t)len_a);
    std::memcpy(ctx->h_lists + off_b, g_b, (size_t)len_b);
    LDX_CUDA(cudaMemcpyAsync(d_buf, ctx->h_lists, off_b + (size_t)len_b, cudaMemcpyHostToDevice, ctx->stream));
    LDX_TRY(launch_lists(ctx, d_buf, len_a, d_buf + off_b, len_b, d_out));
    LDX_CUDA(cudaMemcpyAsync(ctx->h_lists + res_off, d_out, sizeof(ldx_ld_result), cudaMemcpyDeviceToHost, ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, ctx->h_lists + res_off, sizeof(ldx_ld_result));
    if (out->r2_is_int0 == 2) {   // near a rounding tie: the reference's libm pow decides (calc_ld.py:87)
        const double N = (double)out->n_hap;
        const double pa = (double)out->n_a1 / N, qa = (double)out->n_a0 / N;
        const double pb = (double)out->n_b1 / N, qb = (double)out->n_b0 / N;
        const double r2 = libm_pow(out->d, 2.0) / (((pa * qa) * pb) * qb);
        out->r2 = r2;
        out->r2_e4 = host_round4_e4(r2);
        out->r2_is_int0 = 0;
    }
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ counts
static int make_final_ctx(int64_t n_sel, FinalCtx *fc) {
    fc->n_hap = (double)n_sel;
    // Prove the three-operation quotient of div_by_n (ldx_common.cuh) equals the IEEE quotient for
    // every count that can occur.  RN(1/N) always passed in testing (Markstein's theorem); its
    // two neighbours are tried before giving up so that the device needs no fallback division.
    const double rn = 1.0 / (double)n_sel;
    const double cand[3] = {rn, std::nextafter(rn, 0.0), std::nextafter(rn, 2.0)};
    for (double rcp : cand) {
        bool ok = true;
        for (int64_t n = 0; n <= n_sel && ok; ++n) {
            const double x = (double)n, q0 = x * rcp;
            const double r = std::fma(-q0, fc->n_hap, x);
            ok = std::fma(r, rcp, q0) == x / fc->n_hap;
        }
        if (ok) { fc->rcp_n = rcp; return LDX_OK; }
    }
    return set_error(LDX_ERR_DATA, "no exact reciprocal found for this haplotype count");
}

extern "C" int32_t ldx_finalise_counts(ldx_ctx *ctx, int32_t n_hap, const int32_t *n11, const int32_t *n1a,
                                       const int32_t *n1b, int64_t n, double *d, double *dprime, double *r2,
                                       uint32_t *packed) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    LDX_REQUIRE(n >= 0 && (n == 0 || (n11 && n1a && n1b)), "bad count arrays");
    LDX_REQUIRE(n_hap <= (1 << 24), "n_hap out of range");
    if (n_hap <= 0) return set_error(LDX_ERR_EMPTY, "division by zero");
    for (int64_t k = 0; k < n; ++k)
        LDX_REQUIRE(n1a[k] >= 0 && n1a[k] <= n_hap && n1b[k] >= 0 && n1b[k] <= n_hap && n11[k] >= 0 &&
                    n11[k] <= n1a[k] && n11[k] <= n1b[k] && n1a[k] + n1b[k] - n11[k] <= n_hap, "inconsistent counts");
    if (n == 0) return LDX_OK;
    LDX_CUDA(cudaSetDevice(ctx->device));
    LDX_TRY(settle_before_host_call(ctx));
    FinalCtx fc;
    LDX_TRY(make_final_ctx(n_hap, &fc));
    int32_t *d_in; double *d_d = nullptr, *d_dp = nullptr, *d_r2 = nullptr; uint32_t *d_pk;
    LDX_TRY(arena_get(ctx, S_IA, sizeof(int32_t) * 3 * (size_t)n, (void **)&d_in));
    if (d) LDX_TRY(arena_get(ctx, S_D, sizeof(double) * (size_t)n, (void **)&d_d));
    if (dprime) LDX_TRY(arena_get(ctx, S_DP, sizeof(double) * (size_t)n, (void **)&d_dp));
    if (r2) LDX_TRY(arena_get(ctx, S_R2, sizeof(double) * (size_t)n, (void **)&d_r2));
    LDX_TRY(arena_get(ctx, S_PACKED, sizeof(uint32_t) * (size_t)n, (void **)&d_pk));
    cudaStream_t st = ctx->stream;
    LDX_CUDA(cudaMemcpyAsync(d_in, n11, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_in + n, n1a, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_in + 2 * n, n1b, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    LDX_TRY(launch_finalise_counts(ctx, fc, d_in, d_in + n, d_in + 2 * n, n, d_d, d_dp, d_r2, d_pk));
    if (d) LDX_CUDA(cudaMemcpyAsync(d, d_d, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (dprime) LDX_CUDA(cudaMemcpyAsync(dprime, d_dp, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (r2) LDX_CUDA(cudaMemcpyAsync(r2, d_r2, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (packed) LDX_CUDA(cudaMemcpyAsync(packed, d_pk, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    std::vector<FixupRec> recs;
    LDX_TRY(collect_fixups(ctx, recs));   // synchronises
    for (const FixupRec &r : recs) {
        double r2_exact;
        const uint32_t w = host_finalise_rec(r, fc.n_hap, &r2_exact);
        if (packed) packed[r.out_index] = w;
        if (r2) r2[r.out_index] = r2_exact;
    }
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ store
extern "C" int32_t ldx_store_create(ldx_ctx *ctx, int64_t n_variants, int32_t n_hap, ldx_store **store_out) {
    LDX_REQUIRE(ctx && store_out, "NULL argument");
    LDX_REQUIRE(n_variants >= 0 && n_variants < (1ll << 31), "n_variants out of range");
    LDX_REQUIRE(n_hap > 0 && n_hap <= (1 << 24), "n_hap out of range");
    *store_out = nullptr;
    LDX_CUDA(cudaSetDevice(ctx->device));
    ldx_store *s = new (std::nothrow) ldx_store();
    if (!s) return set_error(LDX_ERR_NOMEM, "store allocation failed");
    s->ctx = ctx; s->n_variants = n_variants; s->n_hap = n_hap;
    s->words = (n_hap + 63) / 64;
    s->stride_words = (s->words + 15) / 16 * 16;
    const size_t nv = (size_t)std::max<int64_t>(n_variants, 1);
    cudaError_t e = cudaMalloc(&s->d_planes, nv * s->stride_words * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_planes, 0, nv * s->stride_words * sizeof(uint64_t), ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_mask, s->stride_words * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_mask_user, s->stride_words * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_common, s->stride_words * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_all_slots, s->stride_words * sizeof(uint64_t));
    if (e == cudaSuccess) {
        // the common presence pattern of a fresh store: every one of the n_hap slots (complete diploid rows)
        s->h_common.assign((size_t)s->stride_words, 0);
        for (int h = 0; h < n_hap; ++h) s->h_common[(size_t)(h >> 6)] |= 1ull << (h & 63);
        e = cudaMemcpy(s->d_all_slots, s->h_common.data(), s->stride_words * sizeof(uint64_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_common, s->h_common.data(), s->stride_words * sizeof(uint64_t), cudaMemcpyHostToDevice);
    }
    // STORE_FREQ_PAD zeroed records past the last variant: the tcgen05 engine's direct mode reads whole 128-row tiles
    if (e == cudaSuccess) e = cudaMalloc(&s->d_freq, (nv + STORE_FREQ_PAD) * sizeof(VarFreq));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_freq, 0, (nv + STORE_FREQ_PAD) * sizeof(VarFreq), ctx->stream);
    if (e != cudaSuccess) { ldx_store_destroy(s); cudaGetLastError(); return set_error(LDX_ERR_NOMEM, std::string("store device allocation: ") + cudaGetErrorString(e)); }
    *store_out = s;
    return LDX_OK;
}

extern "C" int32_t ldx_store_destroy(ldx_store *s) {
    if (!s) return LDX_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    s->ctx->rows_cache_valid = false;          // a staged variant list validated against this store means nothing to the next one
    cudaFree(s->d_planes); cudaFree(s->d_mask); cudaFree(s->d_freq);
    cudaFree(s->d_mask_user); cudaFree(s->d_common); cudaFree(s->d_all_slots); cudaFree(s->d_kind); cudaFree(s->d_aux); cudaFree(s->d_gen);
    cudaFree(s->d_row_len); cudaFree(s->d_row_n1);
    cudaFree(s->d_pos0); cudaFree(s->d_end0); cudaFree(s->d_idnum); cudaFree(s->d_eligible);
    if (s->h_mask_stage) cudaFreeHost(s->h_mask_stage);
    if (s->mask_staged) cudaEventDestroy(s->mask_staged);
    delete s;
    return LDX_OK;
}

namespace ldx {
int store_alloc_annotations(ldx_store *s) {
    const size_t nv = (size_t)std::max<int64_t>(s->n_variants, 1);
    if (s->d_pos0) return LDX_OK;
    cudaError_t e = cudaMalloc(&s->d_pos0, nv * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_end0, nv * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_idnum, nv * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_eligible, nv);
    if (e != cudaSuccess) { cudaGetLastError(); return set_error(LDX_ERR_NOMEM, "store: annotation allocation failed"); }
    return LDX_OK;
}

// A store allocated for an upper bound of its rows keeps the first n_variants: smaller arrays when that frees a quarter or more.
template <typename T> static int shrink_array(T *&p, size_t old_n, size_t new_n, size_t unit, cudaStream_t st) {
    if (!p || new_n >= old_n) return LDX_OK;
    T *q = nullptr;
    if (cudaMalloc(&q, std::max<size_t>(new_n, 1) * unit) != cudaSuccess) { cudaGetLastError(); return LDX_OK; }   // keep the large one
    if (new_n && cudaMemcpyAsync(q, p, new_n * unit, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { cudaGetLastError(); cudaFree(q); return LDX_OK; }
    cudaStreamSynchronize(st);
    cudaFree(p);
    p = q;
    return LDX_OK;
}
int store_shrink(ldx_store *s, int64_t n_variants) {
    if (n_variants < 0 || n_variants > s->n_variants) return set_error(LDX_ERR_ARG, "store_shrink: bad row count");
    const size_t old_n = (size_t)s->n_variants, new_n = (size_t)n_variants;
    s->n_variants = n_variants;
    s->mask_set = false;
    s->tmap_ready = false;
    if (new_n * 4 > old_n * 3) return LDX_OK;
    cudaStream_t st = s->ctx->stream;
    shrink_array(s->d_planes, old_n, new_n, (size_t)s->stride_words * sizeof(uint64_t), st);
    shrink_array(s->d_pos0, old_n, new_n, 4, st); shrink_array(s->d_end0, old_n, new_n, 4, st);
    shrink_array(s->d_idnum, old_n, new_n, 8, st); shrink_array(s->d_eligible, old_n, new_n, 1, st);
    {   // frequency records: padded (STORE_FREQ_PAD zeroed entries past the last row)
        VarFreq *q = nullptr;
        if (cudaMalloc(&q, (std::max<size_t>(new_n, 1) + STORE_FREQ_PAD) * sizeof(VarFreq)) == cudaSuccess) {
            cudaMemsetAsync(q, 0, (std::max<size_t>(new_n, 1) + STORE_FREQ_PAD) * sizeof(VarFreq), st);
            cudaStreamSynchronize(st);
            cudaFree(s->d_freq);
            s->d_freq = q;
        } else cudaGetLastError();
    }
    if (s->d_kind) { cudaFree(s->d_kind); s->d_kind = nullptr; s->classify_dirty = true; }
    return LDX_OK;
}
}  // namespace ldx

extern "C" int32_t ldx_store_shape(const ldx_store *s, int64_t *n_variants, int32_t *n_hap, int32_t *stride_words) {
    LDX_REQUIRE(s, "store is NULL");
    if (n_variants) *n_variants = s->n_variants;
    if (n_hap) *n_hap = s->n_hap;
    if (stride_words) *stride_words = s->stride_words;
    return LDX_OK;
}

extern "C" int32_t ldx_store_planes_ptr(const ldx_store *s, void **dev_ptr_out) {
    LDX_REQUIRE(s && dev_ptr_out, "NULL argument");
    *dev_ptr_out = s->d_planes;
    return LDX_OK;
}

static int check_rows(const ldx_store *s, int64_t first_row, int64_t n_rows) {
    LDX_REQUIRE(s, "store is NULL");
    LDX_REQUIRE(first_row >= 0 && n_rows >= 0 && first_row + n_rows <= s->n_variants, "row range outside the store");
    return LDX_OK;
}

extern "C" int32_t ldx_store_pack_gt(ldx_store *s, int64_t first_row, int64_t n_rows, const uint8_t *text,
                                     int64_t text_bytes, const int64_t *row_off, int64_t row_pitch,
                                     int32_t n_samples, uint8_t *row_status) {
    LDX_TRY(check_rows(s, first_row, n_rows));
    LDX_REQUIRE(text && text_bytes > 0, "text is empty");
    LDX_REQUIRE(n_samples > 0 && 2 * n_samples == s->n_hap, "n_samples does not match the store (n_hap = 2 * n_samples)");
    // a row of plain diploid fields is 4 * n_samples - 1 bytes (the last separator need not exist); haploid fields make it shorter
    const int64_t row_bytes = 4ll * n_samples - 1, min_bytes = 2ll * n_samples - 1;
    if (row_off) {
        for (int64_t i = 0; i < n_rows; ++i)
            LDX_REQUIRE(row_off[i] >= 0 && row_off[i] + min_bytes <= text_bytes, "row_off outside text");
    } else {
        LDX_REQUIRE(row_pitch >= min_bytes && (n_rows == 0 || (n_rows - 1) * row_pitch + min_bytes <= text_bytes), "row_pitch/text_bytes mismatch");
    }
    if (n_rows == 0) return LDX_OK;
    ldx_ctx *ctx = s->ctx;
    LDX_CUDA(cudaSetDevice(ctx->device));
    uint8_t *d_text, *d_status; int64_t *d_off = nullptr;
    LDX_TRY(arena_get(ctx, S_TEXT, (size_t)text_bytes + (size_t)row_bytes + 64, (void **)&d_text));   // the fast kernel reads a whole plain row from every offset
    LDX_TRY(arena_get(ctx, S_STATUS, (size_t)n_rows, (void **)&d_status));
    LDX_CUDA(cudaMemcpyAsync(d_text, text, (size_t)text_bytes, cudaMemcpyHostToDevice, ctx->stream));
    LDX_CUDA(cudaMemsetAsync(d_text + text_bytes, '\n', (size_t)row_bytes + 64, ctx->stream));
    if (row_off) {
        LDX_TRY(arena_get(ctx, S_ROWOFF, sizeof(int64_t) * (size_t)n_rows, (void **)&d_off));
        LDX_CUDA(cudaMemcpyAsync(d_off, row_off, sizeof(int64_t) * (size_t)n_rows, cudaMemcpyHostToDevice, ctx->stream));
    }
    LDX_TRY(launch_pack_gt(ctx, d_text, d_off, row_pitch, n_rows, n_samples,
                           s->d_planes + first_row * s->stride_words, s->stride_words, d_status));
    std::vector<uint8_t> h_status((size_t)n_rows);
    LDX_CUDA(cudaMemcpyAsync(h_status.data(), d_status, (size_t)n_rows, cudaMemcpyDeviceToHost, ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));
    s->mask_set = false;   // counts are stale
    // rows with a field outside the plain alphabet: parsed again in full (haploid, missing, other codes, '/'), general route
    LDX_TRY(store_pack_general(s, first_row, n_rows, d_text, text_bytes, d_off, row_pitch, n_samples, h_status.data()));
    if (row_status) std::memcpy(row_status, h_status.data(), (size_t)n_rows);
    return LDX_OK;
}

extern "C" int32_t ldx_store_upload(ldx_store *s, int64_t first_row, int64_t n_rows, const uint64_t *planes) {
    LDX_TRY(check_rows(s, first_row, n_rows));
    LDX_REQUIRE(planes || n_rows == 0, "planes is NULL");
    if (n_rows == 0) return LDX_OK;
    LDX_CUDA(cudaSetDevice(s->ctx->device));
    LDX_CUDA(cudaMemcpyAsync(s->d_planes + first_row * s->stride_words, planes,
                             sizeof(uint64_t) * (size_t)n_rows * s->stride_words, cudaMemcpyHostToDevice, s->ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(s->ctx->stream));
    s->mask_set = false;
    return LDX_OK;
}

// The same without the wait: the copy is only enqueued (a pageable source is staged by the runtime before the call returns; a
// pinned one is read by the copy engine later, so it must stay unchanged until the next blocking call on this context).
extern "C" int32_t ldx_store_upload_async(ldx_store *s, int64_t first_row, int64_t n_rows, const uint64_t *planes) {
    LDX_TRY(check_rows(s, first_row, n_rows));
    LDX_REQUIRE(planes || n_rows == 0, "planes is NULL");
    if (n_rows == 0) return LDX_OK;
    LDX_CUDA(cudaSetDevice(s->ctx->device));
    LDX_CUDA(cudaMemcpyAsync(s->d_planes + first_row * s->stride_words, planes,
                             sizeof(uint64_t) * (size_t)n_rows * s->stride_words, cudaMemcpyHostToDevice, s->ctx->stream));
    s->mask_set = false;
    return LDX_OK;
}

extern "C" int32_t ldx_store_download(const ldx_store *s, int64_t first_row, int64_t n_rows, uint64_t *planes) {
    LDX_TRY(check_rows(s, first_row, n_rows));
    LDX_REQUIRE(planes || n_rows == 0, "planes is NULL");
    if (n_rows == 0) return LDX_OK;
    LDX_CUDA(cudaSetDevice(s->ctx->device));
    LDX_CUDA(cudaMemcpyAsync(planes, s->d_planes + first_row * s->stride_words,
                             sizeof(uint64_t) * (size_t)n_rows * s->stride_words, cudaMemcpyDeviceToHost, s->ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return LDX_OK;
}

extern "C" int32_t ldx_store_set_mask(ldx_store *s, const uint64_t *mask) {
    LDX_REQUIRE(s && mask, "NULL argument");
    LDX_CUDA(cudaSetDevice(s->ctx->device));
    if (s->classify_dirty) LDX_TRY(store_classify_rows(s));       // rows were packed since: the common pattern and every row's kind
    // the selection as given (general rows use it with their own presence planes), and folded with the store's common
    // presence pattern (all slots unless e.g. the males of chrX are haploid): the mask of the fast paths, N = its popcount
    // staged in the store's own pinned buffer: the call enqueues two small copies and the count kernel and returns
    const size_t sw = (size_t)s->stride_words;
    if (!s->h_mask_stage) {
        LDX_CUDA(cudaMallocHost((void **)&s->h_mask_stage, 2 * sw * sizeof(uint64_t)));
        LDX_CUDA(cudaEventCreateWithFlags(&s->mask_staged, cudaEventDisableTiming));
    } else LDX_CUDA(cudaEventSynchronize(s->mask_staged));         // the previous selection has left the buffer
    uint64_t *m = s->h_mask_stage, *mu = s->h_mask_stage + sw;
    std::memset(m, 0, 2 * sw * sizeof(uint64_t));
    int64_t n_sel = 0;
    for (int w = 0; w < s->words; ++w) {
        uint64_t x = mask[w];
        if (w == s->words - 1 && (s->n_hap & 63)) x &= (1ull << (s->n_hap & 63)) - 1;   // ignore pad bits
        mu[w] = x;
        m[w] = x & s->h_common[(size_t)w];
        n_sel += __builtin_popcountll(m[w]);
    }
    if (n_sel == 0) return set_error(LDX_ERR_EMPTY, "division by zero");   // empty sample selection, calc_ld.py:33
    s->n_sel = (int32_t)n_sel;
    LDX_TRY(make_final_ctx(n_sel, &s->fc));
    LDX_CUDA(cudaMemcpyAsync(s->d_mask, m, sizeof(uint64_t) * sw, cudaMemcpyHostToDevice, s->ctx->stream));
    LDX_CUDA(cudaMemcpyAsync(s->d_mask_user, mu, sizeof(uint64_t) * sw, cudaMemcpyHostToDevice, s->ctx->stream));
    LDX_CUDA(cudaEventRecord(s->mask_staged, s->ctx->stream));
    LDX_TRY(launch_variant_freq(s));
    s->mask_set = true;
    return LDX_OK;
}

static int require_mask(const ldx_store *s) {
    LDX_REQUIRE(s, "store is NULL");
    if (!s->mask_set) return set_error(LDX_ERR_STATE, "ldx_store_set_mask() must be called first (and again after loading rows)");
    return LDX_OK;
}

extern "C" int32_t ldx_store_counts(const ldx_store *s, int32_t *n1_out, int32_t *p_e4_out, int32_t *n_hap_sel_out) {
    LDX_TRY(require_mask(s));
    if (n_hap_sel_out) *n_hap_sel_out = s->n_sel;
    if ((n1_out || p_e4_out) && s->n_variants > 0) {
        LDX_CUDA(cudaSetDevice(s->ctx->device));
        std::vector<VarFreq> f((size_t)s->n_variants);
        LDX_CUDA(cudaMemcpyAsync(f.data(), s->d_freq, sizeof(VarFreq) * f.size(), cudaMemcpyDeviceToHost, s->ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(s->ctx->stream));
        for (int64_t v = 0; v < s->n_variants; ++v) {
            if (n1_out) n1_out[v] = f[v].n1;
            if (p_e4_out) p_e4_out[v] = f[v].p_e4;
        }
        if (n1_out && s->d_row_n1)            // general rows carry a code in VarFreq.n1: their true counts are kept by K2
            LDX_CUDA(cudaMemcpy(n1_out, s->d_row_n1, sizeof(int32_t) * (size_t)s->n_variants, cudaMemcpyDeviceToHost));
    }
    return LDX_OK;
}

extern "C" int32_t ldx_store_row_counts(const ldx_store *s, int32_t *n1_out, int32_t *len_out, int32_t *kind_out, int64_t *n_general_out) {
    LDX_TRY(require_mask(s));
    if (n_general_out) *n_general_out = s->n_nonsimple;
    if (s->n_variants == 0) return LDX_OK;
    LDX_CUDA(cudaSetDevice(s->ctx->device));
    const size_t n = (size_t)s->n_variants;
    if (kind_out) {
        if (s->d_kind) LDX_CUDA(cudaMemcpy(kind_out, s->d_kind, 4 * n, cudaMemcpyDeviceToHost));
        else for (size_t v = 0; v < n; ++v) kind_out[v] = -1;
    }
    if (len_out) {
        if (s->d_row_len) LDX_CUDA(cudaMemcpy(len_out, s->d_row_len, 4 * n, cudaMemcpyDeviceToHost));
        else for (size_t v = 0; v < n; ++v) len_out[v] = s->n_sel;
    }
    if (n1_out) LDX_TRY(ldx_store_counts(s, n1_out, nullptr, nullptr));
    return LDX_OK;
}

extern "C" int32_t ldx_store_subset(const ldx_store *src, const int32_t *sel, int32_t n_sel, ldx_store **store_out) {
    if (src) cudaSetDevice(src->ctx->device);        // a fresh host thread starts on device 0
    LDX_REQUIRE(src && sel && store_out, "NULL argument");
    LDX_REQUIRE(n_sel > 0, "empty selection");
    for (int32_t k = 0; k < n_sel; ++k) LDX_REQUIRE(sel[k] >= 0 && sel[k] < src->n_hap, "sel[] outside the source haplotypes");
    ldx_ctx *ctx = src->ctx;
    if (src->classify_dirty) LDX_TRY(store_classify_rows(const_cast<ldx_store *>(src)));      // the common pattern is about to be copied
    ldx_store *dst = nullptr;
    LDX_TRY(ldx_store_create(ctx, src->n_variants, n_sel, &dst));
    int32_t *d_sel;
    int rc = arena_get(ctx, S_MISC, sizeof(int32_t) * (size_t)n_sel, (void **)&d_sel);
    if (rc == LDX_OK && cudaMemcpyAsync(d_sel, sel, sizeof(int32_t) * (size_t)n_sel, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        rc = cuda_fail(cudaGetLastError(), "subset: copy sel");
    if (rc == LDX_OK) rc = launch_subset(src, d_sel, dst);
    if (rc == LDX_OK && !src->aux_rows.empty()) {
        // rows of the general route: their present / ref planes and the store's common pattern through the same column gather
        const int64_t n_aux = (int64_t)src->aux_rows.size();
        if (cudaMalloc(&dst->d_aux, (size_t)n_aux * 2 * dst->stride_words * sizeof(uint64_t)) != cudaSuccess) { cudaGetLastError(); rc = set_error(LDX_ERR_NOMEM, "subset: aux planes"); }
        if (rc == LDX_OK) {
            dst->aux_capacity = n_aux;
            dst->aux_rows = src->aux_rows;
            rc = launch_subset_planes(ctx, src->d_aux, src->stride_words, d_sel, n_sel, 2 * n_aux, dst->d_aux, dst->stride_words);
        }
        if (rc == LDX_OK) rc = launch_subset_planes(ctx, src->d_common, src->stride_words, d_sel, n_sel, 1, dst->d_common, dst->stride_words);
        dst->common_loaded = true;                      // the source's pattern restricted to the columns, not the subset's own middle row
        dst->classify_dirty = true;
    }
    if (rc == LDX_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "subset: sync");
    if (rc == LDX_OK && src->annotated) {
        const size_t nv = (size_t)src->n_variants;
        cudaError_t e = cudaMalloc(&dst->d_pos0, nv * 4);
        if (e == cudaSuccess) e = cudaMalloc(&dst->d_end0, nv * 4);
        if (e == cudaSuccess) e = cudaMalloc(&dst->d_idnum, nv * 8);
        if (e == cudaSuccess) e = cudaMalloc(&dst->d_eligible, nv);
        if (e == cudaSuccess) e = cudaMemcpy(dst->d_pos0, src->d_pos0, nv * 4, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dst->d_end0, src->d_end0, nv * 4, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dst->d_idnum, src->d_idnum, nv * 8, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dst->d_eligible, src->d_eligible, nv, cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) rc = cuda_fail(e, "subset: annotations");
        else dst->annotated = true;
    }
    if (rc == LDX_OK) {
        std::vector<uint64_t> ones(dst->stride_words, ~0ull);
        rc = ldx_store_set_mask(dst, ones.data());
    }
    if (rc != LDX_OK) { ldx_store_destroy(dst); return rc; }
    *store_out = dst;
    return LDX_OK;
}

extern "C" int32_t ldx_store_set_annotations(ldx_store *s, const int32_t *pos0, const int32_t *end0,
                                             const int64_t *idnum, const uint8_t *eligible) {
    LDX_REQUIRE(s && pos0 && end0 && idnum && eligible, "NULL argument");
    LDX_CUDA(cudaSetDevice(s->ctx->device));
    const size_t nv = (size_t)std::max<int64_t>(s->n_variants, 1);
    if (!s->d_pos0) {
        LDX_CUDA(cudaMalloc(&s->d_pos0, nv * 4));
        LDX_CUDA(cudaMalloc(&s->d_end0, nv * 4));
        LDX_CUDA(cudaMalloc(&s->d_idnum, nv * 8));
        LDX_CUDA(cudaMalloc(&s->d_eligible, nv));
    }
    const size_t n = (size_t)s->n_variants;
    cudaStream_t st = s->ctx->stream;
    LDX_CUDA(cudaMemcpyAsync(s->d_pos0, pos0, n * 4, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(s->d_end0, end0, n * 4, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(s->d_idnum, idnum, n * 8, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(s->d_eligible, eligible, n, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaStreamSynchronize(st));
    s->annotated = true;
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ store files
namespace {
struct StoreFileHeader {
    char magic[8];
    int64_t n_variants;
    int32_t n_hap, stride_words, annotated, reserved[9];
};
static_assert(sizeof(StoreFileHeader) == 64, "store file header");
constexpr size_t FILE_PIECE = 64u << 20;

struct Staging {            // the context's pinned bounce buffer + a FILE that is closed on every exit path
    void *buf = nullptr;
    FILE *fh = nullptr;
    ~Staging() { if (fh) fclose(fh); }
};
int staging_buffer(ldx_ctx *ctx, Staging &st) {
    if (!ctx->h_stage) LDX_CUDA(cudaMallocHost(&ctx->h_stage, FILE_PIECE));
    st.buf = ctx->h_stage;
    return LDX_OK;
}

int stream_out(ldx_ctx *ctx, Staging &st, const void *dev, size_t bytes) {
    for (size_t off = 0; off < bytes; off += FILE_PIECE) {
        const size_t n = std::min(FILE_PIECE, bytes - off);
        LDX_CUDA(cudaMemcpyAsync(st.buf, (const uint8_t *)dev + off, n, cudaMemcpyDeviceToHost, ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));
        if (fwrite(st.buf, 1, n, st.fh) != n) return ldx::set_error(LDX_ERR_STATE, "store file: write failed");
    }
    return LDX_OK;
}
int stream_in(ldx_ctx *ctx, Staging &st, void *dev, size_t bytes) {
    for (size_t off = 0; off < bytes; off += FILE_PIECE) {
        const size_t n = std::min(FILE_PIECE, bytes - off);
        if (fread(st.buf, 1, n, st.fh) != n) return ldx::set_error(LDX_ERR_ARG, "store file: truncated");
        LDX_CUDA(cudaMemcpyAsync((uint8_t *)dev + off, st.buf, n, cudaMemcpyHostToDevice, ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return LDX_OK;
}
}  // namespace

extern "C" int32_t ldx_store_save(const ldx_store *s, const char *path) {
    LDX_REQUIRE(s && path, "NULL argument");
    ldx_ctx *ctx = s->ctx;
    LDX_CUDA(cudaSetDevice(ctx->device));
    Staging st;
    LDX_TRY(staging_buffer(ctx, st));
    st.fh = fopen(path, "wb");
    if (!st.fh) return set_error(LDX_ERR_ARG, std::string("store file: cannot create ") + path);
    StoreFileHeader h = {};
    // rows with aux planes (the general route) make it a version-2 file: reserved[0] = their number; the section follows the
    // annotations as {row of every aux slot (int64), the slots' present / ref planes}.  kind[] and the common presence pattern
    // are re-derived when the loaded store's samples are selected, exactly as after an ingest.
    const int64_t n_aux = (int64_t)s->aux_rows.size();
    std::memcpy(h.magic, n_aux ? "LDXSTOR2" : "LDXSTOR1", 8);
    h.n_variants = s->n_variants; h.n_hap = s->n_hap; h.stride_words = s->stride_words; h.annotated = s->annotated ? 1 : 0;
    h.reserved[0] = (int32_t)n_aux;
    if (fwrite(&h, sizeof h, 1, st.fh) != 1) return set_error(LDX_ERR_STATE, "store file: write failed");
    const size_t n = (size_t)s->n_variants;
    LDX_TRY(stream_out(ctx, st, s->d_planes, n * s->stride_words * sizeof(uint64_t)));
    if (s->annotated) {
        LDX_TRY(stream_out(ctx, st, s->d_pos0, n * 4));
        LDX_TRY(stream_out(ctx, st, s->d_end0, n * 4));
        LDX_TRY(stream_out(ctx, st, s->d_idnum, n * 8));
        LDX_TRY(stream_out(ctx, st, s->d_eligible, n));
    }
    if (n_aux) {
        if (fwrite(s->aux_rows.data(), sizeof(int64_t), (size_t)n_aux, st.fh) != (size_t)n_aux) return set_error(LDX_ERR_STATE, "store file: write failed");
        LDX_TRY(stream_out(ctx, st, s->d_aux, (size_t)n_aux * 2 * s->stride_words * sizeof(uint64_t)));
    }
    if (fflush(st.fh) != 0) return set_error(LDX_ERR_STATE, "store file: write failed");
    return LDX_OK;
}

extern "C" int32_t ldx_store_load(ldx_ctx *ctx, const char *path, ldx_store **store_out) {
    LDX_REQUIRE(ctx && path && store_out, "NULL argument");
    *store_out = nullptr;
    LDX_CUDA(cudaSetDevice(ctx->device));
    Staging st;
    st.fh = fopen(path, "rb");
    if (!st.fh) return set_error(LDX_ERR_ARG, std::string("store file: cannot open ") + path);
    StoreFileHeader h;
    if (fread(&h, sizeof h, 1, st.fh) != 1 || (std::memcmp(h.magic, "LDXSTOR1", 8) != 0 && std::memcmp(h.magic, "LDXSTOR2", 8) != 0))
        return set_error(LDX_ERR_ARG, "store file: not an ldx store (bad magic)");
    const int64_t n_aux = h.magic[7] == '2' ? h.reserved[0] : 0;
    LDX_REQUIRE(n_aux >= 0 && n_aux <= h.n_variants, "store file: bad header");
    LDX_REQUIRE(h.n_variants >= 0 && h.n_variants < (1ll << 31) && h.n_hap > 0 && h.n_hap <= (1 << 24), "store file: bad header");
    LDX_TRY(staging_buffer(ctx, st));
    ldx_store *s = nullptr;
    LDX_TRY(ldx_store_create(ctx, h.n_variants, h.n_hap, &s));
    int rc = s->stride_words == h.stride_words ? (int)LDX_OK : set_error(LDX_ERR_ARG, "store file: row pitch of another library version");
    const size_t n = (size_t)s->n_variants;
    if (rc == LDX_OK) rc = stream_in(ctx, st, s->d_planes, n * s->stride_words * sizeof(uint64_t));
    if (rc == LDX_OK && h.annotated) {
        const size_t nv = std::max<size_t>(n, 1);
        cudaError_t e = cudaMalloc(&s->d_pos0, nv * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_end0, nv * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_idnum, nv * 8);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_eligible, nv);
        if (e != cudaSuccess) { cudaGetLastError(); rc = set_error(LDX_ERR_NOMEM, "store file: annotation allocation failed"); }
        if (rc == LDX_OK) rc = stream_in(ctx, st, s->d_pos0, n * 4);
        if (rc == LDX_OK) rc = stream_in(ctx, st, s->d_end0, n * 4);
        if (rc == LDX_OK) rc = stream_in(ctx, st, s->d_idnum, n * 8);
        if (rc == LDX_OK) rc = stream_in(ctx, st, s->d_eligible, n);
        if (rc == LDX_OK) s->annotated = true;
    }
    if (rc == LDX_OK && n_aux) {
        s->aux_rows.resize((size_t)n_aux);
        if (fread(s->aux_rows.data(), sizeof(int64_t), (size_t)n_aux, st.fh) != (size_t)n_aux) rc = set_error(LDX_ERR_ARG, "store file: truncated");
        for (int64_t k = 0; k < n_aux && rc == LDX_OK; ++k)
            if (s->aux_rows[(size_t)k] < -1 || s->aux_rows[(size_t)k] >= s->n_variants) rc = set_error(LDX_ERR_ARG, "store file: bad aux section");
        if (rc == LDX_OK && cudaMalloc(&s->d_aux, (size_t)n_aux * 2 * s->stride_words * sizeof(uint64_t)) != cudaSuccess) {
            cudaGetLastError(); rc = set_error(LDX_ERR_NOMEM, "store file: aux planes");
        }
        if (rc == LDX_OK) { s->aux_capacity = n_aux; rc = stream_in(ctx, st, s->d_aux, (size_t)n_aux * 2 * s->stride_words * sizeof(uint64_t)); }
        s->classify_dirty = true;
    }
    if (rc != LDX_OK) { ldx_store_destroy(s); return rc; }
    *store_out = s;
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ pairs
extern "C" int32_t ldx_pairs(ldx_store *s, const int64_t *ia, const int64_t *ib, int64_t n, int32_t *n11,
                             double *d, double *dprime, double *r2, uint32_t *packed) {
    LDX_TRY(require_mask(s));
    LDX_REQUIRE(n >= 0 && (n == 0 || (ia && ib)), "bad pair list");
    for (int64_t k = 0; k < n; ++k)
        LDX_REQUIRE(ia[k] >= 0 && ia[k] < s->n_variants && ib[k] >= 0 && ib[k] < s->n_variants, "pair index outside the store");
    if (n == 0) return LDX_OK;
    ldx_ctx *ctx = s->ctx;
    LDX_CUDA(cudaSetDevice(ctx->device));
    LDX_TRY(settle_before_host_call(ctx));
    if (n <= 16384) {
        // A short list (ld_lite asks for ONE pair): indices in and every output out through the context's pinned staging -- two copies
        // and one wait instead of up to seven pageable copies, each staged by the runtime.
        const size_t nn = (size_t)n, in_bytes = 16 * nn;
        const size_t o_n11 = in_bytes, o_d = o_n11 + ((4 * nn + 7) & ~(size_t)7), o_dp = o_d + 8 * nn, o_r2 = o_dp + 8 * nn, o_pk = o_r2 + 8 * nn, total = o_pk + 4 * nn;
        if (ctx->h_lists_bytes < total) {
            if (ctx->h_lists) cudaFreeHost(ctx->h_lists);
            ctx->h_lists = nullptr; ctx->h_lists_bytes = 0;
            const size_t cap = std::max<size_t>(total * 2, (size_t)1 << 16);
            LDX_CUDA(cudaMallocHost((void **)&ctx->h_lists, cap));
            ctx->h_lists_bytes = cap;
        }
        uint8_t *d_blk, *h = ctx->h_lists;
        LDX_TRY(arena_get(ctx, S_TEXT, total + 64, (void **)&d_blk));
        std::memcpy(h, ia, 8 * nn);
        std::memcpy(h + 8 * nn, ib, 8 * nn);
        LDX_CUDA(cudaMemcpyAsync(d_blk, h, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
        LDX_TRY(launch_pairs(s, reinterpret_cast<int64_t *>(d_blk), reinterpret_cast<int64_t *>(d_blk + 8 * nn), n,
                             n11 ? reinterpret_cast<int32_t *>(d_blk + o_n11) : nullptr, d ? reinterpret_cast<double *>(d_blk + o_d) : nullptr,
                             dprime ? reinterpret_cast<double *>(d_blk + o_dp) : nullptr, r2 ? reinterpret_cast<double *>(d_blk + o_r2) : nullptr,
                             packed ? reinterpret_cast<uint32_t *>(d_blk + o_pk) : nullptr));
        LDX_CUDA(cudaMemcpyAsync(h + in_bytes, d_blk + in_bytes, total - in_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        std::vector<FixupRec> recs;
        LDX_TRY(collect_fixups(ctx, recs));   // synchronises
        if (n11) std::memcpy(n11, h + o_n11, 4 * nn);
        if (d) std::memcpy(d, h + o_d, 8 * nn);
        if (dprime) std::memcpy(dprime, h + o_dp, 8 * nn);
        if (r2) std::memcpy(r2, h + o_r2, 8 * nn);
        if (packed) std::memcpy(packed, h + o_pk, 4 * nn);
        for (const FixupRec &r : recs) {
            double r2_exact;
            const uint32_t w = host_finalise_rec(r, s->fc.n_hap, &r2_exact);
            if (packed) packed[r.out_index] = w;
            if (r2) r2[r.out_index] = r2_exact;
        }
        return LDX_OK;
    }
    int64_t *d_ia, *d_ib; int32_t *d_n11 = nullptr; double *d_d = nullptr, *d_dp = nullptr, *d_r2 = nullptr; uint32_t *d_pk = nullptr;
    LDX_TRY(arena_get(ctx, S_IA, sizeof(int64_t) * (size_t)n, (void **)&d_ia));
    LDX_TRY(arena_get(ctx, S_IB, sizeof(int64_t) * (size_t)n, (void **)&d_ib));
    if (n11) LDX_TRY(arena_get(ctx, S_N11, sizeof(int32_t) * (size_t)n, (void **)&d_n11));
    if (d) LDX_TRY(arena_get(ctx, S_D, sizeof(double) * (size_t)n, (void **)&d_d));
    if (dprime) LDX_TRY(arena_get(ctx, S_DP, sizeof(double) * (size_t)n, (void **)&d_dp));
    if (r2) LDX_TRY(arena_get(ctx, S_R2, sizeof(double) * (size_t)n, (void **)&d_r2));
    if (packed) LDX_TRY(arena_get(ctx, S_PACKED, sizeof(uint32_t) * (size_t)n, (void **)&d_pk));
    cudaStream_t st = ctx->stream;
    LDX_CUDA(cudaMemcpyAsync(d_ia, ia, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_ib, ib, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    LDX_TRY(launch_pairs(s, d_ia, d_ib, n, d_n11, d_d, d_dp, d_r2, d_pk));
    if (n11) LDX_CUDA(cudaMemcpyAsync(n11, d_n11, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (d) LDX_CUDA(cudaMemcpyAsync(d, d_d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (dprime) LDX_CUDA(cudaMemcpyAsync(dprime, d_dp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (r2) LDX_CUDA(cudaMemcpyAsync(r2, d_r2, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (packed) LDX_CUDA(cudaMemcpyAsync(packed, d_pk, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    std::vector<FixupRec> recs;
    LDX_TRY(collect_fixups(ctx, recs));   // synchronises
    for (const FixupRec &r : recs) {
        double r2_exact;
        const uint32_t w = host_finalise_rec(r, s->fc.n_hap, &r2_exact);
        if (packed) packed[r.out_index] = w;
        if (r2) r2[r.out_index] = r2_exact;
    }
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ window
static int stage_window(ldx_store *s, const int64_t *q_row, const int64_t *lo, const int64_t *hi,
                        const int32_t *ws, const int32_t *we, int64_t nq, int64_t **d_arrays5,
                        int64_t *n_chunks_out, int64_t *n_candidates_out) {
    LDX_TRY(require_mask(s));
    if (!s->annotated) return set_error(LDX_ERR_STATE, "ldx_store_set_annotations() must be called before a window scan");
    LDX_REQUIRE(nq >= 0 && nq < (1ll << 31) && (nq == 0 || (q_row && lo && hi && ws && we)), "bad query arrays");
    std::vector<int64_t> prefix((size_t)nq + 1, 0);
    int64_t cand = 0;
    for (int64_t k = 0; k < nq; ++k) {
        LDX_REQUIRE(q_row[k] >= 0 && q_row[k] < s->n_variants, "q_row outside the store");
        LDX_REQUIRE(lo[k] >= 0 && lo[k] <= hi[k] && hi[k] <= s->n_variants, "candidate range outside the store");
        prefix[k + 1] = prefix[k] + (hi[k] - lo[k] + WINDOW_CHUNK - 1) / WINDOW_CHUNK;
        cand += hi[k] - lo[k];
    }
    *n_chunks_out = prefix[nq];
    *n_candidates_out = cand;
    if (nq == 0 || !d_arrays5) return LDX_OK;           // validation and counts only
    ldx_ctx *ctx = s->ctx;
    LDX_CUDA(cudaSetDevice(ctx->device));
    // one staging block: q_row | lo | hi | prefix (int64) | ws | we (int32)
    const size_t n64 = (size_t)nq * 3 + (size_t)nq + 1;
    const size_t bytes = n64 * 8 + (size_t)nq * 8;
    uint8_t *blk;
    LDX_TRY(arena_get(ctx, S_IA, bytes + 64, (void **)&blk));
    int64_t *d_q = (int64_t *)blk, *d_lo = d_q + nq, *d_hi = d_lo + nq, *d_pref = d_hi + nq;
    int32_t *d_ws = (int32_t *)(d_pref + nq + 1), *d_we = d_ws + nq;
    cudaStream_t st = ctx->stream;
    LDX_CUDA(cudaMemcpyAsync(d_q, q_row, 8 * (size_t)nq, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_lo, lo, 8 * (size_t)nq, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_hi, hi, 8 * (size_t)nq, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_pref, prefix.data(), 8 * ((size_t)nq + 1), cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_ws, ws, 4 * (size_t)nq, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaMemcpyAsync(d_we, we, 4 * (size_t)nq, cudaMemcpyHostToDevice, st));
    LDX_CUDA(cudaStreamSynchronize(st));   // prefix is a local
    d_arrays5[0] = d_q; d_arrays5[1] = d_lo; d_arrays5[2] = d_hi; d_arrays5[3] = d_pref;
    d_arrays5[4] = (int64_t *)d_ws; d_arrays5[5] = (int64_t *)d_we;
    return LDX_OK;
}

// Work list of the multi-query window kernel: queries sorted by (lo, hi); one record per 256-row block of the store that
// some query needs, with the (contiguous, because the bounds are monotone) run of queries whose candidate range meets it;
// the kernel cuts each run into groups of WINDOW_MQ.  O(blocks + queries) on the host.  Returns false when the bounds are not
// monotone in that order (then the single-query kernel is used).
static bool build_mq_blocks(const int64_t *lo, const int64_t *hi, int64_t nq, std::vector<int32_t> &order, std::vector<WindowMqBlock> &blocks,
                            int64_t *n_items_out) {
    order.clear(); blocks.clear();
    *n_items_out = 0;
    order.reserve((size_t)nq);
    bool sorted = true;
    for (int64_t k = 0; k < nq; ++k)
        if (hi[k] > lo[k]) {
            if (!order.empty() && (lo[k] < lo[order.back()] || (lo[k] == lo[order.back()] && hi[k] < hi[order.back()]))) sorted = false;
            order.push_back((int32_t)k);
        }
    if (!sorted)
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return lo[a] != lo[b] ? lo[a] < lo[b] : hi[a] < hi[b]; });
    const size_t n = order.size();
    for (size_t k = 1; k < n; ++k)
        if (hi[order[k]] < hi[order[k - 1]]) return false;
    if (n == 0) return true;
    const int64_t B = WINDOW_CHUNK;
    blocks.reserve((size_t)((hi[order[n - 1]] - lo[order[0]]) / B + 2));
    size_t a = 0, b = 0;
    int64_t n_items = 0;
    for (int64_t blk = lo[order[0]] / B; a < n;) {
        const int64_t begin = blk * B, end = begin + B;
        while (a < n && hi[order[a]] <= begin) ++a;                 // candidate ranges that end before this block
        if (a == n) break;
        if (lo[order[a]] >= end) { blk = lo[order[a]] / B; continue; }   // nobody needs this block: jump to the next needed one
        if (b < a) b = a;
        while (b < n && lo[order[b]] < end) ++b;                    // order[a .. b) meet the block
        blocks.push_back(WindowMqBlock{begin, n_items, (int32_t)a, (int32_t)b});
        n_items += (int64_t)((b - a + WINDOW_MQ - 1) / WINDOW_MQ);
        ++blk;
    }
    *n_items_out = n_items;
    return true;
}

extern "C" int32_t ldx_window_dev(ldx_store *s, const int64_t *q_row, const int64_t *lo, const int64_t *hi,
                                  const int32_t *win_start, const int32_t *win_end, int64_t nq,
                                  int32_t measure, int32_t thres_e4, ldx_hit *dev_hits, int64_t cap,
                                  int64_t *dev_n_hits) {
    if (s) cudaSetDevice(s->ctx->device);
    LDX_REQUIRE(measure == LDX_MEASURE_R2 || measure == LDX_MEASURE_DPRIME, "bad measure");
    LDX_REQUIRE(dev_hits && dev_n_hits && cap >= 0, "bad output arguments");
    LDX_REQUIRE((reinterpret_cast<uintptr_t>(dev_hits) & 15) == 0, "dev_hits must be 16-byte aligned");
    int64_t *arr[6] = {}; int64_t n_chunks = 0, cand = 0;
    LDX_TRY(stage_window(s, q_row, lo, hi, win_start, win_end, nq, nullptr, &n_chunks, &cand));       // validation and counts only
    ldx_ctx *ctx = s->ctx;
    // dev_n_hits: [0] hits, [1] pairs scanned
    LDX_CUDA(cudaMemsetAsync(dev_n_hits, 0, 2 * sizeof(int64_t), ctx->stream));
    if (nq == 0 || n_chunks == 0) return LDX_OK;
    // Several queries with overlapping windows (the usual ld_area job): the multi-query kernel loads every store row once
    // per WINDOW_MQ queries instead of once per query.  LDX_WINDOW_MQ=0 forces the single-query kernel.
    static const bool mq_off = getenv("LDX_WINDOW_MQ") && atoi(getenv("LDX_WINDOW_MQ")) == 0;
    std::vector<int32_t> order;
    std::vector<WindowMqBlock> blocks;
    int64_t n_items = 0;
    const bool use_mq = !mq_off && ctx->window_mq && nq >= 2 && window_mq_supported(s) && build_mq_blocks(lo, hi, nq, order, blocks, &n_items) &&
                        n_items < 0x7fffffffll && n_items * WINDOW_MQ < 3 * n_chunks;   // else: the windows barely overlap
    LDX_TRY(begin_dev_call(ctx));
    if (use_mq) {
        // The work lists -- block records, then one 48-byte record per query in sorted order with everything the host knows about
        // it -- are assembled in the context's pinned staging and travel as ONE copy; nothing waits for it (the event guards the
        // staging against the next call).  The kernels read nothing else about the queries.
        const size_t n_srt = order.size();
        const size_t blk_bytes = (blocks.size() * sizeof(WindowMqBlock) + 15) & ~(size_t)15, rec_bytes = n_srt * sizeof(WindowMqQueryX);
        const size_t total = blk_bytes + rec_bytes;
        if (ctx->h_win_bytes < total) {
            if (ctx->h_win) { cudaEventSynchronize(ctx->win_staged); cudaFreeHost(ctx->h_win); ctx->h_win = nullptr; ctx->h_win_bytes = 0; }
            const size_t want = total + total / 2 + 4096;
            LDX_CUDA(cudaMallocHost((void **)&ctx->h_win, want));
            ctx->h_win_bytes = want;
            if (!ctx->win_staged) LDX_CUDA(cudaEventCreateWithFlags(&ctx->win_staged, cudaEventDisableTiming));
        } else LDX_CUDA(cudaEventSynchronize(ctx->win_staged));           // the previous call's copy has left the staging
        std::memcpy(ctx->h_win, blocks.data(), blocks.size() * sizeof(WindowMqBlock));
        WindowMqQueryX *rec = reinterpret_cast<WindowMqQueryX *>(ctx->h_win + blk_bytes);
        for (size_t k = 0; k < n_srt; ++k) {
            const int32_t q = order[k];
            rec[k] = WindowMqQueryX{0, q, (int32_t)q_row[q], (int32_t)lo[q], (int32_t)hi[q], win_start[q], win_end[q], 0, {0, 0, 0}};
        }
        uint8_t *blk;
        LDX_TRY(arena_get(ctx, S_IB, total + 64, (void **)&blk));
        LDX_CUDA(cudaMemcpyAsync(blk, ctx->h_win, total, cudaMemcpyHostToDevice, ctx->stream));
        LDX_CUDA(cudaEventRecord(ctx->win_staged, ctx->stream));
        LDX_TRY(launch_window_mq(s, nq, blk, (int64_t)blocks.size(), blk + blk_bytes, (int64_t)n_srt, reinterpret_cast<unsigned int *>(blk + total),
                                 measure, thres_e4, dev_hits, cap, reinterpret_cast<unsigned long long *>(dev_n_hits)));
    } else {
        LDX_TRY(stage_window(s, q_row, lo, hi, win_start, win_end, nq, arr, &n_chunks, &cand));
        LDX_TRY(launch_window(s, arr[0], arr[1], arr[2], (const int32_t *)arr[4], (const int32_t *)arr[5], arr[3], nq,
                              n_chunks, measure, thres_e4, dev_hits, cap, reinterpret_cast<unsigned long long *>(dev_n_hits)));
    }
    ++ctx->seq;
    LDX_TRY(launch_publish(ctx));
    end_dev_call(ctx, 2, dev_hits, s->fc.n_hap, measure, 1, thres_e4);
    return LDX_OK;
}

__global__ void scatter_words_kernel(uint8_t *base, const uint64_t *__restrict__ byte_off, const uint32_t *__restrict__ words, size_t n);

// The kept pairs leave the scan kernels in no particular order; the reference emits rows in VCF order per query (ld_area.py:215).
// They are sorted by (query, row) ON THE DEVICE before they cross PCIe: keys (query << row_bits | row) with the hit's position as
// the value through cub's radix sort (only the bits the keys use), then one gather.  3 x 10^7 kept pairs (configs[2] at -z 0) take
// milliseconds instead of the seconds std::sort needs for them on one host core.
__global__ void hit_keys_kernel(const ldx_hit *__restrict__ hits, int64_t n, int row_bits, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 h = __ldg(reinterpret_cast<const uint4 *>(hits + i));          // {query, row, n11, packed}
    keys[i] = ((uint64_t)h.x << row_bits) | (uint64_t)h.y;
    idx[i] = (uint32_t)i;
}
__global__ void hit_gather_kernel(const ldx_hit *__restrict__ src, const uint32_t *__restrict__ idx, int64_t n, ldx_hit *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    *reinterpret_cast<uint4 *>(dst + i) = __ldg(reinterpret_cast<const uint4 *>(src + idx[i]));
}
static int sort_hits_on_device(ldx_ctx *ctx, const ldx_hit *d_hits, int64_t n, int64_t n_variants, int64_t nq, ldx_hit **d_sorted_out) {
    int row_bits = 1, q_bits = 1;
    while ((1ll << row_bits) < n_variants) ++row_bits;
    while ((1ll << q_bits) < nq) ++q_bits;
    uint64_t *d_keys; uint32_t *d_idx; ldx_hit *d_out; void *d_tmp;
    LDX_TRY(arena_get(ctx, S_D, 2 * sizeof(uint64_t) * (size_t)n, (void **)&d_keys));
    LDX_TRY(arena_get(ctx, S_DP, 2 * sizeof(uint32_t) * (size_t)n, (void **)&d_idx));
    LDX_TRY(arena_get(ctx, S_N11, sizeof(ldx_hit) * (size_t)n, (void **)&d_out));
    size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint64_t> kb(d_keys, d_keys + n);
    cub::DoubleBuffer<uint32_t> vb(d_idx, d_idx + n);
    LDX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, row_bits + q_bits, ctx->stream));
    LDX_TRY(arena_get(ctx, S_R2, tmp_bytes + 16, &d_tmp));
    const unsigned grid = (unsigned)((n + 255) / 256);
    hit_keys_kernel<<<grid, 256, 0, ctx->stream>>>(d_hits, n, row_bits, d_keys, d_idx);
    LDX_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, kb, vb, (int)n, 0, row_bits + q_bits, ctx->stream));
    hit_gather_kernel<<<grid, 256, 0, ctx->stream>>>(d_hits, vb.Current(), n, d_out);
    ctx->launches += 2;
    LDX_CUDA(cudaGetLastError());
    *d_sorted_out = d_out;
    return LDX_OK;
}

extern "C" int32_t ldx_window(ldx_store *s, const int64_t *q_row, const int64_t *lo, const int64_t *hi,
                              const int32_t *win_start, const int32_t *win_end, int64_t nq, int32_t measure,
                              int32_t thres_e4, ldx_hit *hits, int64_t cap, int64_t *n_hits, int64_t *n_scanned) {
    if (s) cudaSetDevice(s->ctx->device);
    LDX_REQUIRE(n_hits && cap >= 0 && (hits || cap == 0), "bad output arguments");
    LDX_REQUIRE(s, "store is NULL");
    ldx_ctx *ctx = s->ctx;
    *n_hits = 0;
    if (n_scanned) *n_scanned = 0;
    ldx_hit *d_hits; int64_t *d_cnt;
    LDX_TRY(arena_get(ctx, S_HITS, sizeof(ldx_hit) * (size_t)std::max<int64_t>(cap, 1), (void **)&d_hits));
    LDX_TRY(settle_before_host_call(ctx));
    LDX_TRY(arena_get(ctx, S_MISC, 2 * sizeof(int64_t), (void **)&d_cnt));
    LDX_TRY(ldx_window_dev(s, q_row, lo, hi, win_start, win_end, nq, measure, thres_e4, d_hits, cap, d_cnt));
    ctx->pending_q.clear();                // this call settles its own records below
    int64_t h_cnt[2] = {0, 0};
    LDX_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof h_cnt, cudaMemcpyDeviceToHost, ctx->stream));
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_scanned) *n_scanned = h_cnt[1];
    std::vector<FixupRec> recs;
    if (h_cnt[0] > cap) {
        collect_fixups(ctx, recs);
        *n_hits = h_cnt[0];
        return set_error(LDX_ERR_CAPACITY, "hit buffer too small");
    }
    const int64_t n = h_cnt[0];
    const bool device_sort = n >= 4096 && n < (1ll << 31);
    if (n && !device_sort) LDX_CUDA(cudaMemcpyAsync(hits, d_hits, sizeof(ldx_hit) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    LDX_TRY(collect_fixups(ctx, recs));   // synchronises
    if (device_sort) {
        // the settled words of the near-ties go to the device list first (a handful), then the list is sorted there
        if (!recs.empty()) {
            const size_t nr = recs.size();
            std::vector<uint64_t> stage(nr + (nr + 1) / 2);
            uint32_t *words = reinterpret_cast<uint32_t *>(stage.data() + nr);
            for (size_t i = 0; i < nr; ++i) {
                words[i] = settle_word(recs[i], s->fc.n_hap, measure, 1, thres_e4);
                stage[i] = reinterpret_cast<uint64_t>(d_hits) + (recs[i].out_index & FIX_INDEX_MASK) * sizeof(ldx_hit) + offsetof(ldx_hit, packed);
            }
            uint64_t *d_stage;
            LDX_TRY(arena_get(ctx, S_MISC, stage.size() * sizeof(uint64_t) + 64, (void **)&d_stage));
            LDX_CUDA(cudaMemcpyAsync(d_stage, stage.data(), stage.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
            scatter_words_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, ctx->stream>>>(nullptr, d_stage, reinterpret_cast<const uint32_t *>(d_stage + nr), nr);
            ctx->launches++;
            LDX_CUDA(cudaGetLastError());
            LDX_CUDA(cudaStreamSynchronize(ctx->stream));                 // `stage` is a local
        }
        ldx_hit *d_sorted = nullptr;
        LDX_TRY(sort_hits_on_device(ctx, d_hits, n, s->n_variants, nq, &d_sorted));
        LDX_CUDA(cudaMemcpyAsync(hits, d_sorted, sizeof(ldx_hit) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        LDX_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        for (const FixupRec &r : recs) {
            const uint32_t w = settle_word(r, s->fc.n_hap, measure, 1, thres_e4);
            hits[r.out_index & FIX_INDEX_MASK].packed = w;      // LDX_BELOW_THRES marks hits the exact rounding rejects
        }
    }
    int64_t kept = n;
    if (!recs.empty()) {                                        // (a stable compaction: the order survives)
        kept = 0;
        for (int64_t k = 0; k < n; ++k)
            if (!(hits[k].packed & LDX_BELOW_THRES)) hits[kept++] = hits[k];
    }
    if (!device_sort)
        std::sort(hits, hits + kept, [](const ldx_hit &a, const ldx_hit &b) {
            return a.query != b.query ? a.query < b.query : a.row < b.row;
        });
    *n_hits = kept;
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------ triangle
// The variant lists of a call (one per set), concatenated, on the device.  Repeated calls with the same lists for the same
// stores (the usual case) skip validation and the upload: the cache key records which store (and how many rows of it) every
// list was checked against.  contiguous[k] (may be NULL) = rows[k] is first, first + 1, ...
struct RowList { ldx_store *s; const int64_t *rows; int64_t v; };
static int stage_rows(ldx_ctx *ctx, const RowList *lists, int n, int64_t **d_rows_out, int64_t *offsets, bool *contiguous) {
    int64_t total = 0;
    for (int k = 0; k < n; ++k) {
        LDX_TRY(require_mask(lists[k].s));
        LDX_REQUIRE(lists[k].s->ctx == ctx, "store of another context");
        LDX_REQUIRE(lists[k].v >= 0 && lists[k].v < (1ll << 31) && (lists[k].v == 0 || lists[k].rows), "bad rows");
        offsets[k] = total;
        total += lists[k].v;
    }
    *d_rows_out = nullptr;
    if (total == 0) { for (int k = 0; k < n && contiguous; ++k) contiguous[k] = false; return LDX_OK; }
    LDX_CUDA(cudaSetDevice(ctx->device));
    Arena *a = arena_of(ctx);
    void *before = a->ptr[S_ROWS];
    LDX_TRY(arena_get(ctx, S_ROWS, sizeof(int64_t) * (size_t)total, (void **)d_rows_out));
    std::vector<int64_t> key;
    key.reserve(3 * (size_t)n);
    for (int k = 0; k < n; ++k) { key.push_back((int64_t)reinterpret_cast<intptr_t>(lists[k].s)); key.push_back(lists[k].s->n_variants); key.push_back(lists[k].v); }
    bool same = ctx->rows_cache_valid && before == a->ptr[S_ROWS] && (int64_t)ctx->rows_cache.size() == total && ctx->rows_cache_key == key;
    for (int k = 0; k < n && same; ++k)
        same = lists[k].v == 0 || std::memcmp(ctx->rows_cache.data() + offsets[k], lists[k].rows, sizeof(int64_t) * (size_t)lists[k].v) == 0;
    if (!same) {
        ctx->rows_cache_valid = false;
        for (int k = 0; k < n; ++k)
            for (int64_t i = 0; i < lists[k].v; ++i)
                LDX_REQUIRE(lists[k].rows[i] >= 0 && lists[k].rows[i] < lists[k].s->n_variants, "rows[] outside the store");
        ctx->rows_cache.resize((size_t)total);
        for (int k = 0; k < n; ++k)
            if (lists[k].v) std::memcpy(ctx->rows_cache.data() + offsets[k], lists[k].rows, sizeof(int64_t) * (size_t)lists[k].v);
        // the host copy stays alive (and unchanged) until the next staging, which the stream orders after this copy's consumers
        LDX_CUDA(cudaMemcpyAsync(*d_rows_out, ctx->rows_cache.data(), sizeof(int64_t) * (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
        ctx->rows_cache_key = key;
        ctx->rows_cache_contig.assign((size_t)n, 0);
        for (int k = 0; k < n; ++k) {
            bool c = lists[k].v > 0;
            for (int64_t i = 1; i < lists[k].v && c; ++i) c = lists[k].rows[i] == lists[k].rows[0] + i;
            ctx->rows_cache_contig[(size_t)k] = c ? 1 : 0;
        }
        ctx->rows_cache_valid = true;                 // only now: a failed copy must not leave a cache that claims an upload
    }
    for (int k = 0; k < n && contiguous; ++k) contiguous[k] = ctx->rows_cache_contig[(size_t)k] != 0;
    return LDX_OK;
}

extern "C" int32_t ldx_triangle_rows_dev(ldx_store *s, const int64_t *rows, int64_t v, int64_t row_begin, int64_t row_end,
                                         int32_t measure, int32_t has_thres, int32_t thres_e4, int32_t engine,
                                         uint32_t *dev_packed, int32_t *dev_n11) {
    if (s) cudaSetDevice(s->ctx->device);
    LDX_REQUIRE(measure == LDX_MEASURE_R2 || measure == LDX_MEASURE_DPRIME, "bad measure");
    LDX_REQUIRE(engine >= LDX_ENGINE_AUTO && engine <= LDX_ENGINE_MMA, "bad engine");
    LDX_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= v, "bad row range");
    LDX_REQUIRE(row_begin % 128 == 0, "row_begin must be a multiple of 128");
    LDX_REQUIRE(s, "store is NULL");
    ldx_ctx *ctx = s->ctx;
    int64_t *d_rows, off = 0;
    bool contiguous = false;
    const RowList list = {s, rows, v};
    LDX_TRY(stage_rows(ctx, &list, 1, &d_rows, &off, &contiguous));
    if (row_end < 2 || row_begin == row_end) return LDX_OK;
    if (engine == LDX_ENGINE_MMA && !triangle_mma_available())
        return set_error(LDX_ERR_ARG, "the tcgen05 engine is not available in this build");
    const bool use_mma = engine == LDX_ENGINE_MMA || (engine == LDX_ENGINE_AUTO && triangle_mma_available() && row_end >= ctx->mma_min_v &&
                                                      s->n_sel <= triangle_mma_max_haplotypes());
    // rows[] staging is consumed by the kernel on the same stream (stage_rows keeps a host copy alive)
    const uint32_t seq = ctx->seq + 1 ? ctx->seq + 1 : 1;     // 0 means "nothing published"
    LDX_TRY(begin_dev_call(ctx));
    int rc;
    if (use_mma) {
        const MmaSetDesc d = {s, 0, row_end, row_begin, dev_packed, dev_n11, ctx->fix_tag, rows[0], contiguous};
        rc = launch_triangle_mma(ctx, &d, 1, d_rows, measure, has_thres, thres_e4, seq);
    } else rc = launch_triangle_popc(s, d_rows, row_end, row_begin, measure, has_thres, thres_e4, dev_packed, dev_n11);
    LDX_TRY(rc);
    ctx->seq = seq;
    if (!use_mma) LDX_TRY(launch_publish(ctx));
    end_dev_call(ctx, 1, dev_packed, s->fc.n_hap, measure, has_thres, thres_e4);
    return LDX_OK;
}

// Several variant sets -- the reference's unit of work is one matrix per (source file, chromosome), ld_triangle.py:80-88 and
// :406-408 -- in ONE persistent launch of the tcgen05 engine: a 2,000-variant matrix is a single wave of tiles, bound by
// launch, pipeline-fill and drain latencies; a batch of them keeps every SM busy across sets.
extern "C" int32_t ldx_triangle_batch_dev(ldx_ctx *ctx, const ldx_triangle_set *sets, int32_t n_sets, int32_t measure,
                                          int32_t has_thres, int32_t thres_e4, int32_t engine) {
    if (ctx) cudaSetDevice(ctx->device);
    LDX_REQUIRE(ctx && (sets || n_sets == 0) && n_sets >= 0, "bad set list");
    LDX_REQUIRE(measure == LDX_MEASURE_R2 || measure == LDX_MEASURE_DPRIME, "bad measure");
    LDX_REQUIRE(engine >= LDX_ENGINE_AUTO && engine <= LDX_ENGINE_MMA, "bad engine");
    for (int32_t k = 0; k < n_sets; ++k) {
        LDX_REQUIRE(sets[k].store && sets[k].store->ctx == ctx, "every store of a batch must belong to the calling context");
        LDX_REQUIRE(sets[k].v >= 0 && (sets[k].dev_packed || sets[k].v < 2), "bad set");
    }
    for (int32_t first = 0; first < n_sets;) {
        // a launch takes the sets that follow with the same haplotype count and selection size, up to MMA_MAX_SETS
        const ldx_store *s0 = sets[first].store;
        int32_t n = 1;
        while (first + n < n_sets && n < MMA_MAX_SETS && sets[first + n].store->n_hap == s0->n_hap && sets[first + n].store->n_sel == s0->n_sel &&
               sets[first + n].store->mask_set && s0->mask_set) ++n;
        int64_t v_max = sets[first].v;
        for (int32_t k = 1; k < n; ++k) v_max = std::max(v_max, sets[first + k].v);
        const bool use_mma = n > 1 && s0->mask_set && (engine == LDX_ENGINE_MMA || (engine == LDX_ENGINE_AUTO && triangle_mma_available() && v_max >= 2 &&
                                                                                     s0->n_sel <= triangle_mma_max_haplotypes()));
        if (!use_mma) {          // a lone set or too many haplotypes: the single-set path chooses its engine itself
            for (int32_t k = 0; k < n; ++k)
                LDX_TRY(ldx_triangle_dev(sets[first + k].store, sets[first + k].rows, sets[first + k].v, measure, has_thres, thres_e4, engine,
                                         sets[first + k].dev_packed, sets[first + k].dev_n11));
            first += n;
            continue;
        }
        RowList lists[MMA_MAX_SETS];
        int64_t offs[MMA_MAX_SETS];
        for (int32_t k = 0; k < n; ++k) lists[k] = RowList{sets[first + k].store, sets[first + k].rows, sets[first + k].v};
        int64_t *d_rows;
        LDX_TRY(stage_rows(ctx, lists, n, &d_rows, offs, nullptr));
        MmaSetDesc desc[MMA_MAX_SETS];
        if (ctx->pending_q.size() + (size_t)n >= 4096) LDX_TRY(resolve_pending(ctx, nullptr));
        for (int32_t k = 0; k < n; ++k) {
            const ldx_triangle_set &t = sets[first + k];
            LDX_TRY(begin_dev_call(ctx));
            desc[k] = MmaSetDesc{t.store, offs[k], t.v, 0, t.dev_packed, t.dev_n11, ctx->fix_tag, 0, false};
            end_dev_call(ctx, 1, t.dev_packed, t.store->fc.n_hap, measure, has_thres, thres_e4);
        }
        const uint32_t seq = ctx->seq + 1 ? ctx->seq + 1 : 1;
        LDX_TRY(launch_triangle_mma(ctx, desc, n, d_rows, measure, has_thres, thres_e4, seq));
        ctx->seq = seq;
        first += n;
    }
    return LDX_OK;
}

extern "C" int32_t ldx_triangle_dev(ldx_store *s, const int64_t *rows, int64_t v, int32_t measure,
                                    int32_t has_thres, int32_t thres_e4, int32_t engine,
                                    uint32_t *dev_packed, int32_t *dev_n11) {
    return ldx_triangle_rows_dev(s, rows, v, 0, v, measure, has_thres, thres_e4, engine, dev_packed, dev_n11);
}

extern "C" int32_t ldx_triangle_rows(ldx_store *s, const int64_t *rows, int64_t v, int64_t row_begin, int64_t row_end,
                                     int32_t measure, int32_t has_thres, int32_t thres_e4, int32_t engine,
                                     uint32_t *packed, int32_t *n11) {
    if (s) cudaSetDevice(s->ctx->device);
    LDX_REQUIRE(s, "store is NULL");
    LDX_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= v, "bad row range");
    ldx_ctx *ctx = s->ctx;
    const int64_t n_pairs = (row_end > 1 ? row_end * (row_end - 1) / 2 : 0) - (row_begin > 1 ? row_begin * (row_begin - 1) / 2 : 0);
    uint32_t *d_pk = nullptr; int32_t *d_n11 = nullptr;
    LDX_TRY(settle_before_host_call(ctx));
    if (n_pairs) {
        LDX_TRY(arena_get(ctx, S_PACKED, sizeof(uint32_t) * (size_t)n_pairs, (void **)&d_pk));   // always: fix-ups index it
        if (n11) LDX_TRY(arena_get(ctx, S_N11, sizeof(int32_t) * (size_t)n_pairs, (void **)&d_n11));
    }
    LDX_TRY(ldx_triangle_rows_dev(s, rows, v, row_begin, row_end, measure, has_thres, thres_e4, engine, d_pk, d_n11));
    ctx->pending_q.clear();                // this call settles its own records below
    if (!n_pairs) return LDX_OK;
    if (packed) LDX_CUDA(cudaMemcpyAsync(packed, d_pk, sizeof(uint32_t) * (size_t)n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    if (n11) LDX_CUDA(cudaMemcpyAsync(n11, d_n11, sizeof(int32_t) * (size_t)n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<FixupRec> recs;
    LDX_TRY(collect_fixups(ctx, recs));   // synchronises
    if (packed)
        for (const FixupRec &r : recs) packed[r.out_index & FIX_INDEX_MASK] = settle_word(r, s->fc.n_hap, measure, has_thres, thres_e4);
    return LDX_OK;
}

extern "C" int32_t ldx_triangle(ldx_store *s, const int64_t *rows, int64_t v, int32_t measure,
                                int32_t has_thres, int32_t thres_e4, int32_t engine, uint32_t *packed,
                                int32_t *n11) {
    return ldx_triangle_rows(s, rows, v, 0, v, measure, has_thres, thres_e4, engine, packed, n11);
}

// ------------------------------------------------------------------------------------------ narrow outputs of the all-pairs call
// The drivers only ever print ONE measure of a matrix (ld_triangle.py:230 picks the r_square or the d_prime key); a caller that
// wants the numbers on the host pays PCIe for every byte, so: 2 bytes per pair for the measure asked for, or -- with a threshold
// (-z) -- only the pairs that pass it.
__global__ void narrow_values_kernel(const uint32_t *__restrict__ packed, uint16_t *__restrict__ out, int64_t n, int dprime) {
    // bits 0..13 the value, bit 14 LDX_V16_BELOW, bit 15 LDX_V16_INT0 (the D' half of the word already has this shape)
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i4 + 8 <= n && (reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(packed + i4)), b = __ldcs(reinterpret_cast<const uint4 *>(packed + i4 + 4));
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = dprime ? w[k] >> 16 : (w[k] & 0xbfffu) | ((w[k] >> 16) & 0x4000u);
        *reinterpret_cast<uint4 *>(out + i4) = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    } else {
        for (int64_t i = i4; i < n && i < i4 + 8; ++i) {
            const uint32_t w = packed[i];
            out[i] = (uint16_t)(dprime ? w >> 16 : (w & 0xbfffu) | ((w >> 16) & 0x4000u));
        }
    }
}
static inline uint16_t narrow_word(uint32_t w, int measure) {
    return (uint16_t)(measure == LDX_MEASURE_DPRIME ? w >> 16 : (w & 0xbfffu) | ((w >> 16) & 0x4000u));
}

extern "C" int32_t ldx_triangle_values(ldx_store *s, const int64_t *rows, int64_t v, int64_t row_begin, int64_t row_end, int32_t measure,
                                       int32_t has_thres, int32_t thres_e4, int32_t engine, uint16_t *values) {
    if (s) cudaSetDevice(s->ctx->device);
    LDX_REQUIRE(s && values, "NULL argument");
    LDX_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= v, "bad row range");
    ldx_ctx *ctx = s->ctx;
    const int64_t n_pairs = (row_end > 1 ? row_end * (row_end - 1) / 2 : 0) - (row_begin > 1 ? row_begin * (row_begin - 1) / 2 : 0);
    uint32_t *d_pk = nullptr; uint16_t *d_v16 = nullptr;
    LDX_TRY(settle_before_host_call(ctx));
    if (n_pairs) {
        LDX_TRY(arena_get(ctx, S_PACKED, sizeof(uint32_t) * (size_t)n_pairs, (void **)&d_pk));
        LDX_TRY(arena_get(ctx, S_N11, sizeof(uint16_t) * (size_t)n_pairs, (void **)&d_v16));
    }
    LDX_TRY(ldx_triangle_rows_dev(s, rows, v, row_begin, row_end, measure, has_thres, thres_e4, engine, d_pk, nullptr));
    ctx->pending_q.clear();                // this call settles its own records below
    if (!n_pairs) return LDX_OK;
    narrow_values_kernel<<<(unsigned)((n_pairs + 2047) / 2048), 256, 0, ctx->stream>>>(d_pk, d_v16, n_pairs, measure == LDX_MEASURE_DPRIME);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    LDX_CUDA(cudaMemcpyAsync(values, d_v16, sizeof(uint16_t) * (size_t)n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<FixupRec> recs;
    LDX_TRY(collect_fixups(ctx, recs));   // synchronises
    for (const FixupRec &r : recs) values[r.out_index & FIX_INDEX_MASK] = narrow_word(settle_word(r, s->fc.n_hap, measure, has_thres, thres_e4), measure);
    return LDX_OK;
}

// The pairs that pass the threshold, in matrix order: flags counted per chunk of 4,096 words, one scan over the chunk counts,
// then every chunk writes its survivors behind those of the chunks before it.
constexpr int HIT_CHUNK = 4096, HIT_PER_THREAD = HIT_CHUNK / 256;
__global__ void __launch_bounds__(256) count_passing_kernel(const uint32_t *__restrict__ packed, int64_t n, uint32_t *__restrict__ chunk_count) {
    const int64_t base = (int64_t)blockIdx.x * HIT_CHUNK;
    uint32_t c = 0;
    for (int k = 0; k < HIT_PER_THREAD; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < n) c += !(packed[i] & LDX_BELOW_THRES);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ uint32_t part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = part[0] + part[1] + part[2] + part[3] + part[4] + part[5] + part[6] + part[7];
}
__global__ void __launch_bounds__(1024) scan_chunks_kernel(const uint32_t *__restrict__ chunk_count, int64_t n_chunks, int64_t *__restrict__ chunk_first,
                                                            int64_t *__restrict__ total_out) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < n_chunks; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const int64_t x = i < n_chunks ? chunk_count[i] : 0;
        int64_t incl = x;
        for (int d = 1; d < 32; d <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += y; }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t w = warp_sum[threadIdx.x];
            for (int d = 1; d < 32; d <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= d) w += y; }
            warp_sum[threadIdx.x] = w;                  // inclusive over warps
        }
        __syncthreads();
        const int64_t before = carry + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + incl - x;
        if (i < n_chunks) chunk_first[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(256) write_passing_kernel(const uint32_t *__restrict__ packed, int64_t n, const int64_t *__restrict__ chunk_first,
                                                             ldx_pair_hit *__restrict__ hits, int64_t cap) {
    // thread t owns words base + t*16 .. +15, so that the survivors leave in matrix order
    const int64_t base = (int64_t)blockIdx.x * HIT_CHUNK + (int64_t)threadIdx.x * HIT_PER_THREAD;
    uint32_t w[HIT_PER_THREAD], flags = 0;
#pragma unroll
    for (int k = 0; k < HIT_PER_THREAD; ++k) {
        w[k] = base + k < n ? packed[base + k] : LDX_BELOW_THRES;
        flags |= (uint32_t)!(w[k] & LDX_BELOW_THRES) << k;
    }
    const uint32_t mine = __popc(flags);
    uint32_t incl = mine;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += y; }
    __shared__ uint32_t warp_sum[8];
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = incl - mine;
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) before += warp_sum[k];
    if (!flags) return;
    int64_t at = chunk_first[blockIdx.x] + before;
    // (row, col) of the first word: row = the largest r with r(r-1)/2 <= index
    int64_t row = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)base)) * 0.5);
    while (row * (row - 1) / 2 > base) --row;
    while ((row + 1) * row / 2 <= base) ++row;
    int64_t col = base - row * (row - 1) / 2;
#pragma unroll
    for (int k = 0; k < HIT_PER_THREAD; ++k) {
        if ((flags >> k) & 1u) {
            if (at < cap) hits[at] = ldx_pair_hit{(int32_t)row, (int32_t)col, w[k]};
            ++at;
        }
        if (++col == row) { ++row; col = 0; }
    }
}

extern "C" int32_t ldx_triangle_hits(ldx_store *s, const int64_t *rows, int64_t v, int32_t measure, int32_t thres_e4, int32_t engine,
                                     ldx_pair_hit *hits, int64_t cap, int64_t *n_hits) {
    if (s) cudaSetDevice(s->ctx->device);
    LDX_REQUIRE(s && n_hits && cap >= 0 && (hits || cap == 0), "bad argument");
    *n_hits = 0;
    ldx_ctx *ctx = s->ctx;
    const int64_t n_pairs = v > 1 ? v * (v - 1) / 2 : 0;
    LDX_TRY(settle_before_host_call(ctx));
    if (!n_pairs) { const RowList list = {s, rows, v}; int64_t *d_rows, off; return stage_rows(ctx, &list, 1, &d_rows, &off, nullptr); }
    const int64_t n_chunks = (n_pairs + HIT_CHUNK - 1) / HIT_CHUNK;
    LDX_REQUIRE(n_chunks < (1ll << 31), "too many pairs for one call");
    uint32_t *d_pk = nullptr, *d_count = nullptr; int64_t *d_first = nullptr; ldx_pair_hit *d_hits = nullptr;
    LDX_TRY(arena_get(ctx, S_PACKED, sizeof(uint32_t) * (size_t)n_pairs, (void **)&d_pk));
    LDX_TRY(arena_get(ctx, S_MISC, (size_t)n_chunks * 4 + 16, (void **)&d_count));
    LDX_TRY(arena_get(ctx, S_ROWOFF, (size_t)(n_chunks + 1) * 8, (void **)&d_first));       // [n_chunks] = the total
    LDX_TRY(arena_get(ctx, S_HITS, sizeof(ldx_pair_hit) * (size_t)std::max<int64_t>(cap, 1), (void **)&d_hits));
    LDX_TRY(ldx_triangle_rows_dev(s, rows, v, 0, v, measure, 1, thres_e4, engine, d_pk, nullptr));
    ctx->pending_q.clear();
    count_passing_kernel<<<(unsigned)n_chunks, 256, 0, ctx->stream>>>(d_pk, n_pairs, d_count);
    scan_chunks_kernel<<<1, 1024, 0, ctx->stream>>>(d_count, n_chunks, d_first, d_first + n_chunks);
    write_passing_kernel<<<(unsigned)n_chunks, 256, 0, ctx->stream>>>(d_pk, n_pairs, d_first, d_hits, cap);
    ctx->launches += 3;
    LDX_CUDA(cudaGetLastError());
    int64_t total = 0;
    LDX_CUDA(cudaMemcpyAsync(&total, d_first + n_chunks, 8, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<FixupRec> recs;
    LDX_TRY(collect_fixups(ctx, recs));   // synchronises
    const int64_t n_dev = std::min(total, cap);
    if (n_dev) LDX_CUDA(cudaMemcpy(hits, d_hits, sizeof(ldx_pair_hit) * (size_t)n_dev, cudaMemcpyDeviceToHost));
    // near-ties settled on the host may enter, leave or change the list
    auto index_of = [](const ldx_pair_hit &h) { return (int64_t)h.row * (h.row - 1) / 2 + h.col; };
    std::vector<std::pair<int64_t, uint32_t>> add;
    std::vector<int64_t> drop;
    for (const FixupRec &r : recs) {
        const int64_t idx = (int64_t)(r.out_index & FIX_INDEX_MASK);
        const uint32_t w = settle_word(r, s->fc.n_hap, measure, 1, thres_e4);
        ldx_pair_hit *e = std::lower_bound(hits, hits + n_dev, idx, [&](const ldx_pair_hit &h, int64_t i) { return index_of(h) < i; });
        const bool listed = e != hits + n_dev && index_of(*e) == idx;
        if (listed && !(w & LDX_BELOW_THRES)) e->packed = w;
        else if (listed) drop.push_back(idx);
        else if (!(w & LDX_BELOW_THRES)) add.emplace_back(idx, w);
    }
    int64_t n_out = total;
    if (total > cap) { *n_hits = total + (int64_t)add.size(); return set_error(LDX_ERR_CAPACITY, "hit buffer too small"); }
    if (!drop.empty() || !add.empty()) {
        std::vector<ldx_pair_hit> merged;
        merged.reserve((size_t)n_dev + add.size());
        std::sort(drop.begin(), drop.end());
        for (int64_t k = 0; k < n_dev; ++k) if (!std::binary_search(drop.begin(), drop.end(), index_of(hits[k]))) merged.push_back(hits[k]);
        for (const auto &a : add) {
            int64_t row = (int64_t)((1.0 + std::sqrt(1.0 + 8.0 * (double)a.first)) * 0.5);
            while (row * (row - 1) / 2 > a.first) --row;
            while ((row + 1) * row / 2 <= a.first) ++row;
            merged.push_back(ldx_pair_hit{(int32_t)row, (int32_t)(a.first - row * (row - 1) / 2), a.second});
        }
        std::sort(merged.begin(), merged.end(), [&](const ldx_pair_hit &a, const ldx_pair_hit &b) { return index_of(a) < index_of(b); });
        n_out = (int64_t)merged.size();
        if (n_out > cap) { *n_hits = n_out; return set_error(LDX_ERR_CAPACITY, "hit buffer too small"); }
        std::copy(merged.begin(), merged.end(), hits);
    }
    *n_hits = n_out;
    return LDX_OK;
}

// The all-pairs call and the table writer in one: the words of matrix rows row_begin..row_end-1 go to the ctx's scratch,
// are settled there, and only their text crosses PCIe.
extern "C" int32_t ldx_triangle_table(ldx_store *s, const int64_t *rows, int64_t v, int64_t row_begin, int64_t row_end,
                                      int32_t measure, int32_t has_thres, int32_t thres_e4, int32_t engine,
                                      const char *prefixes, const int64_t *prefix_off, int32_t flags,
                                      char *text, int64_t cap, int64_t *n_bytes) {
    if (s) cudaSetDevice(s->ctx->device);
    LDX_REQUIRE(s && n_bytes, "NULL argument");
    *n_bytes = 0;
    LDX_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= v, "bad row range");
    LDX_REQUIRE((flags & ~LDX_TEXT_OUT_ON_DEVICE) == 0, "bad flags");
    ldx_ctx *ctx = s->ctx;
    const int64_t n_pairs = (row_end > 1 ? row_end * (row_end - 1) / 2 : 0) - (row_begin > 1 ? row_begin * (row_begin - 1) / 2 : 0);
    uint32_t *d_pk = nullptr;
    LDX_TRY(settle_before_host_call(ctx));
    LDX_TRY(arena_get(ctx, S_PACKED, sizeof(uint32_t) * (size_t)std::max<int64_t>(n_pairs, 1), (void **)&d_pk));
    LDX_TRY(ldx_triangle_rows_dev(s, rows, v, row_begin, row_end, measure, has_thres, thres_e4, engine, d_pk, nullptr));
    // ldx_triangle_text resolves the pending call (near-ties settled into d_pk) before it reads the words
    return ldx_triangle_text(ctx, d_pk, v, row_begin, row_end, measure, prefixes, prefix_off, flags | LDX_TEXT_PACKED_ON_DEVICE,
                             text, cap, n_bytes);
}

// ------------------------------------------------------------------------------------------ resolve
__global__ void scatter_words_kernel(uint8_t *base, const uint64_t *__restrict__ byte_off, const uint32_t *__restrict__ words, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) *reinterpret_cast<uint32_t *>(base + byte_off[i]) = words[i];
}

static int resolve_pending(ldx_ctx *ctx, int64_t *n_fixed_out) {
    if (n_fixed_out) *n_fixed_out = 0;
    LDX_CUDA(cudaSetDevice(ctx->device));
    std::vector<FixupRec> recs;
    LDX_TRY(collect_fixups(ctx, recs, true));
    std::vector<ldx_ctx::Pending> q;
    q.swap(ctx->pending_q);
    ctx->fix_tag = 0;
    if (recs.empty() || q.empty()) return LDX_OK;
    // settle on the host (libm pow), then ONE upload and one scatter kernel: a 100,000-variant triangle
    // has ~10^4 near-ties, far too many for a copy + synchronise each
    const size_t n = recs.size();
    std::vector<uint64_t> stage(n + (n + 1) / 2);                 // [n] device addresses | [n] words
    uint32_t *words = reinterpret_cast<uint32_t *>(stage.data() + n);
    for (size_t i = 0; i < n; ++i) {
        const FixupRec &r = recs[i];
        const size_t call = (size_t)(r.out_index >> FIX_TAG_SHIFT);
        if (call == 0 || call > q.size()) return set_error(LDX_ERR_STATE, "fix-up record without a pending call");
        const ldx_ctx::Pending &p = q[call - 1];
        const uint64_t idx = r.out_index & FIX_INDEX_MASK;
        words[i] = settle_word(r, p.n_hap, p.measure, p.has_thres, p.thres_e4);
        stage[i] = reinterpret_cast<uint64_t>(p.dev_out) + (p.kind == 1 ? idx * sizeof(uint32_t) : idx * sizeof(ldx_hit) + offsetof(ldx_hit, packed));
    }
    uint64_t *d_stage;
    LDX_TRY(arena_get(ctx, S_MISC, stage.size() * sizeof(uint64_t), (void **)&d_stage));
    LDX_CUDA(cudaMemcpyAsync(d_stage, stage.data(), stage.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    scatter_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(nullptr, d_stage, reinterpret_cast<const uint32_t *>(d_stage + n), n);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    LDX_CUDA(cudaStreamSynchronize(ctx->stream));                 // `stage` is a local
    if (n_fixed_out) *n_fixed_out = (int64_t)recs.size();
    return LDX_OK;
}

extern "C" int32_t ldx_resolve(ldx_ctx *ctx, int64_t *n_fixed_out) {
    LDX_REQUIRE(ctx, "ctx is NULL");
    return resolve_pending(ctx, n_fixed_out);
}
