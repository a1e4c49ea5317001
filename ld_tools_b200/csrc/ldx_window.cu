// ldx_window.cu -- K4: the ld_area window scan (replaces the loop at ld_area.py:215-249).
//
// For every query variant q and every candidate store row j in [lo_q, hi_q):
//     n11 = popcount(mask & plane[q] & plane[j])                      calc_ld.py:30-32
//     D, D', r2, round(., 4)                                           calc_ld.py:33-97
// fused with the reference's filters: window overlap (pysam fetch semantics, ld_area.py:215-217),
// id class and same-id skip (ld_area.py:222-225) and the ROUNDED-value threshold (ld_area.py:248).
// Survivors are appended to a compact hit list; nothing else is written.
//
// Roofline: HBM.  Each (q, j) pair must read row j once: stride_words*8 bytes (640 B at 5008
// haplotypes, 632 B of them payload).  The query plane lives in registers, the mask is folded
// into it, so the per-pair traffic is exactly one store row.
//
// Work decomposition: a work item is WINDOW_CHUNK (256) consecutive candidate rows of one query.
// Items are numbered through a prefix sum over queries (chunk_prefix); persistent CTAs take items
// grid-stride.  Inside a CTA 8 lanes share a row: lane l reads 16-byte granules l, l+8, ... so
// every group load is one full 128-byte line; eight rows per group are reduced with a 7-shuffle
// transpose so that afterwards thread t owns the count of row base+t and all 256 threads run the
// fp64 finalisation in parallel.
#include "ldx_internal.h"
#include "ldx_fixup.cuh"

namespace ldx {

constexpr int WIN_THREADS = WINDOW_CHUNK;   // one thread per row in the epilogue

struct WindowArgs {
    const uint4 *planes; const uint4 *mask; int32_t stride_u4;
    const VarFreq *freq; FinalCtx fc;
    const int32_t *pos0, *end0; const int64_t *idnum; const uint8_t *eligible;
    const int64_t *q_row, *lo, *hi; const int32_t *win_start, *win_end;
    const int64_t *chunk_prefix; int64_t nq, n_chunks;
    int measure, thres_e4;
    int32_t n_sel;                 // N under the mask
    float screen_t, screen_g;      // single-precision screen (see screen_below): threshold - 1/2 - slack, and the
                                   // reference chain's own error bound; screen_t <= 0 switches the screen off
    ldx_hit *hits; int64_t cap; unsigned long long *counters;   // [0] hits, [1] pairs scanned
    FixupSink fix;
};

// Single-precision screen of the rounded-threshold test (ld_area.py:248).  Almost every candidate of a
// window is far below an LD threshold like r2 >= 0.8; for those the fp64 chain of calc_ld.py:33-97 need not
// run at all.  x = value * 10^4 is evaluated from the exact integers Dn = n11*N - n1a*n1b, ... with a
// relative error below 12 * 2^-24 (same arithmetic and bound as the all-pairs epilogue,
// ldx_triangle_mma.cu), and the reference's own chain is within g of the exact ratio: when
// x + error < T - 1/2 the reference's round(value, 4) * 10^4 is at most T - 1 and the row is dropped.
// A monomorphic pair is int 0 in the reference (calc_ld.py:68-69, :89-90): below any positive threshold.
// Everything else -- including every kept row -- takes the exact path.  Needs N <= 8192 (int32 products).
__device__ __forceinline__ bool screen_below(int32_t n11, int32_t N, int32_t n1a, int32_t n1b, int measure, float t_minus, float g) {
    const int32_t n0a = N - n1a, n0b = N - n1b, P = n1a * n1b, Dn = n11 * N - P;
    const float fD = fabsf(__int2float_rn(Dn));
    float x, err;
    if (measure == LDX_MEASURE_R2) {
        const int32_t da = n1a * n0a, db = n1b * n0b;
        if (da == 0 || db == 0) return true;
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(__int2float_rn(da), __int2float_rn(db))));
        x = __fmul_rn(__fmul_rn(__fmul_rn(fD, 1.0e4f), fD), r);
        err = __fmaf_rn(x, 12.0f * 5.9604644775390625e-08f, g);
    } else {
        const int32_t m = Dn > 0 ? min(n1a * n0b, n0a * n1b) : min(P, n0a * n0b);
        if (m == 0) return true;
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__int2float_rn(m)));
        x = __fmul_rn(__fmul_rn(fD, 1.0e4f), r);
        err = __fmaf_rn(x, 8.0f * 5.9604644775390625e-08f, g);
    }
    return __fadd_rn(x, err) < t_minus;      // false for NaN: falls to the exact path
}

template <int NG>   // NG = 16-byte granules per lane per row; 0 = runtime loop
__global__ void __launch_bounds__(WIN_THREADS, 3)
window_kernel(const WindowArgs A) {
    __shared__ int64_t s_q, s_base;
    const int tid = threadIdx.x, lane8 = tid & 7, group = tid >> 3;
    const int ng = NG ? NG : A.stride_u4 / 8;
    unsigned long long scanned = 0;

    for (int64_t c = blockIdx.x; c < A.n_chunks; c += gridDim.x) {
        if (tid == 0) {   // which query owns work item c: last q with chunk_prefix[q] <= c
            int64_t a = 0, b = A.nq;
            while (b - a > 1) { const int64_t m = (a + b) >> 1; if (A.chunk_prefix[m] <= c) a = m; else b = m; }
            s_q = a;
            s_base = A.lo[a] + (c - A.chunk_prefix[a]) * WINDOW_CHUNK;
        }
        __syncthreads();
        const int64_t q = s_q, base = s_base;
        const int64_t hi = A.hi[q];
        const int64_t qrow = A.q_row[q];
        __syncthreads();   // s_q / s_base may be overwritten by the next iteration

        // ---- counting: group g covers rows base + 8g + i, i = 0..7
        int cnt[8];
        const uint4 *qp = A.planes + qrow * A.stride_u4;
        if (NG) {
            uint4 qm[NG ? NG : 1];
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                const uint4 x = ldg_u4(qp + j * 8 + lane8), m = __ldg(A.mask + j * 8 + lane8);
                qm[j] = make_uint4(x.x & m.x, x.y & m.y, x.z & m.z, x.w & m.w);
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const int64_t r0 = base + group * 8 + i, r1 = r0 + 1;
                const bool v0 = r0 < hi, v1 = r1 < hi;
                const uint4 *p0 = A.planes + (v0 ? r0 : qrow) * A.stride_u4 + lane8;
                const uint4 *p1 = A.planes + (v1 ? r1 : qrow) * A.stride_u4 + lane8;
                uint4 x0[NG ? NG : 1], x1[NG ? NG : 1];
#pragma unroll
                for (int j = 0; j < NG; ++j) { x0[j] = ldg_u4_stream(p0 + j * 8); x1[j] = ldg_u4_stream(p1 + j * 8); }
                int c0 = 0, c1 = 0;
#pragma unroll
                for (int j = 0; j < NG; ++j) { c0 += popc_and_u4(x0[j], qm[j]); c1 += popc_and_u4(x1[j], qm[j]); }
                cnt[i] = c0; cnt[i + 1] = c1;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t r = base + group * 8 + i;
                const uint4 *p = A.planes + (r < hi ? r : qrow) * A.stride_u4 + lane8;
                int cc = 0;
                for (int j = 0; j < ng; ++j) {
                    const uint4 x = ldg_u4_stream(p + j * 8), y = ldg_u4(qp + j * 8 + lane8), m = __ldg(A.mask + j * 8 + lane8);
                    cc += __popc(x.x & y.x & m.x) + __popc(x.y & y.y & m.y) + __popc(x.z & y.z & m.z) + __popc(x.w & y.w & m.w);
                }
                cnt[i] = cc;
            }
        }
        // ---- transpose-reduce over the 8 lanes: lane l ends with the total of row i = l
        {
            const bool up = lane8 & 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int send = up ? cnt[i] : cnt[i + 4];
                const int keep = up ? cnt[i + 4] : cnt[i];
                cnt[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        {
            const bool up = lane8 & 2;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int send = up ? cnt[i] : cnt[i + 2];
                const int keep = up ? cnt[i + 2] : cnt[i];
                cnt[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
        }
        int n11;
        {
            const bool up = lane8 & 1;
            const int send = up ? cnt[0] : cnt[1];
            const int keep = up ? cnt[1] : cnt[0];
            n11 = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }

        // ---- epilogue: thread t owns row base + t
        const int64_t row = base + tid;
        bool pass = false;
        uint32_t packed = 0;
        if (row < hi) {
            const int32_t ws = A.win_start[q], we = A.win_end[q];
            const bool scan = A.pos0[row] < we && A.end0[row] > ws      // fetch overlap, ld_area.py:215-217
                              && A.eligible[row]                           // rs\d+$ and not MULTI_ALLELIC, :223-224
                              && A.idnum[row] != A.idnum[qrow];            // :222
            if (scan) ++scanned;
            if (scan && !(A.screen_t > 0.0f && screen_below(n11, A.n_sel, A.freq[qrow].n1, A.freq[row].n1, A.measure, A.screen_t, A.screen_g))) {
                const VarFreq fa = A.freq[qrow], fb = A.freq[row];         // var_1 = query, var_2 = row (:242)
                const PairFinal f = finalise_pair(n11, fa, fb, A.fc);
                packed = f.packed;
                int32_t m = measure_e4(packed, A.measure);
                if (A.measure == LDX_MEASURE_R2 && (packed & LDX_R2_NEARTIE)) ++m;   // keep; host settles the tie
                pass = m >= A.thres_e4;                                    // rounded value, :248
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, pass);
        if (ballot) {
            const int lane = tid & 31;
            unsigned long long slot0 = 0;
            if (lane == __ffs(ballot) - 1) slot0 = atomicAdd(A.counters, (unsigned long long)__popc(ballot));
            slot0 = __shfl_sync(0xffffffffu, slot0, __ffs(ballot) - 1);
            if (pass) {
                const unsigned long long slot = slot0 + __popc(ballot & ((1u << lane) - 1));
                if ((int64_t)slot < A.cap) {
                    // ldx_hit = {query, row, n11, packed}: one 16-byte store
                    *reinterpret_cast<uint4 *>(A.hits + slot) =
                        make_uint4((uint32_t)q, (uint32_t)row, (uint32_t)n11, packed);
                    if (packed & LDX_R2_NEARTIE)
                        fixup_append(A.fix, slot, n11, A.freq[qrow].n1, A.freq[row].n1, packed);
                }
            }
        }
    }
    // pairs scanned (for the bench's pairs/s figure): one atomic per warp
    for (int o = 16; o; o >>= 1) scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
    if ((tid & 31) == 0 && scanned) atomicAdd(A.counters + 1, scanned);
}

// Persistent grid: resident CTAs per SM (from the occupancy calculator) x SM count, so that every
// CTA is co-resident and the grid-stride loop balances.
template <int NG>
static int launch_window_ng(ldx_ctx *ctx, const WindowArgs &A) {
    static int per_sm = 0;
    if (!per_sm) {
        LDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_kernel<NG>, WIN_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
    }
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > A.n_chunks) grid = A.n_chunks;
    timing_begin(ctx);
    window_kernel<NG><<<(int)grid, WIN_THREADS, 0, ctx->stream>>>(A);
    timing_end(ctx);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "window_kernel");
    return LDX_OK;
}

int launch_window(ldx_store *s, const int64_t *d_qrow, const int64_t *d_lo, const int64_t *d_hi,
                  const int32_t *d_ws, const int32_t *d_we, const int64_t *d_chunk_prefix, int64_t nq,
                  int64_t n_chunks, int measure, int thres_e4, ldx_hit *d_hits, int64_t cap,
                  unsigned long long *d_counters) {
    if (n_chunks <= 0) return LDX_OK;
    ldx_ctx *ctx = s->ctx;
    WindowArgs A;
    A.planes = reinterpret_cast<const uint4 *>(s->d_planes);
    A.mask = reinterpret_cast<const uint4 *>(s->d_mask);
    A.stride_u4 = s->stride_words / 2;
    A.freq = s->d_freq; A.fc = s->fc;
    A.pos0 = s->d_pos0; A.end0 = s->d_end0; A.idnum = s->d_idnum; A.eligible = s->d_eligible;
    A.q_row = d_qrow; A.lo = d_lo; A.hi = d_hi; A.win_start = d_ws; A.win_end = d_we;
    A.chunk_prefix = d_chunk_prefix; A.nq = nq; A.n_chunks = n_chunks;
    A.measure = measure; A.thres_e4 = thres_e4;
    A.n_sel = s->n_sel;
    {
        const double n = (double)s->n_sel, g = 1.0e4 * 16.0 * 1.1102230246251565e-16 * n * n;
        A.screen_g = (float)(g + 1.0e-4);
        A.screen_t = (s->n_sel <= 8192 && thres_e4 > 0) ? (float)((double)thres_e4 - 0.5 - 1.0e-3) : 0.0f;
    }
    A.hits = d_hits; A.cap = cap; A.counters = d_counters;
    A.fix = FixupSink{ctx->d_fix, ctx->d_fix_count, ctx->fix_capacity, ctx->fix_tag};
    switch (A.stride_u4 / 8) {
        case 1: return launch_window_ng<1>(ctx, A);
        case 2: return launch_window_ng<2>(ctx, A);
        case 5: return launch_window_ng<5>(ctx, A);    // 5008 haplotypes
        default: return launch_window_ng<0>(ctx, A);
    }
}

}  // namespace ldx
