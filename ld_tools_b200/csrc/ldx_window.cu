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
    const GenStore *gen;           // the general route (variants with missing calls / haploid samples / other codes), or nullptr
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

// The per-row tail of both window kernels: thread `tid` owns store row `row` of query q (row >= the query's lo by
// construction; rows at or beyond `hi` are not candidates).  Filters, single-precision screen, exact finalisation, rounded
// threshold, and the warp-aggregated append of the kept pair.  Every thread of the warp must call it.
// The tail of a candidate pair that passed the filters (`scan`): general route or screen + exact finalisation, the rounded
// threshold, and the warp-aggregated append.  n1q / n1r: VarFreq.n1 of the query and of the row.  Every thread of the warp calls it.
// The r2 screen with the variants' halves of the denominator prepared beforehand: qa = 10^4 / (n1a * n0a) comes with the query's
// record, rb = 1 / (n1b * n0b) is formed once per block by the thread that owns the row (0 for a monomorphic variant: x = 0, the
// pair is dropped, as calc_ld's int 0 is below any positive threshold).  x = (Dn * Dn) * (qa * rb): two IMAD, one conversion,
// three multiplications -- no reciprocal and no second conversion per pair.  Relative error: two rcp.approx (2^-23 each), the
// roundings of 10^4 * rcp, qa * rb, Dn * Dn, the final product and of Dn itself (twice) = 10 * 2^-24, inside the same 12 * 2^-24
// bound as screen_below.
__device__ __forceinline__ bool screen_below_r2_pre(int32_t n11, int32_t N, int32_t n1a, int32_t n1b, float qa, float rb, float t_minus, float g) {
    const float fD = __int2float_rn(n11 * N - n1a * n1b);
    const float x = __fmul_rn(__fmul_rn(fD, fD), __fmul_rn(qa, rb));
    return __fadd_rn(x, __fmaf_rn(x, 12.0f * 5.9604644775390625e-08f, g)) < t_minus;      // false for NaN: falls to the exact path
}
__device__ __forceinline__ float half_denominator_rcp(int32_t n1, int32_t N, float scale) {
    const int32_t d = n1 * (N - n1);
    if (n1 < 0 || d <= 0) return 0.0f;              // a general-route code or a monomorphic variant
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__int2float_rn(d)));
    return __fmul_rn(scale, r);
}

template <bool PRE = false, typename Counter>
__device__ __forceinline__ void pair_tail(const WindowArgs &A, int64_t q, int64_t qrow, int64_t row, bool scan, int n11_in, int32_t n1q, int32_t n1r,
                                          int tid, Counter &scanned, float qa = 0.0f, float rb = 0.0f) {
    int n11 = n11_in;
    bool pass = false;
    uint32_t packed = 0;
    GenCounts gc;
    gc.n_pair = 0;                                                      // != 0: the pair took the general route
    if (scan) {
        ++scanned;
        if (A.gen && (n1q | n1r) < 0) {                                 // a variant of the general route: rare, one thread does it all
            gc = general_pair_counts(*A.gen, qrow, n1q, row, n1r);      // var_1 = query, var_2 = row (:242)
            packed = finalise_general(gc).packed;
            n11 = gc.n11;
            int32_t m = measure_e4(packed, A.measure);
            if (A.measure == LDX_MEASURE_R2 && (packed & LDX_R2_NEARTIE)) ++m;
            pass = m >= A.thres_e4;
        } else if (!(A.screen_t > 0.0f && (PRE && A.measure == LDX_MEASURE_R2 ? screen_below_r2_pre(n11, A.n_sel, n1q, n1r, qa, rb, A.screen_t, A.screen_g)
                                                                               : screen_below(n11, A.n_sel, n1q, n1r, A.measure, A.screen_t, A.screen_g)))) {
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
                if (packed & LDX_R2_NEARTIE) {
                    if (gc.n_pair != 0) fixup_append_general(A.fix, slot, gc, packed);
                    else fixup_append(A.fix, slot, n11, n1q, n1r, packed);
                }
            }
        }
    }
}

__device__ __forceinline__ void row_epilogue(const WindowArgs &A, int64_t q, int64_t qrow, int64_t row, int64_t hi, int n11, int tid,
                                             unsigned long long &scanned) {
    bool scan = false;
    int32_t n1q = 0, n1r = 0;
    if (row < hi) {
        const int32_t ws = A.win_start[q], we = A.win_end[q];
        scan = A.pos0[row] < we && A.end0[row] > ws                    // fetch overlap, ld_area.py:215-217
               && A.eligible[row]                                      // rs\d+$ and not MULTI_ALLELIC, :223-224
               && A.idnum[row] != A.idnum[qrow];                       // :222
        if (scan) { n1q = A.freq[qrow].n1; n1r = A.freq[row].n1; }
    }
    pair_tail(A, q, qrow, row, scan, n11, n1q, n1r, tid, scanned);
}

// 8 x 8 transpose-reduce over the lanes of a group: in, cnt[i] = this lane's partial count of row i; out, the total of row
// i = lane8.
__device__ __forceinline__ int transpose_reduce8(int (&cnt)[8], int lane8) {
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
    const bool up = lane8 & 1;
    const int send = up ? cnt[0] : cnt[1];
    const int keep = up ? cnt[1] : cnt[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// ------------------------------------------------------------------------------------------ carry-save popcount
// popcount(x & q) over a lane's 20 words (5008 haplotypes: five 16-byte granules per lane) with 6 POPC instead of 20.  POPC
// issues on the XU pipe at 16 lanes/clk/SM, LOP3 on the ALU pipe at 64: a carry-save adder (sum = a^b^c, carry = maj(a,b,c),
// two LOP3) turns three words of weight w into one of weight w and one of weight 2w, and a tree of 14 of them leaves
// 2 + 1 + 2 + 1 words of weight 1, 2, 4, 8 (Harley-Seal).  48 LOP3 + 6 POPC: bound by the ALU pipe at ~26 clk per (row, query)
// and lane group instead of 40 on the XU pipe.
__device__ __forceinline__ void csa(uint32_t &h, uint32_t &l, uint32_t a, uint32_t b, uint32_t c) {
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(l) : "r"(a), "r"(b), "r"(c));      // a ^ b ^ c
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(h) : "r"(a), "r"(b), "r"(c));      // majority
}
__device__ __forceinline__ int popc_and_20(const uint4 (&x)[5], const uint4 *__restrict__ q, int lane8) {
    uint32_t w[20];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const uint4 m = q[j * 8 + lane8];
        w[4 * j] = x[j].x & m.x; w[4 * j + 1] = x[j].y & m.y; w[4 * j + 2] = x[j].z & m.z; w[4 * j + 3] = x[j].w & m.w;
    }
    uint32_t o[8], t[9], f[4], e, hh, ll;
    // weight 1: 20 words -> 2
#pragma unroll
    for (int k = 0; k < 6; ++k) csa(t[k], o[k], w[3 * k], w[3 * k + 1], w[3 * k + 2]);
    o[6] = w[18]; o[7] = w[19];
    csa(t[6], o[0], o[0], o[1], o[2]);
    csa(t[7], o[1], o[3], o[4], o[5]);
    csa(t[8], o[0], o[0], o[1], o[6]);              // weight-1 words left: o[0], o[7]
    // weight 2: 9 words -> 1
    csa(f[0], t[0], t[0], t[1], t[2]);
    csa(f[1], t[1], t[3], t[4], t[5]);
    csa(f[2], t[2], t[6], t[7], t[8]);
    csa(f[3], t[0], t[0], t[1], t[2]);              // weight-2 word left: t[0]
    // weight 4: 4 words -> 2, weight 8: 1
    csa(e, f[0], f[0], f[1], f[2]);                 // weight-4 words left: f[0], f[3]
    (void)hh; (void)ll;
    return __popc(o[0]) + __popc(o[7]) + 2 * __popc(t[0]) + 4 * (__popc(f[0]) + __popc(f[3])) + 8 * __popc(e);
}

// ------------------------------------------------------------------------------------------ multi-query kernel
// The single-query kernel above reads every candidate row once PER QUERY: 9 TB/s of L2 traffic at configs[2], three
// quarters of what the L2 slices deliver, with the POPC pipe 57% busy.  Windows of neighbouring queries overlap almost
// completely (+/-500 kb windows of queries ~35 kb apart), so this kernel turns the loop inside out: a work item is a
// 256-row block of the store and up to MQ queries whose windows cover it; the block's rows are loaded once into registers
// and counted against all MQ query planes, which sit (mask folded in) in shared memory.  L2 traffic per pair drops by MQ
// and the POPC pipe becomes the limit.  Needs the queries sorted with monotone window bounds (the host checks).
constexpr int MQ = 4;
// One record per 256-row block that some query needs: rows base .. base+255, met by the queries at sorted positions
// a .. b-1 (MqQuery records); the kernel walks them in groups of MQ.
struct MqBlock { int64_t base, first_item; int32_t a, b; };
struct MqQuery { int64_t q, qrow, lo, hi; };               // query index of the call, its store row, its candidate range

// One 48-byte record per query with everything the per-pair filters need (filled by mq_extend_kernel before the scan), so that the
// scan kernels' inner loops read shared memory and registers only.
struct MqQueryX { int64_t idnum; int32_t q, qrow, lo, hi, ws, we, n1, pad[3]; };
static_assert(sizeof(MqQueryX) == WINDOW_MQ_EXT_BYTES, "three 16-byte loads");

// The host has filled q, qrow, lo, hi, ws, we; what lives on the device is added here.
__global__ void mq_extend_kernel(const WindowArgs A, int64_t n, MqQueryX *__restrict__ recs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t qrow = recs[i].qrow, n1 = A.freq[qrow].n1;
    recs[i].idnum = A.idnum[qrow];
    recs[i].n1 = n1;
    recs[i].pad[0] = __float_as_int(half_denominator_rcp(n1, A.n_sel, 1.0e4f));      // the query's half of the r2 screen (screen_below_r2_pre)
}

template <int NG, bool L1ROWS>
__global__ void __launch_bounds__(WIN_THREADS, 2)
window_mq_kernel(const WindowArgs A, const MqBlock *__restrict__ blocks, int64_t n_blocks, const MqQueryX *__restrict__ sorted,
                 unsigned int *__restrict__ next_block, int64_t n_rows) {
    // double-buffered per group of queries: their planes (mask folded in) and their records
    __shared__ uint4 qs[2][MQ][NG * 8];
    __shared__ uint4 s_rec[2][MQ][3];                        // MqQueryX records
    __shared__ unsigned int s_j;
    constexpr int PL = MQ * NG * 8;                         // threads that fetch one 16-byte granule of one query plane each
    static_assert(PL <= WIN_THREADS, "plane loaders");
    const int tid = threadIdx.x, lane8 = tid & 7, group = tid >> 3;
    const int k_pl = tid / (NG * 8), g_pl = tid % (NG * 8);
    unsigned long long scanned = 0;
    for (;;) {
        // blocks cost one pass per group of queries that meets them (1 .. ~10): handed out dynamically
        if (tid == 0) s_j = atomicAdd(next_block, 1u);
        __syncthreads();
        const unsigned int j = s_j;
        if (j >= n_blocks) break;
        const MqBlock blk = blocks[j];
        const int n_groups = (blk.b - blk.a + MQ - 1) / MQ;
        // ---- the row this thread owns in every epilogue of the block, and its filter columns: once per block
        const int64_t row = blk.base + tid;
        const bool in_store = row < n_rows;
        const int64_t rc = in_store ? row : n_rows - 1;
        const int32_t pos0 = A.pos0[rc], end0 = A.end0[rc], n1r = A.freq[rc].n1, row32 = (int32_t)row;
        const int64_t idn = A.idnum[rc];
        const bool elig = in_store && A.eligible[rc];
        const float rb = half_denominator_rcp(n1r, A.n_sel, 1.0f);
        // ---- the block's first group, synchronously; the store row of the second group's plane granule for later
        if (tid < MQ * 3) s_rec[0][tid / 3][tid % 3] = __ldg(reinterpret_cast<const uint4 *>(sorted + min(blk.a + tid / 3, blk.b - 1)) + tid % 3);
        int64_t qrow_next = 0;
        if (tid < PL) {
            const int64_t qrow0 = sorted[min(blk.a + k_pl, blk.b - 1)].qrow;
            const uint4 x = ldg_u4(A.planes + qrow0 * A.stride_u4 + g_pl), m = __ldg(A.mask + g_pl);
            qs[0][k_pl][g_pl] = make_uint4(x.x & m.x, x.y & m.y, x.z & m.z, x.w & m.w);
            qrow_next = sorted[min(blk.a + MQ + k_pl, blk.b - 1)].qrow;
        }
        __syncthreads();
        for (int g = 0; g < n_groups; ++g) {
            const int cur = g & 1, q0 = blk.a + g * MQ, nq = min(MQ, blk.b - q0);
            // ---- the NEXT group's plane granule and records, requested now and parked in shared memory after this
            //      group's work: single independent loads (the host pre-sorted the records), so nobody stalls on them
            uint4 nx = make_uint4(0, 0, 0, 0), mk = nx, rec_next = nx;
            const bool more = g + 1 < n_groups;
            int64_t qrow_next2 = 0;
            if (more && tid < PL) {
                nx = ldg_u4(A.planes + qrow_next * A.stride_u4 + g_pl); mk = __ldg(A.mask + g_pl);
                qrow_next2 = sorted[min(q0 + 2 * MQ + k_pl, blk.b - 1)].qrow;
            }
            if (more && tid < MQ * 3) rec_next = __ldg(reinterpret_cast<const uint4 *>(sorted + min(q0 + MQ + tid / 3, blk.b - 1)) + tid % 3);
            // ---- counting: group of lanes covers rows base + 8 * group + i, each loaded once for all queries of this pass
            int cnt[MQ][8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const int64_t r0 = blk.base + group * 8 + i, r1 = r0 + 1;
                const uint4 *p0 = A.planes + (r0 < n_rows ? r0 : n_rows - 1) * A.stride_u4 + lane8;
                const uint4 *p1 = A.planes + (r1 < n_rows ? r1 : n_rows - 1) * A.stride_u4 + lane8;
                uint4 x0[NG], x1[NG];
#pragma unroll
                for (int jj = 0; jj < NG; ++jj) {
                    x0[jj] = L1ROWS ? ldg_u4(p0 + jj * 8) : ldg_u4_stream(p0 + jj * 8);
                    x1[jj] = L1ROWS ? ldg_u4(p1 + jj * 8) : ldg_u4_stream(p1 + jj * 8);
                }
#pragma unroll
                for (int k = 0; k < MQ; ++k) {
                    int c0 = 0, c1 = 0;
                    if (k < nq) {                           // block-uniform
                        if (NG == 5) {
                            c0 = popc_and_20(reinterpret_cast<const uint4 (&)[5]>(x0), &qs[cur][k][0], lane8);
                            c1 = popc_and_20(reinterpret_cast<const uint4 (&)[5]>(x1), &qs[cur][k][0], lane8);
                        } else {
#pragma unroll
                            for (int jj = 0; jj < NG; ++jj) {
                                const uint4 qm = qs[cur][k][jj * 8 + lane8];
                                c0 += popc_and_u4(x0[jj], qm); c1 += popc_and_u4(x1[jj], qm);
                            }
                        }
                    }
                    cnt[k][i] = c0; cnt[k][i + 1] = c1;
                }
            }
            // ---- per query: transpose-reduce, then thread t owns row base + t
#pragma unroll
            for (int k = 0; k < MQ; ++k) {
                if (k < nq) {                               // block-uniform
                    const int n11 = transpose_reduce8(cnt[k], lane8);
                    const uint4 r0 = s_rec[cur][k][0], r1 = s_rec[cur][k][1];           // {idnum lo, hi, q, qrow} {lo, hi, ws, we}
                    const uint4 r2 = s_rec[cur][k][2];                                  // {n1, qa, -, -}
                    const int32_t n1q = (int32_t)r2.x;
                    const int64_t idq = (int64_t)(((unsigned long long)r0.y << 32) | r0.x);
                    const bool scan = row32 >= (int32_t)r1.x && row32 < (int32_t)r1.y       // the query's candidate range
                                      && pos0 < (int32_t)r1.w && end0 > (int32_t)r1.z       // fetch overlap, ld_area.py:215-217
                                      && elig && idn != idq;                                // :223-224, :222
                    pair_tail<true>(A, (int64_t)(int32_t)r0.z, (int64_t)(int32_t)r0.w, row, scan, n11, n1q, n1r, tid, scanned, __uint_as_float(r2.y), rb);
                }
            }
            if (more) {
                if (tid < PL) qs[cur ^ 1][k_pl][g_pl] = make_uint4(nx.x & mk.x, nx.y & mk.y, nx.z & mk.z, nx.w & mk.w);
                if (tid < MQ * 3) s_rec[cur ^ 1][tid / 3][tid % 3] = rec_next;
            }
            qrow_next = qrow_next2;
            __syncthreads();
        }
    }
    for (int o = 16; o; o >>= 1) scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
    if ((tid & 31) == 0 && scanned) atomicAdd(A.counters + 1, scanned);
}

// ------------------------------------------------------------------------------------------ 128-byte rows: one thread per row
// A subset store (ldx_store_subset: e.g. the 1006 EUR haplotypes of configs[2]) has rows of 16 words.  With 8 lanes per row
// the kernel above spends more instructions on the 8 x 8 transposes, the per-(row, query) filter loads and the re-loading of
// the block's rows for every group of queries than on the 4 AND + POPC a lane has to do.  Here a thread OWNS a row for the
// whole block: its 128 bytes and its filter columns (pos0, end0, eligible, idnum, n1) are loaded into registers once, every
// query that meets the block is then one pass over registers against the query's plane in shared memory (broadcast reads), and
// the count needs no reduction across lanes.  The popcount is carry-save per 8 words (4 POPC + 8 LOP3 instead of 8 POPC: POPC
// issues at a quarter of the LOP3 rate), which balances the two pipes.  Per-query columns come in one 48-byte record each
// (MqQueryX, filled by a small kernel before), so nothing in the inner loop waits on global memory.
constexpr int RQ = 8;                                      // queries per group (their planes double-buffered in shared memory)
// popcount(x & q) over 8 words with 4 POPC: three carry-save adders leave two words of weight 1 and three of weight 2, a fourth
// turns those into one of weight 2 and one of weight 4
__device__ __forceinline__ int popc_and_8(const uint4 &xa, const uint4 &xb, const uint4 &qa, const uint4 &qb) {
    uint32_t h0, l0, h1, l1, h2, l2, f, t;
    csa(h0, l0, xa.x & qa.x, xa.y & qa.y, xa.z & qa.z);
    csa(h1, l1, xa.w & qa.w, xb.x & qb.x, xb.y & qb.y);
    csa(h2, l2, l0, l1, xb.z & qb.z);
    csa(f, t, h0, h1, h2);
    return __popc(l2) + __popc(xb.w & qb.w) + 2 * __popc(t) + 4 * __popc(f);
}

template <int NG>                                            // NG = 128-byte units per row (1: up to 1024 haplotypes, 2: up to 2048)
__global__ void __launch_bounds__(WIN_THREADS, 2)
window_rows1_kernel(const WindowArgs A, const MqBlock *__restrict__ blocks, int64_t n_blocks, const MqQueryX *__restrict__ ext,
                    unsigned int *__restrict__ next_block, int64_t n_rows) {
    constexpr int GR = 8 * NG;                               // 16-byte granules per row
    __shared__ uint4 qs[2][RQ][GR];
    __shared__ uint4 s_rec[2][RQ][3];                        // MqQueryX records
    __shared__ unsigned int s_j;
    const int tid = threadIdx.x, k_pl = tid / GR, g_pl = tid % GR;     // loader threads (tid < RQ * GR): granule g_pl of query k_pl
    constexpr int PL = RQ * GR;
    static_assert(PL <= WIN_THREADS, "plane loaders");
    unsigned long long scanned_total = 0;
    for (;;) {
        if (tid == 0) s_j = atomicAdd(next_block, 1u);
        __syncthreads();
        const unsigned int j = s_j;
        if (j >= n_blocks) break;
        const MqBlock blk = blocks[j];
        const int n_groups = (blk.b - blk.a + RQ - 1) / RQ;
        // ---- this thread's row and its filter columns, once per block
        const int64_t row = blk.base + tid;
        const bool in_store = row < n_rows;
        const int64_t rc = in_store ? row : n_rows - 1;
        uint4 x[GR];
#pragma unroll
        for (int i = 0; i < GR; ++i) x[i] = ldg_u4(A.planes + rc * GR + i);
        const int32_t pos0 = A.pos0[rc], end0 = A.end0[rc], n1r = A.freq[rc].n1, row32 = (int32_t)row;
        const int64_t idn = A.idnum[rc];
        const bool elig = in_store && A.eligible[rc];
        const float rb = half_denominator_rcp(n1r, A.n_sel, 1.0f);
        // ---- the block's first group of queries, synchronously; the store row of the second group's plane granule for later
        int32_t qrow_next = 0;
        if (tid < PL) {
            const int32_t qrow0 = ext[min(blk.a + k_pl, blk.b - 1)].qrow;
            const uint4 y = ldg_u4(A.planes + (int64_t)qrow0 * GR + g_pl), m = __ldg(A.mask + g_pl);
            qs[0][k_pl][g_pl] = make_uint4(y.x & m.x, y.y & m.y, y.z & m.z, y.w & m.w);
            qrow_next = ext[min(blk.a + RQ + k_pl, blk.b - 1)].qrow;
        }
        if (tid < RQ * 3) s_rec[0][tid / 3][tid % 3] = __ldg(reinterpret_cast<const uint4 *>(ext + min(blk.a + tid / 3, blk.b - 1)) + tid % 3);
        __syncthreads();
        unsigned int scanned = 0;
        for (int g = 0; g < n_groups; ++g) {
            const int cur = g & 1, q0 = blk.a + g * RQ, nq = min(RQ, blk.b - q0);
            // ---- the NEXT group's plane granules and records, requested now and parked in shared memory after this group's work
            const bool more = g + 1 < n_groups;
            uint4 nx = make_uint4(0, 0, 0, 0), mk = nx, rec_next = nx;
            int32_t qrow_next2 = 0;
            if (more && tid < PL) {
                nx = ldg_u4(A.planes + (int64_t)qrow_next * GR + g_pl); mk = __ldg(A.mask + g_pl);
                qrow_next2 = ext[min(q0 + 2 * RQ + k_pl, blk.b - 1)].qrow;
            }
            if (more && tid < RQ * 3) rec_next = __ldg(reinterpret_cast<const uint4 *>(ext + min(q0 + RQ + tid / 3, blk.b - 1)) + tid % 3);
            // ---- this group's queries against the row in registers, two at a time (independent chains for the scheduler)
            for (int k = 0; k < nq; k += 2) {                          // nq is block-uniform
                const bool two = k + 1 < nq;
                const int k1 = two ? k + 1 : k;
                int n11a = 0, n11b = 0;
#pragma unroll
                for (int h = 0; h < 4 * NG; ++h) {
                    n11a += popc_and_8(x[2 * h], x[2 * h + 1], qs[cur][k][2 * h], qs[cur][k][2 * h + 1]);
                    n11b += popc_and_8(x[2 * h], x[2 * h + 1], qs[cur][k1][2 * h], qs[cur][k1][2 * h + 1]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u == 1 && !two) break;
                    const int kk = u ? k1 : k;
                    const uint4 r0 = s_rec[cur][kk][0], r1 = s_rec[cur][kk][1];        // {idnum lo, hi, q, qrow} {lo, hi, ws, we}
                    const uint4 r2 = s_rec[cur][kk][2];                                 // {n1, qa, -, -}
                    const int32_t n1q = (int32_t)r2.x;
                    const int64_t idq = (int64_t)(((unsigned long long)r0.y << 32) | r0.x);
                    const bool scan = row32 >= (int32_t)r1.x && row32 < (int32_t)r1.y       // the query's candidate range
                                      && pos0 < (int32_t)r1.w && end0 > (int32_t)r1.z       // fetch overlap, ld_area.py:215-217
                                      && elig && idn != idq;                                // :223-224, :222
                    pair_tail<true>(A, (int64_t)(int32_t)r0.z, (int64_t)(int32_t)r0.w, row, scan, u ? n11b : n11a, n1q, n1r, tid, scanned, __uint_as_float(r2.y), rb);
                }
            }
            if (more) {
                if (tid < PL) qs[cur ^ 1][k_pl][g_pl] = make_uint4(nx.x & mk.x, nx.y & mk.y, nx.z & mk.z, nx.w & mk.w);
                if (tid < RQ * 3) s_rec[cur ^ 1][tid / 3][tid % 3] = rec_next;
            }
            qrow_next = qrow_next2;
            __syncthreads();
        }
        scanned_total += scanned;
    }
    for (int o = 16; o; o >>= 1) scanned_total += __shfl_xor_sync(0xffffffffu, scanned_total, o);
    if ((tid & 31) == 0 && scanned_total) atomicAdd(A.counters + 1, scanned_total);
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
        const int n11 = transpose_reduce8(cnt, lane8);

        // ---- epilogue: thread t owns row base + t
        row_epilogue(A, q, qrow, base + tid, hi, n11, tid, scanned);
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

template <int NG>
static int launch_window_mq_ng(ldx_ctx *ctx, const WindowArgs &A, const MqBlock *d_blocks, int64_t n_blocks, const MqQueryX *d_sorted,
                               unsigned int *d_next, int64_t n_rows) {
    static int per_sm = 0;
    static const bool l1rows = !(getenv("LDX_WINDOW_MQ_L1") && atoi(getenv("LDX_WINDOW_MQ_L1")) == 0);
    if (!per_sm) {
        LDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_mq_kernel<NG, true>, WIN_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
    }
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > n_blocks) grid = n_blocks;
    // (the caller brackets the record kernel and this one with one timing pair)
    if (l1rows) window_mq_kernel<NG, true><<<(int)grid, WIN_THREADS, 0, ctx->stream>>>(A, d_blocks, n_blocks, d_sorted, d_next, n_rows);
    else window_mq_kernel<NG, false><<<(int)grid, WIN_THREADS, 0, ctx->stream>>>(A, d_blocks, n_blocks, d_sorted, d_next, n_rows);
    timing_end(ctx);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "window_mq_kernel");
    return LDX_OK;
}

static void fill_window_args(ldx_store *s, WindowArgs &A, const int64_t *d_qrow, const int64_t *d_lo, const int64_t *d_hi, const int32_t *d_ws,
                             const int32_t *d_we, int64_t nq, int measure, int thres_e4, ldx_hit *d_hits, int64_t cap, unsigned long long *d_counters) {
    ldx_ctx *ctx = s->ctx;
    A.planes = reinterpret_cast<const uint4 *>(s->d_planes);
    A.mask = reinterpret_cast<const uint4 *>(s->d_mask);
    A.stride_u4 = s->stride_words / 2;
    A.freq = s->d_freq; A.fc = s->fc;
    A.pos0 = s->d_pos0; A.end0 = s->d_end0; A.idnum = s->d_idnum; A.eligible = s->d_eligible;
    A.q_row = d_qrow; A.lo = d_lo; A.hi = d_hi; A.win_start = d_ws; A.win_end = d_we;
    A.chunk_prefix = nullptr; A.nq = nq; A.n_chunks = 0;
    A.measure = measure; A.thres_e4 = thres_e4;
    A.n_sel = s->n_sel;
    const double n = (double)s->n_sel, g = 1.0e4 * 16.0 * 1.1102230246251565e-16 * n * n;
    A.screen_g = (float)(g + 1.0e-4);
    A.screen_t = (s->n_sel <= 8192 && thres_e4 > 0) ? (float)((double)thres_e4 - 0.5 - 1.0e-3) : 0.0f;
    A.hits = d_hits; A.cap = cap; A.counters = d_counters;
    A.fix = FixupSink{ctx->d_fix, ctx->d_fix_count, ctx->fix_capacity, ctx->fix_tag};
    A.gen = s->n_nonsimple > 0 ? s->d_gen : nullptr;
}

bool window_mq_supported(const ldx_store *s) { const int ng = s->stride_words / 16; return ng >= 1 && ng <= 5; }

// d_blocks: MqBlock records; d_ext: one MqQueryX record per query in sorted order, host part filled (WindowMqBlock / WindowMqQueryX
// on the host side); d_next: one word of device scratch for the dynamic block counter
int launch_window_mq(ldx_store *s, int64_t nq, const void *d_blocks, int64_t n_blocks, void *d_ext, int64_t n_sorted, unsigned int *d_next,
                     int measure, int thres_e4, ldx_hit *d_hits, int64_t cap, unsigned long long *d_counters) {
    if (n_blocks <= 0) return LDX_OK;
    static_assert(sizeof(MqBlock) == sizeof(WindowMqBlock) && sizeof(MqQueryX) == sizeof(WindowMqQueryX) && offsetof(MqQueryX, ws) == offsetof(WindowMqQueryX, ws),
                  "host and device work-list records");
    WindowArgs A;
    fill_window_args(s, A, nullptr, nullptr, nullptr, nullptr, nullptr, nq, measure, thres_e4, d_hits, cap, d_counters);   // the kernels read the records only
    const MqBlock *blocks = reinterpret_cast<const MqBlock *>(d_blocks);
    static const bool rows1_off = getenv("LDX_WINDOW_ROWS1") && atoi(getenv("LDX_WINDOW_ROWS1")) == 0;
    if (!d_ext || n_sorted <= 0 || s->n_variants >= (1ll << 31)) return set_error(LDX_ERR_STATE, "multi-query window kernel: no record scratch");
    ldx_ctx *ctx = s->ctx;
    MqQueryX *ext = reinterpret_cast<MqQueryX *>(d_ext);
    LDX_CUDA(cudaMemsetAsync(d_next, 0, sizeof(unsigned int), ctx->stream));
    timing_begin(ctx);                      // one pair around the record kernel and the scan kernel
    mq_extend_kernel<<<(unsigned)((n_sorted + 255) / 256), 256, 0, ctx->stream>>>(A, n_sorted, ext);
    ctx->launches++;
    LDX_LAUNCHED(ctx, "mq_extend_kernel");
    if ((A.stride_u4 == 8 || A.stride_u4 == 16) && !rows1_off) {     // 128- and 256-byte rows: a thread per row
        static int per_sm[2] = {0, 0};
        const int w = A.stride_u4 == 16;
        if (!per_sm[w]) {
            if (w) LDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[w], window_rows1_kernel<2>, WIN_THREADS, 0));
            else LDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[w], window_rows1_kernel<1>, WIN_THREADS, 0));
            if (per_sm[w] < 1) per_sm[w] = 1;
        }
        const int64_t grid = std::min<int64_t>((int64_t)ctx->sm_count * per_sm[w], n_blocks);
        if (w) window_rows1_kernel<2><<<(int)grid, WIN_THREADS, 0, ctx->stream>>>(A, blocks, n_blocks, ext, d_next, s->n_variants);
        else window_rows1_kernel<1><<<(int)grid, WIN_THREADS, 0, ctx->stream>>>(A, blocks, n_blocks, ext, d_next, s->n_variants);
        timing_end(ctx);
        ctx->launches++;
        LDX_LAUNCHED(ctx, "window_rows1_kernel");
        return LDX_OK;
    }
    switch (A.stride_u4 / 8) {
        case 1: return launch_window_mq_ng<1>(ctx, A, blocks, n_blocks, ext, d_next, s->n_variants);
        case 2: return launch_window_mq_ng<2>(ctx, A, blocks, n_blocks, ext, d_next, s->n_variants);
        case 3: return launch_window_mq_ng<3>(ctx, A, blocks, n_blocks, ext, d_next, s->n_variants);     // e.g. two super-populations gathered: 2049..3072 haplotypes
        case 4: return launch_window_mq_ng<4>(ctx, A, blocks, n_blocks, ext, d_next, s->n_variants);
        case 5: return launch_window_mq_ng<5>(ctx, A, blocks, n_blocks, ext, d_next, s->n_variants);
        default: return set_error(LDX_ERR_STATE, "multi-query window kernel: unsupported row pitch");
    }
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
    A.gen = s->n_nonsimple > 0 ? s->d_gen : nullptr;
    switch (A.stride_u4 / 8) {
        case 1: return launch_window_ng<1>(ctx, A);
        case 2: return launch_window_ng<2>(ctx, A);
        case 5: return launch_window_ng<5>(ctx, A);    // 5008 haplotypes
        default: return launch_window_ng<0>(ctx, A);
    }
}

}  // namespace ldx
