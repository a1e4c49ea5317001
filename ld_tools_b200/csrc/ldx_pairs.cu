// ldx_pairs.cu -- K3: LD of an explicit list of variant pairs, and the list-level calculator.
//
//   pairs_kernel  replaces calc_ld(g1, g2) at ld_lite.py:143 (and any caller holding store rows)
//   lists_kernel  replaces backend/calc_ld.py:3-99 for two raw genotype vectors, including the
//                 reference's quirks: zip truncation (:30-31) and non-0/1 entries (:37-40)
#include "ldx_internal.h"
#include "ldx_fixup.cuh"

namespace ldx {

constexpr int PAIRS_THREADS = 256;

// Eight lanes per pair: lane l loads 16-byte granules l, l+8, ... of both rows (each group load
// covers one 128-byte line), ANDs with the mask, popcounts; a 3-step shuffle tree sums the group.
__global__ void __launch_bounds__(PAIRS_THREADS)
pairs_kernel(const uint4 *__restrict__ planes, const uint4 *__restrict__ mask, int32_t stride_u4,
             const VarFreq *__restrict__ freq, FinalCtx fc, const int64_t *__restrict__ ia,
             const int64_t *__restrict__ ib, int64_t n, int32_t *__restrict__ o_n11, double *__restrict__ o_d,
             double *__restrict__ o_dp, double *__restrict__ o_r2, uint32_t *__restrict__ o_packed, FixupSink fix,
             const GenStore *__restrict__ gen) {
    const int sub = threadIdx.x & 7;
    const int64_t group0 = ((int64_t)blockIdx.x * PAIRS_THREADS + threadIdx.x) >> 3;
    const int64_t n_groups = ((int64_t)gridDim.x * PAIRS_THREADS) >> 3;
    for (int64_t kb = group0 & ~3ll; kb < n; kb += n_groups) {     // warp-uniform bound
        const int64_t k = kb + (group0 & 3);
        const bool valid = k < n;
        const int64_t ra = valid ? ia[k] : 0, rb = valid ? ib[k] : 0;
        const uint4 *pa = planes + ra * stride_u4, *pb = planes + rb * stride_u4;
        int cnt = 0;
        if (valid)
            for (int g = sub; g < stride_u4; g += 8) {
                const uint4 a = ldg_u4(pa + g), b = ldg_u4(pb + g), m = __ldg(mask + g);
                cnt += __popc(a.x & b.x & m.x) + __popc(a.y & b.y & m.y) + __popc(a.z & b.z & m.z) +
                       __popc(a.w & b.w & m.w);
            }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
        if (sub == 0 && valid) {
            const VarFreq fa = freq[ra], fb = freq[rb];
            if (gen && (fa.n1 | fb.n1) < 0) {                 // a variant of the general route: explicit counts, the pairing's own N
                const GenCounts c = general_pair_counts(*gen, ra, fa.n1, rb, fb.n1);
                const GenFinal g = finalise_general(c);
                if (o_n11) o_n11[k] = c.n11;
                if (o_d) o_d[k] = g.d;
                if (o_dp) o_dp[k] = g.dprime;
                if (o_r2) o_r2[k] = g.r2;
                if (o_packed) {
                    o_packed[k] = g.packed;
                    if (g.packed & LDX_R2_NEARTIE) fixup_append_general(fix, (uint64_t)k, c, g.packed);
                }
                continue;
            }
            const PairFinal f = finalise_pair(cnt, fa, fb, fc);
            if (o_n11) o_n11[k] = cnt;
            if (o_d) o_d[k] = f.d;
            if (o_dp) o_dp[k] = f.dprime;
            if (o_r2) o_r2[k] = f.r2;
            if (o_packed) {
                o_packed[k] = f.packed;
                if (f.packed & LDX_R2_NEARTIE) fixup_append(fix, (uint64_t)k, cnt, fa.n1, fb.n1, f.packed);
            }
        }
    }
}

int launch_pairs(ldx_store *s, const int64_t *d_ia, const int64_t *d_ib, int64_t n, int32_t *d_n11,
                 double *d_d, double *d_dp, double *d_r2, uint32_t *d_packed) {
    if (n <= 0) return LDX_OK;
    ldx_ctx *ctx = s->ctx;
    int64_t blocks = (n * 8 + PAIRS_THREADS - 1) / PAIRS_THREADS;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    pairs_kernel<<<(int)blocks, PAIRS_THREADS, 0, ctx->stream>>>(
        reinterpret_cast<const uint4 *>(s->d_planes), reinterpret_cast<const uint4 *>(s->d_mask),
        s->stride_words / 2, s->d_freq, s->fc, d_ia, d_ib, n, d_n11, d_d, d_dp, d_r2, d_packed,
        FixupSink{ctx->d_fix, ctx->d_fix_count, ctx->fix_capacity, ctx->fix_tag}, s->n_nonsimple > 0 ? s->d_gen : nullptr);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------
// calc_ld.py:33-97 for arrays of integer counts (n11, n1 of var_1, n1 of var_2) under a given N:
// the finalisation alone, for callers that already hold counts, and the direct test hook for the
// fp64 path (tests compare millions of count triples with the reference arithmetic).
__global__ void __launch_bounds__(256)
finalise_counts_kernel(FinalCtx fc, const int32_t *__restrict__ n11, const int32_t *__restrict__ n1a,
                       const int32_t *__restrict__ n1b, int64_t n, double *__restrict__ o_d, double *__restrict__ o_dp,
                       double *__restrict__ o_r2, uint32_t *__restrict__ o_packed, FixupSink fix) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        VarFreq fa, fb;
        bool tie;
        fa.n1 = n1a[k]; fb.n1 = n1b[k];
        fa.p = __ddiv_rn((double)fa.n1, fc.n_hap); fa.q = __ddiv_rn((double)((int)fc.n_hap - fa.n1), fc.n_hap);
        fb.p = __ddiv_rn((double)fb.n1, fc.n_hap); fb.q = __ddiv_rn((double)((int)fc.n_hap - fb.n1), fc.n_hap);
        fa.pq = __dmul_rn(fa.p, fa.q); fb.pq = __dmul_rn(fb.p, fb.q);
        fa.p_e4 = (int32_t)round4_e4(fa.p, tie); fb.p_e4 = (int32_t)round4_e4(fb.p, tie);
        const PairFinal f = finalise_pair(n11[k], fa, fb, fc);
        if (o_d) o_d[k] = f.d;
        if (o_dp) o_dp[k] = f.dprime;
        if (o_r2) o_r2[k] = f.r2;
        if (o_packed) {
            o_packed[k] = f.packed;
            if (f.packed & LDX_R2_NEARTIE) fixup_append(fix, (uint64_t)k, n11[k], fa.n1, fb.n1, f.packed);
        }
    }
}

int launch_finalise_counts(ldx_ctx *ctx, const FinalCtx &fc, const int32_t *d_n11, const int32_t *d_n1a,
                           const int32_t *d_n1b, int64_t n, double *d_d, double *d_dp, double *d_r2, uint32_t *d_packed) {
    if (n <= 0) return LDX_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    finalise_counts_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(fc, d_n11, d_n1a, d_n1b, n, d_d, d_dp, d_r2, d_packed,
                                                                  FixupSink{ctx->d_fix, ctx->d_fix_count, ctx->fix_capacity, ctx->fix_tag});
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    return LDX_OK;
}

// ------------------------------------------------------------------------------------------
// List-level calculator: one CTA counts over byte-coded genotypes, thread 0 finalises with the
// general formula (explicit ref counts, so "other" alleles and unequal lengths behave exactly
// like the reference's list.count() calls).
constexpr int LISTS_THREADS = 512;

__device__ __forceinline__ int block_sum(int v, int *scratch) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < LISTS_THREADS / 32; ++w) t += scratch[w];
    return t;
}

__global__ void __launch_bounds__(LISTS_THREADS)
lists_kernel(const uint8_t *__restrict__ ga, int64_t len_a, const uint8_t *__restrict__ gb, int64_t len_b,
             ldx_ld_result *__restrict__ out) {
    __shared__ int scratch[LISTS_THREADS / 32];
    const int64_t n_hap = len_a < len_b ? len_a : len_b;                 // calc_ld.py:30-31
    int c11 = 0, a1 = 0, a0 = 0, b1 = 0, b0 = 0;
    for (int64_t i = threadIdx.x; i < len_a; i += LISTS_THREADS) {
        const uint8_t x = ga[i];
        a1 += x == 1; a0 += x == 0;                                      // :37-38
        if (i < n_hap) c11 += (x == 1) & (gb[i] == 1);                   // :32
    }
    for (int64_t i = threadIdx.x; i < len_b; i += LISTS_THREADS) {
        const uint8_t y = gb[i];
        b1 += y == 1; b0 += y == 0;                                      // :39-40
    }
    c11 = block_sum(c11, scratch); a1 = block_sum(a1, scratch); a0 = block_sum(a0, scratch);
    b1 = block_sum(b1, scratch);   b0 = block_sum(b0, scratch);
    if (threadIdx.x != 0) return;
    ldx_ld_result r;
    r.n_hap = n_hap; r.n_11 = c11; r.n_a1 = a1; r.n_a0 = a0; r.n_b1 = b1; r.n_b0 = b0;
    const double N = (double)n_hap;
    const double f11 = __ddiv_rn((double)c11, N);                        // :33
    const double pa = __ddiv_rn((double)a1, N), qa = __ddiv_rn((double)a0, N);   // :41-42
    const double pb = __ddiv_rn((double)b1, N), qb = __ddiv_rn((double)b0, N);   // :43-44
    const double t = __dmul_rn(pa, pb);
    const double d = __dsub_rn(f11, t);                                  // :50
    double bound;
    if (d >= 0.0) { const double x = __dmul_rn(pa, qb), y = __dmul_rn(qa, pb); bound = (y < x) ? y : x; }
    else { const double x = -t, y = -__dmul_rn(qa, qb); bound = (y > x) ? y : x; }
    bool tie;
    r.d = d; r.dprime = 0.0; r.r2 = 0.0; r.dprime_e4 = 0.0; r.r2_e4 = 0.0;
    r.dprime_is_int0 = 0; r.r2_is_int0 = 0;
    r.p_a = pa; r.p_b = pb;
    r.p_a_e4 = round4_e4_wide(pa, tie); r.p_b_e4 = round4_e4_wide(pb, tie);       // :96-97
    if (bound == 0.0) { r.dprime_is_int0 = 1; r.r2_is_int0 = 1; }       // :68-69, :89-90
    else {
        r.dprime = __ddiv_rn(d, bound);                                  // :67 / :74
        r.dprime_e4 = round4_e4_wide(r.dprime, tie);
        if (r.dprime != 0.0) {
            const double den = __dmul_rn(__dmul_rn(__dmul_rn(pa, qa), pb), qb);   // :87-88
            r.r2 = __ddiv_rn(__dmul_rn(d, d), den);
            r.r2_e4 = round4_e4_wide(r.r2, tie);
            if (tie) r.r2_is_int0 = 2;      // in-flight marker: host settles the tie with libm pow
        } else r.r2_is_int0 = 1;
    }
    *out = r;
}

int launch_lists(ldx_ctx *ctx, const uint8_t *d_ga, int64_t len_a, const uint8_t *d_gb, int64_t len_b,
                 ldx_ld_result *d_out) {
    lists_kernel<<<1, LISTS_THREADS, 0, ctx->stream>>>(d_ga, len_a, d_gb, len_b, d_out);
    ctx->launches++;
    LDX_CUDA(cudaGetLastError());
    return LDX_OK;
}

}  // namespace ldx
