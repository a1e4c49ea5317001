/* TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of ld-tools' LD hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (ld_tools_b200 / libldx.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 * tests/golden/calc_ld_golden.json, which was produced by the unmodified reference
 * backend/calc_ld.py (tests/golden/make_golden.py).
 *
 * Citations are paths relative to /root/reference.  `pow` is glibc's libm pow, the very
 * function CPython's float ** calls, so `d ** 2` (calc_ld.py:87) is reproduced bit for bit
 * on the same host.  Compile WITHOUT -ffast-math and with -ffp-contract=off: the reference's
 * `f11 - p1 * p2` (calc_ld.py:50) is a rounded multiply followed by a rounded subtract.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n_hap, n_11, n_a1, n_a0, n_b1, n_b0;   /* calc_ld.py:31-32, :37-40 */
    double d, dprime, r2, p_a, p_b;                /* before rounding (calc_ld.py:50-90) */
    double r2_rounded, dprime_rounded, p_a_rounded, p_b_rounded;   /* calc_ld.py:94-97 */
    int32_t dprime_is_int0;   /* ZeroDivisionError branch, calc_ld.py:68-69 / :75-76 */
    int32_t r2_is_int0;       /* D' == 0 branch, calc_ld.py:89-90 */
} ldo_result;

/* Python's round(x, 4) on a float: correctly rounded decimal conversion of the EXACT binary
 * value (ties to even), then back to the nearest double.  glibc's printf/strtod are both
 * exact, so "%.4f" -> strtod is that definition spelled with libc.  (calc_ld.py:94-97) */
double ldo_round4(double x) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.4f", x);
    return strtod(buf, NULL);
}

/* `d ** 2` is libm pow in CPython (float_pow -> pow).  gcc folds pow(x, 2.0) into x*x, which
 * differs from glibc's pow by 1 ulp for ~0.1% of inputs, so call it through a volatile
 * pointer (and the Makefile passes -fno-builtin-pow). */
static double (*volatile libm_pow)(double, double) = pow;

/* calc_ld.py:33-97 from integer counts.  Returns -1 for n_hap == 0 (the reference raises an
 * uncaught ZeroDivisionError at :33). */
int ldo_finalise(int64_t n_hap, int64_t n_11, int64_t n_a1, int64_t n_a0, int64_t n_b1,
                 int64_t n_b0, ldo_result *o) {
    if (n_hap <= 0) return -1;
    memset(o, 0, sizeof *o);
    o->n_hap = n_hap; o->n_11 = n_11; o->n_a1 = n_a1; o->n_a0 = n_a0; o->n_b1 = n_b1; o->n_b0 = n_b0;
    const double N = (double)n_hap;
    const double f11 = (double)n_11 / N;                       /* :33 */
    const double p_a = (double)n_a1 / N, q_a = (double)n_a0 / N;   /* :41-42 */
    const double p_b = (double)n_b1 / N, q_b = (double)n_b0 / N;   /* :43-44 */
    const double prod = p_a * p_b;
    const double d = f11 - prod;                               /* :50 */
    double bound, dprime = 0.0;
    if (d >= 0) {                                              /* :63 */
        const double x = p_a * q_b, y = q_a * p_b;
        bound = (y < x) ? y : x;                               /* Python min(x, y), :64-65 */
    } else {                                                   /* :70 */
        const double x = (-p_a) * p_b, y = (-q_a) * q_b;
        bound = (y > x) ? y : x;                               /* Python max(x, y), :71-72 */
    }
    if (bound == 0.0) o->dprime_is_int0 = 1;                   /* :68-69, :75-76 */
    else dprime = d / bound;                                   /* :67, :74 */
    double r2 = 0.0;
    if (!o->dprime_is_int0 && dprime != 0.0)                   /* :86 */
        r2 = libm_pow(d, 2.0) / (p_a * q_a * p_b * q_b);       /* :87-88, left-assoc product */
    else
        o->r2_is_int0 = 1;                                     /* :90 */
    o->d = d; o->dprime = dprime; o->r2 = r2; o->p_a = p_a; o->p_b = p_b;
    o->r2_rounded = ldo_round4(r2);                            /* :94 */
    o->dprime_rounded = ldo_round4(dprime);                    /* :95 */
    o->p_a_rounded = ldo_round4(p_a);                          /* :96 */
    o->p_b_rounded = ldo_round4(p_b);                          /* :97 */
    return 0;
}

/* calc_ld.py:30-40 on byte-coded genotypes: 0 = ref, 1 = alt, anything else = "other"
 * (None / 2 / ...: present in the pairing, absent from both allele counts). */
int ldo_calc_ld_bytes(const uint8_t *g_a, int64_t len_a, const uint8_t *g_b, int64_t len_b,
                      ldo_result *o) {
    const int64_t n_hap = len_a < len_b ? len_a : len_b;       /* zip truncation, :30-31 */
    int64_t n_11 = 0, a1 = 0, a0 = 0, b1 = 0, b0 = 0;
    for (int64_t i = 0; i < n_hap; ++i) n_11 += (g_a[i] == 1) & (g_b[i] == 1);   /* :32 */
    for (int64_t i = 0; i < len_a; ++i) { a1 += g_a[i] == 1; a0 += g_a[i] == 0; }   /* :37-38 */
    for (int64_t i = 0; i < len_b; ++i) { b1 += g_b[i] == 1; b0 += g_b[i] == 0; }   /* :39-40 */
    return ldo_finalise(n_hap, n_11, a1, a0, b1, b0, o);
}

/* ---- bitplane helpers: haplotype h of a variant is bit (h % 64) of word (h / 64). ---- */

int64_t ldo_popc_and3(const uint64_t *mask, const uint64_t *a, const uint64_t *b, int64_t words) {
    int64_t n = 0;
    for (int64_t w = 0; w < words; ++w) n += __builtin_popcountll(mask[w] & a[w] & b[w]);
    return n;
}

void ldo_variant_counts(const uint64_t *planes, int64_t stride, int64_t n_variants,
                        const uint64_t *mask, int64_t words, int32_t *n1) {
    for (int64_t v = 0; v < n_variants; ++v)
        n1[v] = (int32_t)ldo_popc_and3(mask, planes + v * stride, planes + v * stride, words);
}

/* Pair list: counts + finalisation per pair (the ld_lite shape, ld_lite.py:143). */
int ldo_pairs(const uint64_t *planes, int64_t stride, const uint64_t *mask, int64_t words,
              const int64_t *ia, const int64_t *ib, int64_t n_pairs, ldo_result *out) {
    int64_t n_hap = 0;
    for (int64_t w = 0; w < words; ++w) n_hap += __builtin_popcountll(mask[w]);
    for (int64_t k = 0; k < n_pairs; ++k) {
        const uint64_t *a = planes + ia[k] * stride, *b = planes + ib[k] * stride;
        const int64_t n11 = ldo_popc_and3(mask, a, b, words);
        const int64_t a1 = ldo_popc_and3(mask, a, a, words), b1 = ldo_popc_and3(mask, b, b, words);
        if (ldo_finalise(n_hap, n11, a1, n_hap - a1, b1, n_hap - b1, out + k)) return -1;
    }
    return 0;
}

/* Lower triangle (row > col) of a variant list, the ld_triangle.py:133-193 loop shape:
 * var_1 = the ROW variant, var_2 = the COLUMN variant.  Output packed by rows,
 * index = row*(row-1)/2 + col.  rows[] are store row indices, already in matrix order. */
int ldo_triangle(const uint64_t *planes, int64_t stride, const uint64_t *mask, int64_t words,
                 const int64_t *rows, int64_t n_rows, ldo_result *out) {
    int64_t n_hap = 0;
    for (int64_t w = 0; w < words; ++w) n_hap += __builtin_popcountll(mask[w]);
    int32_t *n1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_rows > 0 ? n_rows : 1));
    if (!n1) return -2;
    for (int64_t r = 0; r < n_rows; ++r) {
        const uint64_t *a = planes + rows[r] * stride;
        n1[r] = (int32_t)ldo_popc_and3(mask, a, a, words);
    }
    int rc = 0;
    for (int64_t r = 1; r < n_rows && !rc; ++r) {
        const uint64_t *a = planes + rows[r] * stride;
        for (int64_t c = 0; c < r; ++c) {
            const uint64_t *b = planes + rows[c] * stride;
            const int64_t n11 = ldo_popc_and3(mask, a, b, words);
            if (ldo_finalise(n_hap, n11, n1[r], n_hap - n1[r], n1[c], n_hap - n1[c],
                             out + (r * (r - 1) / 2 + c))) { rc = -1; break; }
        }
    }
    free(n1);
    return rc;
}

/* Window scan, the ld_area.py:215-249 loop shape for ONE query: candidate store rows
 * [lo, hi); a row is scanned iff it overlaps the 0-based half-open window
 * [win_start, win_end) (pysam fetch semantics: pos0 < win_end && end0 > win_start), is
 * class-eligible (rs\d+$ id and not MULTI_ALLELIC, ld_area.py:222-225) and does not share the
 * query's id (:222).  var_1 = query, var_2 = opposing row (:242).  A row is KEPT iff the
 * ROUNDED measure >= thres (:248).  measure: 0 = r_square, 1 = d_prime.
 * Writes kept rows (ascending) to out_rows / out_res; returns the number kept. */
int64_t ldo_window(const uint64_t *planes, int64_t stride, const uint64_t *mask, int64_t words,
                   const int32_t *pos0, const int32_t *end0, const int64_t *idnum,
                   const uint8_t *eligible, int64_t q_row, int64_t lo, int64_t hi,
                   int32_t win_start, int32_t win_end, int measure, double thres,
                   int64_t *out_rows, ldo_result *out_res, int64_t cap) {
    int64_t n_hap = 0, kept = 0;
    for (int64_t w = 0; w < words; ++w) n_hap += __builtin_popcountll(mask[w]);
    const uint64_t *a = planes + q_row * stride;
    const int64_t a1 = ldo_popc_and3(mask, a, a, words);
    for (int64_t j = lo; j < hi; ++j) {
        if (!(pos0[j] < win_end && end0[j] > win_start)) continue;
        if (!eligible[j] || idnum[j] == idnum[q_row]) continue;
        const uint64_t *b = planes + j * stride;
        const int64_t b1 = ldo_popc_and3(mask, b, b, words);
        ldo_result r;
        if (ldo_finalise(n_hap, ldo_popc_and3(mask, a, b, words), a1, n_hap - a1, b1, n_hap - b1, &r))
            return -1;
        const double val = measure == 0 ? r.r2_rounded : r.dprime_rounded;
        if (val < thres) continue;
        if (kept < cap) { out_rows[kept] = j; out_res[kept] = r; }
        ++kept;
    }
    return kept;
}

/* 1000G-style GT text -> bitplanes (the store builder's definition).  Row layout: n_samples
 * fields of 4 bytes "a|b" + one separator byte (TAB, or LF/anything after the last sample);
 * haplotype index = 2*sample + allele slot, matching the flat list the drivers build with
 * `+= rec.samples[name]['GT']` (ld_area.py:182-187).  status[v] = 0 ok, 1 = a byte outside
 * the phased diploid biallelic alphabet ('0'/'1' alleles, '|' separator). */
void ldo_pack_gt(const uint8_t *text, const int64_t *row_off, int64_t n_variants, int32_t n_samples,
                 uint64_t *planes, int64_t stride, uint8_t *status) {
    for (int64_t v = 0; v < n_variants; ++v) {
        const uint8_t *row = text + row_off[v];
        uint64_t *pl = planes + v * stride;
        memset(pl, 0, sizeof(uint64_t) * (size_t)stride);
        uint8_t bad = 0;
        for (int32_t s = 0; s < n_samples; ++s) {
            const uint8_t c0 = row[4 * s], sep = row[4 * s + 1], c1 = row[4 * s + 2];
            if ((c0 != '0' && c0 != '1') || (c1 != '0' && c1 != '1') || sep != '|') bad = 1;
            const int64_t h = 2 * (int64_t)s;
            if (c0 == '1') pl[h >> 6] |= 1ull << (h & 63);
            if (c1 == '1') pl[(h + 1) >> 6] |= 1ull << ((h + 1) & 63);
        }
        status[v] = bad;
    }
}

int64_t ldo_sizeof_result(void) { return (int64_t)sizeof(ldo_result); }

/* The engine's packed result word (include/ldx.h) derived from an oracle result: lets the tests
 * compare millions of pairs array-to-array.  0x8000 = r2 is int 0, 0x80000000 = D' is int 0. */
uint32_t ldo_packed_word(const ldo_result *r) {
    uint32_t w = 0;
    if (r->r2_is_int0) w |= 0x00008000u; else w |= (uint32_t)llround(r->r2_rounded * 10000.0);
    if (r->dprime_is_int0) w |= 0x80000000u; else w |= ((uint32_t)llround(r->dprime_rounded * 10000.0)) << 16;
    return w;
}

int ldo_finalise_packed_many(int64_t n_hap, const int32_t *n11, const int32_t *n1a, const int32_t *n1b,
                             int64_t n, uint32_t *out) {
    for (int64_t k = 0; k < n; ++k) {
        ldo_result r;
        if (ldo_finalise(n_hap, n11[k], n1a[k], n_hap - n1a[k], n1b[k], n_hap - n1b[k], &r)) return -1;
        out[k] = ldo_packed_word(&r);
    }
    return 0;
}
