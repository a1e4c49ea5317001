// ldx_area_format.cu -- the ld_area writers at scale (second half).
//
// Replaces the per-hit Python of ld_area.py:252-283: for every kept (query, opposing variant) pair one line
// of `'\t'.join(map(str, oppos_var_ann))` (:273-274), one element of the JSON list json.dump(..., indent=4) writes
// (:275-283), or one rsID line (:258-260).  Host code (the records' text columns live on the host; the GPU's part of the
// job -- the window scan -- already returned 16 bytes per kept pair): every query's rows are formatted by one of `threads`
// host threads straight into its place in the caller's buffer, after a sizing pass of the same code over a counting sink.
// At -z 0 configs[2] emits ~3 x 10^7 rows; the Python loop it replaces (four text slices and a list per hit) ran at
// ~3 x 10^5 rows/s.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ldx_internal.h"

#define LDX_REQUIRE(cond, msg) do { if (!(cond)) return ldx::set_error(LDX_ERR_ARG, msg); } while (0)

namespace {

struct CountSink {
    int64_t n = 0;
    void put(const char *, size_t len) { n += (int64_t)len; }
    void ch(char) { ++n; }
};
struct WriteSink {
    char *p;
    void put(const char *s, size_t len) { std::memcpy(p, s, len); p += len; }
    void ch(char c) { *p++ = c; }
};

// str(int) for the positions and distances
template <class S> void put_int(S &s, int64_t v) {
    char buf[24];
    int n = 0;
    const bool neg = v < 0;
    uint64_t u = neg ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) buf[n++] = '-';
    while (n) s.ch(buf[--n]);
}

// str(k / 10000.0): the digit arithmetic of ldx_format_e4 (csrc/ldx_format.cu), NUL-padded to 8 bytes
template <class S> void put_e4(S &s, int32_t e4) {
    char b[8];
    ldx_format_e4(e4, b);
    s.put(b, strnlen(b, 8));
}

// the rounded measure as Python prints the object calc_ld returned: the int 0 or a float (calc_ld.py:68-69, :89-90, :94-95)
template <class S> void put_r2(S &s, uint32_t w) { if (w & LDX_R2_INT0) s.ch('0'); else put_e4(s, (int32_t)(w & LDX_R2_MASK)); }
template <class S> void put_dp(S &s, uint32_t w) { if (w & LDX_DP_INT0) s.ch('0'); else put_e4(s, (int32_t)((w & LDX_DP_MASK) >> LDX_DP_SHIFT)); }

struct Field { const char *p; size_t n; };

#define PUT_LIT(sink, lit) (sink).put(lit, sizeof(lit) - 1)

// json.dumps of a str (ensure_ascii=True): quotes, backslash, control characters, non-ASCII as \uXXXX (UTF-8 decoded)
template <class S> void put_json_str(S &s, Field f) {
    static const char hex[] = "0123456789abcdef";
    auto u16 = [&](uint32_t x) { const char u[6] = {'\\', 'u', hex[(x >> 12) & 15], hex[(x >> 8) & 15], hex[(x >> 4) & 15], hex[x & 15]}; s.put(u, 6); };
    s.ch('"');
    for (size_t i = 0; i < f.n; ++i) {
        const unsigned char c = (unsigned char)f.p[i];
        if (c == '"') PUT_LIT(s, "\\\"");
        else if (c == '\\') PUT_LIT(s, "\\\\");
        else if (c == '\n') PUT_LIT(s, "\\n");
        else if (c == '\r') PUT_LIT(s, "\\r");
        else if (c == '\t') PUT_LIT(s, "\\t");
        else if (c == '\b') PUT_LIT(s, "\\b");
        else if (c == '\f') PUT_LIT(s, "\\f");
        else if (c < 0x20) u16(c);
        else if (c < 0x80) s.ch((char)c);                         // DEL included: json.dumps prints it as is
        else {
            uint32_t cp = 0xfffd;                                 // a malformed sequence: what errors='replace' would give
            const int extra = c >= 0xf0 ? 3 : c >= 0xe0 ? 2 : c >= 0xc0 ? 1 : 0;
            if (extra > 0 && i + (size_t)extra < f.n) {
                cp = c & (0x3f >> extra);
                for (int k = 1; k <= extra; ++k) cp = (cp << 6) | ((unsigned char)f.p[i + (size_t)k] & 0x3f);
                i += (size_t)extra;
            }
            if (cp >= 0x10000) { cp -= 0x10000; u16(0xd800 + (cp >> 10)); u16(0xdc00 + (cp & 0x3ff)); }
            else u16(cp);
        }
    }
    s.ch('"');
}

struct Table {
    const uint8_t *blob; const int64_t *off; const ldx_vcf_row *rows; const int32_t *p_e4;
    Field col(int64_t r, int32_t field_off) const {               // a fixed column: from its offset to the next tab
        const char *rec = reinterpret_cast<const char *>(blob + off[r]);
        const size_t len = (size_t)(off[r + 1] - off[r]);
        size_t a = (size_t)field_off, b = a;
        while (b < len && rec[b] != '\t') ++b;
        return Field{rec + a, b - a};
    }
    Field vt(int64_t r) const {                                   // ','.join(rec.info['VT']): the VT key of INFO, verbatim
        const Field info = col(r, rows[r].info_off);
        size_t i = 0;
        while (i < info.n) {
            size_t e = i;
            while (e < info.n && info.p[e] != ';') ++e;
            if (e - i >= 3 && info.p[i] == 'V' && info.p[i + 1] == 'T' && info.p[i + 2] == '=') return Field{info.p + i + 3, e - i - 3};
            i = e + 1;
        }
        return Field{info.p, 0};
    }
};

// ---- the part of a row that depends on the opposing variant alone, formatted once per variant of the call (a variant sits in
// the windows of ~30 queries at configs[2]): TSV "pos\tid\tref\talt\ttype\t", JSON the element up to `"alt_freq": `
template <class S> void format_row_prefix(S &s, const Table &T, int64_t r, int format) {
    const Field id = T.col(r, T.rows[r].id_off);
    if (format == LDX_AREA_RSIDS) { s.put(id.p, id.n); s.ch('\n'); return; }                      // ld_area.py:258-260: the whole line
    const Field ref = T.col(r, T.rows[r].ref_off), alt = T.col(r, T.rows[r].alt_off), vt = T.vt(r);
    if (format == LDX_AREA_TSV) {                                                                  // :264-274
        put_int(s, T.rows[r].pos); s.ch('\t'); s.put(id.p, id.n); s.ch('\t'); s.put(ref.p, ref.n); s.ch('\t'); s.put(alt.p, alt.n); s.ch('\t');
        s.put(vt.p, vt.n); s.ch('\t');
        return;
    }
    // one element of the list json.dump(trg_obj, ..., indent=4) writes (:275-283), with the separator that precedes it
    PUT_LIT(s, ",\n    {\n        \"hg38_pos\": "); put_int(s, T.rows[r].pos);
    PUT_LIT(s, ",\n        \"rsID\": "); put_json_str(s, id);
    PUT_LIT(s, ",\n        \"ref\": "); put_json_str(s, ref);
    PUT_LIT(s, ",\n        \"alt\": "); put_json_str(s, alt);
    PUT_LIT(s, ",\n        \"type\": "); put_json_str(s, vt);
    PUT_LIT(s, ",\n        \"alt_freq\": ");
}

// str(k / 10000.0) for every k a packed word can hold, built once per process: 8 bytes of text + its length
struct E4Table {
    char txt[20001][8];
    uint8_t len[20001];
    E4Table() {
        for (int k = 0; k <= 20000; ++k) {
            if (ldx_format_e4(k, txt[k]) != LDX_OK) std::memset(txt[k], 0, 8);
            len[k] = (uint8_t)strnlen(txt[k], 8);
        }
    }
};
const E4Table &e4_table() { static const E4Table t; return t; }

struct HitFormatter {
    const Table &T; const E4Table &E; int format;
    const uint32_t *pre_len; const uint64_t *pre_off; const char *pre_txt;          // per store row: its prefix text (touched rows only)
    // the rounded measures as Python prints the objects calc_ld returned: the int 0 or a float (calc_ld.py:68-69, :89-90, :94-95)
    static int r2_e4(uint32_t w) { return (w & LDX_R2_INT0) ? -1 : (int)(w & LDX_R2_MASK); }
    static int dp_e4(uint32_t w) { return (w & LDX_DP_INT0) ? -1 : (int)((w & LDX_DP_MASK) >> LDX_DP_SHIFT); }
    // beyond the table (values above 2.0: only a pairing of lists of unequal ploidy produces them): integer part, '.', the four
    // decimals without trailing zeros but at least one -- what Python prints for the double nearest to k * 10^-4
    static size_t big_len(int k) { int f = k % 10000, nd = 4; while (nd > 1 && f % 10 == 0) { f /= 10; --nd; } return int_len(k / 10000) + 1 + (size_t)nd; }
    static char *put_big(char *p, int k) {
        p = put_int(p, k / 10000);
        *p++ = '.';
        int f = k % 10000, nd = 4;
        while (nd > 1 && f % 10 == 0) { f /= 10; --nd; }
        for (int i = nd - 1; i >= 0; --i) { p[i] = (char)('0' + f % 10); f /= 10; }
        return p + nd;
    }
    size_t e4_len(int k) const { return k < 0 ? 1 : k <= 20000 ? E.len[k] : big_len(k); }
    static size_t int_len(int64_t v) { size_t n = v < 0 ? 2 : 1; uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v; while (u >= 10) { u /= 10; ++n; } return n; }
    char *put_e4(char *p, int k) const {
        if (k < 0) { *p++ = '0'; return p; }
        if (k > 20000) return put_big(p, k);
        std::memcpy(p, E.txt[k], 8);
        return p + E.len[k];
    }
    static char *put_int(char *p, int64_t v) {
        char buf[24];
        int n = 0;
        const bool neg = v < 0;
        uint64_t u = neg ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
        do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (neg) *p++ = '-';
        while (n) *p++ = buf[--n];
        return p;
    }
    // ov: optional per-hit overrides {alt_e4, r2_e4, dp_e4}, each < 0 = "as the tables / the packed word say"
    static int pick(const int32_t *ov, int i, int dflt) { return ov && ov[i] >= 0 ? ov[i] : dflt; }
    size_t size(const ldx_hit &h, int64_t qrow, int32_t alt_e4, const int32_t *ov) const {
        const size_t a = pre_len[h.row];
        if (format == LDX_AREA_RSIDS) return a;
        const int64_t dist = (int64_t)T.rows[h.row].pos - (int64_t)T.rows[qrow].pos;                // :272
        const size_t vals = e4_len(pick(ov, 0, alt_e4)) + e4_len(pick(ov, 1, r2_e4(h.packed))) + e4_len(pick(ov, 2, dp_e4(h.packed))) + int_len(dist);
        return a + vals + (format == LDX_AREA_TSV ? 4 : sizeof(",\n        \"r2\": ") - 1 + sizeof(",\n        \"D'\": ") - 1 + sizeof(",\n        \"dist\": ") - 1 + sizeof("\n    }") - 1);
    }
    // NOTE: writes up to 7 bytes past the end of a value (8-byte copies of the e4 texts); the caller's buffer has that slack
    char *write(char *p, const ldx_hit &h, int64_t qrow, int32_t alt_e4_in, const int32_t *ov) const {
        std::memcpy(p, pre_txt + pre_off[h.row], pre_len[h.row]);
        p += pre_len[h.row];
        if (format == LDX_AREA_RSIDS) return p;
        const int64_t dist = (int64_t)T.rows[h.row].pos - (int64_t)T.rows[qrow].pos;
        const int alt_e4 = pick(ov, 0, alt_e4_in), r2v = pick(ov, 1, r2_e4(h.packed)), dpv = pick(ov, 2, dp_e4(h.packed));
        if (format == LDX_AREA_TSV) {
            p = put_e4(p, alt_e4); *p++ = '\t'; p = put_e4(p, r2v); *p++ = '\t'; p = put_e4(p, dpv); *p++ = '\t';
            p = put_int(p, dist); *p++ = '\n';
            return p;
        }
#define LIT(x) do { std::memcpy(p, x, sizeof(x) - 1); p += sizeof(x) - 1; } while (0)
        p = put_e4(p, alt_e4); LIT(",\n        \"r2\": "); p = put_e4(p, r2v); LIT(",\n        \"D'\": "); p = put_e4(p, dpv);
        LIT(",\n        \"dist\": "); p = put_int(p, dist); LIT("\n    }");
#undef LIT
        return p;
    }
};

}  // namespace

extern "C" int32_t ldx_area_format(const ldx_hit *hits, int64_t n_hits, const int64_t *q_row, int64_t nq, const uint8_t *blob,
                                   const int64_t *blob_off, const ldx_vcf_row *rows, int64_t n_rows, const int32_t *p_e4,
                                   const int32_t *overrides, int32_t format, int32_t threads, char **text_out,
                                   int64_t *n_bytes, int64_t *query_off) {
    LDX_REQUIRE(text_out && n_bytes && query_off && q_row && blob && blob_off && rows && p_e4, "NULL argument");
    LDX_REQUIRE(n_hits >= 0 && nq >= 0 && n_rows >= 0 && (hits || n_hits == 0), "bad sizes");
    LDX_REQUIRE(format == LDX_AREA_TSV || format == LDX_AREA_JSON || format == LDX_AREA_RSIDS, "bad format");
    *n_bytes = 0; *text_out = nullptr;
    // the hits of query k: [first[k], first[k + 1]) -- they arrive sorted by (query, row)
    std::vector<int64_t> first((size_t)nq + 1, 0);
    std::vector<uint8_t> touched((size_t)n_rows, 0);
    for (int64_t i = 0; i < n_hits; ++i) {
        LDX_REQUIRE(hits[i].query >= 0 && hits[i].query < nq && hits[i].row >= 0 && hits[i].row < n_rows, "hit outside the query / row tables");
        LDX_REQUIRE(i == 0 || hits[i].query >= hits[i - 1].query, "hits must be sorted by query");
        first[(size_t)hits[i].query + 1]++;
        touched[(size_t)hits[i].row] = 1;
    }
    for (int64_t k = 0; k < nq; ++k) {
        LDX_REQUIRE(q_row[k] >= 0 && q_row[k] < n_rows, "q_row outside the row table");
        first[(size_t)k + 1] += first[(size_t)k];
    }
    for (int64_t r = 0; r < n_rows; ++r) LDX_REQUIRE(!touched[(size_t)r] || (p_e4[r] >= 0 && p_e4[r] <= 20000), "p_e4 out of range");
    const Table T{blob, blob_off, rows, p_e4};
    const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : (int)std::thread::hardware_concurrency(), std::max<int64_t>(1, n_hits / 4096)));
    auto parallel = [&](int64_t n, int64_t chunk, auto &&body) {   // dynamic chunks of [0, n) over the threads
        std::atomic<int64_t> next{0};
        auto work = [&]() {
            for (;;) {
                const int64_t k0 = next.fetch_add(chunk);
                if (k0 >= n) break;
                for (int64_t k = k0; k < std::min(n, k0 + chunk); ++k) body(k);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_thr; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    };
    // ---- every touched variant's prefix, once: sizes, offsets, text
    std::vector<uint32_t> pre_len((size_t)n_rows, 0);
    std::vector<uint64_t> pre_off((size_t)n_rows + 1, 0);
    parallel(n_rows, 4096, [&](int64_t r) {
        if (!touched[(size_t)r]) return;
        CountSink c;
        format_row_prefix(c, T, r, format);
        pre_len[(size_t)r] = (uint32_t)c.n;
    });
    for (int64_t r = 0; r < n_rows; ++r) pre_off[(size_t)r + 1] = pre_off[(size_t)r] + pre_len[(size_t)r];
    std::vector<char> pre_txt((size_t)pre_off[(size_t)n_rows] + 8);
    parallel(n_rows, 4096, [&](int64_t r) {
        if (!touched[(size_t)r]) return;
        WriteSink w{pre_txt.data() + pre_off[(size_t)r]};
        format_row_prefix(w, T, r, format);
    });
    const HitFormatter F{T, e4_table(), format, pre_len.data(), pre_off.data(), pre_txt.data()};
    auto alt_of = [&](int64_t i) { return p_e4[hits[i].row]; };                                      // var_2_alt_freq, calc_ld.py:97 (complete rows)
    auto ov_of = [&](int64_t i) { return overrides ? overrides + 3 * i : nullptr; };
    // ---- pass 1: sizes (table look-ups only)
    parallel(nq, 16, [&](int64_t k) {
        int64_t n = 0;
        for (int64_t i = first[(size_t)k]; i < first[(size_t)k + 1]; ++i) n += (int64_t)F.size(hits[i], q_row[k], alt_of(i), ov_of(i));
        query_off[k + 1] = n;
    });
    query_off[0] = 0;
    for (int64_t k = 0; k < nq; ++k) query_off[k + 1] += query_off[k];
    char *out = static_cast<char *>(std::malloc((size_t)query_off[nq] + 16));
    if (!out) return ldx::set_error(LDX_ERR_NOMEM, "area text allocation failed");
    // ---- pass 2: every query's rows into their place.  The values are copied as whole 8-byte table entries (up to 7 bytes past
    // their end), so a row is built in a small buffer and copied out exactly: no byte outside the query's range is touched.
    parallel(nq, 1, [&](int64_t k) {
        char row[4096];
        char *dst = out + query_off[k];
        for (int64_t i = first[(size_t)k]; i < first[(size_t)k + 1]; ++i) {
            const size_t need = F.size(hits[i], q_row[k], alt_of(i), ov_of(i));
            if (need + 8 <= sizeof row) {
                F.write(row, hits[i], q_row[k], alt_of(i), ov_of(i));
                std::memcpy(dst, row, need);
            } else {                                               // a record with very long alleles: through a heap buffer
                std::vector<char> big(need + 8);
                F.write(big.data(), hits[i], q_row[k], alt_of(i), ov_of(i));
                std::memcpy(dst, big.data(), need);
            }
            dst += need;
        }
    });
    *text_out = out;
    *n_bytes = query_off[nq];
    return LDX_OK;
}
