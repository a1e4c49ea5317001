// ldx_inflate.cu -- host side of the ingest path: <chrom>.vcf.gz -> decompressed text in memory (host code only).
//
// 1000 Genomes VCFs are BGZF files (what tabix indexes, prep_intgen_data.py:138): a series of independent gzip
// members of at most 64 KiB each, every one announcing its compressed size in a "BC" extra field and its
// uncompressed size in its trailer.  That makes the layout of the output known before a single byte is inflated,
// so the members are inflated by all host cores in parallel, each straight into its final place.  A plain gzip
// file (one stream, or several concatenated members without the BC field) is inflated sequentially instead.
// The text then goes to the GPU in one piece (ldx_store_ingest_vcf).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

#include "ldx_internal.h"

#define LDX_REQUIRE(cond, msg) do { if (!(cond)) return ldx::set_error(LDX_ERR_ARG, msg); } while (0)

namespace {

using Member = ldx::BgzfMember;       // {in_off, in_len, out_off, out_len, data_off}; data_off: first deflate byte inside the member

// Parses one gzip member header at p[0..n).  Returns the header length (0 = not a gzip header) and, for a BGZF
// member, its total size in *bsize (else 0).
size_t gzip_header(const uint8_t *p, size_t n, size_t *bsize) {
    *bsize = 0;
    if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) return 0;
    const uint8_t flg = p[3];
    size_t h = 10;
    if (flg & 4) {                                         // FEXTRA
        if (h + 2 > n) return 0;
        const size_t xlen = p[h] | (p[h + 1] << 8);
        const size_t x0 = h + 2;
        if (x0 + xlen > n) return 0;
        for (size_t q = x0; q + 4 <= x0 + xlen;) {
            const size_t slen = p[q + 2] | (p[q + 3] << 8);
            if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= x0 + xlen) *bsize = (size_t)(p[q + 4] | (p[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        h = x0 + xlen;
    }
    if (flg & 8) { while (h < n && p[h]) ++h; ++h; }      // FNAME
    if (flg & 16) { while (h < n && p[h]) ++h; ++h; }     // FCOMMENT
    if (flg & 2) h += 2;                                   // FHCRC
    return h <= n ? h : 0;
}

bool inflate_member(const uint8_t *in, const Member &m, uint8_t *out) {
    z_stream z;
    std::memset(&z, 0, sizeof z);
    if (inflateInit2(&z, -15) != Z_OK) return false;       // raw deflate: the header was parsed above
    z.next_in = const_cast<Bytef *>(in + m.in_off + m.data_off);
    z.avail_in = (uInt)(m.in_len - m.data_off - 8);
    z.next_out = out + m.out_off;
    z.avail_out = (uInt)m.out_len;
    const int rc = inflate(&z, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && z.total_out == m.out_len;
    inflateEnd(&z);
    if (!ok) return false;
    const uint8_t *t = in + m.in_off + m.in_len - 8;       // trailer: CRC32, ISIZE
    const uint32_t crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
    return (uint32_t)crc32(0L, out + m.out_off, (uInt)m.out_len) == crc;
}

}  // namespace

namespace ldx {

// Walks the members of a BGZF file by their announced sizes.  false: not (entirely) BGZF.
bool bgzf_scan(const uint8_t *in, size_t n, std::vector<BgzfMember> &members, size_t *total_out) {
    members.clear();
    size_t total = 0;
    for (size_t off = 0; off < n;) {
        size_t bsize = 0;
        const size_t h = gzip_header(in + off, n - off, &bsize);
        if (!h || !bsize || bsize < h + 8 || off + bsize > n) return false;
        const uint8_t *t = in + off + bsize - 4;
        const size_t isize = t[0] | (t[1] << 8) | (t[2] << 16) | ((size_t)t[3] << 24);
        if (isize > (1u << 16)) return false;
        members.push_back(BgzfMember{off, bsize, total, isize, h});
        total += isize;
        off += bsize;
    }
    *total_out = total;
    return true;
}

// Members [first, last) inflated by `threads` host threads, member k to out + (members[k].out_off - members[first].out_off).
bool bgzf_inflate_range(const uint8_t *in, const std::vector<BgzfMember> &members, size_t first, size_t last, uint8_t *out, int threads) {
    if (first >= last) return true;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min<int>(nt, 64));
    nt = (int)std::min<size_t>((size_t)nt, std::max<size_t>((last - first) / 16, 1));
    uint8_t *base = out - members[first].out_off;
    std::atomic<size_t> next{first};
    std::atomic<bool> failed{false};
    auto work = [&]() {
        for (;;) {
            const size_t b = next.fetch_add(64);               // 64 members (<= 4 MiB of text) per grab
            if (b >= last || failed.load()) return;
            const size_t e = std::min(last, b + 64);
            for (size_t k = b; k < e; ++k)
                if (members[k].out_len && !inflate_member(in, members[k], base)) { failed.store(true); return; }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return !failed.load();
}

int read_whole_file(const char *path, std::vector<uint8_t> &in) {
    FILE *fh = fopen(path, "rb");
    if (!fh) return set_error(LDX_ERR_ARG, std::string("cannot open ") + path);
    fseek(fh, 0, SEEK_END);
    const long sz = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    if (sz < 0) { fclose(fh); return set_error(LDX_ERR_ARG, "cannot size the file"); }
    in.resize((size_t)sz);
    const size_t got = sz ? fread(in.data(), 1, (size_t)sz, fh) : 0;
    fclose(fh);
    if (got != (size_t)sz) return set_error(LDX_ERR_ARG, "short read");
    return LDX_OK;
}

}  // namespace ldx

/* Whole .gz file -> malloc'ed text (*text_out, free with ldx_free_host).  threads <= 0: all host cores.
 * *was_bgzf_out (may be NULL) = 1 when the file was a BGZF series inflated in parallel. */
extern "C" int32_t ldx_inflate_gz_file(const char *path, int32_t threads, uint8_t **text_out, int64_t *text_bytes_out,
                                       int32_t *was_bgzf_out) {
    LDX_REQUIRE(path && text_out && text_bytes_out, "NULL argument");
    *text_out = nullptr; *text_bytes_out = 0;
    if (was_bgzf_out) *was_bgzf_out = 0;
    FILE *fh = fopen(path, "rb");
    if (!fh) return ldx::set_error(LDX_ERR_ARG, std::string("inflate: cannot open ") + path);
    std::vector<uint8_t> in;
    {
        fseek(fh, 0, SEEK_END);
        const long sz = ftell(fh);
        fseek(fh, 0, SEEK_SET);
        if (sz < 0) { fclose(fh); return ldx::set_error(LDX_ERR_ARG, "inflate: cannot size the file"); }
        in.resize((size_t)sz);
        const size_t got = sz ? fread(in.data(), 1, (size_t)sz, fh) : 0;
        fclose(fh);
        if (got != (size_t)sz) return ldx::set_error(LDX_ERR_ARG, "inflate: short read");
    }
    if (in.empty()) return LDX_OK;
    // ---- BGZF?  Walk the members by their announced sizes.
    std::vector<Member> members;
    bool bgzf = true;
    size_t total = 0;
    for (size_t off = 0; off < in.size();) {
        size_t bsize = 0;
        const size_t h = gzip_header(in.data() + off, in.size() - off, &bsize);
        if (!h || !bsize || bsize < h + 8 || off + bsize > in.size()) { bgzf = false; break; }
        const uint8_t *t = in.data() + off + bsize - 4;
        const size_t isize = t[0] | (t[1] << 8) | (t[2] << 16) | ((size_t)t[3] << 24);
        if (isize > (1u << 16)) { bgzf = false; break; }
        members.push_back(Member{off, bsize, total, isize, h});
        total += isize;
        off += bsize;
    }
    if (bgzf) {
        uint8_t *out = static_cast<uint8_t *>(std::malloc(std::max<size_t>(total, 1)));
        if (!out) return ldx::set_error(LDX_ERR_NOMEM, "inflate: out of host memory");
        int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
        nt = std::max(1, std::min<int>(nt, 64));
        nt = (int)std::min<size_t>((size_t)nt, std::max<size_t>(members.size() / 16, 1));
        std::atomic<size_t> next{0};
        std::atomic<bool> failed{false};
        auto work = [&]() {
            for (;;) {
                const size_t b = next.fetch_add(64);               // 64 members (<= 4 MiB of text) per grab
                if (b >= members.size() || failed.load()) return;
                const size_t e = std::min(members.size(), b + 64);
                for (size_t k = b; k < e; ++k)
                    if (members[k].out_len && !inflate_member(in.data(), members[k], out)) { failed.store(true); return; }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        if (failed.load()) { std::free(out); return ldx::set_error(LDX_ERR_ARG, "inflate: corrupt BGZF block (deflate error or CRC mismatch)"); }
        *text_out = out; *text_bytes_out = (int64_t)total;
        if (was_bgzf_out) *was_bgzf_out = 1;
        return LDX_OK;
    }
    // ---- plain gzip: one stream (or concatenated members), sequentially
    size_t cap = std::max<size_t>(in.size() * 8, 1 << 20), used = 0;
    uint8_t *out = static_cast<uint8_t *>(std::malloc(cap));
    if (!out) return ldx::set_error(LDX_ERR_NOMEM, "inflate: out of host memory");
    z_stream z;
    std::memset(&z, 0, sizeof z);
    if (inflateInit2(&z, 15 + 16) != Z_OK) { std::free(out); return ldx::set_error(LDX_ERR_STATE, "inflate: zlib init failed"); }
    const uint8_t *src = in.data();
    size_t remaining = in.size();
    int rc = Z_OK;
    bool done = false;
    auto refill = [&]() {
        if (z.avail_in == 0 && remaining > 0) {
            const size_t chunk = std::min<size_t>(remaining, 1u << 30);
            z.next_in = const_cast<Bytef *>(src); z.avail_in = (uInt)chunk;
            src += chunk; remaining -= chunk;
        }
    };
    while (!done) {
        refill();
        if (used == cap) {
            uint8_t *bigger = static_cast<uint8_t *>(std::realloc(out, cap * 2));
            if (!bigger) { inflateEnd(&z); std::free(out); return ldx::set_error(LDX_ERR_NOMEM, "inflate: out of host memory"); }
            out = bigger; cap *= 2;
        }
        const size_t room = std::min<size_t>(cap - used, 1u << 30);
        z.next_out = out + used;
        z.avail_out = (uInt)room;
        rc = inflate(&z, Z_NO_FLUSH);
        used += room - z.avail_out;
        if (rc == Z_STREAM_END) {
            refill();
            if (z.avail_in >= 2 && z.next_in[0] == 0x1f && z.next_in[1] == 0x8b) {      // a further member follows
                Bytef *nx = z.next_in; const uInt left = z.avail_in;
                if (inflateReset(&z) != Z_OK) { rc = Z_DATA_ERROR; break; }
                z.next_in = nx; z.avail_in = left;
                continue;
            }
            done = true;                                     // end of the last member (trailing bytes, if any, are ignored)
        } else if (rc == Z_BUF_ERROR) {
            if (z.avail_in == 0 && remaining == 0 && z.avail_out != 0) break;           // truncated stream
        } else if (rc != Z_OK) {
            break;
        }
    }
    inflateEnd(&z);
    if (!done) { std::free(out); return ldx::set_error(LDX_ERR_ARG, "inflate: not a complete gzip stream"); }
    *text_out = out; *text_bytes_out = (int64_t)used;
    return LDX_OK;
}

extern "C" int32_t ldx_free_host(void *p) {
    std::free(p);
    return LDX_OK;
}
