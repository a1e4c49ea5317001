"""Synthetic 1000 Genomes-shaped data: phased haplotypes with LD structure, panel, VCF, conversion.db.

The real GRCh38 rsID VCFs the reference was written for are gone upstream (reference README.md:1-2)
and there is no network, so correctness and throughput are judged on data of the same SHAPE:
2504 samples / 5008 haplotypes in 26 populations and 5 super-populations (EUR = 503 samples),
phased diploid GT columns, mostly-rare allele-frequency spectrum, LD blocks that decay with
distance.  Haplotypes are mosaics of a small founder panel (Li-Stephens style copying with
recombination switches), which yields realistic r2 >= 0.8 neighbourhoods and variants that are
monomorphic inside a sub-population.
"""
import os
import sqlite3

import numpy as np

SEED = 20130502


class BgzfWriter:
    """Minimal BGZF writer (the block-gzip container htslib / tabix use): gzip members of at most 65,280 input
    bytes, each with a 'BC' extra field holding its compressed size - 1, closed by the 28-byte empty EOF block."""

    BLOCK = 65280

    def __init__(self, path, level=1):
        import zlib
        self._z, self._fh, self._buf, self._level = zlib, open(path, "wb"), bytearray(), level

    def _emit(self, chunk):
        import struct
        c = self._z.compressobj(self._level, self._z.DEFLATED, -15)
        body = c.compress(bytes(chunk)) + c.flush()
        self._fh.write(struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, len(body) + 25))
        self._fh.write(body + struct.pack("<II", self._z.crc32(bytes(chunk)) & 0xffffffff, len(chunk)))

    def write(self, data):
        self._buf += data
        while len(self._buf) >= self.BLOCK:
            self._emit(self._buf[:self.BLOCK])
            del self._buf[:self.BLOCK]

    def close(self):
        if self._buf:
            self._emit(self._buf)
        self._emit(b"")
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

# 1000G phase 3 super-population sizes: AFR 661, AMR 347, EAS 504, EUR 503, SAS 489
SUPER_POPS = [("AFR", 661, ["YRI", "LWK", "GWD", "MSL", "ESN", "ASW", "ACB"]),
              ("AMR", 347, ["MXL", "PUR", "CLM", "PEL"]),
              ("EAS", 504, ["CHB", "JPT", "CHS", "CDX", "KHV"]),
              ("EUR", 503, ["CEU", "TSI", "FIN", "GBR", "IBS"]),
              ("SAS", 489, ["GIH", "PJL", "BEB", "STU", "ITU"])]


def make_panel(n_samples=2504, seed=SEED):
    """-> list of (sample, pop, super_pop, gender) in VCF column order."""
    rng = np.random.default_rng(seed)
    total = sum(n for _, n, _ in SUPER_POPS)
    rows = []
    for sp, n, pops in SUPER_POPS:
        k = max(1, round(n * n_samples / total)) if n_samples != total else n
        for i in range(k):
            rows.append([pops[i % len(pops)], sp])
    rows = rows[:n_samples]
    while len(rows) < n_samples:
        rows.append(["CEU", "EUR"])
    order = rng.permutation(len(rows))           # populations interleaved across VCF columns
    out = []
    for j, i in enumerate(order):
        pop, sp = rows[i]
        out.append((f"HG{j:05d}", pop, sp, "male" if rng.random() < 0.5 else "female"))
    return out


def synth_haplotypes(n_variants, n_hap, seed=SEED, n_founders=48, switch_rate=0.004, pop_of_hap=None):
    """-> uint8 [n_variants, n_hap] of 0/1 alleles with block-like LD.

    pop_of_hap (optional int array [n_hap]) biases each group towards its own founders so that
    allele frequencies differ between super-populations."""
    rng = np.random.default_rng(seed)
    n_groups = int(pop_of_hap.max()) + 1 if pop_of_hap is not None else 1
    # founder alleles: frequency spectrum ~ Beta(0.25, 1.6): most variants rare, some common
    freq = rng.beta(0.25, 1.6, size=n_variants)
    founders = (rng.random((n_variants, n_founders)) < freq[:, None]).astype(np.uint8)
    out = np.empty((n_variants, n_hap), dtype=np.uint8)
    # each group prefers a window of the founder panel
    pref = np.zeros((n_groups, n_founders))
    for g in range(n_groups):
        centre = (g + 0.5) * n_founders / n_groups
        dist = np.minimum(np.abs(np.arange(n_founders) - centre), n_founders - np.abs(np.arange(n_founders) - centre))
        pref[g] = np.exp(-dist / (n_founders / (1.5 * n_groups))) + 0.02
        pref[g] /= pref[g].sum()
    groups = pop_of_hap if pop_of_hap is not None else np.zeros(n_hap, dtype=np.int64)
    chunk = 512
    for h0 in range(0, n_hap, chunk):
        h1 = min(n_hap, h0 + chunk)
        nh = h1 - h0
        switches = rng.random((nh, n_variants)) < switch_rate
        switches[:, 0] = True
        seg = np.cumsum(switches, axis=1) - 1                      # segment id per (hap, variant)
        n_seg = int(seg.max()) + 1
        choice = np.empty((nh, n_seg), dtype=np.int64)
        for g in range(n_groups):
            rows = np.flatnonzero(groups[h0:h1] == g)
            if rows.size:
                choice[rows] = rng.choice(n_founders, size=(rows.size, n_seg), p=pref[g])
        path = np.take_along_axis(choice, seg, axis=1)             # founder copied at each variant
        out[:, h0:h1] = founders[np.arange(n_variants)[None, :], path].T
    # sprinkle private mutations (singletons/doubletons are the bulk of real 1000G sites)
    n_priv = n_variants // 3
    rows = rng.integers(0, n_variants, size=n_priv)
    cols = rng.integers(0, n_hap, size=n_priv)
    out[rows, cols] ^= 1
    return out


def pack_bits(h01):
    """[V, n_hap] 0/1 -> [V, stride_words] uint64 planes in the store layout (include/ldx.h)."""
    from .engine import stride_words
    h01 = np.ascontiguousarray(h01, dtype=np.uint8)
    n_var, n_hap = h01.shape
    stride = stride_words(n_hap)
    padded = np.zeros((n_var, stride * 64), dtype=np.uint8)
    padded[:, :n_hap] = h01
    return np.packbits(padded, axis=1, bitorder="little").view("<u8").reshape(n_var, stride)


def random_planes(n_variants, n_hap, seed=SEED, density=None):
    """Fast bench input: planes whose variants have a rare-heavy frequency spectrum (no LD model)."""
    from .engine import stride_words
    rng = np.random.default_rng(seed)
    stride = stride_words(n_hap)
    planes = np.zeros((n_variants, stride), dtype="<u8")
    words = (n_hap + 63) // 64
    freq = rng.beta(0.4, 1.2, size=n_variants) if density is None else np.full(n_variants, density)
    # AND of k random words has density 2^-k; mix levels to approximate each variant's frequency
    raw = rng.integers(0, 1 << 63, size=(n_variants, words, 3), dtype=np.uint64) << np.uint64(1)
    raw |= rng.integers(0, 2, size=(n_variants, words, 3), dtype=np.uint64)
    lvl = np.clip(np.round(-np.log2(np.maximum(freq, 1e-4))), 0, 3).astype(int)
    w = raw[:, :, 0].copy()
    w[lvl >= 2] &= raw[lvl >= 2, :, 1]
    w[lvl >= 3] &= raw[lvl >= 3, :, 2]
    w[lvl == 0] |= raw[lvl == 0, :, 1]
    planes[:, :words] = w
    if n_hap & 63:
        planes[:, words - 1] &= np.uint64((1 << (n_hap & 63)) - 1)
    return planes


# ---------------------------------------------------------------------------------------------
# Large stores generated ON the GPU (benchmarks at chromosome / genome scale)
# ---------------------------------------------------------------------------------------------
GEN_BLOCK = 1 << 18        # rows per generator block: the seed depends on (chromosome, block) only
GEN_GROUP = 8              # neighbours sharing a base pattern


class _DevView:
    """A raw device address as a CUDA array: torch.as_tensor() wraps it without a copy."""

    def __init__(self, addr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (addr, False), "version": 3}


def fill_store_grouped(store, dev, chrom, row_begin, row_end):
    """Fill `store` (holding chromosome rows row_begin..row_end-1) on the GPU, straight into its planes.

    Variants come in groups of 8 neighbours sharing a base pattern (alt frequency 1/4) with 1/64 of the haplotypes
    flipped independently: neighbours are in strong LD (r2 ~ 0.85), everything else is not, so an r2 >= 0.8 filter keeps
    a few pairs per query, as on real data.  The generator is seeded per (chromosome, 2^18-row block): every rank of a
    sharded job sees the same genome whatever rows it holds."""
    import torch
    n_rows, stride, n_hap = row_end - row_begin, store.stride_words, store.n_hap
    planes = torch.as_tensor(_DevView(store.planes_ptr, n_rows * stride), device=dev).view(n_rows, stride)
    words = (n_hap + 63) // 64

    def rnd(n, g):      # n x stride uniformly random 64-bit words (two int32 draws per word)
        return torch.randint(-(1 << 31), 1 << 31, (n, 2 * stride), generator=g, device=dev, dtype=torch.int32).view(torch.int64)

    for b in range(row_begin // GEN_BLOCK, (row_end + GEN_BLOCK - 1) // GEN_BLOCK):
        g = torch.Generator(device=dev)
        g.manual_seed(1_000_003 * (chrom + 1) + b)
        base = rnd(GEN_BLOCK // GEN_GROUP, g) & rnd(GEN_BLOCK // GEN_GROUP, g)
        noise = rnd(GEN_BLOCK, g)
        for _ in range(5):
            noise &= rnd(GEN_BLOCK, g)
        rows = base.repeat_interleave(GEN_GROUP, dim=0) ^ noise
        rows[:, words:] = 0
        if n_hap & 63:
            rows[:, words - 1] &= (1 << (n_hap & 63)) - 1
        a, e = max(b * GEN_BLOCK, row_begin), min((b + 1) * GEN_BLOCK, row_end)
        planes[a - row_begin:e - row_begin] = rows[a - b * GEN_BLOCK:e - b * GEN_BLOCK]
        del base, noise, rows
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------
# 1000G-format files: <dir>/<chrom>.vcf.gz, integrated_call_samples panel, conversion.db
# ---------------------------------------------------------------------------------------------

def make_records(n_variants, chrom="22", start=16_050_000, mean_gap=33, seed=SEED):
    """Variant annotations with the oddities the drivers must survive:
    non-rs ids, MULTI_ALLELIC rows, INDELs with long REF, consecutive duplicate (pos, id) rows."""
    rng = np.random.default_rng(seed + 1)
    pos = start + np.cumsum(rng.geometric(1.0 / mean_gap, size=n_variants))
    recs = []
    bases = "ACGT"
    for i in range(n_variants):
        u = rng.random()
        rid, ref, alt, vt, multi = f"rs{1000 + 7 * i}", bases[i % 4], bases[(i + 1 + i // 4) % 4], "SNP", False
        if ref == alt:
            alt = bases[(bases.index(ref) + 2) % 4]
        if u < 0.03:
            rid = "." if u < 0.015 else f"esv{3000 + i}"
        elif u < 0.06:
            multi, alt, vt = True, alt + "," + bases[(bases.index(ref) + 3) % 4], "SNP"
        elif u < 0.12:
            ref, vt = ref + "".join(bases[(i + k) % 4] for k in range(int(rng.integers(1, 40)))), "INDEL"
        recs.append({"chrom": chrom, "pos": int(pos[i]), "id": rid, "ref": ref, "alt": alt, "vt": vt,
                     "multi": multi})
    # consecutive duplicates: same (pos, id) twice in a row (prep_intgen_data.py:170-175 drops both)
    for i in range(5, n_variants - 1, max(97, n_variants // 12)):
        if recs[i]["id"].startswith("rs") and not recs[i]["multi"]:
            recs[i + 1] = dict(recs[i], alt=recs[i]["alt"], vt="INDEL", ref=recs[i]["ref"] + "T")
    return recs


def gt_row_text(h_row):
    """0/1 haplotype vector (2 per sample) -> b'0|1\\t1|0...' without the trailing newline."""
    a = np.asarray(h_row, dtype=np.uint8).reshape(-1, 2)
    buf = np.empty((a.shape[0], 4), dtype=np.uint8)
    buf[:, 0] = a[:, 0] + 48
    buf[:, 1] = 124
    buf[:, 2] = a[:, 1] + 48
    buf[:, 3] = 9
    return buf.tobytes()[:-1]


def write_intgen_dir(path, panel, recs, haps, chrom="22", gt_text=None):
    """Write <path>/<chrom>.vcf.gz, the panel file and conversion.db as prep_intgen_data would
    have left them (prep_intgen_data.py:51-63 samples table, :146-182 variants table + index).
    gt_text (optional): per variant the bytes of its genotype columns (tab-joined sample fields) instead of the plain
    phased diploid text of `haps` -- haploid, missing, unphased ... fields."""
    os.makedirs(path, exist_ok=True)
    names = [p[0] for p in panel]
    with open(os.path.join(path, "integrated_call_samples_v3.20130502.ALL.panel"), "w") as fh:
        fh.write("sample\tpop\tsuper_pop\tgender\n")
        for row in panel:
            fh.write("\t".join(row) + "\n")
    with BgzfWriter(os.path.join(path, f"{chrom}.vcf.gz")) as fh:          # BGZF, like the real 1000G files (any gzip reader reads it)
        fh.write(b"##fileformat=VCFv4.1\n##source=ld_tools_b200.synth\n")
        fh.write(("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n").encode())
        written = []
        for r, h in zip(recs, haps):
            ac = int(h.sum())
            info = (r.get("info_prefix", "") + f"AC={ac};AF={ac / len(h):.6g};AN={len(h)};VT={r['vt']}" + (";MULTI_ALLELIC" if r["multi"] else "")
                    + r.get("info_suffix", ""))
            head = f"{r['chrom']}\t{r['pos']}\t{r['id']}\t{r['ref']}\t{r['alt']}\t100\tPASS\t{info}\tGT\t"
            fh.write(head.encode() + (gt_row_text(h) if gt_text is None else gt_text[len(written)]) + b"\n")
            written.append(1)
    db = os.path.join(path, "conversion.db")
    if os.path.exists(db):
        os.remove(db)
    with sqlite3.connect(db) as conn:
        cur = conn.cursor()
        cur.execute("CREATE TABLE samples (sample TEXT, pop TEXT, super_pop TEXT, gender TEXT)")
        cur.executemany("INSERT INTO samples VALUES (?, ?, ?, ?)", panel)
        cur.execute("CREATE TABLE variants (CHROM TEXT, POS INTEGER, ID TEXT)")
        rows = conversion_rows(recs)
        cur.executemany("INSERT INTO variants VALUES (?, ?, ?)", rows)
        cur.execute("CREATE INDEX id ON variants (ID)")
        conn.commit()
    return db


def conversion_rows(recs):
    """The rows prep_intgen_data.py:163-177 keeps: rs ids only, no MULTI_ALLELIC, and every row of a
    run of consecutively repeated (chrom, pos, id) removed."""
    import re
    keep = [(r["chrom"], r["pos"], r["id"]) for r in recs
            if re.match(r"rs\d+$", r["id"]) and not r["multi"]]
    out, i = [], 0
    while i < len(keep):
        j = i
        while j + 1 < len(keep) and keep[j + 1] == keep[i]:
            j += 1
        if j == i:
            out.append(keep[i])
        i = j + 1
    return out
