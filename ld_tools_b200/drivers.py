"""The three ld-tools drivers re-pointed at the GPU batch API.

What the reference does per pair -- two tabix fetches, 2 x 2504 `rec.samples[name]['GT']` lookups and
one pure-Python calc_ld call (ld_area.py:215-249, ld_triangle.py:133-230, ld_lite.py:109-144) -- is
replaced by: read each chromosome's VCF once, bit-pack its GT columns on the GPU (K1), and make ONE
library call per chromosome (K4 window scan / K5 all-pairs / K3 pairs).  Everything a user sees stays
as the reference writes it: directory and file names, UCSC-style headers, TSV / JSON / rsIDs / matrix
layouts, value formatting (int `0` vs float, round(x, 4)), sample selection by gender / population
through conversion.db.  The writers below restate the reference's formats and cite them; the
golden files under tests/golden/drivers/ were produced by the UNMODIFIED reference drivers
(tests/golden/make_driver_golden.py) and tests/test_drivers_gpu.py compares byte for byte.

There is no CPU path here: every LD number comes out of libldx.so.
"""
import gzip
import json
import os
import re
import sqlite3

import numpy as np

from .engine import Context, HostText, Store, dprime_value, r2_value, threshold_e4
from ._lib import DP_INT0, DP_MASK, DP_SHIFT, R2_INT0, R2_MASK, VCF_ROW_DTYPE, LdxError

R2_MASK_SAT = 16383                 # a 14-bit field of the packed word at its ceiling (general route only)

RS_RE = re.compile(r"rs\d+$")
TEXT_SLAB_BYTES = 256 << 20          # matrix text is formatted and written in slabs of at most this many bytes


# --------------------------------------------------------------------------- conversion.db helpers
def get_sample_names(gend_names, pop_names, convdb_path):
    """backend/get_sample_names.py:5-45, unchanged in behaviour (same SQL, same tuple quirks)."""
    query = f"SELECT sample FROM samples WHERE gender IN {tuple(gend_names)}"
    if tuple(pop_names) != ("ALL",):
        query += f" AND (super_pop IN {tuple(pop_names)} OR pop IN {tuple(pop_names)})"
    query = query.replace(",)", ")")
    with sqlite3.connect(convdb_path) as conn:
        cur = conn.cursor()
        names = [t[0] for t in cur.execute(query)]
        cur.close()
    return names


def create_src_dict(src_dir_path, src_file_name, meta_lines_quan, convdb_path):
    """backend/create_src_dict.py:5-64: leftmost rs\\d+ per line, de-duplicated, looked up in
    conversion.db, grouped by chromosome as [[pos, rsID], ...] in the database's answer order."""
    with open(os.path.join(src_dir_path, src_file_name)) as fh:
        for _ in range(meta_lines_quan):
            fh.readline()
        rs_ids = set()
        for line in fh:
            m = re.search(r"rs\d+\b", line)
            if m:
                rs_ids.add(m.group())
    if not rs_ids:
        return {}
    q = f"SELECT * FROM variants WHERE ID IN {tuple(rs_ids)}".replace(",)", ")")
    out = {}
    with sqlite3.connect(convdb_path) as conn:
        cur = conn.cursor()
        for chrom, pos, rs_id in cur.execute(q):
            out.setdefault(chrom, []).append([pos, rs_id])
        cur.close()
    return out


def _reference_helpers():
    """The reference's own backend/get_sample_names.py and backend/create_src_dict.py when its tree is importable (on sys.path,
    or named by LD_TOOLS_REFERENCE): these two stay as they are in the reference, so the drivers prefer
    the originals and keep the restatements above as the fallback.  -> (get_sample_names, create_src_dict)."""
    import importlib
    import sys
    ref = os.environ.get("LD_TOOLS_REFERENCE")
    added = False
    try:
        if ref and os.path.isfile(os.path.join(ref, "backend", "calc_ld.py")) and ref not in sys.path:
            sys.path.insert(0, ref)
            added = True
        g = importlib.import_module("backend.get_sample_names")
        c = importlib.import_module("backend.create_src_dict")
        here = os.path.dirname(os.path.abspath(g.__file__))
        if os.path.isfile(os.path.join(here, "calc_ld.py")) and os.path.dirname(os.path.abspath(c.__file__)) == here:
            return g.get_sample_names, c.create_src_dict
    except Exception:                                   # no such package, or somebody else's `backend`
        pass
    finally:
        if added:
            sys.path.remove(ref)
    return get_sample_names, create_src_dict


def gender_tuple(gend_names):
    """ld_area.py:47-52."""
    return ("male",) if gend_names == "male" else ("female",) if gend_names == "female" else ("male", "female")


# --------------------------------------------------------------------------- VCF ingest (one pass per chromosome)
class _Column:
    """One text column of the records (ID, REF, ALT or the VT key of INFO), sliced on demand from the fixed columns
    the ingest kept: only the rows a driver reports are ever decoded."""

    def __init__(self, cd, field, vt=False):
        self.cd, self.field, self.vt = cd, field, vt

    def __len__(self):
        return self.cd.rows.shape[0]

    def __getitem__(self, k):
        cd = self.cd
        if k < 0:
            k += len(self)
        a = int(cd._off[k]) + int(cd.rows[self.field][k])
        b = cd._blob.find(b"\t", a, int(cd._off[k + 1]))
        text = cd._blob[a:b if b >= 0 else int(cd._off[k + 1])].decode()
        if not self.vt:
            return text
        vt = [x[3:] for x in text.split(";") if x.startswith("VT=")]        # ld_area.py:233 rec.info['VT']
        return vt[0] if vt else ""

    def __iter__(self):
        return (self[k] for k in range(len(self)))


class ChromData:
    """One <chrom>.vcf.gz read once: annotations on the host, genotypes as a bit-plane store in HBM.

    The packed store is also kept on disk next to the VCF (<chrom>.vcf.gz.ldxstore: planes + window annotations,
    written by ldx_store_save; <chrom>.vcf.gz.ldxmeta.npz: the text columns the writers print), the way the
    reference keeps its conversion.db / tabix indices there (prep_intgen_data.py:138-182): later runs skip the
    inflate + parse + GPU packing and load 632 B per variant instead of 10 KB of text.  A cache whose recorded VCF
    size or mtime differs from the file's is rebuilt; LDX_NO_STORE_CACHE=1 (or cache=False) ignores it."""

    CACHE_VERSION = 3           # 3: record intervals honour INFO/END

    def __init__(self, ctx, vcf_path, cache=True):
        cache = cache and not os.environ.get("LDX_NO_STORE_CACHE")
        if not (cache and self._load_cache(ctx, vcf_path)):
            self._ingest(ctx, vcf_path)
            if cache:
                self._save_cache(vcf_path)
        self.n_variants, self.n_samples = len(self.pos), len(self.samples)
        # the longest interval among the rows a scan can report (an ineligible structural variant's END= may span megabases: it is
        # never a hit, so it must not widen every query's candidate range)
        span = (self.end0 - self.pos0)[self.rows["eligible"] != 0]
        self.max_ref_len = int(span.max()) if len(span) else 1
        self.col_of = {n: i for i, n in enumerate(self.samples)}
        # (pos, id) -> row is resolved on demand (row_of): a position lookup plus an ID comparison of the few records at that
        # position, so that a chromosome with millions of rows costs nothing up front.  Needs ascending positions (every tabix-
        # indexed VCF has them); a file without gets the exhaustive map.
        self._sorted = bool(self.n_variants < 2 or (np.diff(self.pos) >= 0).all())
        self._row_map = None

    # ---- cache
    @staticmethod
    def _cache_paths(vcf_path):
        return vcf_path + ".ldxstore", vcf_path + ".ldxmeta.npz"

    def _load_cache(self, ctx, vcf_path):
        store_path, meta_path = self._cache_paths(vcf_path)
        if not (os.path.exists(store_path) and os.path.exists(meta_path)):
            return False
        try:
            st = os.stat(vcf_path)
            with np.load(meta_path) as z:
                if [int(x) for x in z["source"]] != [self.CACHE_VERSION, st.st_size, st.st_mtime_ns]:
                    return False
                samples = bytes(z["samples"]).decode().split("\n") if len(z["samples"]) else []
                rows = np.frombuffer(bytes(z["rows"]), dtype=VCF_ROW_DTYPE).copy()
                blob, off = bytes(z["blob"]), z["off"].astype(np.int64)
            if off.shape[0] != rows.shape[0] + 1 or (rows.shape[0] and int(off[-1]) != len(blob)):
                return False
            self.store = Store.load(ctx, store_path)
        except (OSError, ValueError, KeyError, LdxError):
            return False
        if self.store.n_variants != rows.shape[0] or self.store.n_hap != 2 * len(samples):
            self.store.close()
            return False
        self._set_columns(samples, rows, blob, off)
        return True

    def _save_cache(self, vcf_path):
        store_path, meta_path = self._cache_paths(vcf_path)
        try:
            st = os.stat(vcf_path)
            tmp_store = store_path + f".tmp{os.getpid()}"
            self.store.save(tmp_store)
            os.replace(tmp_store, store_path)                # readers see the old file or the whole new one, never half of it
            tmp = meta_path + f".tmp{os.getpid()}.npz"
            np.savez(tmp, source=np.array([self.CACHE_VERSION, st.st_size, st.st_mtime_ns], dtype=np.int64),
                     samples=np.frombuffer("\n".join(self.samples).encode(), dtype=np.uint8),
                     rows=np.frombuffer(self.rows.tobytes(), dtype=np.uint8), blob=np.frombuffer(self._blob, dtype=np.uint8),
                     off=self._off)
            os.replace(tmp, meta_path)                     # the meta file appears last and atomically: it validates the pair
        except (OSError, LdxError):
            pass                                           # a read-only data directory: work without the cache

    def _set_columns(self, samples, rows, blob, off):
        """rows: one VCF_ROW_DTYPE record per variant (parsed on the GPU); blob/off: the records' fixed columns."""
        self.samples, self.rows, self._blob, self._off = samples, rows, blob, off
        self._blob_arr = np.frombuffer(blob, dtype=np.uint8) if len(blob) else np.zeros(1, dtype=np.uint8)
        self.pos = rows["pos"].astype(np.int64)
        self.pos0 = (self.pos - 1).astype(np.int32)
        self.end0 = (self.pos0 + rows["ref_len"]).astype(np.int32)
        self.multi = rows["multi"].astype(bool)
        self.ids, self.refs, self.alts = _Column(self, "id_off"), _Column(self, "ref_off"), _Column(self, "alt_off")
        self.vts = _Column(self, "info_off", vt=True)

    # ---- first run: the VCF itself.  The host only inflates the file and reads the #CHROM line; splitting lines and
    #      fields, parsing POS / ID / REF / INFO and packing the genotypes all happen on the GPU, one slab of text at a time
    #      (ldx_store_ingest_vcf_file: a 1000 Genomes chromosome is ~65 GB of text for a 4 GB store).
    SLAB_BYTES = int(os.environ.get("LDX_INGEST_SLAB_MB", "256")) << 20

    @staticmethod
    def _header_samples(vcf_path):
        with gzip.open(vcf_path, "rb") as fh:              # the meta lines and the #CHROM line
            head = fh.read(1 << 24)
        h = 0 if head.startswith(b"#CHROM") else head.find(b"\n#CHROM") + 1
        if h == 0 and not head.startswith(b"#CHROM"):
            raise ValueError(f"{vcf_path}: no #CHROM header line in the first 16 MiB")
        e = head.find(b"\n", h)
        samples = head[h:e if e >= 0 else len(head)].decode().rstrip("\r").split("\t")[9:]
        if not samples:
            raise ValueError(f"{vcf_path}: no sample columns")
        return samples

    def _ingest(self, ctx, vcf_path):
        samples = self._header_samples(vcf_path)
        self.store, rows, blob, off, _ = Store.ingest_vcf_file(ctx, vcf_path, len(samples), slab_bytes=self.SLAB_BYTES)
        # status bit 0 (a genotype field that is not plain "a|b": haploid, missing, other codes, unphased) is fine: such rows
        # take the engine's general route.  What no parser takes -- a line that is not a record, a POS that is not a number,
        # a field with more than two alleles -- is refused by name.
        bad = np.flatnonzero((rows["status"] & 14) != 0)
        if len(bad):
            k = int(bad[0])
            line = blob[off[k]:off[k] + 60].tobytes().decode(errors="replace")
            self.store.close()
            raise ValueError(f"{vcf_path}: record {k} ({line!r}...) is not a VCF record this engine can read "
                             f"(status {int(rows['status'][k])}: 2 = too few columns, 4 = POS, 8 = genotype fields)")
        self._set_columns(samples, rows, blob.tobytes(), off)

    def select_samples(self, sample_names):
        """The haplotypes of the chosen samples; names absent from the VCF are skipped like the reference's
        `except KeyError: continue` (ld_area.py:184-187).  When they are at most half of the store's columns the selected
        columns are gathered into a narrower store once (ldx_store_subset: e.g. 1006 of 5008 haplotypes -> 128-byte rows
        instead of 640-byte rows under a mask) and every scan of this run reads that: `self.scan`."""
        cols = np.array([self.col_of[n] for n in sample_names if n in self.col_of], dtype=np.int64)
        hap = np.sort(np.concatenate([2 * cols, 2 * cols + 1]))
        if getattr(self, "scan", None) is not None and self.scan is not self.store:
            self.scan.close()
        self.store.select_haplotypes(hap)
        self.scan = self.store
        if 0 < len(hap) <= self.store.n_hap // 2 and not os.environ.get("LDX_NO_SUBSET_STORE"):
            self.scan = self.store.subset(hap)
        self.n1, self.p_e4, self.n_hap_sel = self.scan.counts()
        # rows of the general route (missing calls, haploid samples ...): their own list lengths, for the alt frequencies calc_ld
        # reports per PAIR (var_2_alt_freq = n1 / len(zip(...)), calc_ld.py:31,43,97)
        self.row_n1, self.row_len, self.kind, self.n_general = self.scan.row_counts()

    def pair_values_e4(self, row_a, row_b):
        """(round(r2, 4), round(D', 4)) * 10^4 of calc_ld(var_1 = row_a, var_2 = row_b) from the engine's unrounded values, -1 for
        the reference's int 0: for the pairs whose values do not fit the packed word's 14-bit fields (more than 1.6383: a
        pairing of lists of unequal ploidy, e.g. a pseudo-autosomal with an X-specific variant)."""
        out = self.scan.pairs([row_a], [row_b], raw=True)
        w = int(out["packed"][0])
        e4 = lambda x: int(round(round(float(x), 4) * 10000))                     # noqa: E731
        return (-1 if w & R2_INT0 else e4(out["r2"][0])), (-1 if w & DP_INT0 else e4(out["dprime"][0]))

    def pair_alt_e4(self, row_a, row_b):
        """round(var_2_alt_freq, 4) * 10^4 of calc_ld(var_1 = row_a, var_2 = row_b): the alt count of b over the PAIRING's length."""
        n = min(int(self.row_len[row_a]), int(self.row_len[row_b]))
        return int(round(round(int(self.row_n1[row_b]) / n, 4) * 10000)) if n else 0

    def row_of(self, pos, rs_id):
        """Store row of the first record with this position and ID (the drivers `break` at the first match, ld_area.py:150-159)."""
        pos = int(pos)
        if self._sorted:
            k = int(np.searchsorted(self.pos, pos, side="left"))
            while k < self.n_variants and self.pos[k] == pos:
                if self.ids[k] == rs_id:
                    return k
                k += 1
            raise KeyError((pos, rs_id))
        if self._row_map is None:
            self._row_map = {}
            for k, (p, i) in enumerate(zip(self.pos.tolist(), self.ids)):
                self._row_map.setdefault((p, i), k)
        return self._row_map[(pos, rs_id)]

    def close(self):
        if getattr(self, "scan", None) is not None and self.scan is not self.store:
            self.scan.close()
        self.scan = None
        self.store.close()


def _ucsc(key, val):
    """ld_area.py:3-14 build_ucsc_header."""
    if isinstance(val, str):
        val = f'"{val}"'
    elif isinstance(val, tuple):
        val = ",".join(f'"{v}"' for v in val)
    return f"{key}={val}"


# --------------------------------------------------------------------------- fan-out over devices
def _device_list(devices):
    """None -> the current device; an int N -> devices 0..N-1; else the list itself."""
    if devices is None:
        return [-1]
    if isinstance(devices, int):
        return list(range(devices))
    return list(devices)


class _Worker:
    """One device of a job: its context and the chromosomes it has loaded.  The reference fans its jobs out with
    multiprocessing.Pool over source files (ld_area.py:320-342, ld_triangle.py:390-411); here the unit is finer -- a
    (source file, chromosome) table, or a slab of one table's queries / matrix rows -- and the workers are threads, one per
    GPU (the library calls release the GIL)."""

    def __init__(self, device, ctx, intgen_dir_path, sample_names):
        self.device, self.ctx, self.own = device, ctx, ctx is None
        self.intgen_dir_path, self.sample_names = intgen_dir_path, sample_names
        self.chroms = {}

    def chrom(self, chrom):
        if self.ctx is None:
            self.ctx = Context(self.device)
        if chrom not in self.chroms:
            cd = self.chroms[chrom] = ChromData(self.ctx, os.path.join(self.intgen_dir_path, f"{chrom}.vcf.gz"))
            cd.select_samples(self.sample_names)
        return self.chroms[chrom]

    def close(self):
        for cd in self.chroms.values():
            cd.close()
        if self.own and self.ctx is not None:
            self.ctx.close()


def _run_on_devices(workers, jobs_per_worker, fn):
    """fn(worker, job) for every job of every worker, one thread per worker; the first exception is re-raised."""
    import threading
    errors = []

    def run(w, jobs):
        try:
            for job in jobs:
                fn(w, job)
        except BaseException as e:      # noqa: BLE001 -- re-raised in the caller's thread
            errors.append(e)
    if len(workers) == 1:
        run(workers[0], jobs_per_worker[0])
    else:
        ths = [threading.Thread(target=run, args=(w, j)) for w, j in zip(workers, jobs_per_worker)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    if errors:
        raise errors[0]


# --------------------------------------------------------------------------- ld_area
def ld_area(src_dir_path, intgen_dir_path, trg_top_dir_path=None, meta_lines_quan=0, gend_names="both", pop_names="all",
            flank_size=100000, ld_thres_measure="r_square", ld_low_thres=0.8, trg_file_type="tsv", ctx=None, devices=None):
    """ld_area.py as a function: same arguments as its CLI (cli/ld_area_cli_en.py:36-60), same output
    tree (<src>_in_LD/<chrom>/<rsID>_chr<chrom>_<m>_<thres>.<ext>, ld_area.py:82-84,160).
    devices: the GPUs of the job (None = the current one, N = the first N, or a list).  With several, the (source file,
    chromosome) tables are dealt to them by chromosome, and a lone table's queries are cut into slabs balanced by candidate
    pairs (region sharding: every query's file is written by the device that scanned it, nothing is exchanged)."""
    from . import shard
    from ._lib import AREA_JSON, AREA_RSIDS, AREA_TSV
    from .engine import area_format
    src_dir_path, intgen_dir_path = os.path.normpath(src_dir_path), os.path.normpath(intgen_dir_path)
    trg_top = src_dir_path if trg_top_dir_path is None else os.path.normpath(trg_top_dir_path)
    convdb = os.path.join(intgen_dir_path, "conversion.db")
    gends, pops = gender_tuple(gend_names), tuple(pop_names.upper().split(","))
    ref_sample_names, ref_src_dict = _reference_helpers()
    sample_names = ref_sample_names(gends, pops, convdb)
    ext = trg_file_type if trg_file_type in ("tsv", "json") else "txt"
    fmt = {"tsv": AREA_TSV, "json": AREA_JSON}.get(trg_file_type, AREA_RSIDS)
    meta_keys = ["chr", "gends", "pops", "each_flank", f"{ld_thres_measure}_thres"]
    header_row = ["hg38_pos", "rsID", "ref", "alt", "type", "alt_freq", "r2", "D'", "dist"]
    t_e4 = threshold_e4(ld_low_thres)
    devs = _device_list(devices)
    workers = [_Worker(d, ctx if k == 0 else None, intgen_dir_path, sample_names) for k, d in enumerate(devs)]
    # ---- the tables of the job (host work: the source files and conversion.db), directories made as the reference makes them
    tables = []
    for src_file_name in os.listdir(src_dir_path):
        data_by_chrs = ref_src_dict(src_dir_path, src_file_name, meta_lines_quan, convdb)
        trg_dir = os.path.join(trg_top, f"{src_file_name.rsplit('.', maxsplit=1)[0]}_in_LD")
        for chrom, var_rows in data_by_chrs.items():
            chr_dir = os.path.join(trg_dir, chrom)
            os.makedirs(chr_dir)                                             # ld_area.py:123 (not exist_ok)
            tables.append({"chrom": chrom, "var_rows": var_rows, "chr_dir": chr_dir})
    chrom_order = {c: k for k, c in enumerate(dict.fromkeys(t["chrom"] for t in tables))}
    jobs = [[] for _ in workers]
    if len(workers) > 1 and len(chrom_order) == 1 and len(tables) == 1:
        for k in range(len(workers)):                                        # one table: its queries in slabs, one per device
            jobs[k].append(dict(tables[0], slab=(k, len(workers))))
    else:
        for t in tables:                                                     # a chromosome's store lives on one device
            jobs[chrom_order[t["chrom"]] % len(workers)].append(t)

    def run_table(w, t):
        chrom, var_rows, chr_dir = t["chrom"], t["var_rows"], t["chr_dir"]
        cd = w.chrom(chrom)
        meta_vals = [chrom, gends, pops, flank_size, ld_low_thres]
        ucsc = "##" + " ".join(map(_ucsc, meta_keys, meta_vals))
        # ---- every query of the table in ONE window scan (ld_area.py:152-249)
        q_row = np.array([cd.row_of(p, i) for p, i in var_rows], dtype=np.int64)
        q_pos = cd.pos[q_row]
        ws = np.maximum(q_pos - flank_size, 0)                        # :174-176
        we = q_pos + flank_size                                       # :177
        lo = np.searchsorted(cd.pos0, ws - cd.max_ref_len, side="right")
        hi = np.maximum(np.searchsorted(cd.pos0, we, side="left"), lo)
        mine = np.arange(len(q_row))
        if "slab" in t:                                               # this device's share of the queries, by candidate pairs
            k, n = t["slab"]
            order = np.argsort(q_row, kind="stable")
            cum = np.concatenate([[0], np.cumsum((hi - lo)[order])])
            cuts = np.searchsorted(cum, cum[-1] * np.arange(n + 1) / n, side="left")
            cuts[0], cuts[-1] = 0, len(order)
            mine = np.sort(order[cuts[k]:max(cuts[k + 1], cuts[k])])
            if not len(mine):
                return
        hits, _ = cd.scan.window(q_row[mine], lo[mine], hi[mine], ws[mine], we[mine], ld_thres_measure, t_e4)
        # ---- the writers (:200-283): the rows of every query from the library, headers and the query's own line from here
        overrides = None
        if cd.n_general and len(hits):
            # the general route: var_2_alt_freq belongs to the PAIR where one list is shorter (calc_ld.py:31,43), and such a
            # pairing can produce values beyond the packed word's 1.6383: those come from the engine's unrounded values
            overrides = np.full((len(hits), 3), -1, dtype=np.int32)
            qr_of_hit = q_row[mine][hits["query"]]
            for i in np.flatnonzero(cd.row_len[hits["row"]] != cd.row_len[qr_of_hit]):
                overrides[i, 0] = cd.pair_alt_e4(int(qr_of_hit[i]), int(hits["row"][i]))
            sat = ((hits["packed"] & R2_MASK) == R2_MASK_SAT) | (((hits["packed"] & DP_MASK) >> DP_SHIFT) == R2_MASK_SAT)
            for i in np.flatnonzero(sat):
                overrides[i, 1], overrides[i, 2] = cd.pair_values_e4(int(qr_of_hit[i]), int(hits["row"][i]))
        text, qoff = area_format(w.ctx._lib, hits, q_row[mine], cd._blob_arr, cd._off, cd.rows, cd.p_e4, fmt, overrides=overrides)
        for j, k in enumerate(mine):
            if qoff[j + 1] == qoff[j]:
                continue                                             # empty result: file removed, :291-292
            qr = int(q_row[k])
            q_id = cd.ids[qr]
            q_ann = [int(cd.pos[qr]), q_id, cd.refs[qr], cd.alts[qr], cd.vts[qr], cd.p_e4[qr] / 10000.0] + ["quer"] * 3
            path = os.path.join(chr_dir, f"{q_id}_chr{chrom}_{ld_thres_measure[0]}_{str(ld_low_thres)}.{ext}")
            body = text[qoff[j]:qoff[j + 1]].data
            with open(path, "wb") as fh:
                if trg_file_type == "rsids":                          # :201-204, :258-260
                    fh.write((ucsc + "\n#rsID\n" + q_id + "\n").encode())
                    fh.write(body)
                elif trg_file_type == "tsv":                          # :205-208, :273-274
                    fh.write((ucsc + "\n#" + "\t".join(header_row) + "\n" + "\t".join(map(str, q_ann)) + "\n").encode())
                    fh.write(body)
                else:                                                 # :209-211, :275-283
                    head = json.dumps([dict(zip(meta_keys, meta_vals)), dict(zip(header_row, q_ann))], indent=4)
                    fh.write(head[:-2].encode())                      # ... without the closing "\n]"
                    fh.write(body)
                    fh.write(b"\n]")
    try:
        _run_on_devices(workers, jobs, run_table)
    finally:
        for w in workers:
            w.close()


# --------------------------------------------------------------------------- ld_triangle (table output)
BATCH_MAX_VARIANTS = 8192            # matrices up to this size go into one batched launch per device (ldx_triangle_batch_dev)
BATCH_MAX_BYTES = 4 << 30            # ... as long as their packed words fit this much device memory


def ld_triangle(src_dir_path, intgen_dir_path, trg_top_dir_path=None, meta_lines_quan=0, gend_names="both", pop_names="all",
                ld_measure="r_square", ld_low_thres=None, ctx=None, devices=None):
    """ld_triangle.py -o table as a function (cli/ld_triangle_cli_en.py:40-74).  The heatmap outputs are
    Plotly rendering, out of scope; the matrix they draw is this one.
    devices: as for ld_area.  The (source file, chromosome) matrices are dealt to the devices by chromosome; on each device
    the small ones are computed by ONE batched launch (a 2,000-variant matrix alone is a single wave of tiles) and a large
    one slab by slab; a lone large matrix is cut into row slabs that the devices take in turn while the file is written in
    order."""
    src_dir_path, intgen_dir_path = os.path.normpath(src_dir_path), os.path.normpath(intgen_dir_path)
    trg_top = src_dir_path if trg_top_dir_path is None else os.path.normpath(trg_top_dir_path)
    convdb = os.path.join(intgen_dir_path, "conversion.db")
    gends, pops = gender_tuple(gend_names), tuple(pop_names.upper().split(","))
    ref_sample_names, ref_src_dict = _reference_helpers()
    sample_names = ref_sample_names(gends, pops, convdb)
    t_e4 = None if ld_low_thres is None else threshold_e4(ld_low_thres)
    devs = _device_list(devices)
    workers = [_Worker(d, ctx if k == 0 else None, intgen_dir_path, sample_names) for k, d in enumerate(devs)]
    tab = "\t"
    tables = []
    for src_file_name in os.listdir(src_dir_path):
        data_by_chrs = ref_src_dict(src_dir_path, src_file_name, meta_lines_quan, convdb)
        base = src_file_name.rsplit(".", maxsplit=1)[0]
        trg_dir = os.path.join(trg_top, f"{base}_LD_matr")
        for chrom, var_rows in data_by_chrs.items():
            if len(var_rows) < 2:                                        # ld_triangle.py:80
                continue
            os.makedirs(trg_dir, exist_ok=True)
            var_rows.sort(key=lambda row: row[0])                        # :88 (stable)
            tables.append({"chrom": chrom, "var_rows": var_rows, "path": os.path.join(trg_dir, f"{base}_chr{chrom}_{ld_measure[0]}.tsv")})
    chrom_order = {c: k for k, c in enumerate(dict.fromkeys(t["chrom"] for t in tables))}

    def prepare(w, t):
        cd = w.chrom(t["chrom"])
        poss = [str(p) for p, _ in t["var_rows"]]
        ids = [i for _, i in t["var_rows"]]
        t["v"] = len(ids)
        t["rows"] = np.array([cd.row_of(p, i) for p, i in t["var_rows"]], dtype=np.int64)
        t["head"] = (f"##General\tinfo:\t{ld_measure}\tchr{t['chrom']}\t{tab.join(pops)}\t{tab.join(gends)}\n\n"
                     + "rsIDs\t\t" + "\t".join(ids) + "\n" + "\tPositions\t" + "\t".join(poss) + "\n").encode()       # :351-355
        t["prefixes"] = [(i + "\t" + p + "\t").encode() for i, p in zip(ids, poss)]
        return cd

    def slab_rows(v):
        return min(v, max(256, TEXT_SLAB_BYTES // (7 * v) // 256 * 256))

    def exact_cells(cd, t, text, row_begin):
        """Stores with general-route rows only: a cell printed as 1.6383 is a packed field at its ceiling (a pairing of lists of
        unequal ploidy); its value comes from the engine's unrounded numbers instead.  -> the text, patched (bytes)."""
        text = bytes(text)
        if not cd.n_general or b"1.6383" not in text:
            return text
        lines = text.split(b"\n")
        for k, line in enumerate(lines[:-1]):
            if b"1.6383" not in line:
                continue
            cells = line.split(b"\t")
            for c in range(2, len(cells)):
                if cells[c] == b"1.6383":
                    r2v, dpv = cd.pair_values_e4(int(t["rows"][row_begin + k]), int(t["rows"][c - 2]))      # var_1 = row, var_2 = column
                    val = r2v if ld_measure == "r_square" else dpv
                    cells[c] = str(val / 10000.0).encode() if val >= 0 else b"0"
            lines[k] = b"\t".join(cells)
        return b"\n".join(lines)

    def write_slabbed(w, t, cd):
        """The double loop (:133-230) and the V lines of V cells (:356-360): all-pairs kernel, settlement and the writer in one
        library call per slab of rows; only text leaves the GPU."""
        v, slab = t["v"], slab_rows(t["v"])
        buf = np.empty(7 * v * slab + sum(map(len, t["prefixes"])), dtype=np.uint8)        # one buffer for every slab
        with open(t["path"], "wb") as fh:
            fh.write(t["head"])
            for r0 in range(0, v, slab):
                fh.write(exact_cells(cd, t, cd.scan.triangle_table(t["rows"], t["prefixes"], ld_measure, t_e4, row_begin=r0, row_end=min(v, r0 + slab), out=buf).data, r0))

    def run_device(w, mine):
        """A device's tables: the small matrices in batched launches, the large ones slab by slab."""
        small, group, group_bytes = [], [], 0
        for t in mine:
            cd = prepare(w, t)
            if t["v"] > BATCH_MAX_VARIANTS:
                write_slabbed(w, t, cd)
                continue
            nbytes = 4 * (t["v"] * (t["v"] - 1) // 2)
            if group and group_bytes + nbytes > BATCH_MAX_BYTES:
                small.append(group)
                group, group_bytes = [], 0
            group.append((t, cd))
            group_bytes += nbytes
        if group:
            small.append(group)
        for group in small:
            if len(group) == 1:                                          # nothing to batch with: the one-call path (direct mode applies)
                write_slabbed(w, group[0][0], group[0][1])
                continue
            addrs = [w.ctx.dev_alloc(4 * (t["v"] * (t["v"] - 1) // 2)) for t, _ in group]
            try:
                w.ctx.triangle_batch_dev([(cd.scan, t["rows"], a) for (t, cd), a in zip(group, addrs)], measure=ld_measure, thres_e4_=t_e4)
                w.ctx.resolve()
                for (t, cd), a in zip(group, addrs):
                    with open(t["path"], "wb") as fh:
                        fh.write(t["head"])
                        fh.write(exact_cells(cd, t, w.ctx.triangle_text(a, t["v"], ld_measure, t["prefixes"]).data, 0))
            finally:
                for a in addrs:
                    w.ctx.dev_free(a)

    def run_shared(ws, t):
        """One large matrix over several devices: 256-aligned row slabs balanced by pair count, taken by the devices in turn;
        every device formats its slabs' text and the file is written in slab order."""
        import queue
        import threading
        cds = [prepare(w, t) for w in ws]
        v = t["v"]
        from . import shard
        n_slabs = max(len(ws), min(4 * len(ws), (v + 255) // 256))
        ranges = [r for r in shard.triangle_row_ranges(v, n_slabs) if r[1] > r[0]]
        done = [queue.Queue(maxsize=2) for _ in ws]
        errors = []

        def produce(k):
            try:
                for j in range(k, len(ranges), len(ws)):
                    r0, r1 = ranges[j]
                    done[k].put(exact_cells(cds[k], t, cds[k].scan.triangle_table(t["rows"], t["prefixes"], ld_measure, t_e4, row_begin=r0, row_end=r1).data, r0))
            except BaseException as e:      # noqa: BLE001
                errors.append(e)
                done[k].put(None)
        ths = [threading.Thread(target=produce, args=(k,)) for k in range(len(ws))]
        for th in ths:
            th.start()
        with open(t["path"], "wb") as fh:
            fh.write(t["head"])
            for j in range(len(ranges)):
                text = done[j % len(ws)].get()
                if text is None:
                    break
                fh.write(text)
        for th in ths:
            th.join()
        if errors:
            raise errors[0]

    try:
        if len(workers) > 1 and len(tables) == 1 and len(tables[0]["var_rows"]) > BATCH_MAX_VARIANTS:
            run_shared(workers, tables[0])
        else:
            jobs = [[] for _ in workers]
            for t in tables:
                jobs[chrom_order[t["chrom"]] % len(workers)].append(t)
            _run_on_devices(workers, [[j] for j in jobs], run_device)
    finally:
        for w in workers:
            w.close()


# --------------------------------------------------------------------------- ld_lite
def ld_lite(rs_id_1, rs_id_2, intgen_dir_path, gend_names="both", pop_names="all", ctx=None):
    """ld_lite.py as a function: returns the text the reference prints (ld_lite.py:148-159)."""
    from tabulate import tabulate
    own = ctx is None
    ctx = ctx or Context()
    intgen_dir_path = os.path.normpath(intgen_dir_path)
    convdb = os.path.join(intgen_dir_path, "conversion.db")
    info = []
    with sqlite3.connect(convdb) as conn:                                     # ld_lite.py:33-45 check_rs_id
        cur = conn.cursor()
        for rs in (rs_id_1, rs_id_2):
            row = cur.execute("SELECT CHROM, POS FROM variants WHERE ID = ?", (rs,)).fetchone()
            if row is None:
                raise KeyError(f"{rs} is not a biallelic rs variant of the 1000 Genomes cache")
            info.append(row)
        cur.close()
    if info[0][0] != info[1][0]:                                              # :96-97
        raise ValueError(f"{rs_id_1} and {rs_id_2} belong to different chromosomes")
    chrom, pos1, pos2 = info[0][0], info[0][1], info[1][1]
    cd = ChromData(ctx, os.path.join(intgen_dir_path, f"{chrom}.vcf.gz"))
    try:
        cd.select_samples(_reference_helpers()[0](gender_tuple(gend_names), tuple(pop_names.upper().split(",")), convdb))
        r1, r2 = cd.row_of(pos1, rs_id_1), cd.row_of(pos2, rs_id_2)
        out = cd.scan.pairs([r1], [r2], raw=False)
        w = out["packed"][0]
        r2_e4v, dp_e4v = (int(w) & R2_MASK), ((int(w) & DP_MASK) >> DP_SHIFT)
        if cd.n_general and R2_MASK_SAT in (r2_e4v, dp_e4v):          # beyond the packed fields: from the unrounded values
            r2_e4v, dp_e4v = cd.pair_values_e4(r1, r2)
        r2_obj = 0 if (int(w) & R2_INT0) else r2_e4v / 10000.0
        dp_obj = 0 if (int(w) & DP_INT0) else dp_e4v / 10000.0
        first_alt = lambda r: cd.alts[r].split(",")[0]                        # noqa: E731  (intgen_rec.alts[0], :116)
        return tabulate([["chrom", chrom, chrom], ["hg38_pos", pos1, pos2],
                         ["alleles", cd.refs[r1] + "/" + first_alt(r1), cd.refs[r2] + "/" + first_alt(r2)],
                         ["type", cd.vts[r1].split(",")[0], cd.vts[r2].split(",")[0]],      # intgen_rec.info['VT'][0], ld_lite.py:118,131
                         ["alt_freq", cd.pair_alt_e4(r2, r1) / 10000.0, cd.pair_alt_e4(r1, r2) / 10000.0]],      # var_1 / var_2_alt_freq of the pair
                        headers=[tabulate([["r2", r2_obj], ["D'", dp_obj], ["abs_dist", abs(pos1 - pos2)]],
                                          tablefmt="fancy_grid", disable_numparse=True),
                                 f"\n\n\n{rs_id_1}", f"\n\n\n{rs_id_2}"], tablefmt="fancy_grid")
    finally:
        cd.close()
        if own:
            ctx.close()
