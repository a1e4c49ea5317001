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

from .engine import Context, Store, dprime_value, measure_value, r2_value, threshold_e4
from ._lib import BELOW_THRES, LdxError

RS_RE = re.compile(r"rs\d+$")


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


def gender_tuple(gend_names):
    """ld_area.py:47-52."""
    return ("male",) if gend_names == "male" else ("female",) if gend_names == "female" else ("male", "female")


# --------------------------------------------------------------------------- VCF ingest (one pass per chromosome)
class ChromData:
    """One <chrom>.vcf.gz read once: annotations on the host, genotypes as a bit-plane store in HBM.

    The packed store is also kept on disk next to the VCF (<chrom>.vcf.gz.ldxstore: planes + window annotations,
    written by ldx_store_save; <chrom>.vcf.gz.ldxmeta.npz: the text columns the writers print), the way the
    reference keeps its conversion.db / tabix indices there (prep_intgen_data.py:138-182): later runs skip the
    inflate + parse + GPU packing and load 632 B per variant instead of 10 KB of text.  A cache whose recorded VCF
    size or mtime differs from the file's is rebuilt; LDX_NO_STORE_CACHE=1 (or cache=False) ignores it."""

    CACHE_VERSION = 1

    def __init__(self, ctx, vcf_path, cache=True):
        cache = cache and not os.environ.get("LDX_NO_STORE_CACHE")
        if not (cache and self._load_cache(ctx, vcf_path)):
            self._ingest(ctx, vcf_path)
            if cache:
                self._save_cache(vcf_path)
        self.n_variants, self.n_samples = len(self.pos), len(self.samples)
        self.max_ref_len = int((self.end0 - self.pos0).max()) if self.n_variants else 1
        self.col_of = {n: i for i, n in enumerate(self.samples)}
        self._row_of = {}
        for k, (p, i) in enumerate(zip(self.pos.tolist(), self.ids)):
            self._row_of.setdefault((p, i), k)            # first record wins, as the drivers' `break` does

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
                cols = {k: bytes(z[k]).decode().split("\n") if len(z[k]) else [] for k in ("samples", "ids", "refs", "alts", "vts")}
                self.pos = z["pos"].astype(np.int64)
                self.multi = z["multi"].astype(bool).tolist()
            self.samples, self.ids, self.refs, self.alts, self.vts = (cols[k] for k in ("samples", "ids", "refs", "alts", "vts"))
            if not (len(self.ids) == len(self.refs) == len(self.alts) == len(self.vts) == len(self.multi) == len(self.pos)):
                return False
            self.store = Store.load(ctx, store_path)
        except (OSError, ValueError, KeyError, LdxError):
            return False
        if self.store.n_variants != len(self.pos) or self.store.n_hap != 2 * len(self.samples):
            self.store.close()
            return False
        self.pos0 = (self.pos - 1).astype(np.int32)
        self.end0 = (self.pos0 + np.asarray([len(r) for r in self.refs], dtype=np.int32)).astype(np.int32)
        return True

    def _save_cache(self, vcf_path):
        store_path, meta_path = self._cache_paths(vcf_path)
        try:
            st = os.stat(vcf_path)
            self.store.save(store_path)
            blob = {k: np.frombuffer("\n".join(v).encode(), dtype=np.uint8)
                    for k, v in (("samples", self.samples), ("ids", self.ids), ("refs", self.refs), ("alts", self.alts), ("vts", self.vts))}
            tmp = meta_path + ".tmp.npz"
            np.savez(tmp, source=np.array([self.CACHE_VERSION, st.st_size, st.st_mtime_ns], dtype=np.int64), pos=self.pos,
                     multi=np.asarray(self.multi, dtype=np.uint8), **blob)
            os.replace(tmp, meta_path)                     # the meta file appears last and atomically: it validates the pair
        except (OSError, LdxError):
            pass                                           # a read-only data directory: work without the cache

    # ---- first run: the VCF itself
    def _ingest(self, ctx, vcf_path):
        with gzip.open(vcf_path, "rb") as fh:
            raw = fh.read()
        buf = np.frombuffer(raw, dtype=np.uint8)
        nl = np.flatnonzero(buf == 10)
        starts = np.concatenate([[0], nl[:-1] + 1]) if len(nl) else np.zeros(0, np.int64)
        self.samples, self.pos, self.ids, self.refs, self.alts, self.vts, self.multi = [], [], [], [], [], [], []
        gt_off = []
        for s, e in zip(starts.tolist(), nl.tolist()):
            if raw[s:s + 2] == b"##":
                continue
            if raw[s:s + 1] == b"#":
                self.samples = raw[s:e].decode().split("\t")[9:]
                continue
            # the nine fixed columns; the GT columns stay bytes for the GPU packer
            p, fields = s, []
            for _ in range(9):
                q = raw.index(b"\t", p, e)
                fields.append(raw[p:q])
                p = q + 1
            info = fields[7].decode().split(";")
            self.pos.append(int(fields[1]))
            self.ids.append(fields[2].decode())
            self.refs.append(fields[3].decode())
            self.alts.append(fields[4].decode())
            vt = [x[3:] for x in info if x.startswith("VT=")]
            self.vts.append(vt[0] if vt else "")
            self.multi.append("MULTI_ALLELIC" in info)
            gt_off.append(p)
        n_variants, n_samples = len(self.pos), len(self.samples)
        self.pos = np.asarray(self.pos, dtype=np.int64)
        self.pos0 = (self.pos - 1).astype(np.int32)
        self.end0 = (self.pos0 + np.asarray([len(r) for r in self.refs], dtype=np.int32)).astype(np.int32)
        elig = np.array([bool(RS_RE.match(i)) and not m for i, m in zip(self.ids, self.multi)], dtype=np.uint8)
        # same-id test of ld_area.py:222 on integers: rs number, or a unique negative for non-rs ids
        idnum = np.array([int(i[2:]) if RS_RE.match(i) else -1 - k for k, i in enumerate(self.ids)], dtype=np.int64)
        self.store = Store(ctx, n_variants, 2 * n_samples)
        status = self.store.pack_gt(0, buf, n_samples, row_off=np.asarray(gt_off, dtype=np.int64))
        if (status.astype(bool) & elig.astype(bool)).any():              # rows no driver ever pairs may be anything
            bad = int(np.flatnonzero(status.astype(bool) & elig.astype(bool))[0])
            self.store.close()
            raise ValueError(f"{vcf_path}: record {self.ids[bad]} is not phased diploid biallelic (chrX/Y and "
                             "missing calls are outside the engine's domain, reference README.md:72)")
        self.store.set_annotations(self.pos0, self.end0, idnum, elig)

    def select_samples(self, sample_names):
        """The mask plane of the chosen samples; names absent from the VCF are skipped like the
        reference's `except KeyError: continue` (ld_area.py:184-187)."""
        cols = np.array([self.col_of[n] for n in sample_names if n in self.col_of], dtype=np.int64)
        self.store.select_haplotypes(np.concatenate([2 * cols, 2 * cols + 1]))
        self.n1, self.p_e4, self.n_hap_sel = self.store.counts()

    def row_of(self, pos, rs_id):
        return self._row_of[(int(pos), rs_id)]

    def close(self):
        self.store.close()


def _ucsc(key, val):
    """ld_area.py:3-14 build_ucsc_header."""
    if isinstance(val, str):
        val = f'"{val}"'
    elif isinstance(val, tuple):
        val = ",".join(f'"{v}"' for v in val)
    return f"{key}={val}"


# --------------------------------------------------------------------------- ld_area
def ld_area(src_dir_path, intgen_dir_path, trg_top_dir_path=None, meta_lines_quan=0, gend_names="both", pop_names="all",
            flank_size=100000, ld_thres_measure="r_square", ld_low_thres=0.8, trg_file_type="tsv", ctx=None):
    """ld_area.py as a function: same arguments as its CLI (cli/ld_area_cli_en.py:36-60), same output
    tree (<src>_in_LD/<chrom>/<rsID>_chr<chrom>_<m>_<thres>.<ext>, ld_area.py:82-84,160)."""
    own = ctx is None
    ctx = ctx or Context()
    src_dir_path, intgen_dir_path = os.path.normpath(src_dir_path), os.path.normpath(intgen_dir_path)
    trg_top = src_dir_path if trg_top_dir_path is None else os.path.normpath(trg_top_dir_path)
    convdb = os.path.join(intgen_dir_path, "conversion.db")
    gends, pops = gender_tuple(gend_names), tuple(pop_names.upper().split(","))
    sample_names = get_sample_names(gends, pops, convdb)
    ext = trg_file_type if trg_file_type in ("tsv", "json") else "txt"
    meta_keys = ["chr", "gends", "pops", "each_flank", f"{ld_thres_measure}_thres"]
    header_row = ["hg38_pos", "rsID", "ref", "alt", "type", "alt_freq", "r2", "D'", "dist"]
    t_e4 = threshold_e4(ld_low_thres)
    chrom_cache = {}
    try:
        for src_file_name in os.listdir(src_dir_path):
            data_by_chrs = create_src_dict(src_dir_path, src_file_name, meta_lines_quan, convdb)
            trg_dir = os.path.join(trg_top, f"{src_file_name.rsplit('.', maxsplit=1)[0]}_in_LD")
            for chrom, var_rows in data_by_chrs.items():
                chr_dir = os.path.join(trg_dir, chrom)
                os.makedirs(chr_dir)                                         # ld_area.py:123 (not exist_ok)
                if chrom not in chrom_cache:
                    cd = chrom_cache[chrom] = ChromData(ctx, os.path.join(intgen_dir_path, f"{chrom}.vcf.gz"))
                    cd.select_samples(sample_names)
                cd = chrom_cache[chrom]
                meta_vals = [chrom, gends, pops, flank_size, ld_low_thres]
                ucsc = "##" + " ".join(map(_ucsc, meta_keys, meta_vals))
                # ---- every query of the chromosome in ONE window scan (ld_area.py:152-249)
                q_row = np.array([cd.row_of(p, i) for p, i in var_rows], dtype=np.int64)
                q_pos = cd.pos[q_row]
                ws = np.maximum(q_pos - flank_size, 0)                        # :174-176
                we = q_pos + flank_size                                       # :177
                lo = np.searchsorted(cd.pos0, ws - cd.max_ref_len, side="right")
                hi = np.maximum(np.searchsorted(cd.pos0, we, side="left"), lo)
                hits, _ = cd.store.window(q_row, lo, hi, ws, we, ld_thres_measure, t_e4)
                bounds = np.searchsorted(hits["query"], np.arange(len(q_row) + 1))
                for k, (pos, rs_id) in enumerate(var_rows):
                    mine = hits[bounds[k]:bounds[k + 1]]
                    if not len(mine):
                        continue                                             # empty result: file removed, :291-292
                    qr = int(q_row[k])
                    q_ann = [int(cd.pos[qr]), cd.ids[qr], cd.refs[qr], cd.alts[qr], cd.vts[qr], cd.p_e4[qr] / 10000.0] + ["quer"] * 3
                    path = os.path.join(chr_dir, f"{cd.ids[qr]}_chr{chrom}_{ld_thres_measure[0]}_{str(ld_low_thres)}.{ext}")
                    rows = []
                    for h in mine:
                        r = int(h["row"])
                        rows.append([int(cd.pos[r]), cd.ids[r], cd.refs[r], cd.alts[r], cd.vts[r], cd.p_e4[r] / 10000.0,
                                     r2_value(h["packed"]), dprime_value(h["packed"]), int(cd.pos[r] - cd.pos[qr])])   # :264-272
                    with open(path, "w") as fh:
                        if trg_file_type == "rsids":                          # :201-204, :258-260
                            fh.write(ucsc + "\n#rsID\n" + cd.ids[qr] + "\n")
                            fh.writelines(r[1] + "\n" for r in rows)
                        elif trg_file_type == "tsv":                          # :205-208, :273-274
                            fh.write(ucsc + "\n#" + "\t".join(header_row) + "\n")
                            fh.write("\t".join(map(str, q_ann)) + "\n")
                            fh.writelines("\t".join(map(str, r)) + "\n" for r in rows)
                        else:                                                 # :209-211, :275-283
                            obj = [dict(zip(meta_keys, meta_vals)), dict(zip(header_row, q_ann))]
                            obj += [dict(zip(header_row, r)) for r in rows]
                            json.dump(obj, fh, indent=4)
    finally:
        for cd in chrom_cache.values():
            cd.close()
        if own:
            ctx.close()


# --------------------------------------------------------------------------- ld_triangle (table output)
def ld_triangle(src_dir_path, intgen_dir_path, trg_top_dir_path=None, meta_lines_quan=0, gend_names="both", pop_names="all",
                ld_measure="r_square", ld_low_thres=None, ctx=None):
    """ld_triangle.py -o table as a function (cli/ld_triangle_cli_en.py:40-74).  The heatmap outputs are
    Plotly rendering, out of scope; the matrix they draw is this one."""
    own = ctx is None
    ctx = ctx or Context()
    src_dir_path, intgen_dir_path = os.path.normpath(src_dir_path), os.path.normpath(intgen_dir_path)
    trg_top = src_dir_path if trg_top_dir_path is None else os.path.normpath(trg_top_dir_path)
    convdb = os.path.join(intgen_dir_path, "conversion.db")
    gends, pops = gender_tuple(gend_names), tuple(pop_names.upper().split(","))
    sample_names = get_sample_names(gends, pops, convdb)
    t_e4 = None if ld_low_thres is None else threshold_e4(ld_low_thres)
    chrom_cache = {}
    try:
        for src_file_name in os.listdir(src_dir_path):
            data_by_chrs = create_src_dict(src_dir_path, src_file_name, meta_lines_quan, convdb)
            base = src_file_name.rsplit(".", maxsplit=1)[0]
            trg_dir = os.path.join(trg_top, f"{base}_LD_matr")
            for chrom, var_rows in data_by_chrs.items():
                if len(var_rows) < 2:                                        # ld_triangle.py:80
                    continue
                os.makedirs(trg_dir, exist_ok=True)
                var_rows.sort(key=lambda row: row[0])                        # :88 (stable)
                if chrom not in chrom_cache:
                    cd = chrom_cache[chrom] = ChromData(ctx, os.path.join(intgen_dir_path, f"{chrom}.vcf.gz"))
                    cd.select_samples(sample_names)
                cd = chrom_cache[chrom]
                poss = [str(p) for p, _ in var_rows]
                ids = [i for _, i in var_rows]
                v = len(ids)
                rows = np.array([cd.row_of(p, i) for p, i in var_rows], dtype=np.int64)
                packed, _ = cd.store.triangle(rows, measure=ld_measure, thres_e4_=t_e4)      # :133-230 in one call
                tab = "\t"
                with open(os.path.join(trg_dir, f"{base}_chr{chrom}_{ld_measure[0]}.tsv"), "w") as fh:   # :351-360
                    fh.write(f"##General\tinfo:\t{ld_measure}\tchr{chrom}\t{tab.join(pops)}\t{tab.join(gends)}\n\n")
                    fh.write("rsIDs\t\t" + "\t".join(ids) + "\n")
                    fh.write("\tPositions\t" + "\t".join(poss) + "\n")
                    for r in range(v):
                        cells = []
                        for c in range(v):
                            w = packed[r * (r - 1) // 2 + c] if c < r else None
                            cells.append("0" if w is None or (w & BELOW_THRES) else str(measure_value(w, ld_measure)))
                        fh.write(ids[r] + "\t" + poss[r] + "\t" + "\t".join(cells) + "\n")
    finally:
        for cd in chrom_cache.values():
            cd.close()
        if own:
            ctx.close()


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
        cd.select_samples(get_sample_names(gender_tuple(gend_names), tuple(pop_names.upper().split(",")), convdb))
        r1, r2 = cd.row_of(pos1, rs_id_1), cd.row_of(pos2, rs_id_2)
        out = cd.store.pairs([r1], [r2], raw=False)
        w = out["packed"][0]
        first_alt = lambda r: cd.alts[r].split(",")[0]                        # noqa: E731  (intgen_rec.alts[0], :116)
        return tabulate([["chrom", chrom, chrom], ["hg38_pos", pos1, pos2],
                         ["alleles", cd.refs[r1] + "/" + first_alt(r1), cd.refs[r2] + "/" + first_alt(r2)],
                         ["type", cd.vts[r1], cd.vts[r2]],
                         ["alt_freq", cd.p_e4[r1] / 10000.0, cd.p_e4[r2] / 10000.0]],
                        headers=[tabulate([["r2", r2_value(w)], ["D'", dprime_value(w)], ["abs_dist", abs(pos1 - pos2)]],
                                          tablefmt="fancy_grid", disable_numparse=True),
                                 f"\n\n\n{rs_id_1}", f"\n\n\n{rs_id_2}"], tablefmt="fancy_grid")
    finally:
        cd.close()
        if own:
            ctx.close()
