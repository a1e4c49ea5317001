"""The general route: genotype rows that are not complete phased diploid 0/1 rows -- missing calls,
haploid samples (chrX males), other allele codes, unphased fields -- through the STORE path (K1 general parser, K2, K3, K4,
K5 / K5'), against calc_ld on the lists the reference's drivers would build (`+= rec.samples[name]['GT']`, ld_area.py:182-187;
pysam's GT tuples restated in gt_tuple()).  Oracle: the C restatement of calc_ld.py on byte-coded lists (zip truncation,
None / other codes in N but in neither count).  Run on a B200."""
import numpy as np
import pytest

from oracle import ld_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ld_tools_b200 import Context
    c = Context(0)
    yield c
    c.close()


def gt_tuple(field):
    """pysam's rec.samples[name]['GT'] for a VCF sample field (as tests/refshim/pysam restates it)."""
    gt = field.split(":")[0]
    sep = "|" if "|" in gt else "/"
    return tuple(None if a == "." else int(a) for a in gt.split(sep))


def make_rows(rng, n_var, n_samples, style):
    """-> list of rows, each a list of n_samples GT field strings.
    style 'autosome': mostly plain rows, a share with missing calls / unphased / other codes / sub-fields;
    style 'chrx': males (odd samples) haploid in most rows, a 'PAR' block of all-diploid rows, missing calls here and there."""
    rows = []
    male = (np.arange(n_samples) % 3) == 1
    base = rng.random((n_var, 2 * n_samples)) < rng.choice([0.02, 0.2, 0.5, 0.9], size=n_var)[:, None]
    for i in range(1, n_var):                                   # neighbours in LD
        if rng.random() < 0.4:
            base[i] = base[i - 1] ^ (rng.random(2 * n_samples) < 0.01)
    for v in range(n_var):
        a = base[v].astype(int)
        fields = [f"{a[2 * s]}|{a[2 * s + 1]}" for s in range(n_samples)]
        u = rng.random()
        if style == "chrx":
            par = v < n_var // 8 or v >= n_var - n_var // 16
            if not par:
                for s in np.flatnonzero(male):
                    fields[s] = str(a[2 * s])
            if u < 0.15:
                for s in rng.choice(n_samples, 3, replace=False):
                    fields[s] = "." if (not par and male[s]) else rng.choice([".|.", "./.", f".|{a[2 * s]}"])
        else:
            if u < 0.10:
                for s in rng.choice(n_samples, int(rng.integers(1, 6)), replace=False):
                    fields[s] = rng.choice([".|.", "./.", f".|{a[2 * s]}", f"{a[2 * s]}|.", "."])
            elif u < 0.15:
                fields = [f.replace("|", "/") for f in fields]                  # unphased: the same tuples
            elif u < 0.20:
                for s in rng.choice(n_samples, 2, replace=False):
                    fields[s] = rng.choice(["2|0", "1|2", "10|1", "2/2"])       # other allele codes
            elif u < 0.25:
                fields = [f + ":35:0.99" for f in fields]                       # GT is the first sub-field
            elif u < 0.28:
                fields[int(rng.integers(n_samples))] = "1"                      # a lone haploid call
        rows.append(fields)
    return rows


class GList(list):
    """A genotype list with its byte coding for the C oracle cached."""
    _enc = None

    @property
    def enc(self):
        if self._enc is None:
            self._enc = ld_oracle.encode_genotypes(self)
        return self._enc


def lists_for(rows, sel_samples):
    """The reference's flat genotype lists (one per variant) for the selected samples, in column order."""
    out = []
    for fields in rows:
        g = GList()
        for s in sel_samples:
            g += gt_tuple(fields[s])
        out.append(g)
    return out


def oracle_pair(ga, gb):
    res = ld_oracle.calc_ld_bytes(ga.enc, gb.enc)
    return res, int(ld_oracle.packed_of(res)[0])


def store_from_rows(ctx, rows, n_samples, via="pack_gt"):
    from ld_tools_b200 import Store
    text = "\n".join("\t".join(f) for f in rows) + "\n"
    buf = np.frombuffer(text.encode(), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len("\t".join(f)) + 1 for f in rows])[:-1]]).astype(np.int64)
    st = Store(ctx, len(rows), 2 * n_samples)
    status = st.pack_gt(0, buf, n_samples, row_off=off)
    return st, status


@pytest.mark.parametrize("style,n_samples", [("autosome", 150), ("chrx", 97), ("chrx", 310), ("autosome", 700), ("chrx", 1100), ("autosome", 2504)])   # 128- / 256- / 384- / 640-byte rows
def test_general_rows_all_kernels_equal_calc_ld_on_the_lists(ctx, style, n_samples):
    from ld_tools_b200._lib import TUNE_MMA_TILE_N
    from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC, threshold_e4
    rng = np.random.default_rng(n_samples)
    n_var = 300
    rows = make_rows(rng, n_var, n_samples, style)
    st, status = store_from_rows(ctx, rows, n_samples)
    plain = np.array([all(len(f) == 3 and f[1] == "|" and f[0] in "01" and f[2] in "01" for f in r) for r in rows])
    assert ((status & 1) == 0).tolist() == plain.tolist() and not (status & 8).any()
    sel = np.sort(rng.choice(n_samples, int(0.8 * n_samples), replace=False))
    st.select_haplotypes(np.concatenate([2 * sel, 2 * sel + 1]))
    lists = lists_for(rows, sel)
    n1, ln, kind, n_gen = st.row_counts()
    assert ln.tolist() == [len(g) for g in lists] and n1.tolist() == [g.count(1) for g in lists]
    assert n_gen == int((kind != -1).sum()) and 0 < n_gen < n_var
    if style == "chrx":
        assert (kind == -1).sum() > n_var // 2                   # the haploid-male pattern is the store's common one: fast paths
        assert (kind == -2).sum() > 0                            # all-diploid (PAR) rows: general route without aux planes
    # own alt frequency as ld_area prints it for a query (ld_area.py:188-189)
    _, p_e4, _ = st.counts()
    assert p_e4.tolist() == [int(round(round(g.count(1) / len(g), 4) * 10000)) for g in lists]
    # ---- K3: explicit pairs, raw values and counts
    ia = rng.integers(0, n_var, 400)
    ib = rng.integers(0, n_var, 400)
    got = st.pairs(ia, ib)
    for k in range(400):
        res, word = oracle_pair(lists[ia[k]], lists[ib[k]])
        assert got["packed"][k] == word, (k, ia[k], ib[k], hex(got["packed"][k]), hex(word))
        if kind[ia[k]] != -1 or kind[ib[k]] != -1:
            assert got["n11"][k] == res["n_11"]
        assert abs(got["d"][k] - res["d"]) <= 1e-12 and abs(got["dprime"][k] - res["dprime"]) <= 1e-12 and abs(got["r2"][k] - res["r2"]) <= 1e-12
    # ---- K5' and K5 (64- and 128-wide tiles): the whole triangle
    want = np.zeros(n_var * (n_var - 1) // 2, dtype=np.uint32)
    want_n11 = np.zeros_like(want, dtype=np.int32)
    for r in range(1, n_var):
        for c in range(r):
            res, word = oracle_pair(lists[r], lists[c])          # var_1 = row, var_2 = column (ld_triangle.py:193)
            want[r * (r - 1) // 2 + c] = word
            want_n11[r * (r - 1) // 2 + c] = res["n_11"]
    rows_idx = np.arange(n_var)
    packed, n11 = st.triangle(rows_idx, engine=ENGINE_POPC, want_n11=True)
    assert (packed == want).all() and (n11 == want_n11).all()
    t = threshold_e4(0.3)
    for tile in (64, 128):
        ctx.set_tuning(TUNE_MMA_TILE_N, tile)
        try:
            packed, n11 = st.triangle(rows_idx, engine=ENGINE_MMA, want_n11=True)
            flagged, _ = st.triangle(rows_idx[::-1].copy(), measure="d_prime", thres_e4_=t, engine=ENGINE_MMA)
        finally:
            ctx.set_tuning(TUNE_MMA_TILE_N, 0)
        assert (packed == want).all() and (n11 == want_n11).all()
        ref_flagged, _ = st.triangle(rows_idx[::-1].copy(), measure="d_prime", thres_e4_=t, engine=ENGINE_POPC)
        assert (flagged == ref_flagged).all()
    # ---- K4: windows (every row eligible, positions 10 apart): kept sets and words, both kernels (one query / many)
    pos0 = (np.arange(n_var) * 10 + 100).astype(np.int32)
    st.set_annotations(pos0, pos0 + 1, np.arange(n_var, dtype=np.int64), np.ones(n_var, np.uint8))
    from ld_tools_b200 import shard
    for q_row, measure, thres in ((np.array([150]), "r_square", 0.0), (np.arange(5, 295, 7), "d_prime", 0.5), (np.arange(5, 295, 7), "r_square", 0.05)):
        lo, hi, ws, we = shard.window_bounds(pos0, 1, pos0[q_row].astype(np.int64) + 1, 400)
        te = threshold_e4(thres)
        hits, _ = st.window(q_row, lo, hi, ws, we, measure, te)
        for k, q in enumerate(q_row):
            mine = hits[hits["query"] == k]
            exp_rows, exp_words = [], []
            for j in range(int(lo[k]), int(hi[k])):
                if j == q or not (pos0[j] < we[k] and pos0[j] + 1 > ws[k]):
                    continue
                res, word = oracle_pair(lists[q], lists[j])      # var_1 = query, var_2 = window row (ld_area.py:242)
                val = (word & 0x3FFF) if measure == "r_square" else ((word >> 16) & 0x3FFF)
                if val >= te:
                    exp_rows.append(j)
                    exp_words.append(word)
            assert mine["row"].tolist() == exp_rows and mine["packed"].tolist() == exp_words, (k, q)
    st.close()


def test_general_rows_survive_subset_store_and_store_file(ctx, tmp_path):
    """The narrower store of a sample subset and the on-disk store carry the aux planes: same results as the masked store."""
    from ld_tools_b200 import Store
    from ld_tools_b200.engine import ENGINE_MMA
    rng = np.random.default_rng(5)
    n_samples, n_var = 120, 280
    rows = make_rows(rng, n_var, n_samples, "chrx")
    st, _ = store_from_rows(ctx, rows, n_samples)
    sel = np.sort(rng.choice(n_samples, 40, replace=False))
    hap = np.sort(np.concatenate([2 * sel, 2 * sel + 1]))
    st.select_haplotypes(hap)
    want, want_n11 = st.triangle(np.arange(n_var), engine=ENGINE_MMA, want_n11=True)
    lists = lists_for(rows, sel)
    for r, c in ((5, 2), (200, 17), (279, 270), (150, 149)):
        assert want[r * (r - 1) // 2 + c] == oracle_pair(lists[r], lists[c])[1]
    sub = st.subset(hap)
    got, got_n11 = sub.triangle(np.arange(n_var), engine=ENGINE_MMA, want_n11=True)
    assert (got == want).all() and (got_n11 == want_n11).all()
    assert sub.row_counts()[1].tolist() == st.row_counts()[1].tolist()
    sub.close()
    path = str(tmp_path / "x.ldxstore")
    st.save(path)
    back = Store.load(ctx, path)
    back.select_haplotypes(hap)
    got, got_n11 = back.triangle(np.arange(n_var), engine=ENGINE_MMA, want_n11=True)
    assert (got == want).all() and (got_n11 == want_n11).all()
    back.close()
    st.close()


def test_rows_no_parser_takes_are_flagged(ctx):
    rows = [["0|1", "1|1", "0|0"], ["0|1|1", "0|0", "1|1"], ["0|1", "", "1|1"], ["0|1", "1|1"], ["0|x", "1|1", "0|0"], ["0", "1", "."]]
    from ld_tools_b200 import Store
    text = "\n".join("\t".join(f) for f in rows) + "\n"
    off = np.concatenate([[0], np.cumsum([len("\t".join(f)) + 1 for f in rows])[:-1]]).astype(np.int64)
    st = Store(ctx, len(rows), 6)
    status = st.pack_gt(0, np.frombuffer(text.encode(), dtype=np.uint8), 3, row_off=off)
    assert [int(s) for s in status] == [0, 9, 9, 9, 9, 1]
    st.close()
