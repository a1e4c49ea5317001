"""Parity of the CUDA path (through the C ABI, libldx.so) against the CPU oracle and the golden
vectors produced by the reference's own calc_ld.  Run on a B200: `pytest -m gpu`.

Bars (BASELINE.json north_star): counts bit-exact; D / D' / r2 within 1e-12 absolute of the
reference before rounding (they are in fact bit-exact except r2, which can differ in the last
ulp because the reference squares D with libm pow); rounded outputs identical, including the
reference's int-0 vs float types; ld_area pair sets identical.
"""
import numpy as np
import pytest

from conftest import decode_genotypes, decode_raw
from oracle import calc_ld_port, ld_oracle

pytestmark = pytest.mark.gpu

KEYS = ("r_square", "d_prime", "var_1_alt_freq", "var_2_alt_freq")
TOL = 1e-12   # absolute, pre-rounding (north_star)


@pytest.fixture(scope="module")
def ctx():
    from ld_tools_b200 import Context
    c = Context(0)
    yield c
    c.close()


def same_obj(a, b):
    return type(a) is type(b) and repr(a) == repr(b)


def planes_from_counts(cases, n):
    """Two store rows per (n11, n1a, n1b) case with exactly those counts (scattered bit positions)."""
    rng = np.random.default_rng(n)
    h = np.zeros((2 * len(cases), n), dtype=np.uint8)
    for k, (n11, a, b) in enumerate(cases):
        perm = rng.permutation(n)
        oa, ob = a - n11, b - n11
        h[2 * k, perm[:n11 + oa]] = 1
        h[2 * k + 1, perm[:n11]] = 1
        h[2 * k + 1, perm[n11 + oa:n11 + oa + ob]] = 1
    return ld_oracle.pack_bits(h)


# ------------------------------------------------------------------ calc_ld drop-in (lists)

def test_calc_ld_dropin_matches_reference_golden(golden, ctx):
    from ld_tools_b200 import calc_ld
    for name, sa, sb, *out in golden["list_cases"]:
        g1, g2 = decode_genotypes(sa), decode_genotypes(sb)
        if name == "tuple_inputs":
            g1, g2 = tuple(g1), tuple(g2)
        got = calc_ld(g1, g2)
        assert list(got) == list(KEYS)
        for k, want in zip(KEYS, out[:4]):
            assert repr(got[k]) == want, (name, k, got[k], want)


def test_calc_ld_lists_raw_values_and_counts(golden, ctx):
    from ld_tools_b200.calc_ld import encode_genotypes
    for name, sa, sb, *out in golden["list_cases"]:
        g1, g2 = decode_genotypes(sa), decode_genotypes(sb)
        res = ctx.calc_ld_lists(encode_genotypes(g1), encode_genotypes(g2))
        full = calc_ld_port.calc_ld_full(g1, g2)
        for f_dev, f_port in (("n_hap", "n_hap"), ("n_11", "n_11"), ("n_a1", "n_a1"), ("n_a0", "n_a0"),
                              ("n_b1", "n_b1"), ("n_b0", "n_b0")):
            assert int(res[f_dev]) == full[f_port], (name, f_dev)
        raw = [decode_raw(x) for x in out[4:9]]
        assert float(res["d"]).hex() == float(raw[4]).hex(), name             # D bit-exact
        assert float(res["p_a"]).hex() == float(raw[2]).hex() and float(res["p_b"]).hex() == float(raw[3]).hex()
        assert bool(res["dprime_is_int0"]) == isinstance(raw[1], int), name
        assert bool(res["r2_is_int0"]) == isinstance(raw[0], int), name
        if not isinstance(raw[1], int):
            assert float(res["dprime"]).hex() == raw[1].hex(), name             # D' bit-exact
        if not isinstance(raw[0], int):
            assert abs(float(res["r2"]) - raw[0]) <= TOL, name


def test_calc_ld_empty_raises_zero_division(ctx):
    from ld_tools_b200 import calc_ld
    with pytest.raises(ZeroDivisionError):
        calc_ld([], [])
    with pytest.raises(ZeroDivisionError):
        calc_ld([1, 0], [])


def test_calc_ld_numeric_types(ctx):
    """list.count semantics: 1.0 and True count as alt, None / 2 as neither (calc_ld.py:37-40)."""
    from ld_tools_b200 import calc_ld
    g1, g2 = [1.0, 0.0, True, False, 1, 0, None, 2], [1, 0, 1.0, 0, 0.0, True, 1, 1]
    assert calc_ld(g1, g2) == calc_ld_port.calc_ld(g1, g2)
    assert all(same_obj(calc_ld(g1, g2)[k], calc_ld_port.calc_ld(g1, g2)[k]) for k in KEYS)


# ------------------------------------------------------------------ store rows: pairs kernel

def test_pairs_match_reference_golden_counts(golden, ctx):
    from ld_tools_b200 import Store
    from ld_tools_b200.engine import dprime_value, r2_value
    by_n = {}
    for tag, n, n11, a, b, *out in golden["count_cases"]:
        by_n.setdefault(n, []).append(((n11, a, b), out))
    for n, items in by_n.items():
        planes = planes_from_counts([c for c, _ in items], n)
        st = Store.from_planes(ctx, planes, n)
        st.select_all()
        n1, p_e4, n_sel = st.counts()
        assert n_sel == n
        k = np.arange(len(items))
        res = st.pairs(2 * k, 2 * k + 1)
        for i, ((n11, a, b), out) in enumerate(items):
            assert res["n11"][i] == n11 and n1[2 * i] == a and n1[2 * i + 1] == b
            word = res["packed"][i]
            assert repr(r2_value(word)) == out[0], (n, n11, a, b, r2_value(word), out[0])
            assert repr(dprime_value(word)) == out[1], (n, n11, a, b)
            assert repr(int(p_e4[2 * i]) / 10000.0) == out[2] and repr(int(p_e4[2 * i + 1]) / 10000.0) == out[3]
            raw_r2, raw_dp, raw_d = decode_raw(out[4]), decode_raw(out[5]), decode_raw(out[8])
            assert float(res["d"][i]).hex() == float(raw_d).hex()
            if not isinstance(raw_dp, int):
                assert float(res["dprime"][i]).hex() == raw_dp.hex()
            if not isinstance(raw_r2, int):
                assert abs(res["r2"][i] - raw_r2) <= TOL
        st.close()


def test_pairs_with_mask_and_subset_store(ctx):
    from ld_tools_b200 import Store
    rng = np.random.default_rng(21)
    n_hap = 5008
    h = (rng.random((300, n_hap)) < rng.beta(0.3, 1.5, size=(300, 1))).astype(np.uint8)
    h[10] = 0; h[11] = 1; h[12] = h[13]          # monomorphic rows, identical rows
    planes = ld_oracle.pack_bits(h)
    st = Store.from_planes(ctx, planes, n_hap)
    sel = np.sort(rng.choice(n_hap, size=1006, replace=False))
    mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
    st.set_mask(mask)
    n1, p_e4, n_sel = st.counts()
    assert n_sel == 1006
    assert (n1 == ld_oracle.variant_counts(planes, mask, n_hap)).all()
    ia = rng.integers(0, 300, size=4000); ib = rng.integers(0, 300, size=4000)
    ia[:4], ib[:4] = [10, 11, 12, 10], [5, 5, 13, 11]
    got = st.pairs(ia, ib)
    want = ld_oracle.pairs(planes, mask, n_hap, ia, ib)
    assert (got["n11"] == want["n_11"]).all()
    assert (got["packed"] == ld_oracle.packed_of(want)).all()
    assert (got["d"] == want["d"]).all()                       # bit-exact
    assert (got["dprime"] == want["dprime"]).all()             # bit-exact
    assert np.abs(got["r2"] - want["r2"]).max() <= TOL
    # the same answers from a store that physically holds only the selected columns
    sub = st.subset(sel)
    assert sub.n_hap == 1006 and sub.stride_words == 16
    assert (ld_oracle.unpack_bits(sub.download(), 1006) == h[:, sel]).all()
    got2 = sub.pairs(ia, ib)
    for k in ("n11", "packed", "d", "dprime", "r2"):
        assert (got2[k] == got[k]).all(), k
    sub.close(); st.close()


def test_empty_mask_raises_like_reference(ctx):
    from ld_tools_b200 import Store
    st = Store(ctx, 4, 130)
    with pytest.raises(ZeroDivisionError):
        st.set_mask(np.zeros(st.stride_words, dtype="<u8"))
    st.close()


# ------------------------------------------------------------------ K1: GT text packing

def test_pack_gt_text_matches_oracle(ctx):
    from ld_tools_b200 import Store
    from ld_tools_b200.synth import gt_row_text
    rng = np.random.default_rng(5)
    for n_samples in (1, 37, 503, 2504):
        n_var = 23
        gt = (rng.random((n_var, 2 * n_samples)) < 0.3).astype(np.uint8)
        text, offs = b"", []
        for v in range(n_var):
            prefix = f"22\t{100 + v}\trs{v}\tA\tG\t100\tPASS\tAC=1;VT=SNP{'x' * (v % 7)}\tGT\t".encode()
            offs.append(len(text) + len(prefix))
            text += prefix + gt_row_text(gt[v]) + b"\n"
        buf = np.frombuffer(text, dtype=np.uint8).copy()
        buf[offs[2] + 4 * (n_samples // 2)] = ord(".")             # missing allele
        buf[offs[4] + 4 * (n_samples - 1) + 1] = ord("/")          # unphased last sample
        want_planes, want_status = ld_oracle.pack_gt(buf, np.array(offs), n_samples)
        st = Store(ctx, n_var, 2 * n_samples)
        status = st.pack_gt(0, buf, n_samples, row_off=offs)
        assert status.tolist() == want_status.tolist()
        assert (st.download() == want_planes).all()
        st.close()


# ------------------------------------------------------------------ K5': all pairs (popcount engine)

def make_store(ctx, n_var, n_hap, seed, sel_frac=None):
    from ld_tools_b200 import Store
    from ld_tools_b200.synth import synth_haplotypes
    h = synth_haplotypes(n_var, n_hap, seed=seed)
    planes = ld_oracle.pack_bits(h)
    st = Store.from_planes(ctx, planes, n_hap)
    rng = np.random.default_rng(seed)
    sel = np.arange(n_hap) if sel_frac is None else np.sort(rng.choice(n_hap, int(n_hap * sel_frac), replace=False))
    mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
    st.set_mask(mask)
    return st, planes, mask


@pytest.mark.parametrize("n_var,n_hap,sel_frac", [(2, 5008, None), (65, 5008, None), (200, 5008, 0.2),
                                                  (131, 198, None), (150, 6000, None)])
def test_triangle_matches_oracle(ctx, n_var, n_hap, sel_frac):
    from ld_tools_b200.engine import ENGINE_POPC
    st, planes, mask = make_store(ctx, max(n_var, 8), n_hap, seed=100 + n_var, sel_frac=sel_frac)
    rng = np.random.default_rng(n_var)
    rows = rng.permutation(max(n_var, 8))[:n_var]          # arbitrary matrix order
    packed, n11 = st.triangle(rows, engine=ENGINE_POPC, want_n11=True)
    want = ld_oracle.triangle(planes, mask, n_hap, rows)
    assert (n11 == want["n_11"]).all()
    assert (packed == ld_oracle.packed_of(want)).all()
    # orientation: var_1 = row variant (ld_triangle.py:193) -- spot check through the list-level port
    bits = ld_oracle.unpack_bits(planes & mask[None, :], n_hap)
    selcols = np.flatnonzero(ld_oracle.unpack_bits(mask[None, :], n_hap)[0])
    from ld_tools_b200.engine import dprime_value, r2_value, tri_index
    for r, c in [(1, 0), (n_var - 1, 0), (n_var - 1, n_var - 2)]:
        if r <= c:
            continue
        ref = calc_ld_port.calc_ld(list(map(int, bits[rows[r], selcols])), list(map(int, bits[rows[c], selcols])))
        w = packed[tri_index(r, c)]
        assert same_obj(r2_value(w), ref["r_square"]) and same_obj(dprime_value(w), ref["d_prime"])
    st.close()


def test_triangle_threshold_flag(ctx):
    from ld_tools_b200.engine import BELOW_THRES, ENGINE_POPC, dprime_e4, r2_e4, threshold_e4
    st, planes, mask = make_store(ctx, 180, 1006, seed=9)
    rows = np.arange(180)
    for measure, getter in (("r_square", r2_e4), ("d_prime", dprime_e4)):
        t = threshold_e4(0.8)
        packed, _ = st.triangle(rows, measure=measure, thres_e4_=t, engine=ENGINE_POPC)
        base, _ = st.triangle(rows, measure=measure, engine=ENGINE_POPC)
        assert ((packed & ~np.uint32(BELOW_THRES)) == base).all()
        assert (((packed & BELOW_THRES) != 0) == (getter(base) < t)).all()
    st.close()


def test_triangle_full_size_properties(ctx):
    """BASELINE config 2 shape: 2,000 variants x 5008 haplotypes = 1,999,000 pairs."""
    from ld_tools_b200.engine import ENGINE_POPC
    st, planes, mask = make_store(ctx, 2000, 5008, seed=2)
    rows = np.arange(2000)
    packed, n11 = st.triangle(rows, engine=ENGINE_POPC, want_n11=True)
    bits = ld_oracle.unpack_bits(planes, 5008).astype(np.float32)
    gram = (bits @ bits.T).astype(np.int64)                # exact: counts < 2^24
    r, c = np.tril_indices(2000, -1)
    assert (n11 == gram[r, c]).all()
    n1 = np.diag(gram)
    want = ld_oracle.packed_words(5008, n11, n1[r], n1[c])
    assert (packed == want).all()
    st.close()


# ------------------------------------------------------------------ K4: ld_area window scan

def annotated_store(ctx, n_var, n_hap, seed):
    from ld_tools_b200.synth import make_records
    st, planes, mask = make_store(ctx, n_var, n_hap, seed=seed, sel_frac=0.3)
    recs = make_records(n_var, seed=seed, mean_gap=40)
    pos0 = np.array([r["pos"] - 1 for r in recs], dtype=np.int32)
    end0 = pos0 + np.array([len(r["ref"]) for r in recs], dtype=np.int32)
    import re
    elig = np.array([bool(re.match(r"rs\d+$", r["id"])) and not r["multi"] for r in recs], dtype=np.uint8)
    idnum = np.array([int(r["id"][2:]) if re.match(r"rs\d+$", r["id"]) else -1 - i for i, r in enumerate(recs)],
                     dtype=np.int64)
    st.set_annotations(pos0, end0, idnum, elig)
    return st, planes, mask, pos0, end0, idnum, elig


@pytest.mark.parametrize("measure,thres", [("r_square", 0.8), ("d_prime", 0.95), ("r_square", 0.0), ("r_square", 0.05)])
def test_window_matches_oracle_full_scan(ctx, measure, thres):
    from ld_tools_b200.engine import threshold_e4
    n_var, n_hap = 3000, 5008
    st, planes, mask, pos0, end0, idnum, elig = annotated_store(ctx, n_var, n_hap, seed=33)
    rng = np.random.default_rng(1)
    q_rows = np.concatenate([[0, 1, n_var - 1], rng.choice(np.flatnonzero(elig), 37, replace=False)])
    flank = 5000
    pos = pos0 + 1
    low = np.maximum(pos[q_rows] - flank, 0)                # ld_area.py:174-176
    high = pos[q_rows] + flank                              # ld_area.py:177
    # candidate superset from the position index, with slack for long REF alleles on the left
    max_len = int((end0 - pos0).max())
    lo = np.searchsorted(pos0, low - max_len, side="left")
    hi = np.searchsorted(pos0, high, side="left")
    hits, scanned = st.window(q_rows, lo, hi, low, high, measure, threshold_e4(thres))
    mcode = 0 if measure == "r_square" else 1
    total = 0
    for k, q in enumerate(q_rows):
        rows, res = ld_oracle.window(planes, mask, n_hap, pos0, end0, idnum, elig, q, low[k], high[k], mcode, thres)
        mine = hits[hits["query"] == k]
        assert mine["row"].tolist() == rows.tolist(), (k, q)
        assert (mine["n11"] == res["n_11"]).all()
        assert (mine["packed"] == ld_oracle.packed_of(res)).all()
        total += len(rows)
    assert len(hits) == total
    if thres == 0.0:
        assert scanned == len(hits)                         # everything scanned is kept at threshold 0
    st.close()


@pytest.mark.parametrize("n_hap,thres", [(5008, 0.0), (5008, 0.3), (198, 0.2), (1006, 0.05)])
def test_window_many_overlapping_queries_any_order(ctx, n_hap, thres):
    """The multi-query window kernel's territory: hundreds of queries whose windows cover a large part of the store and
    each other, passed in arbitrary order, with repeated queries and an empty candidate range among them; kept rows,
    counts and packed words per query against the oracle's full scan."""
    from ld_tools_b200.engine import threshold_e4
    n_var = 2500
    st, planes, mask, pos0, end0, idnum, elig = annotated_store(ctx, n_var, n_hap, seed=41 + n_hap)
    rng = np.random.default_rng(n_hap)
    q_rows = rng.choice(np.flatnonzero(elig), 300, replace=True)              # unsorted, with repeats
    flank = 20000                                                              # ~1000 of the 2500 rows per window
    pos = pos0 + 1
    low = np.maximum(pos[q_rows] - flank, 0)
    high = pos[q_rows] + flank
    max_len = int((end0 - pos0).max())
    lo = np.searchsorted(pos0, low - max_len, side="left")
    hi = np.searchsorted(pos0, high, side="left")
    hi[7] = lo[7]                                                              # one query without candidates
    hits, scanned = st.window(q_rows, lo, hi, low, high, "r_square", threshold_e4(thres))
    total = 0
    for k, q in enumerate(q_rows):
        rows, res = ld_oracle.window(planes, mask, n_hap, pos0, end0, idnum, elig, q, low[k], high[k], 0, thres, lo=int(lo[k]), hi=int(hi[k]))
        mine = hits[hits["query"] == k]
        assert mine["row"].tolist() == rows.tolist(), (k, q)
        assert (mine["n11"] == res["n_11"]).all() and (mine["packed"] == ld_oracle.packed_of(res)).all()
        total += len(rows)
    assert len(hits) == total and total > 0
    if thres == 0.0:
        assert scanned == len(hits)
    # windows of different widths: the bounds are no longer monotone and the single-query kernel takes over -- same answers
    hi2 = hi.copy()
    hi2[::3] = np.minimum(lo[::3] + 40, hi[::3])
    hits2, _ = st.window(q_rows, lo, hi2, low, high, "r_square", threshold_e4(thres))
    for k in (0, 3, 4, 150, 299):
        rows, res = ld_oracle.window(planes, mask, n_hap, pos0, end0, idnum, elig, q_rows[k], low[k], high[k], 0, thres, lo=int(lo[k]), hi=int(hi2[k]))
        mine = hits2[hits2["query"] == k]
        assert mine["row"].tolist() == rows.tolist() and (mine["packed"] == ld_oracle.packed_of(res)).all()
    st.close()


def test_window_full_size_properties(ctx):
    """BASELINE configs[2] at full size: 1,000 queries, +/-500 kb, r2 >= 0.8, 503 of 2504 samples, 1.1 M variants x 5008
    haplotypes (31 M candidate pairs; the store is generated on the GPU with neighbours in LD).  Properties that need no
    CPU reference at this size: (1) the multi-query and the one-query kernels keep exactly the same pairs with the same
    words; (2) a store of just the selected haplotype columns (ldx_store_subset) gives the same pairs, counts and words as
    the mask over the full store; (3) a second run is identical; and a seeded sample of queries is checked against the
    oracle's finalisation of numpy popcounts."""
    import torch
    from ld_tools_b200 import Store, shard
    from ld_tools_b200._lib import TUNE_WINDOW_MQ
    from ld_tools_b200.engine import threshold_e4
    from ld_tools_b200.synth import fill_store_grouped
    nv, n_hap, flank = 1_100_000, 5008, 500_000
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(2024)
    st = Store(ctx, nv, n_hap)
    fill_store_grouped(st, dev, 22, 0, nv)
    pos0 = np.sort(rng.integers(16_050_000, 51_200_000, size=nv)).astype(np.int32)
    st.set_annotations(pos0, pos0 + 1, np.arange(nv, dtype=np.int64), np.ones(nv, np.uint8))
    samples = np.sort(rng.choice(n_hap // 2, 503, replace=False))
    hap_idx = np.sort(np.concatenate([2 * samples, 2 * samples + 1]))
    st.select_haplotypes(hap_idx)
    q_row = np.sort(rng.choice(nv, 1000, replace=False)).astype(np.int64)
    lo, hi, ws, we = shard.window_bounds(pos0, 1, pos0[q_row].astype(np.int64) + 1, flank)
    t = threshold_e4(0.8)
    hits_mq, scanned_mq = st.window(q_row, lo, hi, ws, we, "r_square", t)
    again, _ = st.window(q_row, lo, hi, ws, we, "r_square", t)
    ctx.set_tuning(TUNE_WINDOW_MQ, 0)
    try:
        hits_1q, scanned_1q = st.window(q_row, lo, hi, ws, we, "r_square", t)
    finally:
        ctx.set_tuning(TUNE_WINDOW_MQ, 1)
    assert scanned_mq == scanned_1q == int((hi - lo).sum()) - 1000          # every candidate but the query itself
    assert len(hits_mq) > 1000 and (hits_mq == hits_1q).all() and (hits_mq == again).all()
    sub = st.subset(hap_idx)
    hits_sub, scanned_sub = sub.window(q_row, lo, hi, ws, we, "r_square", t)
    assert scanned_sub == scanned_mq and (hits_sub == hits_mq).all()
    # a sample of queries against numpy popcounts + the oracle's finalisation
    words = (n_hap + 63) // 64
    bits = np.zeros(st.stride_words * 64, dtype=np.uint8)
    bits[hap_idx] = 1
    mask = np.packbits(bits, bitorder="little").view("<u8")[:words]
    for k in rng.choice(1000, 4, replace=False):
        a, b, q = int(lo[k]), int(hi[k]), int(q_row[k])
        win = st.download(a, b - a)[:, :words] & mask
        qrow = st.download(q, 1)[0, :words] & mask
        n1 = np.bitwise_count(win).sum(axis=1).astype(np.int32)
        n11 = np.bitwise_count(win & qrow[None, :]).sum(axis=1).astype(np.int32)
        want = ld_oracle.packed_words(1006, n11, np.full(b - a, int(np.bitwise_count(qrow).sum()), np.int32), n1)
        keep = ((want & 0x3FFF) >= t) & (np.arange(a, b) != q)
        got = hits_mq[hits_mq["query"] == k]
        assert got["row"].tolist() == (np.flatnonzero(keep) + a).tolist()
        assert (got["n11"] == n11[keep]).all() and (got["packed"] == want[keep]).all()
    sub.close()
    st.close()


@pytest.mark.parametrize("measure,thres", [("r_square", 0.05), ("d_prime", 0.95), ("r_square", 0.0), ("d_prime", 0.3)])
def test_window_thread_per_row_kernel_equals_the_single_query_kernel(ctx, measure, thres):
    """128-byte rows (a subset store, or any store of at most 1024 haplotypes) take the row-per-thread kernel
    (window_rows1_kernel) when several queries overlap: same hits, same scanned count as the one-query-per-pass kernel and as
    the masked full-width store -- with records that fail the filters (ineligible IDs, multi-allelic, the query's own ID at
    another row, long REF alleles at the window edge), windows clipped at both ends of the store and query groups that do
    not fill a group of 8."""
    from ld_tools_b200._lib import TUNE_WINDOW_MQ
    from ld_tools_b200.engine import threshold_e4
    n_var, n_hap = 5000, 5008
    st, planes, mask, pos0, end0, idnum, elig = annotated_store(ctx, n_var, n_hap, seed=71)
    idnum = idnum.copy()
    idnum[100:140:7] = idnum[120]                       # the query's ID again at other rows: skipped (ld_area.py:222)
    st.set_annotations(pos0, end0, idnum, elig)
    rng = np.random.default_rng(3)
    sel = np.sort(rng.choice(n_hap, 1000, replace=False))
    st.select_haplotypes(sel)
    sub = st.subset(sel)
    assert sub.stride_words == 16
    q_rows = np.unique(np.concatenate([[0, 1, 120, n_var - 1], rng.choice(n_var, 61, replace=False)])).astype(np.int64)
    pos = pos0.astype(np.int64) + 1
    flank = 9000
    max_len = int((end0 - pos0).max())
    lo = np.searchsorted(pos0, np.maximum(pos[q_rows] - flank, 0) - max_len, side="left").astype(np.int64)
    hi = np.searchsorted(pos0, pos[q_rows] + flank, side="left").astype(np.int64)
    ws, we = np.maximum(pos[q_rows] - flank, 0).astype(np.int32), (pos[q_rows] + flank).astype(np.int32)
    t = threshold_e4(thres)
    want, scanned_want = st.window(q_rows, lo, hi, ws, we, measure, t)                  # full-width rows under the mask
    got, scanned = sub.window(q_rows, lo, hi, ws, we, measure, t)                       # thread per row
    ctx.set_tuning(TUNE_WINDOW_MQ, 0)
    try:
        one, scanned_one = sub.window(q_rows, lo, hi, ws, we, measure, t)               # one query per pass
    finally:
        ctx.set_tuning(TUNE_WINDOW_MQ, 1)
    assert scanned == scanned_one == scanned_want and scanned > 0
    assert len(got) > 0 and (got == one).all() and (got == want).all()
    # and against the oracle (full-width planes under the selection mask) for a few queries
    sel_mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
    mcode = 0 if measure == "r_square" else 1
    for k in (0, 2, len(q_rows) - 1, len(q_rows) // 2, int(np.flatnonzero(q_rows == 120)[0])):
        rows, res = ld_oracle.window(planes, sel_mask, n_hap, pos0, end0, idnum, elig, int(q_rows[k]), int(ws[k]), int(we[k]), mcode, thres,
                                     lo=int(lo[k]), hi=int(hi[k]))
        mine = got[got["query"] == k]
        assert mine["row"].tolist() == rows.tolist() and (mine["n11"] == res["n_11"]).all() and (mine["packed"] == ld_oracle.packed_of(res)).all()
    sub.close()
    st.close()


def test_window_edge_cases(ctx):
    from ld_tools_b200.engine import threshold_e4
    st, planes, mask, pos0, end0, idnum, elig = annotated_store(ctx, 600, 198, seed=44)
    # no queries
    hits, scanned = st.window([], [], [], [], [], "r_square", 0)
    assert len(hits) == 0 and scanned == 0
    # empty candidate range and a window that contains only the query itself
    hits, _ = st.window([5, 7], [5, 7], [5, 8], [pos0[5], pos0[7]], [pos0[5] + 1, pos0[7] + 1], "r_square", 0)
    assert len(hits) == 0
    # capacity retry path: tiny cap forces LDX_ERR_CAPACITY then a second call
    q = int(np.flatnonzero(elig)[10])
    hits, _ = st.window([q], [0], [600], [0], [int(end0.max()) + 1], "r_square", 0, cap=3)
    rows, _ = ld_oracle.window(planes, mask, 198, pos0, end0, idnum, elig, q, 0, int(end0.max()) + 1, 0, 0.0)
    assert hits["row"].tolist() == rows.tolist() and len(rows) > 3
    st.close()


# ------------------------------------------------------------------ K5: tcgen05 int8 Gram engine

@pytest.mark.parametrize("tile_n", [64, 128])
@pytest.mark.parametrize("n_var,n_hap,sel_frac", [(2, 5008, None), (129, 5008, None), (700, 5008, 0.2),
                                                  (300, 198, None), (513, 6000, None)])
def test_triangle_mma_bit_exact_vs_popcount_and_oracle(ctx, tile_n, n_var, n_hap, sel_frac):
    """The tensor-core engine must reproduce the popcount engine's counts and words bit for bit."""
    from ld_tools_b200._lib import TUNE_MMA_TILE_N
    from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC
    st, planes, mask = make_store(ctx, max(n_var, 8), n_hap, seed=300 + n_var, sel_frac=sel_frac)
    rng = np.random.default_rng(n_var)
    rows = rng.permutation(max(n_var, 8))[:n_var]
    ctx.set_tuning(TUNE_MMA_TILE_N, tile_n)
    try:
        packed, n11 = st.triangle(rows, engine=ENGINE_MMA, want_n11=True)
    finally:
        ctx.set_tuning(TUNE_MMA_TILE_N, 0)
    ref_packed, ref_n11 = st.triangle(rows, engine=ENGINE_POPC, want_n11=True)
    assert (n11 == ref_n11).all()
    assert (packed == ref_packed).all()
    if n_var <= 300:
        want = ld_oracle.triangle(planes, mask, n_hap, rows)
        assert (n11 == want["n_11"]).all() and (packed == ld_oracle.packed_of(want)).all()
    st.close()


@pytest.mark.parametrize("n_var,n_hap", [(700, 5008), (2000, 198), (2000, 5008), (2500, 5008), (2500, 198)])
def test_triangle_mma_deferred_list_overflow_is_settled_in_place(ctx, n_var, n_hap):
    """Pairs the single-precision screen cannot settle go onto a list (CTA-wide in the single-wave kernel, <= 148
    tiles; global for the follow-up kernel otherwise).  With the list capped at 3 entries nearly all of them take
    the overflow paths -- settled by the lane / warp that found them -- and the result must not change."""
    from ld_tools_b200._lib import TUNE_DEFER_CAP
    from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC, threshold_e4
    st, planes, mask = make_store(ctx, n_var, n_hap, seed=900 + n_var)
    rows = np.arange(n_var)
    t = threshold_e4(0.3)
    ref_packed, ref_n11 = st.triangle(rows, measure="d_prime", thres_e4_=t, engine=ENGINE_POPC, want_n11=True)
    ctx.set_tuning(TUNE_DEFER_CAP, 3)
    try:
        packed, n11 = st.triangle(rows, measure="d_prime", thres_e4_=t, engine=ENGINE_MMA, want_n11=True)
        plain, _ = st.triangle(rows, engine=ENGINE_MMA)
    finally:
        ctx.set_tuning(TUNE_DEFER_CAP, 0)
    assert (n11 == ref_n11).all()
    assert (packed == ref_packed).all()
    assert (plain == (ref_packed & ~np.uint32(0x40000000))).all()
    st.close()


@pytest.mark.parametrize("n_var,n_hap,sel_frac", [(513, 5008, None), (2000, 5008, 0.2), (2600, 5008, None), (2600, 198, None)])
def test_triangle_mma_cta_pair_kernel_bit_exact(ctx, n_var, n_hap, sel_frac):
    """LDX_TUNE_MMA_PAIR: 256 x 128 tiles on CTA pairs (tcgen05.mma.cta_group::2; the peer's wideners report to their own
    barrier and forwarder warps relay each completed stage to the leader).  Same counts and words as the popcount engine,
    for single-wave and multi-wave calls."""
    from ld_tools_b200._lib import TUNE_MMA_PAIR, TUNE_MMA_TILE_N
    from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC, threshold_e4
    st, planes, mask = make_store(ctx, n_var, n_hap, seed=1200 + n_var, sel_frac=sel_frac)
    rows = np.random.default_rng(n_var).permutation(n_var)
    t = threshold_e4(0.05)
    ref_packed, ref_n11 = st.triangle(rows, thres_e4_=t, engine=ENGINE_POPC, want_n11=True)
    ctx.set_tuning(TUNE_MMA_PAIR, 1)
    ctx.set_tuning(TUNE_MMA_TILE_N, 128)
    try:
        packed, n11 = st.triangle(rows, thres_e4_=t, engine=ENGINE_MMA, want_n11=True)
        plain, _ = st.triangle(rows, engine=ENGINE_MMA)
        part, _ = st.triangle_rows(rows, 256, n_var, engine=ENGINE_MMA)
    finally:
        ctx.set_tuning(TUNE_MMA_PAIR, 0)
        ctx.set_tuning(TUNE_MMA_TILE_N, 0)
    assert (n11 == ref_n11).all()
    assert (packed == ref_packed).all()
    assert (plain == (ref_packed & ~np.uint32(0x40000000))).all()
    assert (part == plain[256 * 255 // 2:]).all()
    st.close()


def test_triangle_mma_full_size_and_threshold(ctx):
    from ld_tools_b200.engine import BELOW_THRES, ENGINE_MMA, r2_e4, threshold_e4
    st, planes, mask = make_store(ctx, 2000, 5008, seed=2)
    rows = np.arange(2000)
    packed, n11 = st.triangle(rows, engine=ENGINE_MMA, want_n11=True)
    bits = ld_oracle.unpack_bits(planes, 5008).astype(np.float32)
    gram = (bits @ bits.T).astype(np.int64)
    r, c = np.tril_indices(2000, -1)
    assert (n11 == gram[r, c]).all()
    n1 = np.diag(gram)
    assert (packed == ld_oracle.packed_words(5008, n11, n1[r], n1[c])).all()
    t = threshold_e4(0.8)
    flagged, _ = st.triangle(rows, measure="r_square", thres_e4_=t, engine=ENGINE_MMA)
    assert ((flagged & ~np.uint32(BELOW_THRES)) == packed).all()
    assert (((flagged & BELOW_THRES) != 0) == (r2_e4(packed) < t)).all()
    st.close()


# ------------------------------------------------------------------ fp64 finalisation at scale

def test_finalise_counts_golden(golden, ctx):
    """Every golden count case through the count-level entry point: rounded values identical,
    D and D' bit-exact, r2 within 1e-12 (bit-exact unless the reference's pow rounds the other way)."""
    from ld_tools_b200.engine import dprime_value, r2_value
    by_n = {}
    for tag, n, n11, a, b, *out in golden["count_cases"]:
        by_n.setdefault(n, []).append(((n11, a, b), out))
    for n, items in by_n.items():
        c = np.array([x for x, _ in items], dtype=np.int32)
        res = ctx.finalise_counts(n, c[:, 0], c[:, 1], c[:, 2])
        for i, (_, out) in enumerate(items):
            assert repr(r2_value(res["packed"][i])) == out[0] and repr(dprime_value(res["packed"][i])) == out[1]
            raw_r2, raw_dp, raw_d = decode_raw(out[4]), decode_raw(out[5]), decode_raw(out[8])
            assert float(res["d"][i]).hex() == float(raw_d).hex()
            if not isinstance(raw_dp, int):
                assert float(res["dprime"][i]).hex() == raw_dp.hex()
            if not isinstance(raw_r2, int):
                assert abs(res["r2"][i] - raw_r2) <= TOL


@pytest.mark.parametrize("n_hap", [2, 7, 198, 1006, 5008, 131072])
def test_finalise_counts_random_vs_oracle(ctx, n_hap):
    """Millions of random count triples: the branch-free division / rounding path must agree with
    the C oracle (IEEE division, libm pow, printf-rounding) on every packed word."""
    rng = np.random.default_rng(n_hap)
    n = 1_500_000 if n_hap >= 198 else 20000
    a = rng.integers(0, n_hap + 1, size=n)
    b = rng.integers(0, n_hap + 1, size=n)
    rare = rng.random(n) < 0.4
    a[rare] = rng.integers(0, min(n_hap, 30) + 1, size=int(rare.sum()))
    rare = rng.random(n) < 0.3
    b[rare] = rng.integers(0, min(n_hap, 30) + 1, size=int(rare.sum()))
    lo, hi = np.maximum(0, a + b - n_hap), np.minimum(a, b)
    n11 = lo + (rng.random(n) * (hi - lo + 1)).astype(np.int64)
    n11 = np.minimum(n11, hi)
    full = rng.random(n) < 0.2
    n11[full] = hi[full]
    res = ctx.finalise_counts(n_hap, n11, a, b)
    want = ld_oracle.packed_words(n_hap, n11, a, b)
    bad = np.flatnonzero(res["packed"] != want)
    assert bad.size == 0, (bad[:5], n11[bad[:5]], a[bad[:5]], b[bad[:5]])
    sub = rng.choice(n, 3000, replace=False)
    exact = ld_oracle.finalise_many(n_hap, n11[sub], a[sub], b[sub])
    assert (res["d"][sub] == exact["d"]).all()
    assert (res["dprime"][sub] == exact["dprime"]).all()
    assert np.abs(res["r2"][sub] - exact["r2"]).max() <= TOL


# ------------------------------------------------------------------ multi-GPU shard unit: row ranges of the triangle

@pytest.mark.parametrize("engine_name", ["popc", "mma"])
def test_triangle_rows_are_slices_of_the_full_triangle(ctx, engine_name):
    """ldx_triangle_rows (the unit ld_tools_b200/shard.py hands to each GPU) against the oracle and
    against the unsharded call: every range is the contiguous slice [tri(begin), tri(end))."""
    from ld_tools_b200 import shard
    from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC, threshold_e4
    engine = ENGINE_MMA if engine_name == "mma" else ENGINE_POPC
    n_var, n_hap = 700, 5008
    st, planes, mask = make_store(ctx, n_var, n_hap, seed=77, sel_frac=0.6)
    rows = np.random.default_rng(3).permutation(n_var)
    want = ld_oracle.triangle(planes, mask, n_hap, rows)
    want_packed = ld_oracle.packed_of(want)
    full, full_n11 = st.triangle(rows, engine=engine, want_n11=True)
    assert (full == want_packed).all() and (full_n11 == want["n_11"]).all()
    ranges = [(0, 128), (128, 384), (384, 700), (256, 256), (0, 700), (640, 700), (0, 1)]
    for world in (2, 3, 8):
        ranges += shard.triangle_row_ranges(n_var, world)
    for a, b in ranges:
        part, part_n11 = st.triangle_rows(rows, a, b, engine=engine, want_n11=True)
        sl = shard.triangle_slice(a, b)
        assert part.shape[0] == sl.stop - sl.start
        assert (part == want_packed[sl]).all(), (a, b)
        assert (part_n11 == want["n_11"][sl]).all(), (a, b)
    # the threshold flag travels with the slice
    t = threshold_e4(0.3)
    part, _ = st.triangle_rows(rows, 128, 512, thres_e4_=t, engine=engine)
    sl = shard.triangle_slice(128, 512)
    below = (want_packed[sl] & 0x3FFF) < t
    assert (((part & 0x40000000) != 0) == below).all()
    with pytest.raises(Exception):
        st.triangle_rows(rows, 64, 700, engine=engine)          # begin must sit on a 128-row panel
    st.close()


def test_kernel_timing_counts_the_all_pairs_launches(ctx):
    st, planes, mask = make_store(ctx, 300, 5008, seed=5)
    rows = np.arange(300)
    ctx.kernel_timing(True)
    for _ in range(3):
        st.triangle(rows)
    ms, n = ctx.kernel_timing(False)
    assert n == 3 and 0.0 < ms < 50.0
    st.triangle(rows)
    ms, n = ctx.kernel_timing(False)
    assert n == 0 and ms == 0.0
    st.close()


def test_several_device_resident_calls_one_resolve(ctx):
    """Calls enqueued back to back (each with its own result buffer) and settled by ONE ldx_resolve():
    the near-tie records of every call must find their way to the right buffer.  Small haplotype
    counts make exact rounding ties (r2 * 10^4 = k + 1/2) common.  (Same variant count as the 5008-
    haplotype test before it: the cached tile list must not be taken for this call's.)"""
    import torch
    from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC, threshold_e4
    n_var, n_hap = 600, 198
    st, planes, mask = make_store(ctx, n_var, n_hap, seed=12)
    dev = torch.device("cuda", 0)
    calls = []
    rng = np.random.default_rng(8)
    for k, (engine, measure, thres) in enumerate([(ENGINE_MMA, "r_square", None), (ENGINE_POPC, "d_prime", 0.5),
                                                  (ENGINE_MMA, "r_square", 0.1), (ENGINE_MMA, "d_prime", None)]):
        rows = rng.permutation(n_var)[: 300 + 100 * (k % 3)]
        v = len(rows)
        out = torch.zeros(v * (v - 1) // 2, dtype=torch.int32, device=dev)
        t = None if thres is None else threshold_e4(thres)
        torch.cuda.synchronize()
        st.triangle_dev(rows, out.data_ptr(), measure=measure, thres_e4_=t, engine=engine)
        calls.append((rows, out, measure, t))
    n_fixed = ctx.resolve()
    for rows, out, measure, t in calls:
        want = ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, rows))
        if t is not None:
            shift = 0 if measure == "r_square" else 16
            want = want | np.where(((want >> shift) & 0x3FFF) < t, np.uint32(0x40000000), np.uint32(0))
        got = out.cpu().numpy().view(np.uint32)
        assert (got == want).all(), (measure, t, int((got != want).sum()))
    assert n_fixed > 0, "this input is expected to contain rounding near-ties"
    # a host-buffer call right after still sees a clean list
    packed, _ = st.triangle(np.arange(200))
    assert (packed == ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, np.arange(200)))).all()
    st.close()


# ------------------------------------------------------------------ matrix text (ld_triangle.py:351-360 on the GPU)
def _python_body(packed, v, measure, prefixes, row_begin=0, row_end=None):
    """The same lines from the decoded words, with Python's own str() per cell."""
    from ld_tools_b200.engine import BELOW_THRES, measure_value, tri_index
    row_end = v if row_end is None else row_end
    out = []
    for r in range(row_begin, row_end):
        base = tri_index(r, 0) - tri_index(row_begin, 0)
        cells = ["0" if (c >= r or packed[base + c] & BELOW_THRES) else str(measure_value(packed[base + c], measure))
                 for c in range(v)]
        out.append(prefixes[r] + "\t".join(cells).encode() + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("n_hap,measure,thres", [(198, "r_square", None), (198, "d_prime", 0.3), (1006, "r_square", 0.05),
                                                 (14, "d_prime", None)])
def test_triangle_text_matches_reference_writer(ctx, n_hap, measure, thres):
    """ldx_triangle_text against the reference's writer restated (oracle/table_port.py), fed with the dicts the
    oracle's calc_ld returns for every pair: int 0 vs 0.0, the rounded threshold, row > col only, str() per cell."""
    from oracle import table_port
    from ld_tools_b200.engine import ENGINE_POPC, threshold_e4
    n_var = 75
    st, planes, mask = make_store(ctx, n_var, n_hap, seed=300 + n_hap)
    rows = np.random.default_rng(n_hap).permutation(n_var)
    res = ld_oracle.triangle(planes, mask, n_hap, rows)
    vals = [ld_oracle.as_reference_dict(x)[measure] for x in res]
    ids = [f"rs{1000 + 7 * k}" for k in range(n_var)]
    poss = [str(16050000 + 913 * k) for k in range(n_var)]
    want = table_port.matrix_body(lambda r, c: vals[r * (r - 1) // 2 + c], n_var, ids, poss, thres).encode()
    packed, _ = st.triangle(rows, measure=measure, thres_e4_=None if thres is None else threshold_e4(thres), engine=ENGINE_POPC)
    prefixes = [(i + "\t" + p + "\t").encode() for i, p in zip(ids, poss)]
    got = ctx.triangle_text(packed, n_var, measure, prefixes)
    assert got.tobytes() == want
    st.close()


def test_triangle_text_full_size_slabs_and_device_words(ctx):
    """configs[1] shape (2,000 x 2,000 cells): host words = device words = slabs concatenated = Python's str() per cell;
    exact size query; empty prefixes; a matrix wider than one 2,048-cell chunk."""
    import torch
    from ld_tools_b200 import _lib
    from ld_tools_b200.engine import threshold_e4, tri_index
    v = 2500
    st, planes, mask = make_store(ctx, v, 1006, seed=77)
    rows = np.arange(v)
    prefixes = [b"rs%d\t%d\t" % (3 * k + 1, 17000000 + 41 * k) if k % 97 else b"" for k in range(v)]
    for measure, thres in (("r_square", None), ("d_prime", 0.9)):
        packed, _ = st.triangle(rows, measure=measure, thres_e4_=None if thres is None else threshold_e4(thres))
        want = _python_body(packed, v, measure, prefixes)
        whole = ctx.triangle_text(packed, v, measure, prefixes)
        assert whole.tobytes() == want
        # the same words left in HBM by the device-resident call, text kept in HBM too
        dev = torch.zeros(packed.shape[0], dtype=torch.int32, device="cuda:0")
        torch.cuda.synchronize()
        st.triangle_dev(rows, dev.data_ptr(), measure=measure, thres_e4_=None if thres is None else threshold_e4(thres))
        text_dev = torch.zeros(len(want) + 5, dtype=torch.uint8, device="cuda:0")
        torch.cuda.synchronize()
        n = ctx.triangle_text(dev.data_ptr(), v, measure, prefixes, dev_text=(text_dev.data_ptr(), text_dev.numel()))
        assert n == len(want) and text_dev[:n].cpu().numpy().tobytes() == want and int(text_dev[n:].sum()) == 0
        # slabs of rows (unaligned boundaries) concatenate to the whole text
        parts = []
        for r0, r1 in ((0, 1), (1, 2), (2, 701), (701, 2499), (2499, 2500)):
            parts.append(ctx.triangle_text(packed[tri_index(r0, 0):tri_index(r1, 0)], v, measure, prefixes, r0, r1).tobytes())
        assert b"".join(parts) == want
        # size query / too-small buffer
        with pytest.raises(_lib.LdxError) as e:
            ctx.triangle_text(packed, v, measure, prefixes, out=np.zeros(len(want) - 1, dtype=np.uint8))
        assert e.value.code == _lib.ERR_CAPACITY
        exact = np.zeros(len(want), dtype=np.uint8)
        assert ctx.triangle_text(packed, v, measure, prefixes, out=exact).tobytes() == want
    st.close()


def test_triangle_text_tiny_matrices(ctx):
    for v in (1, 2, 3):
        st, planes, mask = make_store(ctx, 8, 198, seed=v)
        packed, _ = st.triangle(np.arange(v))
        prefixes = [b"a\t1\t"] * v
        assert ctx.triangle_text(packed, v, "r_square", prefixes).tobytes() == _python_body(packed, v, "r_square", prefixes)
        st.close()
    assert ctx.triangle_text(np.zeros(0, np.uint32), 5, "r_square", [b""] * 5, 2, 2).shape[0] == 0


@pytest.mark.parametrize("n_var,n_hap,measure,thres", [(700, 198, "r_square", None), (700, 5008, "d_prime", 0.4),
                                                        (130, 1006, "r_square", 0.02), (1, 198, "r_square", None)])
def test_triangle_table_is_triangle_plus_text(ctx, n_var, n_hap, measure, thres):
    """ldx_triangle_table (all-pairs kernel + settlement + writer, words never leave HBM) = the two separate calls,
    whole and in slabs of 256 rows; small haplotype counts make near-ties (settled on the host) common."""
    from ld_tools_b200.engine import threshold_e4
    st, planes, mask = make_store(ctx, max(n_var, 8), n_hap, seed=500 + n_var + n_hap)
    rows = np.random.default_rng(n_var).permutation(max(n_var, 8))[:n_var]
    t = None if thres is None else threshold_e4(thres)
    prefixes = [b"rs%d\t%d\t" % (k, 5 * k) for k in range(n_var)]
    packed, _ = st.triangle(rows, measure=measure, thres_e4_=t)
    want = ctx.triangle_text(packed, n_var, measure, prefixes).tobytes()
    assert want == _python_body(packed, n_var, measure, prefixes)
    assert st.triangle_table(rows, prefixes, measure, t).tobytes() == want
    parts = [st.triangle_table(rows, prefixes, measure, t, row_begin=r0, row_end=min(n_var, r0 + 256)).tobytes()
             for r0 in range(0, n_var, 256)]
    assert b"".join(parts) == want
    st.close()


def test_async_upload_and_no_wait_mask_give_the_same_results(ctx):
    """ldx_store_upload_async + ldx_store_set_mask (which no longer waits for its own copies) + a blocking result call, repeated
    with changing planes and masks from pinned and pageable memory: every step's result is that step's input's -- nothing is
    read from a staging buffer that a later call has already overwritten."""
    import torch
    from ld_tools_b200 import Store
    from ld_tools_b200.synth import random_planes
    n_var, n_hap = 700, 1500
    st = Store(ctx, n_var, n_hap)
    rows = np.arange(n_var, dtype=np.int64)
    rng = np.random.default_rng(5)
    pinned = torch.empty((n_var, st.stride_words), dtype=torch.int64).pin_memory()
    for step in range(6):
        planes = random_planes(n_var, n_hap, seed=100 + step)
        sel = np.sort(rng.choice(n_hap, int(rng.integers(200, n_hap)), replace=False))
        mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
        if step % 2 == 0:
            pinned.numpy().view("<u8")[:] = planes
            st.upload(0, pinned.numpy().view("<u8"), wait=False)
        else:
            st.upload(0, planes.copy(), wait=False)                      # pageable: staged by the runtime before the call returns
        st.set_mask(mask)
        st.set_mask(mask)                                                # twice in a row: the second waits for the first's staging
        got = st.triangle_values(rows, "r_square")
        want = ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, rows))
        assert (got == ((want & 0xBFFF) | ((want >> 16) & 0x4000)).astype(np.uint16)).all(), step
        n1, _, n_sel = st.counts()
        assert n_sel == len(sel) and (n1 == ld_oracle.variant_counts(planes, mask, n_hap)).all()
    st.close()


def test_window_large_hit_lists_come_back_sorted(ctx):
    """With threshold 0 every scanned pair is kept (half a million here): the list is sorted by (query, row) on the device
    before it crosses PCIe -- strictly increasing keys, as many hits as pairs scanned, every query's rows exactly its candidates
    that pass the filters, words equal to the oracle's for sampled hits."""
    n_var, n_hap = 4000, 400
    st, planes, mask, pos0, end0, idnum, elig = annotated_store(ctx, n_var, n_hap, seed=91)
    rng = np.random.default_rng(4)
    q_rows = np.sort(rng.choice(n_var, 150, replace=False)).astype(np.int64)
    lo = np.zeros(len(q_rows), dtype=np.int64)
    hi = np.full(len(q_rows), n_var, dtype=np.int64)
    ws = np.zeros(len(q_rows), dtype=np.int32)
    we = np.full(len(q_rows), 2**31 - 1, dtype=np.int32)
    hits, scanned = st.window(q_rows, lo, hi, ws, we, "r_square", 0)
    assert len(hits) == scanned > 400_000
    key = hits["query"].astype(np.int64) * n_var + hits["row"]
    assert (np.diff(key) > 0).all()
    for k in (0, 77, 149):
        q = int(q_rows[k])
        want_rows = [r for r in range(n_var) if elig[r] and idnum[r] != idnum[q]]
        assert hits["row"][hits["query"] == k].tolist() == want_rows
    pick = rng.choice(len(hits), 300, replace=False)
    ref = ld_oracle.pairs(planes, mask, n_hap, q_rows[hits["query"][pick]], hits["row"][pick].astype(np.int64))
    assert (hits["packed"][pick] == ld_oracle.packed_of(ref)).all() and (hits["n11"][pick] == ref["n_11"]).all()
    st.close()
