"""The tcgen05 all-pairs engine's round-2 modes against the popcount engine and the oracle (bit for bit):
minor-allele operands (variants whose alt allele is the major one enter complemented), direct mode (no gather kernel:
the planes through a TMA tensor map), several variant sets per launch (ldx_triangle_batch_dev).  Run on a B200."""
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


def mixed_frequency_store(ctx, n_var, n_hap, seed, sel_frac=None):
    """Variants over the whole frequency range: rare, common, alt-major (n1 > N/2), exact halves, all-ref and all-alt rows,
    with neighbours in LD (copies with a few flips) so that high r2 / D' values and exact zeros of D occur."""
    from ld_tools_b200 import Store
    rng = np.random.default_rng(seed)
    freq = rng.choice([0.001, 0.01, 0.2, 0.5, 0.8, 0.99, 0.999], size=n_var)
    h = (rng.random((n_var, n_hap)) < freq[:, None]).astype(np.uint8)
    for i in range(1, n_var):
        k = rng.random()
        if k < 0.25:
            h[i] = h[i - 1] ^ (rng.random(n_hap) < 0.002)
        elif k < 0.35:
            h[i] = 1 - h[i - 1]                                 # complementary: D < 0, D' = 1
    h[rng.integers(0, n_var, max(1, n_var // 50))] = 1         # monomorphic alt
    h[rng.integers(0, n_var, max(1, n_var // 50))] = 0         # monomorphic ref
    half = rng.integers(0, n_var)
    h[half] = 0
    h[half, rng.permutation(n_hap)[: n_hap // 2]] = 1           # exactly N/2 (of all haplotypes)
    planes = ld_oracle.pack_bits(h)
    st = Store.from_planes(ctx, planes, n_hap)
    sel = np.arange(n_hap) if sel_frac is None else np.sort(rng.choice(n_hap, int(n_hap * sel_frac), replace=False))
    mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
    st.set_mask(mask)
    return st, planes, mask


MODES = [("tile64", dict(tile=64)), ("tile128", dict(tile=128, pair=0)), ("pair", dict(tile=128, pair=1)),
         ("direct_off", dict(tile=128, pair=0, direct=0))]


def run_mma(ctx, st, rows, mode, **kw):
    from ld_tools_b200._lib import TUNE_MMA_DIRECT, TUNE_MMA_PAIR, TUNE_MMA_TILE_N
    from ld_tools_b200.engine import ENGINE_MMA
    ctx.set_tuning(TUNE_MMA_TILE_N, mode.get("tile", 0))
    ctx.set_tuning(TUNE_MMA_PAIR, mode.get("pair", -1))
    ctx.set_tuning(TUNE_MMA_DIRECT, mode.get("direct", -1))
    try:
        return st.triangle(rows, engine=ENGINE_MMA, **kw)
    finally:
        ctx.set_tuning(TUNE_MMA_TILE_N, 0)
        ctx.set_tuning(TUNE_MMA_PAIR, -1)
        ctx.set_tuning(TUNE_MMA_DIRECT, -1)


@pytest.mark.parametrize("name,mode", MODES)
@pytest.mark.parametrize("n_var,n_hap,sel_frac", [(300, 5008, None), (700, 5008, 0.3), (2000, 1006, None), (2600, 198, None),
                                                  (1500, 8192, None), (2100, 5008, 0.7)])
def test_major_allele_variants_all_kernels(ctx, name, mode, n_var, n_hap, sel_frac):
    """Counts and words equal the popcount engine's for variants of every frequency class, in matrix order (gather path)
    and on contiguous rows (direct path when the call is one wave), with and without a threshold."""
    from ld_tools_b200.engine import BELOW_THRES, ENGINE_POPC, threshold_e4
    st, planes, mask = mixed_frequency_store(ctx, n_var, n_hap, seed=n_var + n_hap, sel_frac=sel_frac)
    t = threshold_e4(0.2)
    for rows in (np.arange(n_var), np.random.default_rng(n_var).permutation(n_var)):
        ref_packed, ref_n11 = st.triangle(rows, measure="d_prime", thres_e4_=t, engine=ENGINE_POPC, want_n11=True)
        packed, n11 = run_mma(ctx, st, rows, mode, measure="d_prime", thres_e4_=t, want_n11=True)
        assert (n11 == ref_n11).all()
        assert (packed == ref_packed).all()
        plain, _ = run_mma(ctx, st, rows, mode)
        assert (plain == (ref_packed & ~np.uint32(BELOW_THRES))).all()
    if n_var <= 700:                                                               # `rows` / `plain`: the permuted pass
        want = ld_oracle.triangle(planes, mask, n_hap, rows)
        assert (plain == ld_oracle.packed_of(want)).all()
    st.close()


@pytest.mark.parametrize("start,n_var,n_store", [(0, 2000, 2000), (137, 1800, 2500), (600, 1400, 2000), (1, 1025, 1026), (5, 300, 4000)])
@pytest.mark.parametrize("sel_frac", [None, 0.4])
def test_direct_mode_equals_gather_mode(ctx, start, n_var, n_store, sel_frac):
    """rows = start .. start + n_var - 1 of a larger store: the direct kernel (tensor map, rows past the store's end are
    zero-filled by the TMA unit) and the gather kernel give the same words; both equal the popcount engine's."""
    from ld_tools_b200.engine import ENGINE_POPC
    st, planes, mask = mixed_frequency_store(ctx, n_store, 5008, seed=start + n_var, sel_frac=sel_frac)
    rows = np.arange(start, start + n_var)
    ref, ref_n11 = st.triangle(rows, engine=ENGINE_POPC, want_n11=True)
    launches0 = ctx.launch_count
    direct, n11 = run_mma(ctx, st, rows, dict(tile=128, pair=0, direct=1), want_n11=True)
    launches_direct = ctx.launch_count - launches0
    gathered, _ = run_mma(ctx, st, rows, dict(tile=128, pair=0, direct=0))
    launches_gather = ctx.launch_count - launches0 - launches_direct
    assert (direct == ref).all() and (gathered == ref).all() and (n11 == ref_n11).all()
    assert launches_direct == 1 and launches_gather == 2          # one wave: no gather kernel, no follow-up kernel
    # a slice of the same triangle (row ranges are what the multi-GPU path shards by)
    from ld_tools_b200._lib import TUNE_MMA_DIRECT, TUNE_MMA_TILE_N
    from ld_tools_b200.engine import ENGINE_MMA
    if n_var > 256:
        ctx.set_tuning(TUNE_MMA_TILE_N, 128)
        try:
            part, _ = st.triangle_rows(rows, 256, n_var, engine=ENGINE_MMA)
        finally:
            ctx.set_tuning(TUNE_MMA_TILE_N, 0)
        assert (part == ref[256 * 255 // 2:]).all()
    st.close()


def test_batch_equals_single_calls(ctx):
    """ldx_triangle_batch_dev: sets of different sizes from two stores (same sample selection size) in one launch; every
    set's words and counts equal a single-set call's; one ldx_resolve() settles the near-ties of all of them."""
    import torch
    from ld_tools_b200.engine import ENGINE_POPC, threshold_e4
    st_a, _, _ = mixed_frequency_store(ctx, 2300, 5008, seed=1)
    st_b, _, _ = mixed_frequency_store(ctx, 900, 5008, seed=2)
    rng = np.random.default_rng(0)
    sets = [(st_a, np.arange(2000)), (st_b, rng.permutation(900)[:700]), (st_a, rng.permutation(2300)[:1300]), (st_b, np.arange(3)),
            (st_a, np.arange(100, 101)), (st_b, np.arange(0, 513)), (st_a, np.arange(2300)), (st_a, np.arange(0))]
    dev = torch.device("cuda", 0)
    for measure, thres in (("r_square", None), ("d_prime", threshold_e4(0.5))):
        outs = [torch.zeros(max(len(r) * (len(r) - 1) // 2, 1), dtype=torch.int32, device=dev) for _, r in sets]
        n11s = [torch.zeros(max(len(r) * (len(r) - 1) // 2, 1), dtype=torch.int32, device=dev) for _, r in sets]
        launches0 = ctx.launch_count
        ctx.triangle_batch_dev([(s, r, o.data_ptr(), q.data_ptr()) for (s, r), o, q in zip(sets, outs, n11s)], measure=measure, thres_e4_=thres)
        ctx.resolve()
        torch.cuda.synchronize()
        assert ctx.launch_count - launches0 <= 4          # gather + all-pairs + deferred pairs (+ the near-tie scatter)
        for (s, r), o, q in zip(sets, outs, n11s):
            n = len(r) * (len(r) - 1) // 2
            ref, ref_n11 = s.triangle(r, measure=measure, thres_e4_=thres, engine=ENGINE_POPC, want_n11=True)
            assert (o.cpu().numpy().view(np.uint32)[:n] == ref).all(), (len(r), measure)
            assert (q.cpu().numpy()[:n] == ref_n11).all(), (len(r), measure)
    st_a.close()
    st_b.close()


def test_batch_of_many_small_sets_and_mixed_selections(ctx):
    """More sets than one launch takes (32), and stores whose selections differ in size (separate launches)."""
    import torch
    from ld_tools_b200.engine import ENGINE_POPC
    st_a, _, _ = mixed_frequency_store(ctx, 600, 1006, seed=11)
    st_b, _, _ = mixed_frequency_store(ctx, 600, 1006, seed=12, sel_frac=0.5)
    rng = np.random.default_rng(3)
    sets = []
    for k in range(40):
        s = st_b if k in (7, 8, 30) else st_a
        sets.append((s, rng.permutation(600)[: int(rng.integers(2, 400))]))
    dev = torch.device("cuda", 0)
    outs = [torch.zeros(len(r) * (len(r) - 1) // 2, dtype=torch.int32, device=dev) for _, r in sets]
    ctx.triangle_batch_dev([(s, r, o.data_ptr()) for (s, r), o in zip(sets, outs)])
    ctx.resolve()
    torch.cuda.synchronize()
    for (s, r), o in zip(sets, outs):
        ref, _ = s.triangle(r, engine=ENGINE_POPC)
        assert (o.cpu().numpy().view(np.uint32) == ref).all(), len(r)
    st_a.close()
    st_b.close()


def test_same_rows_for_a_smaller_store_are_checked_again(ctx):
    """the staged row list is cached per (store, size); the same list for a smaller store must fail cleanly."""
    from ld_tools_b200 import LdxError
    from ld_tools_b200.engine import ENGINE_MMA
    big, _, _ = mixed_frequency_store(ctx, 800, 198, seed=21)
    small, _, _ = mixed_frequency_store(ctx, 400, 198, seed=22)
    rows = np.arange(800)
    big.triangle(rows, engine=ENGINE_MMA)
    with pytest.raises(LdxError):
        small.triangle(rows, engine=ENGINE_MMA)
    ok, _ = small.triangle(np.arange(400), engine=ENGINE_MMA)
    assert ok.shape[0] == 400 * 399 // 2
    big.close()
    small.close()


# ------------------------------------------------------------------ narrow host outputs (ldx_triangle_values / ldx_triangle_hits)
def _narrow(words, measure):
    words = words.astype(np.uint32)
    if measure == "d_prime":
        return (words >> 16).astype(np.uint16)
    return ((words & 0xBFFF) | ((words >> 16) & 0x4000)).astype(np.uint16)


@pytest.mark.parametrize("n_var,n_hap", [(700, 198), (2000, 5008), (131, 70)])
def test_two_byte_values_and_threshold_hit_lists_equal_the_packed_triangle(ctx, n_var, n_hap):
    """2 bytes per pair of one measure, and only the pairs above a threshold, are the packed triangle narrowed / filtered --
    with small haplotype counts (198) the near-ties settled on the host enter, leave and change both outputs."""
    from ld_tools_b200._lib import BELOW_THRES, PAIR_HIT_DTYPE, V16_BELOW, V16_INT0, V16_VALUE
    from ld_tools_b200.engine import ENGINE_AUTO, ENGINE_MMA, ENGINE_POPC, measure_value, threshold_e4, tri_index
    if n_hap == 198:                                  # founder haplotypes: few distinct count tables, exact rounding ties among them
        from ld_tools_b200 import Store
        from ld_tools_b200.synth import synth_haplotypes
        planes = ld_oracle.pack_bits(synth_haplotypes(n_var, n_hap, seed=12))
        st = Store.from_planes(ctx, planes, n_hap)
        st.select_all()
    else:
        st, planes, mask = mixed_frequency_store(ctx, n_var, n_hap, seed=n_var)
    rng = np.random.default_rng(n_var)
    rows = np.sort(rng.permutation(n_var)[: n_var - 3])
    v = len(rows)
    for measure, thres, engine in [("r_square", None, ENGINE_AUTO), ("d_prime", None, ENGINE_MMA), ("r_square", 0.2, ENGINE_MMA),
                                   ("d_prime", 0.9, ENGINE_POPC), ("r_square", 0.0001, ENGINE_AUTO)]:
        t = None if thres is None else threshold_e4(thres)
        packed, _ = st.triangle(rows, measure, t, engine=engine)
        vals = st.triangle_values(rows, measure, t, engine=engine)
        assert vals.dtype == np.uint16 and (vals == _narrow(packed, measure)).all()
        # the decoded number a caller prints is the same from either form
        for i in rng.integers(0, len(vals), 50):
            w16 = int(vals[i])
            got = 0 if w16 & V16_INT0 else (w16 & V16_VALUE) / 10000
            assert got == measure_value(int(packed[i]), measure) and bool(w16 & V16_BELOW) == bool(int(packed[i]) & BELOW_THRES)
        if v > 256:                                   # a row range of the same triangle
            part = st.triangle_values(rows, measure, t, engine=engine, row_begin=128, row_end=v - 5)
            assert (part == vals[tri_index(128, 0):tri_index(v - 5, 0)]).all()
        if t is not None:
            hits = st.triangle_hits(rows, measure, t, engine=engine)
            keep = np.flatnonzero((packed & BELOW_THRES) == 0)
            r = np.floor((1 + np.sqrt(1 + 8.0 * keep)) / 2).astype(np.int64)
            r -= (r * (r - 1) // 2 > keep)
            r += ((r + 1) * r // 2 <= keep)
            want = np.zeros(len(keep), dtype=PAIR_HIT_DTYPE)
            want["row"], want["col"], want["packed"] = r, keep - r * (r - 1) // 2, packed[keep]
            assert len(hits) == len(want) and hits.tobytes() == want.tobytes(), (measure, thres, len(hits), len(want))
            assert len(want) > 0
            small = np.zeros(max(len(want) - 1, 1), dtype=PAIR_HIT_DTYPE)        # a buffer one entry short says how many there are
            with pytest.raises(Exception):
                st.triangle_hits(rows, measure, t, engine=engine, out=small)
    # near-ties did occur in the small-N case (otherwise the host-side patching above was not exercised)
    if n_hap == 198:
        import torch
        out = torch.zeros(v * (v - 1) // 2, dtype=torch.int32, device="cuda:0")
        st.triangle_dev(rows, out.data_ptr(), thres_e4_=threshold_e4(0.2))
        assert ctx.resolve() > 0
    assert len(st.triangle_values(rows[:1])) == 0 and len(st.triangle_hits(rows[:1], "r_square", 5000)) == 0
    st.close()
