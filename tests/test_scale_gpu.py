"""Full-size parity of the BASELINE configs the multi-GPU path is built for, inside `pytest -m gpu` (one B200):

  configs[3]  100,000-variant all-pairs triangle (5e9 pairs, CTA-pair tcgen05 kernel): 10^6 seeded pairs against numpy
              popcounts + the oracle's finalisation, the (alt, alt) count checksum of EVERY 128-row panel against a CPU
              restatement, and a rank's row range (shard.triangle_row_ranges, 8 ranks) against the same slice of the whole
  configs[4]  a genome-shaped ld_area job (22 chromosome stores, +/-1 Mb windows) cut by shard.genome_pieces: the pieces of
              4 ranks, run one after the other, reproduce the unsharded hit lists; whole windows against the oracle
"""
import numpy as np
import pytest

from oracle import ld_oracle

pytestmark = pytest.mark.gpu
N_HAP = 5008


@pytest.fixture(scope="module")
def ctx():
    from ld_tools_b200 import Context
    c = Context(0)
    yield c
    c.close()


def panel_checksums_cpu(planes, n_hap, tile=128):
    """sum over rows r of a 128-row panel and columns c < r of n11[r][c], for every panel, without forming the pairs:
    n11[r][c] = <a_r, a_c>, so the rows' sums against everything before the panel are a_r . (column counts so far) and the
    pairs inside the panel are the strict lower triangle of a 128 x 128 Gram matrix."""
    bits = ld_oracle.unpack_bits(planes, n_hap)
    v = bits.shape[0]
    counts = np.zeros(n_hap, dtype=np.int64)
    out = []
    for p0 in range(0, v, tile):
        b = bits[p0:p0 + tile].astype(np.float32)
        before = int((b.astype(np.int64) @ counts).sum())
        gram = (b @ b.T).astype(np.int64)
        out.append(before + int(np.tril(gram, -1).sum()))
        counts += bits[p0:p0 + tile].sum(axis=0, dtype=np.int64)
    return np.array(out, dtype=np.int64)


def test_config3_100k_variant_triangle_sampled_pairs_and_panel_checksums(ctx):
    import torch
    from ld_tools_b200 import Store, shard
    from ld_tools_b200.engine import ENGINE_AUTO
    from ld_tools_b200.synth import random_planes
    v = 100_000
    dev = torch.device("cuda", 0)
    planes = random_planes(v, N_HAP, seed=4)
    st = Store.from_planes(ctx, planes, N_HAP)
    st.select_all()
    rows = np.arange(v, dtype=np.int64)
    n_pairs = v * (v - 1) // 2
    packed = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    n11 = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    ctx.kernel_timing(True)
    st.triangle_dev(rows, packed.data_ptr(), dev_n11=n11.data_ptr(), engine=ENGINE_AUTO)
    n_fixed = ctx.resolve()
    torch.cuda.synchronize()
    ms, launches = ctx.kernel_timing(False)
    assert launches == 1 and n_fixed >= 0
    # ---- 10^6 seeded pairs: counts and packed words
    rng = np.random.default_rng(2026)
    n_chk = 1_000_000
    r = rng.integers(1, v, size=n_chk)
    c = (rng.random(n_chk) * r).astype(np.int64)
    r[:2000], c[:2000] = np.arange(v - 2000, v), np.arange(v - 2001, v - 1)        # the diagonal's neighbours in the last panels
    idx = torch.from_numpy(r * (r - 1) // 2 + c).to(dev)
    words = planes[:, : (N_HAP + 63) // 64]
    n1 = np.bitwise_count(words).sum(axis=1).astype(np.int32)
    want_n11 = np.bitwise_count(words[r] & words[c]).sum(axis=1).astype(np.int32)
    assert (n11[idx].cpu().numpy() == want_n11).all()
    assert (packed[idx].cpu().numpy().view(np.uint32) == ld_oracle.packed_words(N_HAP, want_n11, n1[r], n1[c])).all()
    # ---- (alt, alt) count checksum of every 128-row panel of the triangle (782 of them)
    want_sums = panel_checksums_cpu(planes, N_HAP)
    got_sums = np.array([int(n11[shard.tri(p0):shard.tri(min(p0 + 128, v))].sum(dtype=torch.int64).item()) for p0 in range(0, v, 128)], dtype=np.int64)
    assert (got_sums == want_sums).all(), np.flatnonzero(got_sums != want_sums)[:10]
    # ---- one rank's row range of an 8-GPU job is the same slice of the whole triangle, bit for bit
    del n11
    begin, end = shard.triangle_row_ranges(v, 8)[5]
    part = torch.empty(shard.tri(end) - shard.tri(begin), dtype=torch.int32, device=dev)
    st.triangle_rows_dev(rows, begin, end, part.data_ptr(), engine=ENGINE_AUTO)
    ctx.resolve()
    torch.cuda.synchronize()
    assert torch.equal(part, packed[shard.tri(begin):shard.tri(end)])
    st.close()


def test_config4_genome_pieces_reproduce_the_unsharded_scan(ctx):
    """A genome-shaped job small enough for a test (22 chromosomes, 2.2 M variants, 3,000 queries, +/-1 Mb): every rank's
    pieces hold only the rows its queries' windows reach; the union of 4 ranks' hits equals the unsharded scan of every
    chromosome; sampled whole windows equal the oracle."""
    import torch
    from ld_tools_b200 import Store, shard
    from ld_tools_b200._lib import BELOW_THRES, HIT_DTYPE, R2_MASK
    from ld_tools_b200.engine import threshold_e4
    from ld_tools_b200.synth import fill_store_grouped
    dev = torch.device("cuda", 0)
    chr_mb = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09,
              133.28, 114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82]
    n_variants, n_queries, flank, world = 2_200_000, 3000, 1_000_000, 4
    thres = threshold_e4(0.8)
    chroms = []
    for c, mb in enumerate(chr_mb):
        nv = int(round(n_variants * mb / sum(chr_mb)))
        nq = max(1, int(round(n_queries * mb / sum(chr_mb))))
        rng = np.random.default_rng(9000 + c)
        pos0 = np.sort(rng.integers(10_000, int(mb * 1e6), size=nv, dtype=np.int64)).astype(np.int32)
        q_row = np.sort(rng.choice(nv, nq, replace=False)).astype(np.int64)
        lo, hi, ws, we = shard.window_bounds(pos0, 1, pos0[q_row].astype(np.int64) + 1, flank)
        chroms.append({"nv": nv, "pos0": pos0, "q_row": q_row, "lo": lo, "hi": hi, "ws": ws, "we": we})
    q_first = np.concatenate([[0], np.cumsum([len(ch["q_row"]) for ch in chroms])])

    def scan(c, rb, re, qa, qb):
        """Rows rb..re-1 of chromosome c as a store; its queries qa..qb-1 -> hits in job-wide numbering."""
        ch = chroms[c]
        st = Store(ctx, re - rb, N_HAP)
        fill_store_grouped(st, dev, c, rb, re)
        pos0 = ch["pos0"][rb:re]
        st.set_annotations(pos0, pos0 + 1, (np.int64(c) << 32) + np.arange(rb, re, dtype=np.int64), np.ones(re - rb, np.uint8))
        st.select_all()
        hits, scanned = st.window(ch["q_row"][qa:qb] - rb, ch["lo"][qa:qb] - rb, ch["hi"][qa:qb] - rb, ch["ws"][qa:qb], ch["we"][qa:qb], "r_square", thres)
        h = np.array(hits, dtype=HIT_DTYPE, copy=True)
        h["query"] += qa + q_first[c]
        h["row"] += rb
        return st, h, scanned

    # ---- the sharded job: each rank's pieces, one rank after the other
    pieces_by_rank = shard.genome_pieces(chroms, world)
    sharded, work = [], []
    for rank in range(world):
        scanned_rank = 0
        for pc in pieces_by_rank[rank]:
            st, h, scanned = scan(pc["chrom"], pc["row_begin"], pc["row_end"], pc["qa"], pc["qb"])
            assert pc["row_end"] - pc["row_begin"] <= chroms[pc["chrom"]]["nv"]
            st.close()
            sharded.append(h)
            scanned_rank += scanned
        work.append(scanned_rank)
    sharded = np.concatenate(sharded)
    sharded = sharded[np.lexsort((sharded["row"], sharded["query"]))]
    assert max(work) <= 1.15 * (sum(work) / world)               # pieces are balanced by candidate pairs
    # ---- the unsharded job, chromosome by chromosome; whole windows of sampled queries against the oracle
    whole = []
    rng = np.random.default_rng(1)
    words = (N_HAP + 63) // 64
    for c, ch in enumerate(chroms):
        st, h, _ = scan(c, 0, ch["nv"], 0, len(ch["q_row"]))
        whole.append(h)
        if c % 7 == 0:
            k = int(rng.integers(len(ch["q_row"])))
            a, b, q = int(ch["lo"][k]), int(ch["hi"][k]), int(ch["q_row"][k])
            win = st.download(a, b - a)[:, :words]
            qr = st.download(q, 1)[0, :words]
            n1 = np.bitwise_count(win).sum(axis=1).astype(np.int32)
            n11 = np.bitwise_count(win & qr[None, :]).sum(axis=1).astype(np.int32)
            want_w = ld_oracle.packed_words(N_HAP, n11, np.full(b - a, int(np.bitwise_count(qr).sum()), dtype=np.int32), n1)
            keep = ((want_w & R2_MASK) >= thres) & (np.arange(a, b) != q)
            got = h[h["query"] == k + q_first[c]]
            assert got["row"].tolist() == (np.flatnonzero(keep) + a).tolist()
            assert (got["n11"] == n11[keep]).all() and ((got["packed"] & ~np.uint32(BELOW_THRES)) == want_w[keep]).all()
        st.close()
    whole = np.concatenate(whole)
    whole = whole[np.lexsort((whole["row"], whole["query"]))]
    assert len(whole) > n_queries and (sharded == whole).all()
