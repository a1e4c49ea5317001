"""Randomised differential run of the kernels on a B200: random shapes, masks, orders and thresholds, every path against
another path and against the oracle (bit for bit).  Not part of the test suite (it runs for as long as it is told to):

    python tools/fuzz_gpu.py [--seconds 60] [--seed 1]

  window    multi-query kernels (row-per-thread on 128-byte rows, 8-lanes-per-row on wider ones) vs the one-query kernel
            vs the oracle's full scan, random annotations (ineligible rows, repeated IDs, long intervals), both measures
  triangle  tcgen05 engine (tile 64 / 128 / pair / direct on-off) vs the popcount engine vs the oracle, arbitrary row orders,
            thresholds; 2-byte values and threshold hit lists vs the packed words; batches of random sets
  pairs     ldx_pairs vs the oracle
  general   random GT text with haploid / missing / unphased / other-allele fields: K1's general parser and the general route of
            the pair, all-pairs (both engines) and window kernels vs calc_ld on the lists the reference's drivers would build
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401

from ld_tools_b200 import Context, Store  # noqa: E402
from ld_tools_b200._lib import (BELOW_THRES, PAIR_HIT_DTYPE, TUNE_MMA_DIRECT, TUNE_MMA_PAIR, TUNE_MMA_TILE_N,  # noqa: E402
                                TUNE_WINDOW_MQ)
from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC, threshold_e4  # noqa: E402
from oracle import ld_oracle  # noqa: E402


def random_store(ctx, rng, n_var, n_hap):
    freq = rng.choice([0.0, 0.002, 0.05, 0.3, 0.5, 0.7, 0.97, 1.0], size=n_var, p=[0.02, 0.1, 0.2, 0.2, 0.16, 0.2, 0.1, 0.02])
    h = (rng.random((n_var, n_hap)) < freq[:, None]).astype(np.uint8)
    for i in range(1, n_var):
        u = rng.random()
        if u < 0.3:
            h[i] = h[i - 1] ^ (rng.random(n_hap) < rng.choice([0.0, 0.003, 0.05]))
        elif u < 0.35:
            h[i] = 1 - h[i - 1]
    planes = ld_oracle.pack_bits(h)
    st = Store.from_planes(ctx, planes, n_hap)
    k = int(rng.integers(1, n_hap + 1)) if rng.random() < 0.7 else n_hap
    sel = np.sort(rng.choice(n_hap, k, replace=False))
    mask = ld_oracle.mask_from_haplotypes(sel, n_hap)
    st.set_mask(mask)
    return st, planes, mask, sel


def fuzz_window(ctx, rng):
    n_hap = int(rng.choice([int(rng.integers(2, 1025)), int(rng.integers(1025, 2049)), int(rng.integers(2049, 4097)), 5008, int(rng.integers(4097, 7000)),
                            int(rng.integers(2, 300))]))
    n_var = int(rng.integers(50, 3000))
    st, planes, mask, sel = random_store(ctx, rng, n_var, n_hap)
    pos0 = np.sort(rng.integers(1000, 1000 + 40 * n_var, size=n_var)).astype(np.int32)
    end0 = (pos0 + np.where(rng.random(n_var) < 0.1, rng.integers(2, 60, size=n_var), 1)).astype(np.int32)
    idnum = np.arange(n_var, dtype=np.int64) + 7
    dup = rng.choice(n_var, max(1, n_var // 40), replace=False)
    idnum[dup] = idnum[(dup + 3) % n_var]
    elig = (rng.random(n_var) < 0.9).astype(np.uint8)
    st.set_annotations(pos0, end0, idnum, elig)
    nq = int(rng.integers(1, 80))
    q_row = np.sort(rng.choice(n_var, min(nq, n_var), replace=False)).astype(np.int64)
    flank = int(rng.choice([200, 2000, 20000]))
    ws = np.maximum(pos0[q_row].astype(np.int64) + 1 - flank, 0).astype(np.int32)
    we = (pos0[q_row].astype(np.int64) + 1 + flank).astype(np.int32)
    max_len = int((end0 - pos0).max())
    lo = np.searchsorted(pos0, ws.astype(np.int64) - max_len, side="right").astype(np.int64)
    hi = np.maximum(np.searchsorted(pos0, we, side="left"), lo).astype(np.int64)
    measure = str(rng.choice(["r_square", "d_prime"]))
    thres = float(rng.choice([0.0, 0.0001, 0.05, 0.5, 0.8, 1.0]))
    t = threshold_e4(thres)
    got, scanned = st.window(q_row, lo, hi, ws, we, measure, t)
    ctx.set_tuning(TUNE_WINDOW_MQ, 0)
    try:
        one, scanned1 = st.window(q_row, lo, hi, ws, we, measure, t)
    finally:
        ctx.set_tuning(TUNE_WINDOW_MQ, 1)
    assert scanned == scanned1 and len(got) == len(one) and (got == one).all(), ("window mq vs 1q", n_hap, n_var, nq, measure, thres)
    if len(sel) <= 1024 and rng.random() < 0.7:                 # the same through a subset store (row-per-thread kernel)
        sub = st.subset(sel)
        sub.set_annotations(pos0, end0, idnum, elig)
        g2, s2 = sub.window(q_row, lo, hi, ws, we, measure, t)
        assert s2 == scanned and len(g2) == len(got) and (g2 == got).all(), ("window subset", n_hap, len(sel), n_var, nq, measure, thres)
        sub.close()
    for k in rng.choice(len(q_row), min(3, len(q_row)), replace=False):
        rows, res = ld_oracle.window(planes, mask, n_hap, pos0, end0, idnum, elig, int(q_row[k]), int(ws[k]), int(we[k]),
                                     0 if measure == "r_square" else 1, thres, lo=int(lo[k]), hi=int(hi[k]))
        mine = got[got["query"] == k]
        assert mine["row"].tolist() == rows.tolist() and (mine["packed"] == ld_oracle.packed_of(res)).all(), ("window oracle", n_hap, n_var, int(k), measure, thres)
    st.close()
    return f"window n_hap={n_hap} sel={len(sel)} v={n_var} nq={len(q_row)} {measure} {thres}"


def fuzz_triangle(ctx, rng):
    n_hap = int(rng.choice([int(rng.integers(2, 400)), int(rng.integers(400, 8193)), 5008]))
    n_var = int(rng.integers(2, 1400))
    st, planes, mask, sel = random_store(ctx, rng, n_var, n_hap)
    v = int(rng.integers(2, n_var + 1))
    rows = rng.permutation(n_var)[:v].astype(np.int64)
    if rng.random() < 0.3:
        rows = np.sort(rows)
    if rng.random() < 0.2:
        a = int(rng.integers(0, n_var - v + 1))
        rows = np.arange(a, a + v, dtype=np.int64)          # contiguous: direct mode
    measure = str(rng.choice(["r_square", "d_prime"]))
    t = None if rng.random() < 0.5 else threshold_e4(float(rng.choice([0.0001, 0.1, 0.5, 0.9])))
    want = ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, rows))
    if t is not None:
        shift = 0 if measure == "r_square" else 16
        want = want | np.where(((want >> shift) & 0x3FFF) < t, np.uint32(BELOW_THRES), np.uint32(0))
    popc, _ = st.triangle(rows, measure, t, engine=ENGINE_POPC)
    assert (popc == want).all(), ("popc", n_hap, len(sel), v, measure, t)
    ctx.set_tuning(TUNE_MMA_TILE_N, int(rng.choice([0, 64, 128])))
    ctx.set_tuning(TUNE_MMA_PAIR, int(rng.choice([-1, 0, 1])))
    ctx.set_tuning(TUNE_MMA_DIRECT, int(rng.choice([-1, 0, 1])))
    try:
        mma, _ = st.triangle(rows, measure, t, engine=ENGINE_MMA)
        assert (mma == want).all(), ("mma", n_hap, len(sel), v, measure, t, int((mma != want).sum()))
        vals = st.triangle_values(rows, measure, t, engine=ENGINE_MMA)
        narrow = (want >> 16).astype(np.uint16) if measure == "d_prime" else ((want & 0xBFFF) | ((want >> 16) & 0x4000)).astype(np.uint16)
        assert (vals == narrow).all(), ("values16", n_hap, v, measure, t)
        if t is not None:
            hits = st.triangle_hits(rows, measure, t, engine=ENGINE_MMA)
            keep = np.flatnonzero((want & BELOW_THRES) == 0)
            assert len(hits) == len(keep), ("hits count", n_hap, v, measure, t, len(hits), len(keep))
            idx = hits["row"].astype(np.int64) * (hits["row"] - 1) // 2 + hits["col"]
            assert (idx == keep).all() and (hits["packed"] == want[keep]).all(), ("hits", n_hap, v, measure, t)
    finally:
        ctx.set_tuning(TUNE_MMA_TILE_N, 0); ctx.set_tuning(TUNE_MMA_PAIR, -1); ctx.set_tuning(TUNE_MMA_DIRECT, -1)
    # a batch of random sets of the same store
    import torch
    sets = [rng.permutation(n_var)[: int(rng.integers(2, n_var + 1))].astype(np.int64) for _ in range(int(rng.integers(2, 6)))]
    outs = [torch.zeros(max(len(r) * (len(r) - 1) // 2, 1), dtype=torch.int32, device="cuda:0") for r in sets]
    ctx.triangle_batch_dev([(st, r, o.data_ptr()) for r, o in zip(sets, outs)], measure, t)
    ctx.resolve()
    torch.cuda.synchronize()
    for r, o in zip(sets, outs):
        w = ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, r))
        if t is not None:
            shift = 0 if measure == "r_square" else 16
            w = w | np.where(((w >> shift) & 0x3FFF) < t, np.uint32(BELOW_THRES), np.uint32(0))
        n = len(r) * (len(r) - 1) // 2
        assert (o[:n].cpu().numpy().view(np.uint32) == w).all(), ("batch", n_hap, len(r), measure, t)
    # random pairs
    ia, ib = rng.integers(0, n_var, 500), rng.integers(0, n_var, 500)
    out = st.pairs(ia, ib)
    ref = ld_oracle.pairs(planes, mask, n_hap, ia, ib)
    assert (out["packed"] == ld_oracle.packed_of(ref)).all() and (out["n11"] == ref["n_11"]).all(), ("pairs", n_hap, n_var)
    st.close()
    return f"triangle n_hap={n_hap} sel={len(sel)} v={v} {measure} {t}"


def fuzz_general(ctx, rng):
    """Genotype text with haploid / missing / unphased / other-allele fields through K1's general parser and every kernel's
    general route, against calc_ld on the lists the reference's drivers would build (helpers of tests/test_general_gpu.py)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import test_general_gpu as tg
    from ld_tools_b200 import shard
    n_samples = int(rng.choice([int(rng.integers(8, 60)), int(rng.integers(60, 520)), int(rng.integers(520, 1300))]))
    n_var = int(rng.integers(20, 140))
    style = str(rng.choice(["autosome", "chrx"]))
    rows = tg.make_rows(rng, n_var, n_samples, style)
    st, status = tg.store_from_rows(ctx, rows, n_samples)
    assert not (status & 8).any()
    k = int(rng.integers(1, n_samples + 1))
    sel = np.sort(rng.choice(n_samples, k, replace=False))
    st.select_haplotypes(np.concatenate([2 * sel, 2 * sel + 1]))
    lists = tg.lists_for(rows, sel)
    if any(len(g) == 0 for g in lists):                      # an empty pairing is a ZeroDivisionError in the reference: not this tool's subject
        st.close()
        return f"general skipped (an empty list) samples={n_samples}"
    n1, ln, kind, n_gen = st.row_counts()
    assert ln.tolist() == [len(g) for g in lists] and n1.tolist() == [g.count(1) for g in lists], ("row_counts", n_samples, style)
    want = np.zeros(n_var * (n_var - 1) // 2, dtype=np.uint32)
    for r in range(1, n_var):
        for c in range(r):
            want[r * (r - 1) // 2 + c] = tg.oracle_pair(lists[r], lists[c])[1]
    idx = np.arange(n_var)
    for engine in (ENGINE_POPC, ENGINE_MMA):
        got, _ = st.triangle(idx, engine=engine)
        assert (got == want).all(), ("general triangle", engine, n_samples, k, style, int((got != want).sum()))
    ia, ib = rng.integers(0, n_var, 200), rng.integers(0, n_var, 200)
    got = st.pairs(ia, ib)
    for j in range(200):
        res, word = tg.oracle_pair(lists[ia[j]], lists[ib[j]])
        assert got["packed"][j] == word and abs(got["r2"][j] - res["r2"]) <= 1e-12 and abs(got["dprime"][j] - res["dprime"]) <= 1e-12, ("general pairs", n_samples, style, j)
    pos0 = (np.arange(n_var) * 10 + 100).astype(np.int32)
    st.set_annotations(pos0, pos0 + 1, np.arange(n_var, dtype=np.int64), np.ones(n_var, np.uint8))
    q_row = np.sort(rng.choice(n_var, min(12, n_var), replace=False)).astype(np.int64)
    lo, hi, ws, we = shard.window_bounds(pos0, 1, pos0[q_row].astype(np.int64) + 1, 300)
    measure = str(rng.choice(["r_square", "d_prime"]))
    te = threshold_e4(float(rng.choice([0.0, 0.05, 0.5])))
    hits, _ = st.window(q_row, lo, hi, ws, we, measure, te)
    for kq, q in enumerate(q_row):
        mine = hits[hits["query"] == kq]
        exp_rows, exp_words = [], []
        for j in range(int(lo[kq]), int(hi[kq])):
            if j == q or not (pos0[j] < we[kq] and pos0[j] + 1 > ws[kq]):
                continue
            word = tg.oracle_pair(lists[q], lists[j])[1]
            val = (word & 0x3FFF) if measure == "r_square" else ((word >> 16) & 0x3FFF)
            if val >= te:
                exp_rows.append(j)
                exp_words.append(word)
        assert mine["row"].tolist() == exp_rows and mine["packed"].tolist() == exp_words, ("general window", n_samples, style, int(q), measure, te)
    st.close()
    return f"general {style} samples={n_samples} sel={k} v={n_var} rows with aux planes={int((kind >= 0).sum())}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    ctx = Context(0)
    rng = np.random.default_rng(a.seed)
    t0, n = time.time(), 0
    while time.time() - t0 < a.seconds:
        case_seed = int(rng.integers(1 << 62))
        r = np.random.default_rng(case_seed)
        try:
            what = (fuzz_window, fuzz_triangle, fuzz_general)[n % 3](ctx, r)
        except AssertionError as e:
            print(f"FAIL case {n} seed {case_seed}: {e.args}", flush=True)
            raise
        n += 1
        if n % 20 == 0:
            print(f"{n} cases, last: {what}", flush=True)
    print(f"fuzz ok: {n} cases in {time.time() - t0:.0f} s")
    ctx.close()


if __name__ == "__main__":
    main()
