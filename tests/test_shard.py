"""Multi-GPU partition logic on the CPU: the shard arithmetic is pure index
math, checked here against the oracle; the one collective (gather of kept hits) runs under gloo
with world_size 2.  The GPU halves of the same paths are in test_parity_gpu.py."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ld_tools_b200 import shard  # noqa: E402
from ld_tools_b200._lib import HIT_DTYPE  # noqa: E402
from ld_tools_b200.synth import make_records, synth_haplotypes  # noqa: E402
from oracle import ld_oracle  # noqa: E402


# ------------------------------------------------------------------ ld_triangle: row ranges
@pytest.mark.parametrize("v,world", [(2000, 1), (2000, 2), (2000, 8), (100_000, 8), (100_000, 4), (130, 8), (1, 2), (0, 4),
                                     (5000, 3)])
def test_triangle_row_ranges_cover_and_align(v, world):
    rr = shard.triangle_row_ranges(v, world)
    assert len(rr) == world
    assert rr[0][0] == 0 and rr[-1][1] == v
    for k, (a, b) in enumerate(rr):
        assert 0 <= a <= b <= v
        if k:
            assert a == rr[k - 1][1] and a % shard.TRI_ALIGN == 0
    # the slices tile the packed triangle exactly
    assert sum(shard.tri(b) - shard.tri(a) for a, b in rr) == shard.tri(v)


def test_triangle_row_ranges_balance_at_config4():
    """BASELINE configs[3]: 100,000 variants on 8 GPUs -- tile counts within 2% of each other."""
    rr = shard.triangle_row_ranges(100_000, 8)
    tiles = []
    for a, b in rr:
        panels = np.arange(a // 128, (b + 127) // 128)
        tiles.append(int((panels + 1).sum()))
    assert max(tiles) / (sum(tiles) / 8) < 1.02, tiles


def test_triangle_shards_concatenate_to_the_oracle_triangle():
    """Rank k computes the triangle of rows[:end] restricted to rows >= begin; the concatenation of
    the slices is the full packed triangle (what ldx_triangle_rows implements on the GPU)."""
    n_var, n_hap = 300, 198
    h = synth_haplotypes(n_var, n_hap, seed=5)
    planes = ld_oracle.pack_bits(h)
    mask = ld_oracle.mask_from_haplotypes(np.arange(n_hap), n_hap)
    rows = np.arange(n_var)
    full = ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, rows))
    parts = []
    for a, b in shard.triangle_row_ranges(n_var, 3):
        if b < 2 or a == b:
            parts.append(np.zeros(0, np.uint32))
            continue
        sub = ld_oracle.packed_of(ld_oracle.triangle(planes, mask, n_hap, rows[:b]))
        parts.append(sub[shard.tri(a):])
    assert (np.concatenate(parts) == full).all()


# ------------------------------------------------------------------ ld_area: region slabs
def annotated(n_var, n_hap, seed):
    h = synth_haplotypes(n_var, n_hap, seed=seed)
    planes = ld_oracle.pack_bits(h)
    mask = ld_oracle.mask_from_haplotypes(np.arange(n_hap), n_hap)
    recs = make_records(n_var, seed=seed, mean_gap=40)
    pos0 = np.array([r["pos"] - 1 for r in recs], dtype=np.int32)
    end0 = pos0 + np.array([len(r["ref"]) for r in recs], dtype=np.int32)
    elig = np.array([bool(re.match(r"rs\d+$", r["id"])) and not r["multi"] for r in recs], dtype=np.uint8)
    idnum = np.array([int(r["id"][2:]) if re.match(r"rs\d+$", r["id"]) else -1 - i for i, r in enumerate(recs)], dtype=np.int64)
    return planes, mask, pos0, end0, idnum, elig


def oracle_hits(planes, mask, n_hap, pos0, end0, idnum, elig, q_row, lo, hi, ws, we, thres):
    out = []
    for k in range(len(q_row)):
        rows, res = ld_oracle.window(planes, mask, n_hap, pos0, end0, idnum, elig, int(q_row[k]), int(ws[k]), int(we[k]), 0, thres,
                                     lo=int(lo[k]), hi=int(hi[k]))
        rec = np.zeros(len(rows), dtype=HIT_DTYPE)
        rec["query"], rec["row"], rec["n11"], rec["packed"] = k, rows, res["n_11"], ld_oracle.packed_of(res)
        out.append(rec)
    return np.concatenate(out) if out else np.zeros(0, HIT_DTYPE)


def area_job(seed=9, n_var=1500, n_hap=198, n_q=40, flank=4000):
    planes, mask, pos0, end0, idnum, elig = annotated(n_var, n_hap, seed)
    rng = np.random.default_rng(seed)
    q_row = np.sort(rng.choice(np.flatnonzero(elig), n_q, replace=False))
    q_row = rng.permutation(q_row)                       # create_src_dict order is NOT position order
    q_pos = pos0[q_row].astype(np.int64) + 1
    max_len = int((end0 - pos0).max())
    lo, hi, ws, we = shard.window_bounds(pos0, max_len, q_pos, flank)
    return dict(planes=planes, mask=mask, n_hap=n_hap, pos0=pos0, end0=end0, idnum=idnum, elig=elig, q_row=q_row, q_pos=q_pos,
                lo=lo, hi=hi, ws=ws, we=we, max_len=max_len, flank=flank)


def rank_hits(job, slab, thres=0.2):
    """What one rank computes: its slab + halo as a LOCAL store, its queries rebased, hits globalised."""
    a, b = slab["row_begin"], slab["row_end"]
    q, lo, hi = shard.rebase_queries(slab, job["q_row"], job["lo"], job["hi"])
    idx = slab["queries"]
    local = oracle_hits(job["planes"][a:b], job["mask"], job["n_hap"], job["pos0"][a:b], job["end0"][a:b], job["idnum"][a:b],
                        job["elig"][a:b], q, lo, hi, job["ws"][idx], job["we"][idx], thres)
    return shard.globalise_hits(local, slab)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_area_slabs_reproduce_the_unsharded_scan(world):
    job = area_job()
    want = oracle_hits(job["planes"], job["mask"], job["n_hap"], job["pos0"], job["end0"], job["idnum"], job["elig"], job["q_row"],
                       job["lo"], job["hi"], job["ws"], job["we"], 0.2)
    want = want[np.lexsort((want["row"], want["query"]))]
    slabs = shard.area_slabs(job["pos0"], job["max_len"], job["q_row"], job["q_pos"], job["flank"], world)
    assert sorted(np.concatenate([s["queries"] for s in slabs]).tolist()) == list(range(len(job["q_row"])))
    got = np.concatenate([rank_hits(job, s) for s in slabs])
    got = got[np.lexsort((got["row"], got["query"]))]
    assert len(want) > 0 and (got == want).all()
    # every slab holds its queries' candidate rows: the halo is exactly what the windows need
    for s in slabs:
        idx = s["queries"]
        if idx.size:
            assert s["row_begin"] <= job["lo"][idx].min() and job["hi"][idx].max() <= s["row_end"]
    if world > 1:   # the halo keeps the per-rank stores well below the whole store
        assert max(s["row_end"] - s["row_begin"] for s in slabs) < len(job["pos0"])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_genome_pieces_partition_a_multi_chromosome_job(world):
    """BASELINE configs[4] in miniature: the genome-wide query list is cut across chromosome boundaries; every
    query lands on exactly one rank, whose row range of that chromosome covers the query's whole window."""
    rng = np.random.default_rng(world)
    chroms = []
    for c, nv in enumerate([5000, 1200, 3100, 40]):
        pos0 = np.sort(rng.integers(0, 40 * nv, size=nv))
        q_row = np.sort(rng.choice(nv, max(1, nv // 50), replace=False))
        lo, hi, _, _ = shard.window_bounds(pos0, 1, pos0[q_row] + 1, 2000)
        chroms.append({"q_row": q_row, "lo": lo, "hi": hi})
    plan = shard.genome_pieces(chroms, world)
    assert len(plan) == world
    seen = [np.zeros(len(ch["q_row"]), dtype=int) for ch in chroms]
    work = []
    for pieces in plan:
        w = 0
        assert [p["chrom"] for p in pieces] == sorted(p["chrom"] for p in pieces)      # contiguous in genome order
        for p in pieces:
            ch = chroms[p["chrom"]]
            seen[p["chrom"]][p["qa"]:p["qb"]] += 1
            assert p["row_begin"] <= ch["lo"][p["qa"]:p["qb"]].min() and p["row_end"] >= ch["hi"][p["qa"]:p["qb"]].max()
            assert p["row_begin"] <= ch["q_row"][p["qa"]] and p["row_end"] > ch["q_row"][p["qb"] - 1]
            w += int((ch["hi"] - ch["lo"])[p["qa"]:p["qb"]].sum())
        work.append(w)
    assert all((s == 1).all() for s in seen)
    total = sum(int((ch["hi"] - ch["lo"]).sum()) for ch in chroms)
    assert sum(work) == total
    biggest = max(int((ch["hi"] - ch["lo"]).max()) for ch in chroms)
    assert max(work) <= total / world + biggest                                         # balanced to within one window


def test_window_bounds_match_reference_flank_rule():
    pos0 = np.array([9, 99, 100, 149, 150, 151, 400], dtype=np.int64)
    lo, hi, ws, we = shard.window_bounds(pos0, 3, np.array([150, 5]), 50)
    assert ws.tolist() == [100, 0] and we.tolist() == [200, 55]          # ld_area.py:174-177, clamped at 0
    assert hi.tolist() == [6, 1]                                         # rows with pos0 < win_end
    assert lo.tolist() == [1, 0]                                         # pos0 > win_start - max_len


# ------------------------------------------------------------------ the one collective, under gloo
def _gloo_worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    job = area_job()
    slabs = shard.area_slabs(job["pos0"], job["max_len"], job["q_row"], job["q_pos"], job["flank"], world)
    mine = rank_hits(job, slabs[rank])
    allh = shard.gather_hits(mine)
    np.save(os.path.join(tmpdir, f"hits_{rank}.npy"), allh)
    # the tensor form of the same collective (what bench.py times on NCCL): records stay tensors, sorted by (query, row)
    import torch
    from ld_tools_b200._lib import HIT_DTYPE
    local = torch.from_numpy(np.ascontiguousarray(mine).view(np.int32).reshape(-1, 4).copy())      # rank_hits: already job-wide numbering
    ten = shard.gather_hits_tensor(local).numpy().reshape(-1).view(HIT_DTYPE)
    np.save(os.path.join(tmpdir, f"hits_tensor_{rank}.npy"), ten)
    # With the queries numbered in position order (as genome_pieces numbers them), rank r's queries all come before rank
    # r + 1's: the cheaper form -- every rank sorts its own records, no global sort -- returns the same list
    pos_rank = torch.from_numpy(np.argsort(np.argsort(job["q_row"])).astype(np.int32))
    renum = local.clone()
    renum[:, 0] = pos_rank[local[:, 0].long()]
    full = shard.gather_hits_tensor(renum)
    shuffled = renum[torch.randperm(renum.shape[0], generator=torch.Generator().manual_seed(rank))]
    fast = shard.gather_hits_tensor(shuffled, ranks_own_ordered_query_ranges=True)
    assert torch.equal(full, fast) and full.shape[0] == ten.shape[0]
    dist.barrier()
    dist.destroy_process_group()


def test_gather_hits_world2_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    job = area_job()
    want = oracle_hits(job["planes"], job["mask"], job["n_hap"], job["pos0"], job["end0"], job["idnum"], job["elig"], job["q_row"],
                       job["lo"], job["hi"], job["ws"], job["we"], 0.2)
    want = want[np.lexsort((want["row"], want["query"]))]
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"hits_{rank}.npy"))
        assert (got == want).all()
        assert (np.load(os.path.join(str(tmp_path), f"hits_tensor_{rank}.npy")) == want).all()
