"""Multi-GPU partitioning of the LD hot path: pure index math plus one gather.

The path shards into independent units with no exchange during compute:

* ld_triangle (ld_triangle.py:133-230): the lower triangle is cut into contiguous ROW RANGES whose
  tile counts are balanced.  Row r owns r pairs, so equal-work boundaries grow like sqrt(k / world);
  each range is a contiguous slice [tri(begin), tri(end)) of the packed triangle, so gathering the
  shards is plain concatenation.  Every rank holds the whole variant set (64 MB of bit planes at
  100,000 variants) and calls Store.triangle_rows(_dev) on its range.
* ld_area (ld_area.py:152-292): the position-sorted store is cut into contiguous genomic SLABS
  balanced by the number of candidate pairs; a rank loads its slab plus a halo of one flank on both
  sides and scans the queries whose position falls into the slab.  The kept hits (variable length)
  are the only thing exchanged: one all_gather of counts, one of padded records (NCCL on GPU
  tensors, gloo on CPU tensors in the tests).
  A whole-genome job (one store per chromosome) is cut the same way across chromosome boundaries
  (genome_pieces): a rank holds, per chromosome it touches, only the rows its queries' windows reach.
* ld_lite: replicas only.
"""
import numpy as np

from ._lib import HIT_DTYPE

TRI_ALIGN = 256          # row ranges start on the CTA-pair kernel's 256-row panels (ldx_triangle_rows itself needs 128)


def tri(r):
    """Pairs (row, col < row) owned by matrix rows 0..r-1 = offset of row r in the packed triangle."""
    r = int(r)
    return r * (r - 1) // 2 if r > 0 else 0


def triangle_row_ranges(v, world, align=TRI_ALIGN, tile=128):
    """[(begin, end)] * world covering rows 0..v, begin % align == 0, balanced by tile count.

    Cost model: the all-pairs kernel works in 128 x 128 tiles and a tile costs the same K loop
    whether it is full or cut by the diagonal, so the 128-row panel s costs s + 1 tiles.  Boundaries
    are the `align`-row panel indices where the cumulative cost crosses k / world."""
    v, world = int(v), int(world)
    assert world >= 1 and align % tile == 0
    sub = align // tile                                                  # 128-row panels per boundary unit
    n_sub = (v + tile - 1) // tile
    n_panels = (n_sub + sub - 1) // sub
    sub_cost = np.arange(1, n_panels * sub + 1, dtype=np.float64)
    sub_cost[n_sub:] = 0.0                                               # panels beyond the matrix
    cost = sub_cost.reshape(n_panels, sub).sum(axis=1) if n_panels else np.zeros(0)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        b = int(np.searchsorted(cum, target, side="left"))
        # choose the nearer of the two panel boundaries around the target
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[min(b, n_panels)] - target):
            b -= 1
        b = min(max(b, bounds[-1]), n_panels)
        bounds.append(b)
    bounds.append(n_panels)
    rows = [min(b * align, v - v % align) for b in bounds]               # interior boundaries stay aligned even past the end
    rows[-1] = v
    return [(rows[k], rows[k + 1]) for k in range(world)]


def triangle_slice(begin, end):
    """Slice of the packed lower triangle produced by rows begin..end-1."""
    return slice(tri(begin), tri(end))


def window_bounds(pos0, end0_max_len, q_pos, flank):
    """Candidate row ranges [lo, hi) and 0-based half-open windows of ld_area queries.

    pos0: sorted POS-1 of the store rows; q_pos: 1-based POS of the queries (ld_area.py:174-177:
    low = max(0, pos - flank), high = pos + flank, then fetch(chrom, low, high)).  A record
    overlaps the window iff pos0 < high and end0 > low; since end0 <= pos0 + end0_max_len, rows
    with pos0 <= low - end0_max_len can never overlap: lo is the first row above that."""
    pos0 = np.asarray(pos0, dtype=np.int64)
    q_pos = np.asarray(q_pos, dtype=np.int64)
    ws = np.maximum(q_pos - flank, 0)
    we = q_pos + flank
    lo = np.searchsorted(pos0, ws - int(end0_max_len), side="right")
    hi = np.searchsorted(pos0, we, side="left")
    hi = np.maximum(hi, lo)
    return lo.astype(np.int64), hi.astype(np.int64), ws.astype(np.int32), we.astype(np.int32)


def area_slabs(pos0, end0_max_len, q_row, q_pos, flank, world):
    """Region sharding of an ld_area job.  -> list of dicts, one per rank:

        row_begin, row_end    store rows the rank must hold (its slab plus the halo)
        own_begin, own_end    the slab proper: queries with q_row in [own_begin, own_end) are the rank's
        queries               indices into the job's query arrays

    Slab boundaries balance the number of candidate pairs (sum of hi - lo over the owned queries),
    not base pairs: query and variant density vary along the chromosome."""
    pos0 = np.asarray(pos0, dtype=np.int64)
    q_row = np.asarray(q_row, dtype=np.int64)
    n = pos0.shape[0]
    lo, hi, _, _ = window_bounds(pos0, end0_max_len, q_pos, flank)
    order = np.argsort(q_row, kind="stable")
    work = (hi - lo)[order].astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(work)])
    out = []
    cuts = [0]
    for k in range(1, world):
        cuts.append(int(np.searchsorted(cum, cum[-1] * k / world, side="left")))
    cuts.append(order.shape[0])
    cuts = [min(max(c, 0), order.shape[0]) for c in cuts]
    for k in range(world):
        a, b = cuts[k], max(cuts[k + 1], cuts[k])
        mine = order[a:b]
        own_begin = int(q_row[order[a]]) if a < order.shape[0] and k > 0 else 0
        own_end = int(q_row[order[b]]) if b < order.shape[0] and k < world - 1 else n
        if mine.size:
            row_begin, row_end = int(lo[mine].min()), int(hi[mine].max())
            row_begin = min(row_begin, int(q_row[mine].min()))
            row_end = max(row_end, int(q_row[mine].max()) + 1)
        else:
            row_begin = row_end = own_begin
        out.append({"row_begin": row_begin, "row_end": row_end, "own_begin": own_begin, "own_end": own_end,
                    "queries": mine})
    return out


def genome_pieces(chroms, world):
    """Region sharding of a whole-genome ld_area job (one store per chromosome, ld_area.py:152 loops over them).

    chroms: per chromosome a dict with the position-sorted query arrays q_row, lo, hi (window_bounds()).  The
    genome-wide query list, ordered by (chromosome, position), is cut into `world` contiguous pieces with equal
    candidate-pair counts.  -> per rank a list of dicts, one per chromosome the rank touches:

        chrom               index into chroms
        qa, qb              the rank's queries of that chromosome: q_row[qa:qb]
        row_begin, row_end  the chromosome rows the rank must hold (slab + halo of its queries' windows)
    """
    work = np.concatenate([np.asarray(ch["hi"], dtype=np.int64) - np.asarray(ch["lo"], dtype=np.int64) for ch in chroms]).astype(np.float64)
    first = np.concatenate([[0], np.cumsum([len(ch["q_row"]) for ch in chroms])]).astype(np.int64)
    cum = np.concatenate([[0.0], np.cumsum(work)])
    cuts = [int(np.searchsorted(cum, cum[-1] * k / world, side="left")) for k in range(world)] + [len(work)]
    out = []
    for k in range(world):
        a, b = cuts[k], max(cuts[k + 1], cuts[k])
        pieces = []
        for c, ch in enumerate(chroms):
            qa, qb = int(max(a, first[c]) - first[c]), int(min(b, first[c + 1]) - first[c])
            if qb <= qa:
                continue
            q_row, lo, hi = np.asarray(ch["q_row"]), np.asarray(ch["lo"]), np.asarray(ch["hi"])
            pieces.append({"chrom": c, "qa": qa, "qb": qb,
                           "row_begin": int(min(lo[qa:qb].min(), q_row[qa:qb].min())),
                           "row_end": int(max(hi[qa:qb].max(), q_row[qa:qb].max() + 1))})
        out.append(pieces)
    return out


def rebase_queries(slab, q_row, lo, hi):
    """Query arrays of one rank relative to its local store (global row r -> r - row_begin)."""
    idx = slab["queries"]
    off = slab["row_begin"]
    return (np.asarray(q_row)[idx] - off, np.asarray(lo)[idx] - off, np.asarray(hi)[idx] - off)


def globalise_hits(hits, slab):
    """Local hit records -> job-wide numbering (query index of the job, store row of the full store)."""
    out = np.array(hits, dtype=HIT_DTYPE, copy=True)
    if out.shape[0]:
        out["query"] = slab["queries"][out["query"]]
        out["row"] += slab["row_begin"]
    return out


def gather_hits(hits, group=None, device=None):
    """All ranks' hit lists, concatenated and sorted by (query, row), on every rank.

    The one collective of the path: all_gather of the counts, then all_gather of the records padded
    to the longest list (16-byte records as int32 x 4).  With the NCCL backend pass the rank's CUDA
    device; with gloo leave device=None (CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
    n = torch.tensor([hits.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros((cap, 4), dtype=torch.int32, device=device)
    if hits.shape[0]:
        buf[:hits.shape[0]] = torch.from_numpy(hits.view(np.int32).reshape(-1, 4)).to(buf.device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    rec = [p[:c].cpu().numpy().reshape(-1).view(HIT_DTYPE) for p, c in zip(parts, counts) if c]
    allh = np.concatenate(rec) if rec else np.zeros(0, dtype=HIT_DTYPE)
    return allh[np.lexsort((allh["row"], allh["query"]))]


def gather_hits_tensor(local, group=None, ranks_own_ordered_query_ranges=False):
    """The same collective with the records left where they are: `local` is an int32 tensor [n, 4] (ldx_hit records as
    {query, row, n11, packed}) on the rank's CUDA device (NCCL) or on the CPU (gloo); returns every rank's records
    concatenated and sorted by (query, row) as a tensor on the same device.  Every rank sorts ITS records (the work that
    shrinks with the number of ranks), then one all_gather of the counts and one of the records padded to the longest list;
    nothing but the counts crosses to the host.  ranks_own_ordered_query_ranges: rank r's queries all come before rank
    r + 1's (genome_pieces, area_slabs) -- then the concatenation in rank order is the sorted list and no global sort runs."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = local.reshape(-1, 4).contiguous()
    work = None
    if world > 1:                                           # the counts travel while this rank sorts its records
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        counts_t = torch.empty(world, dtype=torch.int64, device=local.device)
        work = dist.all_gather_into_tensor(counts_t, n, group=group, async_op=True)
    if local.shape[0] > 1:
        local = local[torch.argsort((local[:, 0].to(torch.int64) << 32) | local[:, 1].to(torch.int64))]
    if world > 1:
        work.wait()
        counts = counts_t.tolist()                             # the one host synchronisation of the gather
        cap = max(max(counts), 1)
        buf = torch.zeros((cap, 4), dtype=torch.int32, device=local.device)
        buf[:local.shape[0]] = local
        parts = torch.empty((world * cap, 4), dtype=torch.int32, device=local.device)      # rank r's records at rows r * cap ...
        dist.all_gather_into_tensor(parts, buf, group=group)
        allh = torch.cat([parts[r * cap:r * cap + c] for r, c in enumerate(counts)])
        if not ranks_own_ordered_query_ranges:
            allh = allh[torch.argsort((allh[:, 0].to(torch.int64) << 32) | allh[:, 1].to(torch.int64))]
    else:
        allh = local
    return allh
