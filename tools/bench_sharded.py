"""BASELINE configs[3]: one large all-pairs triangle, row-range sharded over the GPUs of a node
(strong scaling).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_sharded.py --variants 100000 [--reps 3] [--check 200000]

Every rank holds the whole variant set (64 MB of bit planes at 100,000 variants), computes the row
range ld_tools_b200.shard.triangle_row_ranges() gives it with ONE ldx_triangle_rows_dev call and
keeps its slice of the packed triangle in HBM (20 GB / N at 100,000 variants).  No data-path
collective: the only communication is the max-over-ranks of the device time.  Parity at full size
is checked on a seeded sample of pairs per rank: (alt, alt) counts recomputed with numpy popcounts
and finalised by the oracle's C restatement must equal the words in HBM bit for bit.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=100_000)
    ap.add_argument("--n-hap", type=int, default=5008)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=200_000, help="pairs per rank verified against the oracle")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ld_tools_b200 import Context, Store, shard
    from ld_tools_b200.engine import ENGINE_AUTO
    from ld_tools_b200.synth import random_planes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    v, n_hap = args.variants, args.n_hap
    planes = random_planes(v, n_hap, seed=4)            # the same variant set on every rank
    st = Store.from_planes(ctx, planes, n_hap)
    st.select_all()
    rows = np.arange(v, dtype=np.int64)
    begin, end = shard.triangle_row_ranges(v, world)[rank]
    n_mine = shard.tri(end) - shard.tri(begin)
    out = torch.empty(max(n_mine, 1), dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        st.triangle_rows_dev(rows, begin, end, out.data_ptr(), engine=ENGINE_AUTO)
        ctx.resolve()

    step()
    barrier()
    ctx.kernel_timing(True)
    times = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        step()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))

    kern_ms, kern_n = ctx.kernel_timing(False)
    # ---- parity on a seeded sample of this rank's pairs
    from oracle import ld_oracle
    rng = np.random.default_rng(100 + rank)
    n_chk = min(args.check, n_mine)
    ok = True
    if n_chk:
        r = rng.integers(max(begin, 1), end, size=n_chk)
        c = (rng.random(n_chk) * r).astype(np.int64)
        idx = r * (r - 1) // 2 + c - shard.tri(begin)
        got = out[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32)
        words = planes[:, : (n_hap + 63) // 64]
        n1 = np.bitwise_count(words).sum(axis=1).astype(np.int32)
        n11 = np.bitwise_count(words[r] & words[c]).sum(axis=1).astype(np.int32)
        want = ld_oracle.packed_words(n_hap, n11, n1[r], n1[c])
        ok = bool((got == want).all())
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    best = min(times)
    total_pairs = v * (v - 1) // 2
    if rank == 0:
        print(json.dumps({"workload": f"ld_triangle large: {v} variants x {n_hap} haplotypes, {total_pairs} pairs, row-range sharded",
                          "n_gpus": world, "scaling": "strong", "ms": best, "ms_all": times, "value": total_pairs / (best * 1e-3),
                          "unit": "pairs/s", "all_pairs_kernel_ms_rank0": kern_ms / max(kern_n, 1), "rows_of_rank0": [begin, end], "sample_checked_per_rank": n_chk,
                          "parity_sample_ok": bool(flag.item())}), flush=True)
    st.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
