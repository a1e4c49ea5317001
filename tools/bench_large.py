"""Steady-state throughput of the all-pairs engines on large variant sets (kernel-only, CUDA events).

    python tools/bench_large.py [V ...] [--tiles 64,128] [--engine mma|popc] [--reps 3] [--trace]

Prints pairs/s and the int8 tensor-pipe fraction for each (V, tile).  With --trace it also dumps
the per-chunk pipeline stamps of CTA 0 (see ldx_debug_trace in include/ldx.h)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ld_tools_b200 import Context, Store  # noqa: E402
from ld_tools_b200._lib import TUNE_MMA_TILE_N, ptr  # noqa: E402
from ld_tools_b200.engine import ENGINE_MMA, ENGINE_POPC  # noqa: E402
from ld_tools_b200.synth import random_planes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("v", nargs="*", type=int, default=[2000, 8192, 32768])
    ap.add_argument("--tiles", default="64,128")
    ap.add_argument("--engine", default="mma")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--n-hap", type=int, default=5008)
    ap.add_argument("--trace", action="store_true")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    peak_i8 = 2 * peaks["bf16_tflops"] * 1e12
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if args.trace:
        ctx._lib.ldx_debug_trace(ctx._h, 1, None)
    engine = ENGINE_MMA if args.engine == "mma" else ENGINE_POPC
    for v in args.v:
        planes = random_planes(v, args.n_hap)
        st = Store.from_planes(ctx, planes, args.n_hap)
        st.select_all()
        rows = np.arange(v)
        n_pairs = v * (v - 1) // 2
        out = torch.empty(n_pairs, dtype=torch.int32, device=dev)
        for tile in [int(t) for t in args.tiles.split(",")] if engine == ENGINE_MMA else [0]:
            if tile:
                ctx.set_tuning(TUNE_MMA_TILE_N, tile)
            st.triangle_dev(rows, out.data_ptr(), engine=engine)
            ctx.resolve()
            best = 1e30
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                st.triangle_dev(rows, out.data_ptr(), engine=engine)
                e1.record(stream)
                ctx.resolve()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            pps = n_pairs / best
            print(f"V={v} tile={tile} {best * 1e3:.3f} ms  {pps:.3e} pairs/s  int8 frac {pps * 2 * args.n_hap / peak_i8:.3f}", flush=True)
            if args.trace and engine == ENGINE_MMA:
                s = np.zeros(4096, dtype=np.uint64)
                ctx._lib.ldx_debug_trace(ctx._h, 1, ptr(s))
                s = s.astype(np.int64)
                t0 = s[0]
                print("  kernel entry %.2f us before the prologue's end; all roles done at %.2f us" % ((t0 - s[7]) / 1e3, (s[56] - t0) / 1e3))
                print("  tile stamps (us after prologue): " + " ".join("%.2f" % ((x - t0) / 1e3) for x in s[1:7] if x))
                if s[39]:
                    print("  CTA 1, epilogue warp 0 (us after its accumulator was ready): " + " ".join("%s %.2f" % (n, (s[k] - s[39]) / 1e3) for n, k in
                          [("ld0>", 40), ("ld0<", 41), ("stored0", 42), ("ld1>", 43), ("ld1<", 44), ("stored1", 45), ("done", 46)] if s[k]))
                c = np.concatenate([s[512:512 + 8 * 192].reshape(192, 8), s[2048:2048 + 2 * 192].reshape(192, 2)], axis=1)
                c = c[c[:, 0] > 0]
                if len(c):
                    e0 = c[:, 0].min()
                    cols = [0, 1, 2, 3, 6, 7, 8, 4] if c[:, 6].max() > 0 else [0, 1, 2, 3, 4]
                    print("  per-CTA life cycle (us after the first CTA's entry): entry / prologue done / accumulator ready / epilogue done / "
                          + ("barrier reached / passed / settled / " if len(cols) > 5 else "") + "all done")
                    for q in (0, 10, 50, 90, 100):
                        print("    p%-3d  %s" % (q, "  ".join("%6.2f" % (np.percentile(c[:, k] - e0, q) / 1e3) for k in cols)))
                    pc = s[2048 + 2 * 192:2048 + 3 * 192][:len(c)]
                    print("    deferred pairs per CTA: min %d median %d max %d (CTA %d)" % (pc.min(), np.median(pc), pc.max(), int(np.argmax(pc))))
                    for last in np.argsort(c[:, 4])[-3:]:
                        print("    slow CTA %d on SM %d: %s" % (last, c[last, 5], "  ".join("%6.2f" % ((c[last, k] - e0) / 1e3) for k in cols)))
                for g in range(0, 48, 1):
                    if not s[64 + g]:
                        break
                    print("  g=%2d prod %.2f bits %.2f opfree %.2f stored %.2f widened %.2f mma %.2f" %
                          (g, (s[128 + g] - t0) / 1e3, (s[192 + g] - t0) / 1e3, (s[256 + g] - t0) / 1e3, (s[320 + g] - t0) / 1e3,
                           (s[8 + g] - t0) / 1e3, (s[64 + g] - t0) / 1e3))
        del out
        st.close()
    ctx.close()


if __name__ == "__main__":
    main()
