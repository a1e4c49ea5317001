"""Kernel-level timings of the all-pairs engine's modes (CUDA events on the launching stream, after warm-up):

    python tools/bench_modes.py [--reps 20]

  single   one 2,000-variant set: direct mode (no gather kernel) vs gather mode, whole call and all-pairs kernel alone
  batch    8 / 16 / 32 sets of 2,000 variants in one launch (ldx_triangle_batch_dev)
  large    8,192 and 32,768 variants (multi-wave, CTA pairs)
Prints one JSON object per line; fractions are of the int8 tensor roofline (2 x measured cuBLAS bf16)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ld_tools_b200 import Context, Store  # noqa: E402
from ld_tools_b200._lib import TUNE_MMA_DIRECT  # noqa: E402
from ld_tools_b200.engine import ENGINE_MMA  # noqa: E402
from ld_tools_b200.synth import random_planes  # noqa: E402

N_HAP = 5008


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--large", default="8192,32768")
    ap.add_argument("--batches", default="8,16,32")
    args = ap.parse_args()
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
        peak_i8 = 2 * peaks["bf16_tflops"] * 1e12
    except Exception:
        peak_i8 = 2 * 1590e12
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, reps, flush_l2=True):
        fn(); ctx.resolve()
        whole, kern = [], []
        for _ in range(reps):
            if flush_l2:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.resolve()
            torch.cuda.synchronize()
            whole.append(e0.elapsed_time(e1) * 1e-3)
        ctx.kernel_timing(True)
        for _ in range(reps):
            if flush_l2:
                flush.zero_()
            fn()
            ctx.resolve()
        ms, n = ctx.kernel_timing(False)
        return float(np.median(whole)), float(np.min(whole)), ms * 1e-3 / max(n, 1)

    v = 2000
    n_pairs = v * (v - 1) // 2
    stores = [Store.from_planes(ctx, random_planes(v, N_HAP, seed=40 + k), N_HAP) for k in range(4)]
    for s in stores:
        s.select_all()
    rows = np.arange(v)
    out = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    for direct in (1, 0):
        ctx.set_tuning(TUNE_MMA_DIRECT, direct)
        med, best, kern = timed(lambda: stores[0].triangle_dev(rows, out.data_ptr(), engine=ENGINE_MMA), args.reps)
        print(json.dumps({"case": "single", "variants": v, "direct": direct, "call_us_median": med * 1e6, "call_us_best": best * 1e6,
                          "allpairs_kernel_us": kern * 1e6, "pairs_per_s_call": n_pairs / med,
                          "roofline_frac_kernel": n_pairs * 2 * N_HAP / kern / peak_i8, "roofline_frac_call": n_pairs * 2 * N_HAP / med / peak_i8}), flush=True)
    ctx.set_tuning(TUNE_MMA_DIRECT, -1)
    for nb in [int(x) for x in args.batches.split(",") if x]:
        outs = [torch.empty(n_pairs, dtype=torch.int32, device=dev) for _ in range(nb)]
        sets = [(stores[k % len(stores)], rows, o.data_ptr()) for k, o in enumerate(outs)]
        med, best, kern = timed(lambda: ctx.triangle_batch_dev(sets, engine=ENGINE_MMA), max(args.reps // 2, 3))
        print(json.dumps({"case": "batch", "sets": nb, "variants": v, "call_us_median": med * 1e6, "call_us_best": best * 1e6,
                          "allpairs_kernel_us": kern * 1e6, "pairs_per_s_call": nb * n_pairs / med,
                          "roofline_frac_kernel": nb * n_pairs * 2 * N_HAP / kern / peak_i8,
                          "roofline_frac_call": nb * n_pairs * 2 * N_HAP / med / peak_i8}), flush=True)
        del outs
    for s in stores:
        s.close()
    for vb in [int(x) for x in args.large.split(",") if x]:
        st = Store.from_planes(ctx, random_planes(vb, N_HAP, seed=4), N_HAP)
        st.select_all()
        rb = np.arange(vb)
        npb = vb * (vb - 1) // 2
        ob = torch.empty(npb, dtype=torch.int32, device=dev)
        med, best, kern = timed(lambda: st.triangle_dev(rb, ob.data_ptr(), engine=ENGINE_MMA), 3, flush_l2=False)
        print(json.dumps({"case": "large", "variants": vb, "call_ms_median": med * 1e3, "allpairs_kernel_ms": kern * 1e3,
                          "pairs_per_s_kernel": npb / kern, "roofline_frac_kernel": npb * 2 * N_HAP / kern / peak_i8}), flush=True)
        del ob
        st.close()
    ctx.close()


if __name__ == "__main__":
    main()
