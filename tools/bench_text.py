"""Table writer at scale: the V x V matrix text of ld_triangle.py:351-360 formatted on B200.

    python tools/bench_text.py [V ...]        (default 2000 20000)

Per V: the all-pairs call leaves the packed triangle in HBM; timed are (a) ldx_triangle_text with words and text in
HBM (both kernels + the scan; the text kernel alone from the library's event pairs), (b) the same with the text
brought to a host buffer, and, beside them, (c) the reference's own writer loop -- '\\t'.join(map(str, row)) over
lists of Python objects -- on a bounded sample of rows on one host core.  The GPU text of the sampled rows is
checked against (c) byte for byte.  HBM bytes counted per launch of the text kernel: 4 B per lower-triangle cell
read + the text written.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ld_tools_b200 import Context, Store  # noqa: E402
from ld_tools_b200.engine import BELOW_THRES, measure_value, tri_index  # noqa: E402
from ld_tools_b200.synth import random_planes  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2000, 20000]
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    ctx = Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    out = []
    for v in sizes:
        n_hap = 5008
        planes = random_planes(v, n_hap, seed=v)
        st = Store.from_planes(ctx, planes, n_hap)
        st.select_all()
        rows = np.arange(v)
        n_pairs = v * (v - 1) // 2
        dev = torch.zeros(n_pairs, dtype=torch.int32, device="cuda:0")
        torch.cuda.synchronize()
        st.triangle_dev(rows, dev.data_ptr())
        ctx.resolve()
        prefixes = [b"rs%d\t%d\t" % (100 + 3 * k, 16000000 + 37 * k) for k in range(v)]
        cap = 7 * v * v + sum(map(len, prefixes))
        text_dev = torch.empty(cap, dtype=torch.uint8, device="cuda:0")
        n = ctx.triangle_text(dev.data_ptr(), v, "r_square", prefixes, dev_text=(text_dev.data_ptr(), cap))     # warm-up

        def timed(fn, reps=5):
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                fn()
                e1.record(stream)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best

        ctx.kernel_timing(True)
        reps = 5
        ms_dev = timed(lambda: ctx.triangle_text(dev.data_ptr(), v, "r_square", prefixes, dev_text=(text_dev.data_ptr(), cap)), reps)
        k_ms, k_n = ctx.kernel_timing(False)
        k_ms /= max(k_n, 1)
        host = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
        ms_host = timed(lambda: ctx.triangle_text(dev.data_ptr(), v, "r_square", prefixes, out=host), 3)
        # ---- store -> text on the host in one call (all-pairs kernel + settlement + writer + the copy of the text)
        two_calls = host[:n].tobytes()
        same = st.triangle_table(rows, prefixes, out=host).tobytes() == two_calls
        ms_table = timed(lambda: st.triangle_table(rows, prefixes, out=host), 3)
        # ---- the reference's writer on a bounded sample of rows (one host core), and parity on those rows
        packed = dev.cpu().numpy().view(np.uint32)
        sample = sorted(set(np.linspace(0, v - 1, num=min(v, 200 if v <= 4000 else 24), dtype=np.int64).tolist()))
        objs = []
        for r in sample:
            base = tri_index(r, 0)
            objs.append([0 if (c >= r or packed[base + c] & BELOW_THRES) else measure_value(packed[base + c], "r_square") for c in range(v)])
        t0 = time.perf_counter()
        lines = [prefixes[r].decode() + "\t".join(map(str, row)) + "\n" for r, row in zip(sample, objs)]      # ld_triangle.py:356-360
        py_s = time.perf_counter() - t0
        text = host[:n].tobytes().split(b"\n")
        ok = all(text[r] + b"\n" == ln.encode() for r, ln in zip(sample, lines))
        bytes_k = 4 * n_pairs + n
        out.append({"variants": v, "cells": v * v, "text_bytes": int(n),
                    "device_call_ms": ms_dev, "text_kernel_ms": k_ms, "text_kernel_launches_timed": int(k_n),
                    "text_kernel_GBps": bytes_k / k_ms / 1e6, "hbm_peak_GBps": hbm, "frac_of_hbm_peak": bytes_k / k_ms / 1e6 / hbm,
                    "cells_per_s_device": v * v / ms_dev * 1e3, "to_host_call_ms": ms_host, "cells_per_s_to_host": v * v / ms_host * 1e3,
                    "text_GBps_to_host": n / ms_host / 1e6, "triangle_table_call_ms": ms_table, "triangle_table_equals_two_calls": same,
                    "python_writer_cells_per_s_1core": len(sample) * v / py_s, "python_rows_sampled": len(sample),
                    "sample_rows_identical": bool(ok)})
        print(json.dumps(out[-1]), flush=True)
        assert ok
        st.close()
        del dev, text_dev
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
