"""Where an end-to-end step's time goes: wall time of each blocking host call of one lane, then whole-job rate by lane count."""
import json
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from ld_tools_b200 import Context, Store  # noqa: E402
from ld_tools_b200.engine import ENGINE_AUTO  # noqa: E402
from ld_tools_b200.synth import pack_bits, synth_haplotypes  # noqa: E402

V, N = 2000, 5008
planes = pack_bits(synth_haplotypes(V, N, seed=1))
pin = torch.from_numpy(planes.view(np.int64)).pin_memory()
host = pin.numpy().view("<u8")
mask = np.full(planes.shape[1], np.uint64(0xFFFFFFFFFFFFFFFF))
mask[(N + 63) // 64:] = 0
if N % 64:
    mask[N // 64] = np.uint64((1 << (N % 64)) - 1)
rows = np.arange(V, dtype=np.int64)
n_pairs = V * (V - 1) // 2
res = {}


def lane_setup(k):
    ctx = Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    st = Store(ctx, V, N)
    out = torch.empty(n_pairs, dtype=torch.int16).pin_memory().numpy().view(np.uint16)
    return ctx, st, out


ctx, st, out = lane_setup(0)
for name, fn in [("upload", lambda: st.upload(0, host)), ("set_mask", lambda: st.set_mask(mask)),
                 ("triangle_values", lambda: st.triangle_values(rows, "r_square", engine=ENGINE_AUTO, out=out))]:
    st.upload(0, host); st.set_mask(mask)
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(300):
        fn()
    res[name + "_us"] = (time.perf_counter() - t0) / 300 * 1e6

lanes = [(ctx, st, out)] + [lane_setup(k) for k in range(1, 8)]
for n in (1, 2, 3, 4, 6, 8):
    gate = threading.Barrier(n + 1)
    steps = 100

    def main(k):
        c, s, o = lanes[k]
        for _ in range(2):
            gate.wait()
            for _ in range(steps):
                s.upload(0, host, wait=False); s.set_mask(mask); s.triangle_values(rows, "r_square", engine=ENGINE_AUTO, out=o)
            gate.wait()
    th = [threading.Thread(target=main, args=(k,)) for k in range(n)]
    for t in th:
        t.start()
    gate.wait(); gate.wait()
    t0 = time.perf_counter()
    gate.wait(); gate.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    res[f"lanes{n}_us_per_step"] = dt / (n * steps) * 1e6
print(json.dumps(res, indent=1))
