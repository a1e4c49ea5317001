"""Short profiling target for ncu: a few all-pairs calls of one shape.

    python tools/prof_target.py V [--direct -1|0|1] [--batch N] [--calls 4]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ld_tools_b200 import Context, Store  # noqa: E402
from ld_tools_b200._lib import TUNE_MMA_DIRECT  # noqa: E402
from ld_tools_b200.engine import ENGINE_MMA  # noqa: E402
from ld_tools_b200.synth import random_planes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("v", type=int)
ap.add_argument("--direct", type=int, default=-1)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--calls", type=int, default=4)
ap.add_argument("--n-hap", type=int, default=5008)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = Context(0)
ctx.set_tuning(TUNE_MMA_DIRECT, a.direct)
st = Store.from_planes(ctx, random_planes(a.v, a.n_hap, seed=4), a.n_hap)
st.select_all()
rows = np.arange(a.v)
n = a.v * (a.v - 1) // 2
outs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(max(a.batch, 1))]
for _ in range(a.calls):
    if a.batch:
        ctx.triangle_batch_dev([(st, rows, o.data_ptr()) for o in outs], engine=ENGINE_MMA)
    else:
        st.triangle_dev(rows, outs[0].data_ptr(), engine=ENGINE_MMA)
    ctx.resolve()
torch.cuda.synchronize()
print("ok")
