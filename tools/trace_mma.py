import sys, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo')
from ld_tools_b200 import Context, Store
from ld_tools_b200._lib import TUNE_MMA_TILE_N, ptr
from ld_tools_b200.engine import ENGINE_MMA
from ld_tools_b200.synth import random_planes
ctx = Context(0)
planes = random_planes(2000, 5008)
st = Store.from_planes(ctx, planes, 5008); st.select_all()
rows = np.arange(2000)
ctx._lib.ldx_debug_trace(ctx._h, 1, None)
for tile in (64, 128, 256):
    ctx.set_tuning(TUNE_MMA_TILE_N, tile)
    for rep in range(3):
        st.triangle(rows, engine=ENGINE_MMA)
    s = np.zeros(256, dtype=np.uint64)
    ctx._lib.ldx_debug_trace(ctx._h, 1, ptr(s))
    s = s.astype(np.int64)
    print("tile", tile, "mainloop1 %.1f us" % ((s[1]-s[0])/1e3), "epi1 %.1f us" % ((s[2]-s[1])/1e3),
          "ready2 %+.1f us after epi1" % ((s[3]-s[2])/1e3) if s[3] else "", "epi2 %.1f us" % ((s[4]-s[3])/1e3) if s[4] else "")
    print("   widener chunk 20 cycles: bit_wait %d, lds+op_wait %d, expand %d, fence %d, sync+arrive %d" % tuple(s[240:245]))
    if tile == 1280:
        t0 = s[0]
        for g in range(0, 44):
            print(g, "prod %.2f" % ((s[128+g]-t0)/1e3), "bits %.2f" % ((s[192+g]-t0)/1e3), "widened %.2f" % ((s[8+g]-t0)/1e3), "mma %.2f" % ((s[64+g]-t0)/1e3))
