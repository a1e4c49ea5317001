import sys, numpy as np, torch
sys.path.insert(0, ".")
from ld_tools_b200 import Context, Store
ctx = Context(0); n_s = 2504; nv = 50000
rng = np.random.default_rng(1)
text = np.empty((nv, 4 * n_s), dtype=np.uint8); bits = rng.integers(0, 2, size=(nv, n_s, 2), dtype=np.uint8)
text[:, 0::4] = bits[:, :, 0] + 48; text[:, 1::4] = 124; text[:, 2::4] = bits[:, :, 1] + 48; text[:, 3::4] = 9
st = Store(ctx, nv, 2 * n_s)
flat = torch.from_numpy(text.reshape(-1)).pin_memory().numpy()
st.pack_gt(0, flat, n_s, row_pitch=4 * n_s)
ctx.kernel_timing(True)
for _ in range(4): st.pack_gt(0, flat, n_s, row_pitch=4 * n_s)
ms, n = ctx.kernel_timing(False)
print("K1 kernel: %.1f us per launch over %d launches, %.0f GB/s of text + planes" % (ms / n * 1e3, n, nv * (4 * n_s + 632) / (ms / n * 1e-3) / 1e9))
got = st.download(0, 64); want = np.packbits(bits[:64].reshape(64, -1), axis=1, bitorder="little")
assert (got.view(np.uint8)[:, :want.shape[1]] == want).all()
