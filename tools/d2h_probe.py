"""Plain pinned D2H copy rates by copy size and stream count (what the box's PCIe link gives, apart from this library)."""
import json
import torch

dev = torch.device("cuda", 0)
res = {}
for mb in (4, 8, 64):
    for n_streams in (1, 3):
        n = mb << 18
        src = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(n_streams)]
        dst = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(n_streams)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        for rep in range(2):
            torch.cuda.synchronize()
            e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
            e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
            for s, a in zip(streams, e0):
                a.record(s)
            for _ in range(20):
                for s, a, b in zip(streams, src, dst):
                    with torch.cuda.stream(s):
                        b.copy_(a, non_blocking=True)
            for s, a in zip(streams, e1):
                a.record(s)
            torch.cuda.synchronize()
            ms = max(a.elapsed_time(b) for a, b in zip(e0, e1))
        res[f"{mb}MB_x{n_streams}"] = 20 * n_streams * n * 4 / (ms * 1e-3) / 1e9
print(json.dumps(res, indent=1))
