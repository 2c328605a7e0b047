#!/usr/bin/env python3
"""Raw pinned host->device copy bandwidth on this box (the floor of bench.py's e2e leg)."""
import time
import torch
n = 311_354_880 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.uniform_()
d = torch.empty(n, dtype=torch.float64, device="cuda")
for parts in (1, 4, 21, 84):
    cuts = [n * i // parts for i in range(parts + 1)]
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for a, b in zip(cuts[:-1], cuts[1:]):
            d[a:b].copy_(h[a:b], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"{parts} copies: {dt*1e3:.3f} ms  {n*8/dt/1e9:.1f} GB/s")
