#!/usr/bin/env python3
"""Raw pinned-memory copy bandwidth of the box (what bounds bench.py's `e2e`): H2D alone, D2H alone,
both directions at once on two streams."""
import time

import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return (n / dt / 1e9 if h2d else 0.0), (n / 2 / dt / 1e9 if d2h else 0.0)


run(True, True, 2)
print("H2D alone        %.1f GB/s" % run(True, False)[0])
print("D2H alone        %.1f GB/s" % run(False, True)[1])
a, b = run(True, True)
print("both directions  H2D %.1f GB/s + D2H %.1f GB/s (1 GiB in, 0.5 GiB out per round)" % (a, b))
