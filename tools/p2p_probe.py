#!/usr/bin/env python3
"""NVLink peer-write probe (one process, all visible GPUs): how fast can GPUs 1..N-1 write into GPU 0 at once,
with destination-aligned 16-byte stores versus 4-byte-aligned destinations (what 12-byte match tuples at an
arbitrary base give)?  torch's copy kernel vectorises only when source and destination are 16-byte aligned."""
import sys
import time

import torch

n = torch.cuda.device_count()
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
words = mib * (1 << 20) // 4
dst = [torch.empty(words + 64, dtype=torch.int32, device="cuda:0") for _ in range(n)]
src = []
for r in range(1, n):
    with torch.cuda.device(r):
        src.append(torch.arange(words + 64, dtype=torch.int32, device="cuda:%d" % r))
        if torch.cuda.can_device_access_peer(r, 0):
            pass


def run(senders, d_off, s_off, reps=3):
    streams = [torch.cuda.Stream(device=r) for r in senders]
    best = 0.0
    for _ in range(reps + 1):
        for d in range(n):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for r, st in zip(senders, streams):
            with torch.cuda.device(r), torch.cuda.stream(st):
                dst[r][d_off: d_off + words].copy_(src[r - 1][s_off: s_off + words], non_blocking=True)
        for d in range(n):
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        best = max(best, len(senders) * words * 4 / dt / 1e9)
    return best


print("GPUs:", n, " %d MiB per sender" % mib)
for senders in ([1], list(range(1, n))) if n > 2 else ([1],):
    for d_off, s_off, what in ((0, 0, "aligned dst, aligned src"), (1, 1, "dst +4 B, src +4 B (relatively aligned)"),
                               (1, 0, "dst +4 B, src aligned"), (0, 1, "dst aligned, src +4 B")):
        print("%d sender(s) -> GPU 0, %-42s %7.1f GB/s into GPU 0" % (len(senders), what, run(senders, d_off, s_off)))
