"""Sharding of a haystack batch across ranks, and the one exchange step of the path.

Haystacks are independent units: the automaton is replicated, a batch is cut
into contiguous, byte-balanced haystack ranges, every rank scans its range, and the per-rank results
(matches + per-haystack offsets) are gathered to rank 0, which rebases the offsets.  Rank order ==
haystack order, so the gathered result is bit-identical to the single-device result.

Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def byte_balanced_ranges(offs, world):
    """offs: n+1 ascending byte offsets.  Returns world+1 haystack boundaries such that every
    rank's byte count is as close as possible to total/world (contiguous ranges)."""
    offs = np.asarray(offs, dtype=np.uint64)
    n = len(offs) - 1
    total = int(offs[-1] - offs[0])
    bounds = [0]
    for r in range(1, world):
        target = int(offs[0]) + total * r // world
        j = int(np.searchsorted(offs, np.uint64(target), side="left"))
        j = min(max(j, bounds[-1]), n)
        # pick the closer of j-1 / j
        if j > bounds[-1] and j <= n and abs(int(offs[j - 1]) - target) < abs(int(offs[min(j, n)]) - target):
            j -= 1
        bounds.append(j)
    bounds.append(n)
    return bounds


def gather_results(matches, offsets, dst=0, group=None):
    """matches: (k, 3) int32 tensor of this rank's tuples, offsets: (n_r + 1,) int64 tensor relative to
    this rank's first match.  Returns (all_matches, all_offsets) on rank ``dst`` (None elsewhere):
    the concatenation in rank order with rebased offsets."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = matches.device
    meta = torch.tensor([matches.shape[0], offsets.numel()], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    counts = [int(m[0]) for m in metas]
    nofs = [int(m[1]) for m in metas]
    max_c, max_o = max(max(counts), 1), max(nofs)
    pad_m = torch.zeros((max_c, 3), dtype=matches.dtype, device=dev)
    pad_m[: matches.shape[0]] = matches
    pad_o = torch.zeros(max_o, dtype=offsets.dtype, device=dev)
    pad_o[: offsets.numel()] = offsets
    if rank == dst:
        gm = [torch.empty_like(pad_m) for _ in range(world)]
        go = [torch.empty_like(pad_o) for _ in range(world)]
    else:
        gm = go = None
    dist.gather(pad_m, gm, dst=dst, group=group)
    dist.gather(pad_o, go, dst=dst, group=group)
    if rank != dst:
        return None, None
    all_m = torch.cat([gm[r][: counts[r]] for r in range(world)])
    parts, base = [], 0
    for r in range(world):
        o = go[r][: nofs[r]] + base
        parts.append(o[:-1] if r + 1 < world else o)
        base += counts[r]
    return all_m, torch.cat(parts)
