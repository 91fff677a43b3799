"""Sharding of a haystack batch across ranks, and the one exchange step of the path.

Haystacks are independent units: the automaton is replicated, a batch is cut
into contiguous, byte-balanced haystack ranges, every rank scans its range, and the per-rank results
(matches + per-haystack offsets) are gathered to rank 0, which rebases the offsets.  Rank order ==
haystack order, so the gathered result is bit-identical to the single-device result.

Two forms of the exchange step:
  * ``PeerGroup`` (GPUs of one node): no collective at all -- every rank's placement kernel stores its matches
    straight into rank 0's dense buffer over NVLink peer memory (dach_group_*, include/daachorse_b200.h); the only
    thing torch.distributed carries is the one-time exchange of the IPC handles;
  * ``gather_results``: the portable form on any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes as C

import numpy as np


class PeerGroup:
    """One rank's end of a shard group.  ``world`` ranks, each scanning ``n_local`` haystacks of a batch of
    ``n_total``; rank 0 owns the dense result (``match_cap`` matches).  ``exchange(blob) -> list of blobs`` carries the
    handles between the ranks: "dist" = torch.distributed.all_gather_object, None = the caller passes the
    blobs of all ranks (``.handle`` of each, rank order) to ``connect`` itself."""

    def __init__(self, rank, world, device, match_cap, n_total, exchange="dist"):
        from . import _lib
        from .automaton import _check

        self._L = _lib.load()
        self._check = _check
        self.rank, self.world, self.device = rank, world, int(device)
        self.match_cap, self.n_total = int(match_cap), int(n_total)
        self._h = C.c_void_p()
        _check(self._L.dach_group_create(rank, world, self.device, self.match_cap, self.n_total, C.byref(self._h)))
        mine = C.create_string_buffer(_lib.GROUP_HANDLE_BYTES)
        _check(self._L.dach_group_export(self._h, mine))
        self.handle = mine.raw
        if exchange == "dist":
            import torch.distributed as dist

            def exchange(blob):
                got = [None] * world
                dist.all_gather_object(got, blob)
                return got
        if exchange is not None:
            self.connect(exchange(self.handle))

    def connect(self, blobs):
        from . import _lib

        assert len(blobs) == self.world and all(len(b) == _lib.GROUP_HANDLE_BYTES for b in blobs)
        self._check(self._L.dach_group_connect(self._h, b"".join(blobs)))

    def place(self, job, hay_base, last, stream=None):
        """The exchange step of ``job`` (a daachorse_b200.Job that has scanned this rank's shard).  On ranks other than
        0 the call blocks until the rank's scan is done and the lower ranks' counts are in (its matches then leave
        through a copy engine): enqueue the next scan first; one thread driving several ranks places in rank order."""
        st = job._stream(stream, self.device)
        self._check(self._L.dach_group_place(self._h, job._h, int(hay_base), 1 if last else 0, st))

    def finish(self, stream=None):
        """Rank 0: waits until every rank's matches of the step have landed, returns the total; others: 0."""
        import torch

        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream.cuda_stream)
        tot = C.c_uint64()
        self._check(self._L.dach_group_finish(self._h, C.byref(tot), st))
        return int(tot.value)

    def result(self, total):
        """Rank 0: (matches (total, 3) int32, offsets (n_total + 1,) int64) as torch views of the result buffers."""
        import torch

        po, pf = C.c_void_p(), C.c_void_p()
        self._check(self._L.dach_group_result(self._h, C.byref(po), C.byref(pf)))
        return (_as_tensor(po.value, (int(total), 3), torch.int32, self.device, self),
                _as_tensor(pf.value, (self.n_total + 1,), torch.int64, self.device, self))

    def close(self):
        if self._h is not None and self._h.value:
            self._L.dach_group_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaArray:
    """__cuda_array_interface__ view of library-owned device memory (keeps the owner alive)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}
        self._owner = owner


def _as_tensor(ptr, shape, dtype, device, owner):
    import torch

    if 0 in shape:
        return torch.empty(shape, dtype=dtype, device="cuda:%d" % device)
    typestr = {torch.int32: "<i4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_CudaArray(ptr, shape, typestr, owner), device="cuda:%d" % device)


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
