"""Multi-rank host logic on CPU: world size 2 over gloo (no GPU, no oracle scan on the ranks'
side needed -- the oracle only provides the per-haystack results being sharded)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_api as O
from daachorse_b200 import shard


def test_byte_balanced_ranges():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        lens = rng.integers(0, 1000, size=257)
        offs = np.zeros(258, dtype=np.uint64)
        offs[1:] = np.cumsum(lens)
        b = shard.byte_balanced_ranges(offs, world)
        assert b[0] == 0 and b[-1] == 257 and all(x <= y for x, y in zip(b, b[1:])) and len(b) == world + 1
        sizes = [int(offs[b[r + 1]] - offs[b[r]]) for r in range(world)]
        assert max(sizes) - min(sizes) <= 2000
    # degenerate: fewer haystacks than ranks, empty batch
    assert shard.byte_balanced_ranges(np.array([0, 5], dtype=np.uint64), 4)[-1] == 1
    assert shard.byte_balanced_ranges(np.array([0], dtype=np.uint64), 2) == [0, 0, 0]


def _worker(rank, world, port, text, offs, ref_m, ref_o, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = shard.byte_balanced_ranges(offs, world)
        lo, hi = b[rank], b[rank + 1]
        # this rank's slice of the reference result stands in for its device scan
        m0, m1 = int(ref_o[lo]), int(ref_o[hi])
        my_m = torch.from_numpy(ref_m[m0:m1].copy())
        my_o = torch.from_numpy((ref_o[lo:hi + 1] - ref_o[lo]).astype(np.int64))
        am, ao = shard.gather_results(my_m, my_o, dst=0)
        if rank == 0:
            q.put((am.numpy().tobytes() == ref_m.tobytes(), bool(np.array_equal(ao.numpy(), ref_o.astype(np.int64)))))
        else:
            assert am is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_reassembles_the_single_device_result(world):
    rng = np.random.default_rng(4)
    pats = [bytes(rng.integers(97, 101, size=int(rng.integers(1, 5))).tolist()) for _ in range(40)]
    pma = O.OraclePma.build(pats)
    lens = rng.integers(0, 300, size=101)
    lens[:3] = 0
    offs = np.zeros(102, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    text = rng.integers(97, 102, size=int(offs[-1])).astype(np.uint8)
    r = pma.scan_batch(O.FIND_OVERLAPPING, text, offs, want_matches=True)
    ref_m = np.stack([r["matches"]["start"], r["matches"]["end"], r["matches"]["value"]], axis=1).astype(np.int32)
    ref_o = np.concatenate([[0], np.cumsum(r["counts"])]).astype(np.int64)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, text, offs, ref_m, ref_o, q)) for rk in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_m, ok_o = q.get(timeout=10)
    assert ok_m and ok_o
