#!/usr/bin/env python3
"""A compact pass over every kernel for compute-sanitizer (memcheck / racecheck / initcheck are 10-100x slower than
native, so this is a subset of tests/test_gpu_parity.py that still launches every kernel the library has):

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py

Golden vectors (all iterators x bytewise / charwise), random batches on every kernel option, text buffers at odd
addresses and with no slack after the last byte (the 8-byte text loads must not touch anything outside), stream
chunks, asynchronous jobs, a two-rank shard group on one device.  Every result is checked against the oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import daachorse_b200 as D
import oracle_api as O
from daachorse_b200 import shard

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "search_tests.json"), encoding="utf-8"))
MODE = {"find_iter": D.FIND, "find_overlapping_iter": D.FIND_OVERLAPPING,
        "find_overlapping_no_suffix_iter": D.FIND_OVERLAPPING_NO_SUFFIX, "leftmost_find_iter": D.LEFTMOST_FIND}
ORC = {D.FIND: O.FIND, D.FIND_OVERLAPPING: O.FIND_OVERLAPPING, D.FIND_OVERLAPPING_NO_SUFFIX: O.FIND_OVERLAPPING_NO_SUFFIX,
       D.LEFTMOST_FIND: O.LEFTMOST_FIND}
KIND = {"Standard": 0, "LeftmostLongest": 1, "LeftmostFirst": 2}
n_scans = 0


def golden():
    global n_scans
    for variant, iterator, coll, kind in GOLD["configs"]:
        if iterator not in MODE:
            continue
        cw = variant == "charwise"
        B = D.CharwiseDoubleArrayAhoCorasickBuilder if cw else D.DoubleArrayAhoCorasickBuilder
        for g in GOLD["collections"][coll]:
            for t in GOLD["groups"][g][::3]:
                pma = B.new().match_kind(KIND[kind]).build(t["patterns"])
                got = [(m.value(), m.start(), m.end()) for m in getattr(pma, iterator)(t["haystack"])]
                assert got == [tuple(x) for x in t["matches"]], t["name"]
                n_scans += 1


def random_case(seed, cw, kind):
    rng = np.random.default_rng(seed)
    table = ["a", "b", "é", "あ", "𝄞", "z"]
    pats = sorted({"".join(table[int(i)] for i in rng.integers(0, 4, size=int(rng.integers(1, 6)))) for _ in range(60)})
    hays = ["".join(table[int(i)] for i in rng.integers(0, 5, size=int(rng.integers(0, 300)))).encode() for _ in range(96)]
    if not cw:
        pats = [p.encode() for p in pats]
    offs = np.zeros(len(hays) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(h) for h in hays])
    text = np.frombuffer(b"".join(hays), dtype=np.uint8)
    B = D.CharwiseDoubleArrayAhoCorasickBuilder if cw else D.DoubleArrayAhoCorasickBuilder
    return B.new().match_kind(kind).build(pats), O.OraclePma.build(pats, charwise=cw, match_kind=kind), text, offs


def random_batches():
    global n_scans
    dev = torch.device("cuda", 0)
    for cw in (False, True):
        for kind in (0, 1):
            pma, opma, text, offs = random_case(10 * kind + cw, cw, kind)
            for mode in ([D.LEFTMOST_FIND] if kind else [D.FIND, D.FIND_OVERLAPPING, D.FIND_OVERLAPPING_NO_SUFFIX]):
                ref = opma.scan_batch(ORC[mode], text, offs, want_matches=True)
                for opts in ({"kernel": 3}, {"kernel": 4}, {"kernel": 3, "hot_entries": 512}, {"kernel": 2}, {"kernel": 1}, {"kernel": 0},
                             {"kernel": 3, "seg_len": 64}, {"kernel": 4, "threads": 256}):
                    for k, v in opts.items():
                        pma.set_option(k, v)
                    r = pma.scan_batch_host(mode, text, offs)
                    assert r.matches.tobytes() == ref["matches"].tobytes(), (cw, kind, mode, opts)
                    n_scans += 1
                    pma.set_option("seg_len", 0)
                    pma.set_option("hot_entries", 0)
                    pma.set_option("threads", 1024)
                pma.set_option("kernel", 3)
                # device-resident text at an odd address, the last haystack ending exactly at the end of the allocation
                pad = 3
                buf = torch.empty(text.size + pad, dtype=torch.uint8, device=dev)
                buf[pad:] = torch.from_numpy(text.copy()).to(dev)
                r = pma.scan_batch_device(mode, buf[pad:], torch.from_numpy(offs.astype(np.int64)).to(dev))
                m = r.matches.cpu().numpy().astype(np.uint32)
                assert m.tobytes() == ref["matches"].tobytes(), ("odd address", cw, kind, mode)
                n_scans += 1


def streams_jobs_groups():
    global n_scans
    dev = torch.device("cuda", 0)
    pma, opma, text, offs = random_case(77, False, 0)
    t = torch.from_numpy(text.copy()).to(dev)
    o = torch.from_numpy(offs.astype(np.int64)).to(dev)
    n = len(offs) - 1
    whole = pma.scan_batch_device(D.FIND_OVERLAPPING, t, o)
    # stream chunks: two rounds
    state = torch.zeros(n, dtype=torch.int32, device=dev)
    half = (offs[:-1] + (offs[1:] - offs[:-1]) // 2).astype(np.int64)
    for lo, hi in ((offs[:-1].astype(np.int64), half), (half, offs[1:].astype(np.int64))):
        lens = hi - lo
        co = np.zeros(n + 1, dtype=np.int64)
        co[1:] = np.cumsum(lens)
        ct = np.concatenate([text[int(a): int(b)] for a, b in zip(lo, hi)]) if co[-1] else np.zeros(0, np.uint8)
        tt = torch.from_numpy(ct.copy()).to(dev) if len(ct) else torch.zeros(16, dtype=torch.uint8, device=dev)[:0]
        pma.scan_stream_device(D.FIND_OVERLAPPING, tt, torch.from_numpy(co).to(dev), state)
        n_scans += 1
    for i in range(0, n, 7):
        assert int(state[i].item()) == opma.state_after(text[int(offs[i]): int(offs[i + 1])].tobytes())
    # jobs on two streams
    jobs = [pma.job(0), pma.job(0)]
    sts = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    outs = [torch.zeros((whole.matches.shape[0] + 3, 3), dtype=torch.int32, device=dev) for _ in range(2)]
    oos = [torch.zeros(n + 1, dtype=torch.int64, device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    for k in range(2):
        jobs[k].scan(D.FIND_OVERLAPPING, t, o, outs[k].shape[0], stream=sts[k])
        jobs[k].place(outs[k], oos[k], stream=sts[1 - k])
    for k in range(2):
        assert jobs[k].wait() == whole.matches.shape[0]
        assert torch.equal(outs[k][: whole.matches.shape[0]], whole.matches) and torch.equal(oos[k], whole.offsets)
        n_scans += 1
    # a two-rank shard group on one device
    bounds = shard.byte_balanced_ranges(offs, 2)
    cap = int(whole.matches.shape[0]) + 16
    groups = [shard.PeerGroup(r, 2, 0, cap, n, exchange=None) for r in range(2)]
    for g in groups:
        g.connect([x.handle for x in groups])
    for step in range(2):
        for r in (0, 1):  # the sanitizer serialises kernels: a rank that waits for a lower rank must be issued after it
            lo, hi = bounds[r], bounds[r + 1]
            tt = t[int(offs[lo]): int(offs[hi])]
            oo = (o[lo: hi + 1] - o[lo]).contiguous()
            jobs[r].scan(D.FIND_OVERLAPPING, tt, oo, cap, stream=sts[r])
            groups[r].place(jobs[r], lo, r == 1, stream=sts[r])
        groups[1].finish(stream=sts[1])
        total = groups[0].finish(stream=sts[0])
        m, oo = groups[0].result(total)
        assert total == whole.matches.shape[0] and torch.equal(m, whole.matches) and torch.equal(oo, whole.offsets)
        n_scans += 2
    for g in groups:
        g.close()


if __name__ == "__main__":
    golden()
    random_batches()
    streams_jobs_groups()
    torch.cuda.synchronize()
    print("sanitize.py: %d scans, all equal to the oracle" % n_scans)
