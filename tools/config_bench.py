#!/usr/bin/env python3
"""Device-resident throughput of every BASELINE.json config with the current kernels (one GPU).
Not the driver's bench (that is bench.py on the headline config); this fills BASELINE.md's table."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import daachorse_b200 as D
from daachorse_b200 import synth as S


def timed(pma, mode, text_t, offs_t, reps=3):
    r = pma.scan_batch_device(mode, text_t, offs_t)
    out = torch.empty((r.matches.shape[0] + 1024, 3), dtype=torch.int32, device=text_t.device)
    oo = torch.empty(offs_t.numel(), dtype=torch.int64, device=text_t.device)
    ms, ks = [], []
    for _ in range(reps):
        r = pma.scan_batch_device(mode, text_t, offs_t, out=out, out_offs=oo)
        st = pma.stats()
        ms.append(st["total_ms"])
        ks.append(st["scan_kernel_ms"])
    nb = text_t.numel()
    return {"GBps_step": nb / (np.mean(ms) * 1e-3) / 1e9, "GBps_kernel": nb / (np.mean(ks) * 1e-3) / 1e9,
            "ms_step": float(np.mean(ms)), "matches_per_byte": r.matches.shape[0] / nb, "bytes": nb}


def batch(cfg, ps, n, pool_bytes, pad_chars=False):
    pool, b = S.make_pool(cfg, ps, pool_bytes)
    starts = S.window_starts(b, len(pool), n, cfg["hay_len"])
    pool_t = torch.from_numpy(pool).cuda()
    text_t, offs_t = S.materialise_on_device(pool_t, torch.from_numpy(starts).cuda(), cfg["hay_len"])
    if pad_chars:  # cut the trailing partial char of every window (C4)
        rows = text_t.view(n, cfg["hay_len"])
        L = cfg["hay_len"]
        done = torch.zeros(n, dtype=torch.bool, device=rows.device)
        for t in range(1, 4):
            bt = rows[:, L - t]
            is_cont = (bt & 0xC0) == 0x80
            need = torch.where(bt < 0x80, 1, torch.where(bt < 0xE0, 2, torch.where(bt < 0xF0, 3, 4)))
            cut = (~done) & (~is_cont) & (need > t)
            for k in range(1, t + 1):
                rows[cut, L - k] = 0x20
            done |= ~is_cont
    return text_t, offs_t


res = {}
LM_ONLY = 'lm' in sys.argv[1:]  # only the leftmost rows
if not LM_ONLY:
  # C2: 10k ASCII patterns, overlapping, 256K x 256 B
  cfg = S.config("C2"); ps = S.make_patterns(cfg)
  pma = D.DoubleArrayAhoCorasick.new(ps.as_list())
  t, o = batch(cfg, ps, cfg["n_haystacks"], cfg["pool_bytes"])
  res["C2 bytewise 10k ASCII, find_overlapping_iter, 256Ki x 256 B"] = timed(pma, D.FIND_OVERLAPPING, t, o)
  del t, o, pma
  # C3: find_iter and overlapping on 1 GiB
  cfg = S.config("C3", 0.25); ps = S.make_patterns(cfg)
  pma = D.DoubleArrayAhoCorasick.new(ps.as_list())
  t, o = batch(cfg, ps, cfg["n_haystacks"], 64 << 20)
  res["C3 bytewise 675k, find_iter, 256Ki x 4 KiB"] = timed(pma, D.FIND, t, o)
  res["C3 bytewise 675k, find_overlapping_no_suffix_iter, 256Ki x 4 KiB"] = timed(pma, D.FIND_OVERLAPPING_NO_SUFFIX, t, o)
  del t, o
  # C5 shape: same automaton family with 1M patterns, few long records
  cfg5 = S.config("C5"); ps5 = S.make_patterns(cfg5)
  pma5 = D.DoubleArrayAhoCorasick.new(ps5.as_list())
  cfg5s = dict(cfg5); cfg5s["hay_len"] = 1 << 20
  t, o = batch(cfg5s, ps5, 1024, 64 << 20)
  res["C5 shape: bytewise 1M patterns, find_overlapping_iter, 1024 x 1 MiB records"] = timed(pma5, D.FIND_OVERLAPPING, t, o)
  del t, o, pma5
# C4: charwise 100k CJK, leftmost longest, 512K x 1 KiB
cfg = S.config("C4"); ps = S.make_patterns(cfg)
pmc = D.CharwiseDoubleArrayAhoCorasickBuilder.new().match_kind(D.MatchKind.LeftmostLongest).build([p.decode() for p in ps.as_list()])
t, o = batch(cfg, ps, cfg["n_haystacks"], cfg["pool_bytes"], pad_chars=True)
res["C4 charwise 100k CJK LeftmostLongest, leftmost_find_iter, 512Ki x 1 KiB"] = timed(pmc, D.LEFTMOST_FIND, t, o)
if LM_ONLY:
    pmc.set_option("kernel", 0)
    res["C4 charwise, lane-per-haystack kernel"] = timed(pmc, D.LEFTMOST_FIND, t, o)
    pmc.set_option("kernel", 1)
    for th in (768, 512):
        pmc.set_option("threads", th)
        res["C4 charwise, lane machine, %d threads" % th] = timed(pmc, D.LEFTMOST_FIND, t, o)
    pms = D.CharwiseDoubleArrayAhoCorasick.new([p.decode() for p in ps.as_list()])
    res["(same data) charwise Standard, find_overlapping_iter"] = timed(pms, D.FIND_OVERLAPPING, t, o)
    res["(same data) charwise Standard, find_iter"] = timed(pms, D.FIND, t, o)
    pms.set_option("kernel", 0)
    res["(same data) charwise Standard, find_overlapping_iter, lane-per-haystack kernel"] = timed(pms, D.FIND_OVERLAPPING, t, o)
    del pms
pmb = D.DoubleArrayAhoCorasickBuilder.new().match_kind(D.MatchKind.LeftmostLongest).build(ps.as_list())
res["(same data) bytewise LeftmostLongest, leftmost_find_iter"] = timed(pmb, D.LEFTMOST_FIND, t, o)
if LM_ONLY:
    pmb.set_option("kernel", 0)
    res["(same data) bytewise LeftmostLongest, lane-per-haystack kernel"] = timed(pmb, D.LEFTMOST_FIND, t, o)
    pmb.set_option("kernel", 1)
    for th in (512, 256):
        pmb.set_option("threads", th)
        res["(same data) bytewise LeftmostLongest, lane machine, %d threads" % th] = timed(pmb, D.LEFTMOST_FIND, t, o)
for k, v in res.items():
    print("%-78s step %7.1f GB/s  kernel %7.1f GB/s  %6.2f ms  m/B %.3f" % (k, v["GBps_step"], v["GBps_kernel"], v["ms_step"], v["matches_per_byte"]))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "config_bench.json"), "w"), indent=1)
