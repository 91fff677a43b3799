"""Timing probe: device-resident C3 slice with different forced segment lengths."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import daachorse_b200 as D
from daachorse_b200 import synth as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = S.config("C3"); ps = S.make_patterns(cfg)
pool, b = S.make_pool(cfg, ps, 64 << 20)
starts = S.window_starts(b, len(pool), n, 4096)
pool_t = torch.from_numpy(pool).cuda()
text_t, offs_t = S.materialise_on_device(pool_t, torch.from_numpy(starts).cuda(), 4096)
pma = D.DoubleArrayAhoCorasick.new(ps.as_list())
base = None
for seg in (-1, 4096, 2048, 1792, 1024, 512, 256):
    pma.set_option("seg_len", seg)
    r = pma.scan_batch_device(D.FIND_OVERLAPPING, text_t, offs_t)
    r = pma.scan_batch_device(D.FIND_OVERLAPPING, text_t, offs_t)
    st = pma.stats()
    tot = int(r.matches.shape[0])
    if base is None:
        base = r.matches.clone()
    same = bool(torch.equal(base, r.matches))
    print("n=%d seg=%5d kernel %.2f ms pipeline %.2f ms  total %d same %s  -> %.1f GB/s" % (
        n, seg, st["scan_kernel_ms"], st["total_ms"], tot, same, text_t.numel() / st["total_ms"] / 1e6), flush=True)
