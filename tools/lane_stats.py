#!/usr/bin/env python3
"""Per-byte event counts of the bytewise Standard lane machine on samples of the bench workloads, from
the CPU emulation of the kernels' lane code (tests/emu, StdMachine3 with its DACH_STAT counters).
No GPU needed.  Usage: python tools/lane_stats.py [C2|C3|C5 ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import emu_api as E
import oracle_api as O
from daachorse_b200 import synth as S

NAMES = "steps probes hits miss_known miss_f2root learns root_falls root_stay sig_skips pushes cache_hits".split()
WHAT = {"steps": "lane iterations", "probes": "record fetches that are probes", "hits": "probes that hit (landings)",
        "miss_known": "own-child probes that missed (signature false positives)", "miss_f2root": "failure probes that missed, next: ROOT",
        "learns": "fetches of a failure state's record", "root_falls": "probes of ROOT's children", "root_stay": "landings in ROOT",
        "sig_skips": "bytes whose own-child probe the signature saved", "pushes": "landings on a state with outputs"}


def run(name, n_hay=256):
    cfg = S.config(name)
    ps = S.make_patterns(cfg)
    pool, b = S.make_pool(cfg, ps, 16 << 20)
    opma = O.OraclePma.build_packed(ps.blob, ps.offs)
    wire = opma.serialize()
    hay_len = min(cfg["hay_len"], 1 << 14)
    starts = S.window_starts(b, len(pool), n_hay, hay_len)
    text, offs = S.materialise_host(pool, starts, hay_len)
    buf = (C.c_ulonglong * len(NAMES))()
    lib = E.lib()
    E.scan(wire, False, 1, text, offs, kernel=3, out_cap=1 << 24)  # sizes the output; counters reset below
    lib.emu_stats(buf, 1)
    rc, m, oo, need = E.scan(wire, False, 1, text, offs, kernel=3, out_cap=max(int(need_cap(text)), 1 << 16))
    lib.emu_stats(buf, 1)
    nb = len(text)
    print("== %s: %d patterns, %d haystacks x %d B, %.4f matches/byte" % (name, len(ps), n_hay, hay_len, need / nb))
    for k, v in zip(NAMES, buf):
        if k in WHAT:
            print("   %-12s %8.4f per byte   %s" % (k, v / nb, WHAT[k]))


def need_cap(text):
    return len(text)  # >= 1 match per byte is far more than any config produces


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["C2", "C3"]):
        run(nm)
