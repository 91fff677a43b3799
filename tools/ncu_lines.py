#!/usr/bin/env python3
"""Warp instructions executed per CUDA source line of a captured kernel (ncu --import-source on, -lineinfo):
where do the issue slots go?  Usage: tools/ncu_lines.py x.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[2]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
iline, isrc, iinst, ithr, isamp = 0, 1, ix["Instructions Executed"], ix["Thread Instructions Executed"], ix["# Samples"]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


agg = collections.OrderedDict()
for r in rows[3:]:
    if len(r) <= iinst:
        continue
    key = (r[iline], r[isrc].strip())
    a = agg.setdefault(key, [0, 0, 0, 0])
    a[0] += num(r[iinst])
    a[1] += num(r[ithr])
    a[2] += num(r[isamp])
    a[3] += 1
tot = sum(a[0] for a in agg.values()) or 1
print("total warp instructions %d, SASS instructions %d" % (tot, sum(a[3] for a in agg.values())))
print("%6s %7s %6s %5s %5s  %s" % ("share", "Minst", "lanes", "sass", "samp%", "line: source"))
ts = sum(a[2] for a in agg.values()) or 1
for (ln, s), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %7.1f %6.1f %5d %5.1f  %s: %s" % (100 * a[0] / tot, a[0] / 1e6, a[1] / max(a[0], 1), a[3], 100 * a[2] / ts, ln, s[:110]))
