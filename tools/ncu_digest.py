#!/usr/bin/env python3
"""Digest of an .ncu-rep for the scan kernel: headline metrics, instruction mix by active-lane
bucket, stall reasons, hottest SASS.  Usage: tools/ncu_digest.py gpurun_out/x.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
want = ('Duration', 'Executed Ipc Active', 'Issue Slots Busy', 'L1/TEX Hit Rate', 'L2 Hit Rate',
        'Avg. Active Threads Per Warp', 'Eligible Warps Per Scheduler', 'Registers Per Thread', 'Achieved Occupancy',
        'Warp Cycles Per Issued Instruction', 'Executed Instructions', 'DRAM Throughput', 'Branch Efficiency',
        'Dynamic Shared Memory Per Block', 'Theoretical Occupancy')
for r in csv.reader(io.StringIO(det)):
    if len(r) > 14 and r[12] in want:
        print("%-40s %-12s %s" % (r[12], r[13], r[14]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
if len(rows) > 2:
    d = dict(zip(rows[0], rows[2]))
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
              'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'smsp__inst_executed.sum',
              'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
              'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
              'smsp__issue_active.avg.pct_of_peak_sustained_active'):
        if k in d:
            print("%-55s %s" % (k, d[k]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
b = collections.Counter()
bs = collections.Counter()
for r in data:
    n = int(r[ix['Instructions Executed']] or 0)
    try:
        t = float(r[ix['Avg. Threads Executed']])
    except Exception:
        t = 0
    k = 'a:>=28' if t >= 28 else 'b:20-28' if t >= 20 else 'c:10-20' if t >= 10 else 'd:4-10' if t >= 4 else 'e:<4'
    b[k] += n
    bs[k] += int(r[ix['# Samples']] or 0)
tot = sum(b.values())
ts = max(1, sum(bs.values()))
print("total warp instructions", tot)
for k in sorted(b):
    print("  lanes %-8s %12d  %5.1f%% inst  %5.1f%% samples" % (k, b[k], 100 * b[k] / tot, 100 * bs[k] / ts))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in st}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]:
    print("  %-28s %8d %5.1f%%" % (h, v, 100 * v / ts))
print("hottest SASS:")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("  %s %7s %11s %5s  %s" % (r[ix['Address']][-5:], r[ix['# Samples']], r[ix['Instructions Executed']],
                                      r[ix['Avg. Threads Executed']][:5], r[ix['Source']][:80]))
