#!/usr/bin/env python
"""Per-region instruction counts of the first kernel in an .ncu-rep (source page): consecutive SASS
instructions with the same execution count are merged; prints warp-level and thread-level
instruction totals per region so divergence (threads/warp) and the phase split are visible.
usage: python profiles/ncu_regions.py rep.ncu-rep [min_share_percent]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = [(int(r[ie]), int(r[te]), int(r[ss]), r[1].strip()) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(d[0] for d in data)
tots = sum(d[2] for d in data)
print("total warp-instr %d, samples %d, sass lines %d" % (tot, tots, len(data)))
regions = []
for i, d in enumerate(data):
    if regions and regions[-1][2] == d[0]:
        regions[-1][1] = i
        regions[-1][3] += d[0]; regions[-1][4] += d[1]; regions[-1][5] += d[2]
    else:
        regions.append([i, i, d[0], d[0], d[1], d[2], d[3]])
for a, b, c, w, t, s, first in regions:
    if w * 100.0 / tot >= min_share or s * 100.0 / max(tots, 1) >= min_share:
        print("sass %4d-%4d  exec/instr %9d  warp-instr %5.1f%%  samples %5.1f%%  threads/warp %4.1f  %s" % (
            a, b, c, w * 100.0 / tot, s * 100.0 / max(tots, 1), t / max(w, 1), first[:48]))

# share of samples / instructions between consecutive CTA barriers (= the kernel's phases)
print("-- between barriers --")
start = 0
acc_w = acc_s = 0
for i, d in enumerate(data):
    acc_w += d[0]; acc_s += d[2]
    if "BAR.SYNC" in d[3] or i == len(data) - 1:
        if acc_w:
            print("sass %4d-%4d  warp-instr %5.1f%%  samples %5.1f%%" % (start, i, acc_w * 100.0 / tot, acc_s * 100.0 / max(tots, 1)))
        start = i + 1
        acc_w = acc_s = 0
