#!/usr/bin/env python
"""Per-CUDA-source-line samples / instructions of the first kernel in an .ncu-rep (needs -lineinfo and
--import-source on).  usage: python profiles/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
files = {}
cur = None
agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr and len(r) > i_e and r[0].isdigit() and r[2] == "-":
        try:
            agg.append((int(r[i_s]), int(r[i_e]), cur.split("/")[-1], int(r[0]), r[1].strip()))
        except ValueError:
            pass
ts = sum(a[0] for a in agg) or 1
ti = sum(a[1] for a in agg) or 1
print("samples %d, warp-instr %d" % (ts, ti))
for s, e, f, ln, src in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s / ts, 100.0 * e / ti, f, ln, src[:90]))
