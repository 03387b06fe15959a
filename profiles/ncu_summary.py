#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per captured launch and the hottest source lines.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
        "smsp__inst_executed_pipe_fp64.sum", "sm__cycles_elapsed.max"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                print("   %-78s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = [(float(r[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print("   stall %-72s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "")
                                           .replace("_per_issue_active.ratio", ""), v))


def source(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] in ("#", "Line #", "Source") or (r and "Source" in r))
    hdr = rows[start]
    try:
        si = hdr.index("Source")
        ii = hdr.index("# Instructions Executed") if "# Instructions Executed" in hdr else hdr.index("Instructions Executed")
    except ValueError:
        print(hdr)
        return
    wi = hdr.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in hdr else None
    data = []
    for r in rows[start + 1:]:
        if len(r) != len(hdr):
            continue
        try:
            data.append((float(r[ii] or 0), float(r[wi] or 0) if wi is not None else 0, r[0], r[si]))
        except ValueError:
            pass
    tot = sum(d[0] for d in data) or 1
    tots = sum(d[1] for d in data) or 1
    print("== hottest source lines (inst share, stall-sample share)")
    for d in sorted(data, key=lambda d: -d[1])[:top]:
        print("   %5.1f%% %5.1f%%  L%-5s %s" % (100 * d[0] / tot, 100 * d[1] / tots, d[2], d[3].strip()[:110]))


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    if "--source" in sys.argv:
        source(rep, int(sys.argv[sys.argv.index("--source") + 1]))
