#!/usr/bin/env python
"""CPU timing of the REFERENCE's own compiled code on the hot path  (TEST INFRASTRUCTURE: only
bench.py's cpu_baseline / --impl reference legs run this; never the product path).

Per frame, the geometry -> neighbour list -> jump rate stages with the reference's compiled
AtomBox from oracle/_ref (built from /root/reference/mdlmc/cython_exts by oracle/build_ref.py):

    d = atombox.length_all_to_all(frame, frame)          # PBCHelper.pyx:88-95, all N x N lengths
    keep = (d <= cutoff + buffer) & (d != 0)             # topology.py:67-69
    omega = a / (1 + exp((d[keep] - b) / c))             # jumprate_generators.py:33-34

This is the reference's compiled kernel driven in its fastest form: the reference's own
get_topology_bruteforce (topology.py:55-72) calls atombox.length once per pair from a Python
double loop and is ~30x slower still.  Frames are independent, so `procs`
worker processes take frames in parallel (the reference itself is single-threaded).

    python oracle/ref_bench.py --workload C2 --frames 64 --procs 16   ->  one JSON line
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

_state = {}


def _init(workload):
    from oracle import ref_import
    from cmdlmc_b200 import synth
    mod = ref_import.import_ref_atombox()
    w = synth.workload(workload)
    cell = np.asarray(w.cell, dtype=float)
    _state["box"] = mod.AtomBoxCubic(cell) if cell.size == 3 else mod.AtomBoxMonoclinic(cell)
    _state["w"] = w


def _work(frames):
    box, w = _state["box"], _state["w"]
    rc = w.cutoff + w.buffer
    a, b, c = w.rate_params[:3]
    pairs = 0
    total = 0.0
    for fr in frames:
        d = np.asarray(box.length_all_to_all(fr, fr))
        keep = (d <= rc) & (d != 0)
        x = d[keep]
        if w.rate_kind == "Fermi":
            om = a / (1 + np.exp((x - b) / c))
        else:   # legacy activation-energy rate (specification text, parity unpinned)
            A, ea, eb, d0, T = w.rate_params
            u = np.maximum(x - d0, 1e-300)
            om = np.where(x > d0, A * np.exp(-(ea * u / np.sqrt(eb + 1.0 / (u * u))) / (8.617333262e-5 * T)), A)
        pairs += int(keep.sum())
        total += float(om.sum())
    return pairs, total


def run(workload, frames, procs, repeats=1):
    from cmdlmc_b200 import synth
    w = synth.workload(workload)
    fr = synth.trajectory(w, frames)
    chunks = [c for c in np.array_split(fr, procs) if len(c)]
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(len(chunks), initializer=_init, initargs=(workload,)) as pool:
        pool.map(_work, [c[:1] for c in chunks])          # warm: imports, page-in
        for _ in range(repeats):
            t = time.perf_counter()
            res = pool.map(_work, chunks)
            times.append(time.perf_counter() - t)
    best = min(times)
    n = w.n_oxygen
    return {"workload": workload, "frames": int(frames), "procs": len(chunks), "seconds": best,
            "seconds_all": times,
            "o_pairs_per_frame": n * (n - 1) // 2,
            "pairs_per_s": frames * n * (n - 1) / 2 / best,
            "directed_pairs": int(sum(r[0] for r in res)),
            "rate_sum": float(sum(r[1] for r in res))}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--repeats", type=int, default=1)
    a = ap.parse_args()
    print(json.dumps(run(a.workload, a.frames, a.procs, a.repeats)))
