#!/usr/bin/env python
"""Throughput of the hot path on the five BASELINE.json configs (one GPU, resident inputs).

Not the bench line (bench.py measures C2, the config the metric is quoted on): a table for
DESIGN.md / profiles/ showing the same kernels at the other named sizes.
    python tools/config_table.py > profiles/<round>_configs.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cmdlmc_b200 as cm  # noqa: E402
from cmdlmc_b200 import runtime, synth  # noqa: E402
from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX  # noqa: E402
from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE, MODE_VERLET  # noqa: E402

runtime.init(0)
runtime.use_torch_stream()
FRAMES = {"C1": 16384, "C2": 16384, "C3": 1024, "C4": 4096, "C5": 32}
REPLICAS = {"C1": 1024, "C2": 1024, "C3": 256, "C4": 1024}


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None or ms < best else best
    return best


out = []
for cfg, B in FRAMES.items():
    w = synth.workload(cfg)
    n = w.n_oxygen
    d = torch.from_numpy(synth.trajectory(w, B)).cuda()
    box = cm.AtomBoxCubic(w.cell) if w.is_ortho else cm.AtomBoxMonoclinic(w.cell)
    rate = cm.Fermi(*w.rate_params) if w.rate_kind == "Fermi" else cm.ActivationEnergy(*w.rate_params)
    row = {"config": cfg, "n_oxygen": n, "frames": B, "cell": "orthorhombic" if w.is_ortho else "general",
           "rate": w.rate_kind}
    for mode, name in ((MODE_BRUTEFORCE, "bruteforce"), (MODE_VERLET, "verlet")):
        topo = DeviceTopology(box, n, w.cutoff, w.buffer, mode, rate, 0)
        topo.build_dev(d.data_ptr(), B)          # sizes + allocates
        ms = timed(lambda: topo.build_dev(d.data_ptr(), B))
        counts, reb, _ = topo.frame_info()
        row[name] = {"ms": ms, "path": "cell-list" if topo.path else "dense",
                     "rebuilds": int(reb.sum()), "pairs_per_frame": float(counts.mean()),
                     "frames_x_O_pairs_per_s": B * n * (n - 1) / 2 / ms * 1e3,
                     "listed_pair_frames_per_s": float(counts.sum()) / ms * 1e3,
                     "list_bytes_per_s_GB": float(counts.sum()) * 24 / ms / 1e6}
        if mode == MODE_VERLET and cfg in REPLICAS:
            R = REPLICAS[cfg]
            lat = np.stack([synth.initial_lattice(n, w.n_protons, 100 + r)[0] for r in range(R)])
            best = None
            for it in range(3):
                kmc = DeviceKMC(box, lat, w.time_step, RNG_PHILOX, seed=5 + it)
                ms_k = timed(lambda: kmc.advance(topo), reps=1)
                best = ms_k if best is None or ms_k < best else best
                st = kmc.state()
            row["kmc_philox"] = {"replicas": R, "ms": best, "events": int(st["n_events"].sum()),
                                 "site_updates_per_s": float(st["site_updates"].sum()) / best * 1e3}
        del topo
    out.append(row)
    print(json.dumps(row), flush=True)
    del d
    torch.cuda.empty_cache()
