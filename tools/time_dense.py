#!/usr/bin/env python
"""Kernel time of the brute-force list build (k_pairs_dense or the cell list) on resident frames.

    python tools/time_dense.py [CFG] [FRAMES] [REPS] [PATH]
Prints one JSON line: ms per launch, frames x O-pairs / s, list bytes per second.  Small enough to
sit under `ncu -k regex:k_pairs_dense`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cmdlmc_b200 as cm  # noqa: E402
from cmdlmc_b200 import _abi, runtime, synth  # noqa: E402
from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
path = int(sys.argv[4]) if len(sys.argv) > 4 else -1
runtime.init(0)
runtime.use_torch_stream()
w = synth.workload(cfg)
n = w.n_oxygen
cell = np.asarray(w.cell, dtype=float)
box = cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)
rate = cm.Fermi(*w.rate_params) if w.rate_kind == "Fermi" else cm.ActivationEnergy(*w.rate_params)
d = torch.from_numpy(synth.trajectory(w, B)).cuda()
topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, 0, path=path)
for _ in range(3):
    topo.build_dev(d.data_ptr(), B)
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    topo.build_dev(d.data_ptr(), B)
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
counts = topo.frame_info()[0]
best = min(ms)
alg = 24.0 * n * B + 24.0 * float(counts.sum())
print(json.dumps({"cfg": cfg, "frames": B, "n": n, "path": topo.path, "ms": ms, "best_ms": best,
                  "pairs_per_s": B * n * (n - 1) / 2 / (best * 1e-3),
                  "directed_pairs_per_frame": float(counts.mean()), "stride": topo.stride,
                  "algorithmic_gb_s": alg / (best * 1e-3) / 1e9, "skin_stats": topo.skin_stats() if hasattr(_abi.lib(), "cmd_topo_skin_stats") else None}))
