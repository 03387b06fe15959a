#!/usr/bin/env python
"""Kernel time of the streaming (Philox) KMC kernel on resident Verlet lists:
python tools/time_kmc_stream.py [CFG] [REPLICAS] [FRAMES]   -> site-updates/s"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cmdlmc_b200 as cm  # noqa: E402
from cmdlmc_b200 import runtime, synth  # noqa: E402
from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX  # noqa: E402
from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
F = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
runtime.init(0)
runtime.use_torch_stream()
w = synth.workload(cfg)
cell = np.asarray(w.cell, float)
box = cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)
rate = cm.Fermi(*w.rate_params)
d = torch.from_numpy(synth.trajectory(w, F)).cuda()
topo = DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, MODE_VERLET, rate, 0)
topo.build_dev(d.data_ptr(), F)
counts = topo.frame_info()[0]
lat = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 100 + r)[0] for r in range(R)])
best = None
for rep in range(3):
    kmc = DeviceKMC(box, lat, w.time_step, RNG_PHILOX, seed=5 + rep)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    kmc.advance(topo)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    best = ms if best is None or ms < best else best
    ev = int(kmc.state()["n_events"].sum())
print("%s replicas %d frames %d: %.3f ms, %.4g site-updates/s, %d events (%s)" % (
    cfg, R, F, best, R * float(counts.sum()) / (best * 1e-3), ev,
    " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("CMDLMC_B200_KMC"))))
