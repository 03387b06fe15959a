"""Brute-force (every frame rebuilt) neighbour lists of a large system through the cell-list path:
python tools/time_cell.py [C3|C5] [frames]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import cmdlmc_b200 as cm
from cmdlmc_b200 import runtime, synth
from cmdlmc_b200.topology import DeviceTopology
cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
runtime.init(0); runtime.use_torch_stream()
w = synth.workload(cfg); n = w.n_oxygen
d = torch.from_numpy(synth.trajectory(w, B)).cuda()
cell = np.asarray(w.cell, float)
box = cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)
rate = cm.Fermi(*w.rate_params) if w.rate_kind == "Fermi" else cm.ActivationEnergy(*w.rate_params)
topo = DeviceTopology(box, n, w.cutoff, w.buffer, 0, rate, 0)
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); topo.build_dev(d.data_ptr(), B); b.record(); torch.cuda.synchronize()
    print("cell list %s %d frames: %.3f ms (%.1f us per frame), %.0f pairs/frame" % (
        cfg, B, a.elapsed_time(b), a.elapsed_time(b) * 1e3 / B, topo.frame_info()[0].mean()))
