"""Time the Verlet pipeline (cmd_topo_build_dev, MODE_VERLET) of one BASELINE config on resident
frames: python tools/time_verlet.py [C1..C5] [frames]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import cmdlmc_b200 as cm
from cmdlmc_b200 import runtime, synth
from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
runtime.init(0); runtime.use_torch_stream()
w = synth.workload(cfg); n = w.n_oxygen
d = torch.from_numpy(synth.trajectory(w, B)).cuda()
cell = np.asarray(w.cell, float)
box = cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)
rate = cm.Fermi(*w.rate_params) if w.rate_kind == "Fermi" else cm.ActivationEnergy(*w.rate_params)
topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, 0)
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); topo.build_dev(d.data_ptr(), B); b.record(); torch.cuda.synchronize()
    counts, rebuilt, _ = topo.frame_info()
    print("verlet %s %d frames: %.3f ms, %d rebuilds, %.0f pairs/frame" % (
        cfg, B, a.elapsed_time(b), int(rebuilt.sum()), counts.mean()))
