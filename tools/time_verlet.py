import sys
sys.path.insert(0, '.')
import numpy as np, torch
import cmdlmc_b200 as cm
from cmdlmc_b200 import runtime, synth
from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET
runtime.init(0); runtime.use_torch_stream()
w = synth.workload("C2"); B = 16384; n = w.n_oxygen
d = torch.from_numpy(synth.trajectory(w, B)).cuda()
box = cm.AtomBoxMonoclinic(w.cell); rate = cm.Fermi(*w.rate_params)
topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, rate, 0)
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); topo.build_dev(d.data_ptr(), B); b.record(); torch.cuda.synchronize()
    print("verlet C2 16384 frames: %.3f ms" % a.elapsed_time(b))
