"""Kernel time of an exact-replay KMC run with few replicas: the one-CTA-per-replica kernel against
the warp-per-replica kernel (CMDLMC_B200_KMC_SOLO=0).  python tools/time_kmc_replay.py [C1|C2] [frames] [replicas]"""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import cmdlmc_b200 as cm
from cmdlmc_b200 import runtime, synth
from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
runtime.init(0); runtime.use_torch_stream()
w = synth.workload(cfg); n = w.n_oxygen
d = torch.from_numpy(synth.trajectory(w, B)).cuda()
box = cm.AtomBoxCubic(w.cell) if w.is_ortho else cm.AtomBoxMonoclinic(w.cell)
topo = DeviceTopology(box, n, w.cutoff, w.buffer, MODE_VERLET, cm.Fermi(*w.rate_params), 0)
topo.build_dev(d.data_ptr(), B)
lat0 = np.stack([synth.initial_lattice(n, w.n_protons, 5 + r)[0] for r in range(R)])
u = np.stack([np.random.RandomState(9 + r).random_sample(16 * B + 1000) for r in range(R)])
for name, env in (("solo", None), ("warp", "0")):
    if env is None: os.environ.pop("CMDLMC_B200_KMC_SOLO", None)
    else: os.environ["CMDLMC_B200_KMC_SOLO"] = env
    for rep in range(2):
        k = DeviceKMC(box, lat0, w.time_step, RNG_REPLAY)
        k.set_replay_stream(u)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); k.advance(topo); b.record(); torch.cuda.synchronize()
        st = k.state()
    ms = a.elapsed_time(b)
    print("%s %s: %d frames x %d replicas  %.2f ms  %.2f us/frame  events/replica %.0f  fallbacks %d" % (
        cfg, name, B, R, ms, ms * 1e3 / B, st["n_events"].mean(), k.selection_fallbacks()))
    lib = cm._abi.lib()
    if os.environ.get("SOLO_PROFILE") and name == "solo":
        import ctypes as C
        lib.cmd_kmc_debug_counter.restype = C.c_int64
        lib.cmd_kmc_debug_counter.argtypes = [C.c_void_p, C.c_int]
        names = ["frame start", "pass A", "scan", "tree+compact", "leaf sums", "combine", "scalar->move",
                 "move select", "move tail", "scalar tail"]
        cyc = [lib.cmd_kmc_debug_counter(k._handle, 2 + i) for i in range(10)]
        tc, tn = (lib.cmd_kmc_debug_counter(k._handle, 2 + i) for i in (10, 11))
        print("   SM clock during the kernel: %.0f MHz" % (tc * 1e3 / max(tn, 1)))
        tot = sum(cyc)
        names += ["  mv re-mask", "  mv scan"]
        cyc += [lib.cmd_kmc_debug_counter(k._handle, 2 + i) for i in (12, 13)]
        for nm, cy in zip(names, cyc):
            print("   %-14s %8.2f us/frame  %5.1f%%" % (nm, cy / (tc * 1e3 / max(tn, 1)) / B, 100.0 * cy / max(tot, 1)))
