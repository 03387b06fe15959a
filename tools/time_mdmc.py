"""(every case three times: cudaMalloc / cudaFree of the ~1 GB list arrays cost 0.05-0.7 s on this
virtualised box and dominate single runs -- read the best of the three)
End-to-end `mdmc`-style run (one replica, reference RNG protocol = exact replay mode):
ArrayTrajectory -> NeighborTopology (Verlet) -> Fermi -> KMCLattice -> ObservablesOutput."""
import gc, sys, time
sys.path.insert(0, '.')
import numpy as np
import cmdlmc_b200 as cm
from cmdlmc_b200 import synth
from cmdlmc_b200.kmc import KMCLattice, ObservablesOutput
from cmdlmc_b200.topology import NeighborTopology
from cmdlmc_b200.trajectory import ArrayTrajectory
for cfg, nfr in (("C1", 20000), ("C2", 20000)):
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    names = np.array(["O"] * w.n_oxygen)
    box = cm.AtomBoxCubic(w.cell) if w.is_ortho else cm.AtomBoxMonoclinic(w.cell)
    for rng in ("replay", "philox") * 3:
        np.random.seed(3)
        t0 = time.perf_counter()
        top = NeighborTopology(ArrayTrajectory(frames, names, time_step=w.time_step), box, donor_atoms="O",
                               cutoff=w.cutoff, buffer=w.buffer)
        kmc = KMCLattice(top, atom_box=box, jumprate_function=cm.Fermi(*w.rate_params), lattice_size=w.n_oxygen,
                         proton_number=w.n_protons, donor_atoms="O", time_step=w.time_step, rng=rng, chunk_size=4096)
        rows = list(ObservablesOutput(kmc, 1000, 100))
        dt = time.perf_counter() - t0
        print(cfg, rng, "frames", nfr, "rows", len(rows), "events", len(kmc.event_log["frame"]), "%.2f s" % dt, "%.0f frames/s" % (nfr / dt), flush=True)
        # teardown (cudaFree of ~1 GB of lists takes 0.04-0.7 s on this virtualised box) stays
        # outside the next run's clock
        del rows, kmc, top
        gc.collect()

# the reference's integration config (tests/integration/mdlmc_run.py:37-70): C1 with its P atoms,
# AngleTopology + FermiAngle
from cmdlmc_b200.topology import AngleTopology
w = synth.workload("C1")
nfr = 20000
frames = synth.trajectory(w, nfr + 1, with_extra=True)
names = np.array(["O"] * w.n_oxygen + ["P"] * w.n_extra)
box = cm.AtomBoxCubic(w.cell)
for rng in ("replay", "philox") * 3:
    np.random.seed(3)
    t0 = time.perf_counter()
    top = AngleTopology(ArrayTrajectory(frames, names, time_step=w.time_step), box, donor_atoms="O",
                        extra_atoms="P", group_size=w.group_size, cutoff=w.cutoff, buffer=w.buffer)
    kmc = KMCLattice(top, atom_box=box, jumprate_function=cm.FermiAngle(*w.rate_params, np.pi / 2),
                     lattice_size=w.n_oxygen, proton_number=w.n_protons, donor_atoms="O",
                     time_step=w.time_step, rng=rng, chunk_size=4096)
    rows = list(ObservablesOutput(kmc, 1000, 100))
    dt = time.perf_counter() - t0
    print("C1+angle", rng, "frames", nfr, "rows", len(rows), "events", len(kmc.event_log["frame"]), "%.2f s" % dt, "%.0f frames/s" % (nfr / dt), flush=True)
    del rows, kmc, top
    gc.collect()

