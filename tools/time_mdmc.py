"""End-to-end `mdmc`-style run (one replica, reference RNG protocol = exact replay mode):
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
    for rng in ("replay", "philox"):
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
