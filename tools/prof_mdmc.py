import sys, time, cProfile, pstats
sys.path.insert(0, '.')
import numpy as np
import cmdlmc_b200 as cm
from cmdlmc_b200 import synth
from cmdlmc_b200.kmc import KMCLattice, ObservablesOutput
from cmdlmc_b200.topology import NeighborTopology
from cmdlmc_b200.trajectory import ArrayTrajectory
w = synth.workload("C2"); nfr = 20000
frames = synth.trajectory(w, nfr)
names = np.array(["O"] * w.n_oxygen)
box = cm.AtomBoxMonoclinic(w.cell)
def run():
    np.random.seed(3)
    top = NeighborTopology(ArrayTrajectory(frames, names, time_step=w.time_step), box, donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer)
    kmc = KMCLattice(top, atom_box=box, jumprate_function=cm.Fermi(*w.rate_params), lattice_size=w.n_oxygen,
                     proton_number=w.n_protons, donor_atoms="O", time_step=w.time_step, rng="replay", chunk_size=4096)
    return list(ObservablesOutput(kmc, 1000, 100))
run()
t0 = time.perf_counter(); run(); print("second run %.2f s" % (time.perf_counter() - t0))
cProfile.run("run()", "/tmp/p.prof")
pstats.Stats("/tmp/p.prof").sort_stats("cumulative").print_stats(22)
