"""Worker of tests/test_gpu_multirank.py: one rank per GPU under torchrun (NCCL).  Frame-block
sharded pair-distance histogram + replica-sharded KMC ensemble; rank 0 writes the reduced result."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cmdlmc_b200 as cm  # noqa: E402
from cmdlmc_b200 import parallel, runtime, synth  # noqa: E402
from cmdlmc_b200.ensemble import run_kmc_ensemble  # noqa: E402
from cmdlmc_b200.topology import MODE_VERLET  # noqa: E402


def main(out_path):
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    runtime.init(local)
    rank, world = parallel.rank_world()
    w = synth.workload("C4")
    nfr, R = 160, 12
    box = cm.AtomBoxMonoclinic(w.cell)
    rate = cm.Fermi(*w.rate_params)
    # frame blocks: each rank generates only the frames it needs
    sh = parallel.ShardedTopology(box, w.n_oxygen, w.cutoff, w.buffer, MODE_VERLET, rate,
                                  lambda a, b: synth.trajectory(w, b - a, start=a), nfr, chunk=50)
    hist = np.zeros(50, np.int64)
    pairs = np.zeros(1, np.int64)
    for first, topo in sh.blocks():
        topo.distance_histogram(0.0, 5.0, 50, out=hist)
        pairs[0] += topo.frame_info()[0].sum()
    tot = parallel.allreduce_sum({"hist": hist, "pairs": pairs})
    # replicas
    ens = run_kmc_ensemble(box, lambda a, b: synth.trajectory(w, b - a, start=a), nfr,
                           n_sites=w.n_oxygen, n_protons=w.n_protons, cutoff=w.cutoff,
                           buffer=w.buffer, jumprate=rate, time_step=w.time_step, n_replicas=R,
                           seed=21, reset_frequency=80, print_frequency=20, chunk=64,
                           histogram=(0.0, 5.0, 50))
    if rank == 0:
        json.dump({"world": world, "hist": tot["hist"].tolist(), "pairs": int(tot["pairs"][0]),
                   "events": ens["events"], "jump_hist": ens["jump_hist"].tolist(),
                   "n_replicas": ens["n_replicas"], "occupancy": ens["occupancy_counts"].tolist(),
                   "msd_mean": ens["observables"]["mean"].tolist(),
                   "msd_sem": ens["observables"]["sem"].tolist()}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
