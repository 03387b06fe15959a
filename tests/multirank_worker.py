"""Worker of tests/test_gpu_multirank.py under torchrun.  Frame-block sharded pair-distance
histogram + replica-sharded KMC ensemble; rank 0 writes the reduced result.

    multirank_worker.py OUT nccl   one rank per GPU; torch.distributed on NCCL AND the library's own
                                   communicator (cmd_comm_init): statistics through
                                   cmd_stats_allreduce, Verlet frame blocks seeded from the
                                   all-gathered step lengths (cmd_allgather_dev)
    multirank_worker.py OUT gloo   every rank on cuda:0 (a one-GPU box): torch.distributed on gloo,
                                   no library communicator -- the coordinate-walking skip path"""
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


def main(out_path, backend="nccl"):
    local = int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    runtime.init(local)
    rank, world = parallel.rank_world()
    comm = parallel.comm_init() if backend == "nccl" else {"world": 1}
    w = synth.workload("C4")
    nfr, R = 160, 12
    box = cm.AtomBoxMonoclinic(w.cell)
    rate = cm.Fermi(*w.rate_params)
    # frame blocks: each rank generates only the frames it needs
    sh = parallel.ShardedTopology(box, w.n_oxygen, w.cutoff, w.buffer, MODE_VERLET, rate,
                                  lambda a, b: synth.trajectory(w, b - a, start=a), nfr, chunk=50)
    hist = np.zeros(50, np.int64)
    pairs = np.zeros(2, np.int64)
    rsum = np.zeros(1)
    for first, topo in sh.blocks():
        topo.distance_histogram(0.0, 5.0, 50, out=hist)
        cnt, reb, rs = topo.frame_info()
        pairs[0] += cnt.sum()
        pairs[1] += reb.sum()
        rsum[0] += rs.sum()
    tot = parallel.allreduce_sum({"hist": hist, "pairs": pairs, "rsum": rsum})
    # replicas
    ens = run_kmc_ensemble(box, lambda a, b: synth.trajectory(w, b - a, start=a), nfr,
                           n_sites=w.n_oxygen, n_protons=w.n_protons, cutoff=w.cutoff,
                           buffer=w.buffer, jumprate=rate, time_step=w.time_step, n_replicas=R,
                           seed=21, reset_frequency=80, print_frequency=20, chunk=64,
                           histogram=(0.0, 5.0, 50))
    if rank == 0:
        json.dump({"world": world, "hist": tot["hist"].tolist(), "pairs": int(tot["pairs"][0]),
                   "rebuilds": int(tot["pairs"][1]), "rate_sum": float(tot["rsum"][0]),
                   "comm_world": comm.get("world"), "nccl_version": comm.get("nccl_version"),
                   "events": ens["events"], "jump_hist": ens["jump_hist"].tolist(),
                   "n_replicas": ens["n_replicas"], "occupancy": ens["occupancy_counts"].tolist(),
                   "msd_mean": ens["observables"]["mean"].tolist(),
                   "msd_sem": ens["observables"]["sem"].tolist()}, open(out_path, "w"))
    dist.barrier()
    parallel.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "nccl")
