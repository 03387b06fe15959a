#!/usr/bin/env python
"""CPU timing of the stochastic stage (metric M2) with the oracle's C port  (TEST INFRASTRUCTURE:
only bench.py's cpu_baseline leg runs this; never the product path).

Per worker process, what one reference `KMCLattice` iteration does per frame
(mdlmc/LMC/MDMC.py:77-171 driving mdlmc/topo/topology.py:80-114 and jumprate_generators.py:33-34):
Verlet refresh / rebuild of the neighbour list, jump rates, allowed-transition mask, time stepping
and proton moves in exact-replay mode.  Every process walks its own replica (own lattice, own
uniform stream) over the same synthetic trajectory -- the reference is single-threaded, "all
cores" means that many independent runs side by side.

    python oracle/kmc_bench.py --workload C2 --frames 256 --procs 16   ->  one JSON line
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def _work(args):
    workload, frames_n, replica = args
    from oracle import oracle as orc
    from cmdlmc_b200 import synth
    w = synth.workload(workload)
    box = orc.OracleBox(w.cell)
    frames = synth.trajectory(w, frames_n)
    t0 = time.perf_counter()
    fptr = [0]
    start, dest, omega = [], [], []
    for row, col, dist, _ in orc.verlet_generator(box, frames, w.cutoff, w.buffer):
        start.append(row)
        dest.append(col)
        omega.append(orc.rates(w.rate_kind, w.rate_params, dist))
        fptr.append(fptr[-1] + len(row))
    start, dest, omega = np.concatenate(start), np.concatenate(dest), np.concatenate(omega)
    t1 = time.perf_counter()
    lattice, _ = synth.initial_lattice(w.n_oxygen, w.n_protons, 4000 + replica)
    u = np.random.RandomState(9000 + replica).random_sample(64 * frames_n + 1000)
    ev = orc.kmc_replay(np.array(fptr), start, dest, omega, lattice, w.time_step, u, 32 * frames_n + 400)
    t2 = time.perf_counter()
    return dict(site_updates=int(fptr[-1]), events=int(ev["n_events"]), topo_s=t1 - t0, kmc_s=t2 - t1,
                lattice=lattice.tolist() if replica == 0 else None)


def run(workload, frames, procs):
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_work, [(workload, 4, r) for r in range(procs)])       # warm: imports, page-in
        t = time.perf_counter()
        res = pool.map(_work, [(workload, frames, r) for r in range(procs)])
        wall = time.perf_counter() - t
    su = sum(r["site_updates"] for r in res)
    one = res[0]
    return dict(workload=workload, frames=frames, procs=procs, seconds=wall,
                site_updates=su, events=sum(r["events"] for r in res),
                site_updates_per_s=su / wall,
                frames_per_s_per_core=frames / (one["topo_s"] + one["kmc_s"]),
                one_core_site_updates_per_s=one["site_updates"] / (one["topo_s"] + one["kmc_s"]),
                one_core_kmc_only_site_updates_per_s=one["site_updates"] / max(one["kmc_s"], 1e-9),
                topology_share=one["topo_s"] / (one["topo_s"] + one["kmc_s"]),
                replica0_lattice=one["lattice"], replica0_events=one["events"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    print(json.dumps(run(a.workload, a.frames, a.procs)))
