#!/usr/bin/env python
"""Smoke-sized pass of every kernel family, small enough to sit under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py [what ...]
    compute-sanitizer --tool racecheck python tools/sanitize_run.py [what ...]

what: dense cell verlet cluster kmc_stream kmc_solo lmc (default: all).  Prints one OK line per
family; results are only sanity-checked here (the parity tests live in tests/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import cmdlmc_b200 as cm  # noqa: E402
from cmdlmc_b200 import runtime, synth  # noqa: E402
from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX, RNG_REPLAY  # noqa: E402
from cmdlmc_b200.lmc import DeviceLMC  # noqa: E402
from cmdlmc_b200.topology import DeviceTopology, build_with_retry  # noqa: E402

what = sys.argv[1:] or ["dense", "cell", "verlet", "cluster", "kmc_stream", "kmc_solo", "lmc"]
runtime.init(0)


def box_of(w):
    cell = np.asarray(w.cell, float)
    return cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)


def rate_of(w):
    return cm.Fermi(*w.rate_params) if w.rate_kind == "Fermi" else cm.ActivationEnergy(*w.rate_params)


def topo_of(w, frames, mode, path=-1):
    box = box_of(w)
    return build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, mode,
                                                       rate_of(w), cap, path=path), frames)


if "dense" in what:      # k_pairs_dense: packed-half filter (C2, C1) and the FP32 filter (forced)
    for cfg in ("C2", "C1"):
        w = synth.workload(cfg)
        t = topo_of(w, synth.trajectory(w, 6), 0, path=0)
        assert (t.frame_info()[0] > 0).all()
    print("OK dense")
if "cell" in what:       # k_cell_*: C3 (triclinic) and a slice of C5
    w = synth.workload("C3")
    t = topo_of(w, synth.trajectory(w, 3), 0, path=1)
    assert (t.frame_info()[0] > 0).all()
    w = synth.workload("C5")
    t = topo_of(w, synth.trajectory(w, 1), 0)
    assert (t.frame_info()[0] > 0).all()
    print("OK cell")
if "verlet" in what:     # k_dr, k_sched_*, k_refresh, k_carry
    w = synth.workload("C2")
    t = topo_of(w, synth.trajectory(w, 24), 1)
    assert t.frame_info()[1][0]
    print("OK verlet")
if "cluster" in what:    # k_schedule_cluster (DSMEM exchange) + split refresh
    n, nfr = 8192, 6
    L = (n / 0.0334) ** (1.0 / 3.0)
    w = synth.Workload("big", np.array([L, L, L]), n, 0, 1, nfr, 0.5, 3.0, 2.0, "Fermi",
                       (0.06, 2.3, 0.1), 11, group_size=0)
    t = topo_of(w, synth.trajectory(w, nfr, amplitude=0.6, noise=0.05), 1)
    assert t.frame_info()[1][0]
    print("OK cluster")
if "kmc_stream" in what or "kmc_solo" in what or "lmc" in what:
    w = synth.workload("C1")
    nfr = 24
    topo = topo_of(w, synth.trajectory(w, nfr), 1)
    counts = topo.frame_info()[0]
    box = box_of(w)
    if "kmc_stream" in what:   # TMA ring + mbarriers, warp per replica
        lat = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 7 + r)[0] for r in range(12)])
        k = DeviceKMC(box, lat, w.time_step, RNG_PHILOX, seed=3)
        k.advance(topo)
        st = k.state()
        assert ((st["lattices"] > 0).sum(axis=1) == w.n_protons).all()
        print("OK kmc_stream")
    if "kmc_solo" in what:     # one CTA per replica, exact arithmetic, host-fed uniforms
        lat = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 7 + r)[0] for r in range(2)])
        k = DeviceKMC(box, lat, w.time_step, RNG_REPLAY)
        u = np.random.RandomState(5).uniform(size=(2, 2 * (16 * nfr + 64)))
        k.set_replay_stream(u)
        k.advance(topo)
        st = k.state()
        assert ((st["lattices"] > 0).sum(axis=1) == w.n_protons).all()
        print("OK kmc_solo")
    if "lmc" in what:          # k_lmc_sweep
        lat = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 9 + r)[0] for r in range(6)])
        dev = DeviceLMC(lat, seed=4)
        dev.advance(topo, w.time_step, 1)
        st = dev.state()
        assert ((st["lattices"] > 0).sum(axis=1) == w.n_protons).all()
        print("OK lmc")
