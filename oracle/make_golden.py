#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference (TEST INFRASTRUCTURE).

Runs only in the build container: imports the reference's compiled Cython classes from
oracle/_ref and its Python layers in place from /root/reference (oracle/ref_import.py), drives
them on seeded synthetic inputs and stores inputs + outputs as small fixtures.  The fixtures
travel to the GPU box; the reference's Python does not.

    python oracle/make_golden.py            # regenerates every fixture

Fixtures (all produced by reference code, none by the oracle restatement):
    geometry.npz     AtomBoxCubic / AtomBoxMonoclinic length, distance, angle, length_all_to_all,
                     next_neighbor on random cells and the known-answer inputs of
                     tests/cython_exts/LMC/test_AtomBox.py
    topology.npz     NeighborTopology.get_topology_bruteforce + topology_verlet_list_generator
    kmc.npz          KMCLattice event traces, consumed uniform stream, observables_output tuples
    fastforward.npz  KMCLattice.fastforward_to_next_jump on constant / sinusoidal rate streams
    angle.npz        AngleTopology colvars + FermiAngle rates on C1 with its P atoms
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from cmdlmc_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class MockTrajectory:
    """Same fake as tests/topo/test_topology.py:23-29."""

    def __init__(self, frames, time_step, names):
        from mdlmc.IO.trajectory_parser import Frame
        self.time_step = time_step
        self._frames = frames
        self._names = names
        self._Frame = Frame

    def __iter__(self):
        for k, pos in enumerate(self._frames):
            yield self._Frame(self._names, pos, time=k * self.time_step)


def make_box(cell):
    from mdlmc.cython_exts.LMC.PBCHelper import AtomBoxCubic, AtomBoxMonoclinic
    cell = np.asarray(cell, dtype=float)
    return AtomBoxCubic(cell) if cell.size == 3 else AtomBoxMonoclinic(cell)


def gen_geometry():
    rng = np.random.RandomState(1234)
    out = {}
    cells = {
        "ortho": np.array([10.0, 11.5, 12.25]),
        "cubic10": np.array([10.0, 10.0, 10.0]),
        "diag9": np.array([10.0, 0, 0, 0, 10.0, 0, 0, 0, 10.0]),
        "mono": synth.workload("C2").cell,
        "tri": np.array([12.0, 0, 0, 2.0, 11.0, 0, 1.0, 3.0, 9.0]),
        "tri2": synth.workload("C3").cell,
    }
    n = 400
    for name, cell in cells.items():
        box = make_box(cell)
        scale = 3.0 * np.abs(cell).max()
        a = rng.uniform(-scale, scale, size=(n, 3))
        b = rng.uniform(-scale, scale, size=(n, 3))
        c = rng.uniform(-scale, scale, size=(n, 3))
        out[name + "_cell"] = cell
        out[name + "_a"], out[name + "_b"], out[name + "_c"] = a, b, c
        out[name + "_length"] = box.length(a, b)
        out[name + "_distance"] = box.distance(a, b)
        out[name + "_angle"] = np.array([box.angle(a[i], b[i], c[i]) for i in range(n)])
        out[name + "_all"] = box.length_all_to_all(a[:40], b[:50])
        nn = [box.next_neighbor(a[i], b[:60]) for i in range(20)]
        out[name + "_nn_idx"] = np.array([x[0] for x in nn], dtype=np.int32)
        out[name + "_nn_dist"] = np.array([x[1] for x in nn])
    # tie behaviour: +-L/2 unchanged in the ortho path, flipped by round()
    z = np.zeros((2, 3))
    t = np.array([[5.0, 5.0, 5.0], [-5.0, -5.0, -5.0]])
    out["tie_ortho"] = make_box(cells["cubic10"]).distance(z, t)
    out["tie_general"] = make_box(cells["diag9"]).distance(z, t)
    # extended box (box_multiplier), PBCHelper.pyx:34-53
    from mdlmc.cython_exts.LMC.PBCHelper import AtomBoxCubic
    bx = AtomBoxCubic(np.array([10.0, 10, 10]), box_multiplier=(2, 3, 4))
    fr = rng.uniform(0, 10, size=(5, 3))
    out["ext_frame"] = fr
    out["ext_pos"] = np.array([bx.position_extended_box(i, fr) for i in range(5 * 24)])
    out["ext_pbc"] = np.asarray(bx.periodic_boundaries_extended)
    np.savez_compressed(os.path.join(GOLD, "geometry.npz"), **out)
    print("geometry.npz", len(out), "arrays")


def gen_topology():
    from mdlmc.topo.topology import NeighborTopology
    out = {}
    for cfg, nfr in (("C1", 120), ("C2", 40)):
        w = synth.workload(cfg)
        box = make_box(w.cell)
        frames = synth.trajectory(w, nfr)
        names = np.array(["O"] * w.n_oxygen)
        topo = NeighborTopology(MockTrajectory(frames, w.time_step, names), box, donor_atoms="O",
                                cutoff=w.cutoff, buffer=w.buffer)
        r0, c0, d0 = topo.get_topology_bruteforce(frames[0])
        out[cfg + "_bf_row"], out[cfg + "_bf_col"], out[cfg + "_bf_dist"] = r0, c0, d0
        counts, dsum, rebuilt, keep = [], [], [], {}
        prev = None
        for k, (row, col, dist, _) in enumerate(topo.topology_verlet_list_generator()):
            counts.append(len(row))
            dsum.append(float(np.sum(dist)))
            reb = prev is None or len(row) != len(prev[0]) or not (
                np.array_equal(row, prev[0]) and np.array_equal(col, prev[1]))
            rebuilt.append(reb)
            prev = (row, col)
            if k in (0, 1, nfr // 2, nfr - 1):
                keep[k] = (row.copy(), col.copy(), dist.copy())
        out[cfg + "_verlet_counts"] = np.array(counts)
        out[cfg + "_verlet_dsum"] = np.array(dsum)
        out[cfg + "_verlet_changed"] = np.array(rebuilt)
        for k, (row, col, dist) in keep.items():
            out["%s_verlet_f%d_row" % (cfg, k)] = row
            out["%s_verlet_f%d_col" % (cfg, k)] = col
            out["%s_verlet_f%d_dist" % (cfg, k)] = dist
        out[cfg + "_nframes"] = np.array(nfr)
    # known-answer input of tests/topo/test_topology.py:32-65
    box = make_box(np.array([10.0, 10, 10]))
    pos = np.array([[0.0, 0, 0], [1.5, 0, 0], [3.0, 0, 0], [6.0, 0, 0], [9.0, 0, 0]])
    topo = NeighborTopology(MockTrajectory([pos], 0.5, np.array(["O"] * 5)), box, cutoff=2.0,
                            buffer=0, donor_atoms="O")
    r, c, d = topo.get_topology_bruteforce(pos)
    out["kat_pos"], out["kat_row"], out["kat_col"], out["kat_dist"] = pos, r, c, d
    np.savez_compressed(os.path.join(GOLD, "topology.npz"), **out)
    print("topology.npz", len(out), "arrays")


def gen_angle():
    """AngleTopology (topology.py:124-167) + FermiAngle (jumprate_generators.py:37-43) of the
    REFERENCE on the C1 integration config with its P atoms (tests/integration/mdlmc_run.py:41-62)."""
    from mdlmc.topo.topology import AngleTopology
    from mdlmc.LMC.jumprate_generators import FermiAngle
    w = synth.workload("C1")
    nfr = 12
    # AngleTopology._determine_groups pulls the first frame out of the (one-shot) cached
    # trajectory generator (topology.py:43,145): the iteration starts at trajectory frame 1
    frames = synth.trajectory(w, nfr + 1, with_extra=True)
    names = np.array(["O"] * w.n_oxygen + ["P"] * w.n_extra)
    traj = MockTrajectory(frames, w.time_step, names)
    top = AngleTopology(traj, make_box(w.cell), donor_atoms="O", extra_atoms="P",
                        group_size=w.group_size, cutoff=w.cutoff, buffer=w.buffer)
    rate = FermiAngle(*w.rate_params, np.pi / 2)
    out = {"nframes": nfr, "group": np.array([top.map_O_to_P[o] for o in range(w.n_oxygen)])}
    counts, asum, rsum = [], [], []
    for k, (start, dest, dist, angle) in enumerate(top):
        if k == 0:
            out.update(start0=start, dest0=dest, dist0=dist, angle0=angle, rate0=rate(dist, angle))
        counts.append(len(start))
        asum.append(angle.sum())
        rsum.append(rate(dist, angle).sum())
        if k == nfr - 1:
            out.update(angle_last=angle, rate_last=rate(dist, angle))
            break
    out.update(counts=np.array(counts), angle_sum=np.array(asum), rate_sum=np.array(rsum))
    assert len(counts) == nfr and "angle_last" in out
    out["cached_frames_after_init"] = 1   # frame 0 sits in the frame cache
    np.savez_compressed(os.path.join(GOLD, "angle.npz"), **out)
    print("angle.npz:", counts[:4], "masked pairs in frame 0:", int((out["rate0"] == 0).sum()))


def run_reference_kmc(w, frames, seed, n_events, reset_frequency=None, print_frequency=None,
                      make_topology=None, names=None, jumprate=None):
    """Drives the reference KMCLattice; records events, lattices and (optionally) observables."""
    from mdlmc.topo.topology import NeighborTopology
    from mdlmc.LMC.MDMC import KMCLattice
    from mdlmc.LMC.jumprate_generators import Fermi
    box = make_box(w.cell)
    if names is None:
        names = np.array(["O"] * w.n_oxygen)
    if make_topology is not None:
        topo = make_topology(MockTrajectory(frames, w.time_step, names), box)
    else:
        topo = NeighborTopology(MockTrajectory(frames, w.time_step, names), box, donor_atoms="O",
                                cutoff=w.cutoff, buffer=w.buffer)
    np.random.seed(seed)
    kmc = KMCLattice(topo, atom_box=box,
                     jumprate_function=jumprate if jumprate is not None else Fermi(*w.rate_params),
                     lattice_size=w.n_oxygen, proton_number=w.n_protons, donor_atoms="O",
                     time_step=w.time_step)
    lattice0 = kmc.lattice.copy()
    events, ff = [], []
    orig_move = kmc.move_proton
    orig_ff = KMCLattice.fastforward_to_next_jump

    def move(start, dest, rates, lattice):
        before = kmc.lattice.copy()
        proton = orig_move(start, dest, rates, lattice)
        after = kmc.lattice
        s = int(np.where((before == proton) & (after == 0))[0][0])
        d = int(np.where((after == proton) & (before == 0))[0][0])
        events.append((s, d, int(proton)))
        return proton

    def fastforward(jumprates, dt):
        for item in orig_ff(jumprates, dt):
            ff.append(item)
            yield item

    kmc.move_proton = move
    kmc.fastforward_to_next_jump = fastforward
    obs = []
    frame_times = []
    lattice_marks = []
    try:
        if reset_frequency is None:
            for n, t, frame in kmc:
                frame_times.append((n, t))
                # what xyz_output (MDMC.py:173-177) would append to this frame: the occupied sites
                lattice_marks.append((float(np.dot(kmc.lattice, np.arange(1, len(kmc.lattice) + 1))),
                                      float(frame["O"].atom_positions[kmc.occupied_sites].sum())))
                if len(events) >= n_events:
                    break
        else:
            for n, t, msd, auto in kmc.observables_output(reset_frequency, print_frequency):
                obs.append((n, t, msd[0], msd[1], msd[2], auto))
                if len(events) >= n_events:
                    break
    except RuntimeError:  # generator raised StopIteration: end of trajectory
        pass
    ne = min(len(events), len(ff))
    ev = np.array(events[:ne], dtype=np.int64).reshape(-1, 3)
    ffa = np.array([(s, df, t) for s, df, t in ff[:ne]], dtype=float).reshape(-1, 3)
    rs = np.random.RandomState(seed)
    lat = np.zeros(w.n_oxygen, dtype=np.int32)
    lat[:w.n_protons] = range(1, w.n_protons + 1)
    rs.shuffle(lat)
    assert np.array_equal(lat, lattice0)
    u = rs.random_sample(2 * (len(ff) + 2))
    return dict(lattice0=lattice0, ev_start=ev[:, 0], ev_dest=ev[:, 1], ev_proton=ev[:, 2],
                ev_frame=ffa[:, 0].astype(np.int64), ev_dframe=ffa[:, 1].astype(np.int64),
                ev_time=ffa[:, 2], u=u, obs=np.array(obs, dtype=float).reshape(-1, 6),
                frame_times=np.array(frame_times, dtype=float).reshape(-1, 2),
                lattice_marks=np.array(lattice_marks, dtype=float).reshape(-1, 2),
                lattice_final=kmc.lattice.copy())


def gen_angle_kmc():
    """The reference KMCLattice on AngleTopology + FermiAngle (the integration config,
    tests/integration/mdlmc_run.py:37-70): _determine_groups leaves trajectory frame 0 in the frame
    cache, so continuous_output numbers it 0 and the KMC walks the trajectory from frame 1
    (topology.py:142-146, MDMC.py:92-94).  Event trace, (frame number, time) of every yielded frame
    with a mark of the lattice xyz_output would see, and observables_output rows."""
    from mdlmc.topo.topology import AngleTopology
    from mdlmc.LMC.jumprate_generators import FermiAngle
    w = synth.workload("C1")
    nfr = 260
    frames = synth.trajectory(w, nfr + 1, with_extra=True)
    names = np.array(["O"] * w.n_oxygen + ["P"] * w.n_extra)

    def mk(traj, box):
        return AngleTopology(traj, box, donor_atoms="O", extra_atoms="P", group_size=w.group_size,
                             cutoff=w.cutoff, buffer=w.buffer)
    rate = FermiAngle(*w.rate_params, np.pi / 2)
    out = {"nframes": nfr + 1, "seed": 17}
    res = run_reference_kmc(w, frames, 17, 10 ** 9, make_topology=mk, names=names, jumprate=rate)
    for k, v in res.items():
        out["cont_" + k] = v
    res = run_reference_kmc(w, frames, 17, 10 ** 9, reset_frequency=100, print_frequency=10,
                            make_topology=mk, names=names, jumprate=rate)
    for k in ("obs", "ev_frame", "ev_start", "ev_dest", "lattice_final", "u", "lattice0"):
        out["obs_" + k] = res[k]
    np.savez_compressed(os.path.join(GOLD, "angle_kmc.npz"), **out)
    print("angle_kmc.npz: events", len(out["cont_ev_start"]), "frames yielded", len(out["cont_frame_times"]),
          "first frame numbers", out["cont_frame_times"][:3, 0], "observable rows", len(out["obs_obs"]))


HYD_RELU = dict(a=0.9, b=2.35, d0=2.5, left_bound=2.2, right_bound=3.2)
HYD_RELAX = 6.0
HYD_PROTONS = 24
HYD_FERMI = (0.3, 2.45, 0.12)


def gen_hydronium():
    """HydroniumTopology (topology.py:170-257) with ReLUTransformation + DistanceInterpolator and
    with an InterpolatedTransformation, driven by the reference KMCLattice on C2-sized frames:
    event traces (rates depend on the residence time of every proton since its last jump)."""
    import copy
    from mdlmc.topo.topology import (HydroniumTopology, ReLUTransformation, DistanceInterpolator,
                                     InterpolatedTransformation)
    out = {}
    w = copy.deepcopy(synth.workload("C2"))
    w.n_protons = HYD_PROTONS
    w.rate_params = HYD_FERMI
    nfr = 150
    frames = synth.trajectory(w, nfr)
    xs = np.linspace(2.0, 3.4, 15)
    ys = 2.3 + 0.8 * (xs - 2.0) ** 1.5
    out["interp_x"], out["interp_y"] = xs, ys
    variants = {
        "relu": lambda traj, box: HydroniumTopology(
            traj, box, donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer,
            distance_transformation_function=ReLUTransformation(**HYD_RELU),
            distance_interpolator=DistanceInterpolator(HYD_RELAX)),
        "interp": lambda traj, box: HydroniumTopology(
            traj, box, donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer,
            distance_transformation_function=InterpolatedTransformation(xs, ys),
            distance_interpolator=None),
    }
    for name, mk in variants.items():
        res = run_reference_kmc(w, frames, 31, n_events=10 ** 9, make_topology=mk)
        for k, v in res.items():
            out["%s_%s" % (name, k)] = v
        print("hydronium", name, "events:", len(res["ev_start"]))
    # colvars of one frame for a fixed lattice / jump-time state (host-level parity)
    box = make_box(w.cell)
    names = np.array(["O"] * w.n_oxygen)
    top = variants["relu"](MockTrajectory(frames[:3], w.time_step, names), box)
    lattice = np.zeros(w.n_oxygen, np.int32)
    lattice[::17][:HYD_PROTONS] = np.arange(1, HYD_PROTONS + 1)
    top.take_lattice_reference(lattice)
    top._time_of_last_jump_vec[::2] = np.linspace(0.0, 0.7, len(top._time_of_last_jump_vec[::2]))
    cols = list(top)
    out["colvar_lattice"] = lattice
    out["colvar_tlast"] = top._time_of_last_jump_vec.copy()
    for k, (s, d, dist) in enumerate(cols):
        out["colvar%d_start" % k], out["colvar%d_dest" % k], out["colvar%d_dist" % k] = s, d, dist
    np.savez_compressed(os.path.join(GOLD, "hydronium.npz"), **out)


def gen_kmc():
    out = {}
    for cfg, nfr, seed in (("C1", 400, 11), ("C2", 60, 12)):
        w = synth.workload(cfg)
        frames = synth.trajectory(w, nfr)
        res = run_reference_kmc(w, frames, seed, n_events=10 ** 9)
        for k, v in res.items():
            out["%s_trace_%s" % (cfg, k)] = v
        res = run_reference_kmc(w, frames, seed + 100, n_events=10 ** 9, reset_frequency=100,
                                print_frequency=10)
        for k, v in res.items():
            out["%s_obs_%s" % (cfg, k)] = v
        out[cfg + "_nframes"] = np.array(nfr)
        out[cfg + "_seed"] = np.array(seed)
        print(cfg, "events", len(out[cfg + "_trace_ev_time"]), "obs rows",
              len(out[cfg + "_obs_obs"]))
    np.savez_compressed(os.path.join(GOLD, "kmc.npz"), **out)
    print("kmc.npz", len(out), "arrays")


def gen_fastforward():
    """tests/LMC/test_MDMC.py:10-93 style rate streams through the reference time stepper."""
    from itertools import cycle
    from mdlmc.LMC.MDMC import KMCLattice
    out = {}
    k = 0
    for dt in (0.1, 0.5, 1.3):
        for omega in (0.03, 0.06, 0.13):
            np.random.seed(0)
            gen = KMCLattice.fastforward_to_next_jump(cycle([omega]), dt)
            rows = [next(gen) for _ in range(100)]
            out["const_%d" % k] = np.array(rows, dtype=float)
            out["const_%d_par" % k] = np.array([dt, omega])
            k += 1
    rates = (0.06 + 0.02 * np.sin(np.linspace(0, 200 * np.pi, 10000)))
    np.random.seed(5)
    gen = KMCLattice.fastforward_to_next_jump(cycle(rates), 0.5)
    out["sin_rows"] = np.array([next(gen) for _ in range(2000)], dtype=float)
    out["sin_rates"] = rates
    out["sin_u"] = np.random.RandomState(5).random_sample(2000)
    out["const_u"] = np.random.RandomState(0).random_sample(100)
    np.savez_compressed(os.path.join(GOLD, "fastforward.npz"), **out)
    print("fastforward.npz", len(out), "arrays")


if __name__ == "__main__":
    ref_import.import_ref_python()
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("geometry", gen_geometry), ("topology", gen_topology),
                     ("fastforward", gen_fastforward), ("kmc", gen_kmc), ("angle", gen_angle), ("angle_kmc", gen_angle_kmc),
                     ("hydronium", gen_hydronium)):
        if not only or name in only:
            fn()
