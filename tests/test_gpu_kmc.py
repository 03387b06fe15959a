"""GPU parity, rows A10-A13: the KMC kernel in exact-replay mode against (1) the reference's own
KMCLattice traces (golden, generated with np.random.seed) and (2) the CPU oracle on larger
seeded cases -- proton-occupancy trajectories BIT-EXACT, times 1e-12 relative -- and in Philox
mode statistically (error bars stated in the test)."""
import numpy as np
import pytest

from cmdlmc_b200 import synth

pytestmark = pytest.mark.gpu


def make_box(cell):
    import cmdlmc_b200 as cm
    cell = np.asarray(cell, dtype=float)
    return cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)


def make_kmc(w, frames, chunk_size=1024, rng="replay", seed=0):
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import KMCLattice
    from cmdlmc_b200.topology import NeighborTopology
    from cmdlmc_b200.trajectory import ArrayTrajectory
    box = make_box(w.cell)
    traj = ArrayTrajectory(frames, np.array(["O"] * w.n_oxygen), time_step=w.time_step)
    topo = NeighborTopology(traj, box, donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer)
    return KMCLattice(topo, atom_box=box, jumprate_function=cm.Fermi(*w.rate_params),
                      lattice_size=w.n_oxygen, proton_number=w.n_protons, donor_atoms="O",
                      time_step=w.time_step, rng=rng, seed=seed, chunk_size=chunk_size)


@pytest.mark.parametrize("cfg,chunk", [("C1", 1024), ("C1", 37), ("C2", 1024), ("C2", 16)])
def test_replay_vs_reference_golden(golden, cfg, chunk):
    """np.random.seed(s); KMCLattice(...)  ==  the reference run with the same seed."""
    g = golden("kmc")
    w = synth.workload(cfg)
    frames = synth.trajectory(w, int(g[cfg + "_nframes"]))
    np.random.seed(int(g[cfg + "_seed"]))
    kmc = make_kmc(w, frames, chunk_size=chunk)
    np.testing.assert_array_equal(kmc.lattice, g[cfg + "_trace_lattice0"])
    got = [(n, t) for n, t, _ in kmc]
    ev = kmc.event_log
    want_n = len(g[cfg + "_trace_ev_time"])
    # the reference dies with RuntimeError at the end of the trajectory after its last complete
    # event; we stop cleanly at the same place (possibly logging same-frame events it never
    # reached is impossible: both stop when the next frame is missing)
    n = min(want_n, len(ev["time"]))
    assert n >= want_n - 1 and n > 100
    np.testing.assert_array_equal(ev["frame"][:n], g[cfg + "_trace_ev_frame"][:n])
    np.testing.assert_array_equal(ev["start"][:n], g[cfg + "_trace_ev_start"][:n])
    np.testing.assert_array_equal(ev["dest"][:n], g[cfg + "_trace_ev_dest"][:n])
    np.testing.assert_array_equal(ev["proton"][:n], g[cfg + "_trace_ev_proton"][:n])
    np.testing.assert_allclose(ev["time"][:n], g[cfg + "_trace_ev_time"][:n], rtol=1e-12)
    ft = g[cfg + "_trace_frame_times"]
    m = min(len(ft), len(got))
    assert m >= len(ft) - 1
    np.testing.assert_array_equal([x[0] for x in got[:m]], ft[:m, 0].astype(int))
    np.testing.assert_allclose([x[1] for x in got[:m]], ft[:m, 1], rtol=1e-12)
    if len(ev["time"]) == want_n:
        np.testing.assert_array_equal(kmc.lattice, g[cfg + "_trace_lattice_final"])


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_observables_vs_reference_golden(golden, cfg):
    g = golden("kmc")
    w = synth.workload(cfg)
    frames = synth.trajectory(w, int(g[cfg + "_nframes"]))
    np.random.seed(int(g[cfg + "_seed"]) + 100)
    kmc = make_kmc(w, frames, chunk_size=64)
    rows = list(kmc.observables_output(100, 10))
    want = g[cfg + "_obs_obs"]
    assert len(want) > 0 and len(rows) >= len(want) - 1
    for (f, t, msd, auto), wr in zip(rows, want):
        assert f == int(wr[0])
        assert t == pytest.approx(wr[1], rel=1e-12)
        np.testing.assert_allclose(msd, wr[2:5], rtol=1e-9, atol=1e-12)
        assert auto == int(wr[5])


def device_topology(w, frames, mode=1):
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, mode,
                                                       rate, cap), frames)
    return box, topo


def topo_to_host(topo):
    counts, _, _ = topo.frame_info()
    fptr = np.concatenate([[0], np.cumsum(counts)])
    st, de, om = [], [], []
    for f in range(len(counts)):
        s, d, _, o = topo.get_frame(f, int(counts[f]))
        st.append(s); de.append(d); om.append(o)
    return fptr, np.concatenate(st), np.concatenate(de), np.concatenate(om)


@pytest.mark.parametrize("cfg,nfr,nrep", [("C1", 1500, 8), ("C4", 300, 5)])
def test_multi_replica_replay_vs_oracle(orc, cfg, nfr, nrep):
    """Several replicas, each with its own legacy RandomState stream, against the oracle run
    on the very rates the device produced: occupancy trajectories bit-exact."""
    from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    box, topo = device_topology(w, frames)
    fptr, start, dest, omega = topo_to_host(topo)
    lattices, streams = [], []
    nu = 2 * (16 * nfr + 64)
    for r in range(nrep):
        lat, rng = synth.initial_lattice(w.n_oxygen, w.n_protons, 1000 + r)
        lattices.append(lat)
        streams.append(rng.random_sample(nu))
    dev = DeviceKMC(box, np.array(lattices), w.time_step, RNG_REPLAY)
    dev.set_event_log(nu // 2)
    dev.set_replay_stream(np.array(streams))
    dev.advance(topo)
    st = dev.state()
    total_ev = 0
    for r in range(nrep):
        lat = lattices[r].copy()
        want = orc.kmc_replay(fptr, start, dest, omega, lat, w.time_step, streams[r], nu // 2)
        ev = dev.events(r)
        assert want["n_events"] == len(ev["time"]) == st["n_events"][r]
        assert want["n_events"] > 50
        np.testing.assert_array_equal(ev["frame"], want["frame"])
        np.testing.assert_array_equal(ev["start"], want["start"])
        np.testing.assert_array_equal(ev["dest"], want["dest"])
        np.testing.assert_array_equal(ev["proton"], want["proton"])
        np.testing.assert_array_equal(ev["time"], want["time"])      # bit-identical event times
        np.testing.assert_array_equal(st["lattices"][r], lat)
        total_ev += want["n_events"]
    assert (st["site_updates"] > 0).all()
    assert dev.tie_count() <= max(2, total_ev // 10000)   # near-tied decisions are rare


def test_philox_statistics_vs_oracle(orc):
    """Philox mode: mean event count and mean squared hop count over 128 GPU replicas agree with
    128 CPU-oracle replicas within 3 combined standard errors; runs are reproducible in the
    seed and differ between seeds."""
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX
    w = synth.workload("C1")
    nfr, nrep = 600, 128
    frames = synth.trajectory(w, nfr)
    box, topo = device_topology(w, frames)
    fptr, start, dest, omega = topo_to_host(topo)
    lattices = np.array([synth.initial_lattice(w.n_oxygen, w.n_protons, 50 + r)[0]
                         for r in range(nrep)])

    def run(seed):
        dev = DeviceKMC(box, lattices, w.time_step, RNG_PHILOX, seed=seed)
        dev.set_event_log(16 * nfr)
        dev.advance(topo)
        return dev, dev.state()

    dev, st = run(7)
    _, st_b = run(7)
    _, st_c = run(8)
    np.testing.assert_array_equal(st["lattices"], st_b["lattices"])
    np.testing.assert_array_equal(st["n_events"], st_b["n_events"])
    assert not np.array_equal(st["lattices"], st_c["lattices"])
    assert ((st["lattices"] > 0).sum(axis=1) == w.n_protons).all()      # protons conserved
    for r in range(4):                                                   # labels are a permutation
        assert sorted(st["lattices"][r][st["lattices"][r] > 0]) == list(range(1, w.n_protons + 1))
    ref_counts, ref_back = [], []
    for r in range(nrep):
        lat = lattices[r].copy()
        u = np.random.RandomState(9000 + r).random_sample(2 * 16 * nfr)
        res = orc.kmc_replay(fptr, start, dest, omega, lat, w.time_step, u, 16 * nfr)
        ref_counts.append(res["n_events"])
        ref_back.append(np.mean(res["start"][1:] == res["dest"][:-1]))
    gpu_counts = st["n_events"].astype(float)
    gpu_back = []
    for r in range(nrep):
        ev = dev.events(r)
        gpu_back.append(np.mean(ev["start"][1:] == ev["dest"][:-1]))
    for a, b in ((gpu_counts, np.array(ref_counts, float)), (np.array(gpu_back), np.array(ref_back))):
        se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
        assert abs(a.mean() - b.mean()) < 3 * se + 1e-12, (a.mean(), b.mean(), se)


def test_xyz_output_and_errors():
    import cmdlmc_b200 as cm
    w = synth.workload("C1")
    frames = synth.trajectory(w, 40)
    np.random.seed(3)
    kmc = make_kmc(w, frames)
    out = list(kmc.xyz_output("H"))
    assert len(out) > 0
    for fr in out:
        assert fr.atom_number == w.n_oxygen + w.n_protons
        assert (fr.atom_names[-w.n_protons:] == "H").all()
    # a fully occupied lattice has no allowed transition: every frame's total rate is 0, the
    # reference scans to the end of the trajectory without an event (and dies there with
    # RuntimeError: generator raised StopIteration); we end the iteration cleanly, no event
    np.random.seed(3)
    w2 = synth.workload("C1")
    w2.n_protons = w2.n_oxygen
    kmc2 = make_kmc(w2, frames)
    assert list(kmc2) == []
    assert len(kmc2.event_log["time"]) == 0


def test_mdmc_driver_end_to_end(tmp_path):
    """`mdmc config_load file.ini` (cmdlmc_b200.main, sections and keys of mdlmc/main.py:73-155) on
    an xyz file: same rows as building the objects by hand with the same global seed."""
    import io
    import cmdlmc_b200 as cm
    from cmdlmc_b200 import main
    from cmdlmc_b200.kmc import KMCLattice, ObservablesOutput
    from cmdlmc_b200.topology import NeighborTopology
    from cmdlmc_b200.trajectory import ArrayTrajectory
    w = synth.workload("C1")
    nfr = 400
    frames = synth.trajectory(w, nfr, with_extra=True)
    names = ["O"] * w.n_oxygen + ["P"] * w.n_extra
    xyz = tmp_path / "traj.xyz"
    with open(xyz, "w") as f:
        for fr in frames:
            f.write("%d\n\n" % len(names))
            for nm, p in zip(names, fr):
                f.write("%s %.17g %.17g %.17g\n" % (nm, p[0], p[1], p[2]))
    ini = tmp_path / "run.ini"
    ini.write_text("""
[Trajectory]
type = XYZTrajectory
filename = %s
time_step = %r
[AtomBox]
type = AtomBoxCubic
periodic_boundaries = [%s]
[NeighborTopology]
type = NeighborTopology
donor_atoms = O
cutoff = %r
buffer = %r
[JumpRate]
type = Fermi
a = %r
b = %r
c = %r
[KMCLattice]
lattice_size = %d
proton_number = %d
donor_atoms = O
time_step = %r
[Output]
type = ObservablesOutput
reset_frequency = 100
print_frequency = 10
""" % (xyz, w.time_step, ", ".join(repr(float(x)) for x in w.cell), w.cutoff, w.buffer,
       w.rate_params[0], w.rate_params[1], w.rate_params[2], w.n_oxygen, w.n_protons, w.time_step))
    np.random.seed(11)
    buf = io.StringIO()
    main.run(str(ini), out=buf)
    lines = buf.getvalue().strip().splitlines()
    assert len(lines) > 5
    np.random.seed(11)
    box = cm.AtomBoxCubic(w.cell)
    rate = cm.Fermi(*w.rate_params)
    top = NeighborTopology(ArrayTrajectory(frames, np.array(names), time_step=w.time_step), box,
                           donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer)
    kmc = KMCLattice(top, atom_box=box, jumprate_function=rate, lattice_size=w.n_oxygen,
                     proton_number=w.n_protons, donor_atoms="O", time_step=w.time_step)
    want = [str(x) for x in ObservablesOutput(kmc, 100, 10)]
    assert lines == want
    assert lines[0].startswith("(10, ") and "array([" in lines[0]


@pytest.mark.parametrize("rng_mode", ["replay", "philox"])
def test_occupancy_histogram_matches_event_log(orc, rng_mode):
    """Occupancy histogram (frames a site was seen occupied) == the same quantity rebuilt from the
    event log, and sums to protons x frames."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX, RNG_REPLAY
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET, build_with_retry
    w = synth.workload("C1")
    nfr, R = 500, 3
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                       MODE_VERLET, cm.Fermi(*w.rate_params), cap), frames)
    lat0 = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 70 + r)[0] for r in range(R)])
    kmc = DeviceKMC(box, lat0, w.time_step, RNG_REPLAY if rng_mode == "replay" else RNG_PHILOX, seed=4)
    kmc.enable_occupancy()
    kmc.set_event_log(20000)
    if rng_mode == "replay":
        kmc.set_replay_stream(np.random.RandomState(8).random_sample((R, 40000)))
    kmc.advance(topo)
    counts, replica_frames = kmc.occupancy()
    st = kmc.state()
    assert replica_frames == R * nfr and counts.sum() == w.n_protons * R * nfr
    want = np.zeros(w.n_oxygen, np.int64)
    for r in range(R):
        ev = kmc.events(r)
        assert len(ev["frame"]) == st["n_events"][r] > 5
        occ = lat0[r] > 0
        prev = 0
        for f, s, d in zip(ev["frame"], ev["start"], ev["dest"]):
            want[occ] += int(f) + 1 - prev
            prev = int(f) + 1
            occ[s], occ[d] = False, True
        want[occ] += nfr - prev
    np.testing.assert_array_equal(counts, want)


def test_replay_many_replicas_equals_single_replica_runs():
    """More replay replicas than SMs (several warps per CTA, exact scratch in global memory) must
    give, replica by replica, what a one-replica run (scratch in shared memory) gives."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET, build_with_retry
    w = synth.workload("C1")
    nfr, R = 120, 320
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                       MODE_VERLET, cm.Fermi(*w.rate_params), cap), frames)
    lat0 = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 500 + (r % 7))[0] for r in range(R)])
    u = np.stack([np.random.RandomState(900 + (r % 5)).random_sample(4000) for r in range(R)])
    big = DeviceKMC(box, lat0, w.time_step, RNG_REPLAY)
    big.set_replay_stream(u)
    big.set_event_log(2000)
    big.advance(topo)
    st = big.state()
    assert (st["n_events"] > 10).all()
    for r in (0, 1, 150, 319):
        one = DeviceKMC(box, lat0[r:r + 1], w.time_step, RNG_REPLAY)
        one.set_replay_stream(u[r:r + 1])
        one.set_event_log(2000)
        one.advance(topo)
        s1 = one.state()
        np.testing.assert_array_equal(s1["lattices"][0], st["lattices"][r])
        assert s1["time"][0] == st["time"][r] and s1["n_events"][0] == st["n_events"][r]
        e1, eb = one.events(0), big.events(r)
        for key in ("frame", "start", "dest", "proton", "time"):
            np.testing.assert_array_equal(e1[key], eb[key])
    # replicas with the same lattice and stream are identical
    same = [r for r in range(R) if r % 35 == 3]
    for r in same[1:]:
        np.testing.assert_array_equal(st["lattices"][r], st["lattices"][same[0]])


def _replay_run(w, topo, box, lat0, u, positions=None):
    from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
    k = DeviceKMC(box, lat0, w.time_step, RNG_REPLAY)
    k.set_replay_stream(u)
    k.set_event_log(20000)
    if positions is not None:
        k.set_observables(50, 10)
    k.advance(topo, positions)
    st = k.state()
    out = {"lattices": st["lattices"].copy(), "time": st["time"].copy(), "n_events": st["n_events"].copy(),
           "events": [k.events(r) for r in range(lat0.shape[0])],
           "rows": [k.observables(r) for r in range(lat0.shape[0])] if positions is not None else None,
           "fallbacks": k.selection_fallbacks()}
    return out


@pytest.mark.parametrize("cfg,nfr", [("C1", 400), ("C2", 150), ("C3", 24)])
def test_one_cta_per_replica_kernel_equals_warp_kernel(monkeypatch, cfg, nfr):
    """Few exact replicas run on one CTA each (parallel prefix selection); the result has to be the
    warp-per-replica kernel's (sequential np.cumsum) bit for bit -- also when every selection is
    forced through the sequential fallback."""
    import torch
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET, build_with_retry
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    # C3 (35 k listed pairs per frame): the compacted arrays do not fit shared memory, the kernel
    # works out of its global scratch
    rate = cm.Fermi(*w.rate_params) if w.rate_kind == "Fermi" else cm.ActivationEnergy(*w.rate_params)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                       MODE_VERLET, rate, cap), frames)
    R = 3
    lat0 = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 40 + r)[0] for r in range(R)])
    u = np.stack([np.random.RandomState(70 + r).random_sample(40000) for r in range(R)])
    pos = torch.from_numpy(frames).cuda()
    runs = {}
    for name, env in (("solo", {}), ("warp", {"CMDLMC_B200_KMC_SOLO": "0"}),
                      ("fallback", {"CMDLMC_B200_KMC_SELECT_MARGIN": "1e40"})):
        for key in ("CMDLMC_B200_KMC_SOLO", "CMDLMC_B200_KMC_SELECT_MARGIN"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        runs[name] = _replay_run(w, topo, box, lat0, u, pos.data_ptr())
    ref = runs["warp"]
    assert (ref["n_events"] > 20).all()
    assert runs["solo"]["fallbacks"] == 0 and ref["fallbacks"] == 0
    assert runs["fallback"]["fallbacks"] == int(ref["n_events"].sum())
    for name in ("solo", "fallback"):
        got = runs[name]
        np.testing.assert_array_equal(got["lattices"], ref["lattices"])
        np.testing.assert_array_equal(got["time"], ref["time"])
        np.testing.assert_array_equal(got["n_events"], ref["n_events"])
        for r in range(R):
            for key in ("frame", "start", "dest", "proton", "time", "dist"):
                np.testing.assert_array_equal(got["events"][r][key], ref["events"][r][key])
            # rows: frame, event time stamp and autocorrelation exactly; the MSD sums are formed by
            # 512 threads instead of 32 lanes, i.e. in another order
            np.testing.assert_array_equal(got["rows"][r][:, [0, 1, 5]], ref["rows"][r][:, [0, 1, 5]])
            np.testing.assert_allclose(got["rows"][r][:, 2:5], ref["rows"][r][:, 2:5], rtol=1e-12, atol=0)


def _cluster_topology(frames, rate, cutoff):
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    box = cm.AtomBoxCubic(np.array([60.0, 60.0, 60.0]))
    n = frames.shape[1]
    topo = build_with_retry(lambda cap: DeviceTopology(box, n, cutoff, 0.0, 0, rate, cap), frames)
    return box, topo


def _grid16(spacing):
    """16 sites on a 2 x 2 x 4 grid; with spacing 1 every pair is inside a 4 A cutoff."""
    g = np.stack(np.meshgrid(np.arange(2), np.arange(2), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
    return 20.0 + spacing * g.astype(float)


# The reference tests fastforward_to_next_jump on bare rate generators (tests/LMC/test_MDMC.py:
# 20-25,59,83).  On the device the rate source is a topology: 16 sites, 8 protons and a distance-
# independent rate a = omega / 64 (Exponential(a, 0)) give 8 * 8 = 64 allowed transitions in every
# configuration, i.e. the total rate omega in every frame.  (A one-proton toy lattice would not do:
# two events inside one frame leave the frame's cached transition list empty and the reference's
# move_proton raises IndexError there, MDMC.py:110 -- the device halts the replica the same way.)
LAT16 = np.array([[1, 2, 3, 4, 5, 6, 7, 8] + [0] * 8], dtype=np.int32)


def test_reference_fastforward_constant_rates():
    """tests/LMC/test_MDMC.py:10-51 on the device: with the same rate in every frame the
    time-dependent KMC reproduces the constant-rate KMC (same uniforms) to 1e-7, and the event's
    frame counter is int(t // dt)."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
    frames = np.tile(_grid16(1.0), (4096, 1, 1))
    for dt in (0.1, 0.5, 1.3):
        for omega in (0.03, 0.06, 0.13):
            box, topo = _cluster_topology(frames, cm.Exponential(omega / 64, 0.0), 4.0)
            assert (topo.frame_info()[0] == 16 * 15).all()
            u = np.random.RandomState(7).random_sample((1, 400))
            u[0, 0:200:2] = np.random.RandomState(0).random_sample(100)
            kmc = DeviceKMC(box, LAT16, dt, RNG_REPLAY)
            kmc.set_replay_stream(u)
            kmc.set_event_log(256)
            for _ in range(64):
                kmc.advance(topo)
                if kmc.state()["n_events"][0] >= 100:
                    break
            ev = kmc.events(0)
            assert len(ev["time"]) >= 100
            want = np.cumsum(-np.log(1 - np.random.RandomState(0).random_sample(100)) / omega)
            np.testing.assert_allclose(ev["time"][:100], want, atol=1e-7, rtol=0)
            np.testing.assert_array_equal((ev["time"][:100] // dt).astype(np.int64), ev["frame"][:100])


def test_reference_variable_rates_average():
    """tests/LMC/test_MDMC.py:54-73: sinusoidal rates, the mean event rate is the mean rate.  Four
    sites on a regular tetrahedron whose edge changes from frame to frame, two protons: always four
    allowed transitions of the same rate exp(-edge).  16 Philox replicas x ~2500 events: 0.5 %
    standard error, asserted within 2 %."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX
    t = np.linspace(0, 200 * np.pi, 10000)
    rates = 0.006 + 0.002 * np.sin(t)
    edge = -np.log(rates / 4)
    tet = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=float) / (2 * np.sqrt(2))
    frames = 30.0 + edge[:, None, None] * tet[None]
    box, topo = _cluster_topology(frames, cm.Exponential(1.0, -1.0), 7.5)
    assert (topo.frame_info()[0] == 12).all()
    lat0 = np.tile(np.array([[1, 2, 0, 0]], dtype=np.int32), (16, 1))
    kmc = DeviceKMC(box, lat0, 0.5, RNG_PHILOX, seed=3)
    for _ in range(84):
        kmc.advance(topo)
    st = kmc.state()
    assert (st["n_events"] > 2000).all()
    got = st["n_events"].sum() / (84 * 10000 * 0.5 * 16)
    assert abs(got - rates.mean()) / rates.mean() < 0.02


def test_reference_variable_rates_index():
    """tests/LMC/test_MDMC.py:76-93: the rate is non-zero in one frame of 117 only (the sites are
    out of range of each other in all the others); every event lands on that frame."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
    length, hot, reps = 117, 73, 64
    one = np.stack([_grid16(1.0) if f == hot else _grid16(9.0) for f in range(length)])
    box, topo = _cluster_topology(np.tile(one, (reps, 1, 1)), cm.Exponential(0.17 / 64, 0.0), 4.0)
    counts = topo.frame_info()[0]
    assert (counts.reshape(reps, length)[:, hot] == 240).all() and counts.sum() == 240 * reps
    kmc = DeviceKMC(box, LAT16, 0.22, RNG_REPLAY)
    kmc.set_replay_stream(np.random.RandomState(1).random_sample((1, 400)))
    kmc.set_event_log(256)
    for _ in range(200):
        kmc.advance(topo)
        if kmc.state()["n_events"][0] >= 101:
            break
    ev = kmc.events(0)
    assert len(ev["frame"]) >= 101
    assert (ev["frame"][:101] % length == hot).all()


@pytest.mark.parametrize("seed", range(10))
def test_randomised_small_lattices_replay_vs_oracle(orc, seed):
    """Random small systems (3-48 sites at about 0.05 per cubic Angstrom, any filling, orthorhombic or triclinic cells, brute-force or
    Verlet lists, frames with very few or no listed pairs): replay traces against the oracle, bit for
    bit, including the event at which a run dies of the reference's IndexError (no transition left
    in the cached frame, MDMC.py:110)."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    rng = np.random.RandomState(100 + seed)
    n = int(rng.randint(3, 49))
    nprot = int(rng.randint(1, n))
    L = np.maximum(6.5, (n / 0.05) ** (1.0 / 3.0) * rng.uniform(0.85, 1.2, 3))
    if rng.rand() < 0.5:
        cell, hm = L.copy(), np.diag(L)
    else:
        hm = np.array([[L[0], 0, 0], [rng.uniform(-0.3, 0.3) * L[1], L[1], 0],
                       [rng.uniform(-0.3, 0.3) * L[2], rng.uniform(-0.3, 0.3) * L[2], L[2]]])
        cell = hm.reshape(-1).copy()
    cutoff = float(rng.uniform(2.2, 3.2))
    buffer = float(rng.choice([0.0, 0.8]))
    mode = 1 if buffer > 0 else 0
    nfr = int(rng.randint(40, 160))
    pos0 = rng.rand(n, 3) @ hm
    frames = pos0[None] + np.cumsum(rng.normal(0, 0.04, (nfr, n, 3)), axis=0)
    rate = cm.Fermi(float(rng.uniform(0.05, 0.5)), float(rng.uniform(2.0, 3.2)), float(rng.uniform(0.05, 0.3)))
    dt = float(rng.uniform(0.2, 2.0))
    nrep = int(rng.randint(1, 4))
    box = make_box(cell)
    topo = build_with_retry(lambda cap: DeviceTopology(box, n, cutoff, buffer, mode, rate, cap), frames)
    fptr, start, dest, omega = topo_to_host(topo)
    lattices, streams = [], []
    nu = 2 * (32 * nfr + 64)
    for r in range(nrep):
        lat = np.zeros(n, np.int32)
        lat[:nprot] = np.arange(1, nprot + 1)
        rr = np.random.RandomState(5000 + 10 * seed + r)
        rr.shuffle(lat)
        lattices.append(lat)
        streams.append(rr.random_sample(nu))
    dev = DeviceKMC(box, np.array(lattices), dt, RNG_REPLAY)
    dev.set_event_log(nu // 2)
    dev.set_replay_stream(np.array(streams))
    dev.advance(topo)
    st = dev.state()
    for r in range(nrep):
        lat = lattices[r].copy()
        want = orc.kmc_replay(fptr, start, dest, omega, lat, dt, streams[r], nu // 2)
        ev = dev.events(r)
        assert want["n_events"] == len(ev["time"]) == st["n_events"][r]
        for key in ("frame", "start", "dest", "proton", "time"):
            np.testing.assert_array_equal(ev[key], want[key])
        np.testing.assert_array_equal(st["lattices"][r], lat)
    print("seed", seed, "sites", n, "protons", nprot, "frames", nfr, "mode", mode, "events", st["n_events"].tolist(),
          "pairs/frame", float(np.diff(fptr).mean()))
    assert st["n_events"].sum() > 0


@pytest.mark.parametrize("cfg,nfr", [("C1", 800), ("C4", 400)])
def test_philox_msd_and_jump_counts_vs_oracle_replicas(orc, cfg, nfr):
    """north_star correctness, part 4: in Philox mode the MSD PER AXIS (output.py:35-49) at every
    printed row and the jump counts of 128 GPU replicas (k_kmc_stream, observables on the device)
    agree with 128 CPU replicas of the oracle (exact-replay arithmetic on legacy RandomState
    streams) within 3 * sqrt(se_gpu^2 + se_cpu^2).  Seeds are fixed: the outcome is deterministic."""
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX
    w = synth.workload(cfg)
    nrep, pf = 128, nfr // 4
    frames = synth.trajectory(w, nfr)
    box, topo = device_topology(w, frames)
    fptr, start, dest, omega = topo_to_host(topo)
    obox = orc.OracleBox(w.cell)
    lattices = np.array([synth.initial_lattice(w.n_oxygen, w.n_protons, 300 + r)[0] for r in range(nrep)])
    dev = DeviceKMC(box, lattices, w.time_step, RNG_PHILOX, seed=2024)
    dev.set_observables(nfr + 1, pf)          # no reset inside the run
    dev.advance(topo, topo.positions_ptr())
    st = dev.state()
    gpu_rows = [dev.observables(r) for r in range(nrep)]
    cpu_rows, cpu_counts = [], []
    for r in range(nrep):
        lat = lattices[r].copy()
        u = np.random.RandomState(77000 + r).random_sample(2 * (24 * nfr + 64))
        res = orc.kmc_replay(fptr, start, dest, omega, lat, w.time_step, u, 24 * nfr + 64)
        cpu_counts.append(res["n_events"])
        obs = orc.observables(obox, frames, lattices[r], res, nfr + 1, pf)
        cpu_rows.append(np.array([[f, t, m[0], m[1], m[2], a] for f, t, m, a in obs]))
    n_rows = min(min(len(x) for x in gpu_rows), min(len(x) for x in cpu_rows))
    assert n_rows >= 2          # the last printed frame may not have been flushed by an event
    checked = 0
    for k in range(n_rows):
        g = np.array([x[k] for x in gpu_rows])
        c = np.array([x[k] for x in cpu_rows])
        assert (g[:, 0] == c[:, 0]).all() and g[0, 0] == (k + 1) * pf
        for axis in (2, 3, 4):
            a, b = g[:, axis], c[:, axis]
            se = np.sqrt(a.var(ddof=1) / nrep + b.var(ddof=1) / nrep)
            assert a.mean() > 0 and abs(a.mean() - b.mean()) < 3 * se, (cfg, k, axis, a.mean(), b.mean(), se)
            checked += 1
        a, b = g[:, 5], c[:, 5]             # covalent autocorrelation
        se = np.sqrt(a.var(ddof=1) / nrep + b.var(ddof=1) / nrep)
        assert abs(a.mean() - b.mean()) < 3 * se + 1e-12, (cfg, k, "autocorr", a.mean(), b.mean(), se)
    a, b = st["n_events"].astype(float), np.array(cpu_counts, float)
    se = np.sqrt(a.var(ddof=1) / nrep + b.var(ddof=1) / nrep)
    assert abs(a.mean() - b.mean()) < 3 * se, (cfg, "jump counts", a.mean(), b.mean(), se)
    assert checked >= 6


@pytest.mark.parametrize("engine", ["kmc", "lmc"])
def test_c4_verified_replay_subset_64_replicas(orc, engine):
    """BASELINE config 4's verification leg at its stated size: 64 replicas on the 384-O lattice
    over 1000 frames in exact-replay mode, every replica bit-identical with the CPU oracle
    (KMC: the reference's RandomState protocol; legacy LMC sweep: GSL MT19937 streams)."""
    w = synth.workload("C4")
    nfr, nrep = 1000, 64
    frames = synth.trajectory(w, nfr)
    box, topo = device_topology(w, frames)
    fptr, start, dest, omega = topo_to_host(topo)
    if engine == "kmc":
        from cmdlmc_b200.kmc import DeviceKMC, RNG_REPLAY
        nu = 2 * (24 * nfr + 64)
        lattices, streams = [], []
        for r in range(nrep):
            lat, rng = synth.initial_lattice(w.n_oxygen, w.n_protons, 1000 + r)
            lattices.append(lat)
            streams.append(rng.random_sample(nu))
        dev = DeviceKMC(box, np.array(lattices), w.time_step, RNG_REPLAY)
        dev.set_event_log(nu // 2)
        dev.set_replay_stream(np.array(streams))
        dev.advance(topo)
        assert dev.events_dropped() == 0
        st = dev.state()
        total = 0
        for r in range(nrep):
            lat = lattices[r].copy()
            want = orc.kmc_replay(fptr, start, dest, omega, lat, w.time_step, streams[r], nu // 2)
            ev = dev.events(r)
            assert want["n_events"] == len(ev["time"]) == st["n_events"][r] and want["n_events"] > 1000
            for key in ("frame", "start", "dest", "proton", "time"):
                np.testing.assert_array_equal(ev[key], want[key], err_msg="replica %d %s" % (r, key))
            np.testing.assert_array_equal(st["lattices"][r], lat)
            total += want["n_events"]
        assert total > 64 * 1000
    else:
        # four blocks of 250 frames (the pregenerated streams of 64 replicas x 1000 frames would
        # take 5 GB of host memory at once); lattices carry over, every block gets fresh streams
        from cmdlmc_b200.lmc import DeviceLMC, RNG_REPLAY, gsl_streams
        lattices = np.array([synth.initial_lattice(w.n_oxygen, w.n_protons, 1000 + r)[0] for r in range(nrep)])
        want = lattices.copy()
        dev = DeviceLMC(lattices, RNG_REPLAY)
        jumps_want = np.zeros(nrep, np.int64)
        for blk in range(4):
            if blk:
                box, topo = device_topology(w, synth.trajectory(w, 250, start=250 * blk))
                fptr, start, dest, omega = topo_to_host(topo)
            else:
                fptr, start, dest, omega = fptr[:251], start[:fptr[250]], dest[:fptr[250]], omega[:fptr[250]]
                box, topo = device_topology(w, frames[:250])
            counts = np.diff(fptr)
            streams = [gsl_streams(500 + r + 1000 * blk, counts) for r in range(nrep)]
            pick, acc = np.stack([x[0] for x in streams]), np.stack([x[1] for x in streams])
            del streams
            dev.set_replay_stream(pick, acc)
            dev.advance(topo, w.time_step, 1)
            prob = omega * w.time_step
            for r in range(nrep):
                pos = 0
                for f in range(250):
                    a, b = fptr[f], fptr[f + 1]
                    jumps_want[r] += orc.lmc_sweep(start[a:b], dest[a:b], prob[a:b], want[r],
                                                   pick[r, pos:pos + b - a], acc[r, pos:pos + b - a])
                    pos += b - a
            st = dev.state()
            np.testing.assert_array_equal(st["lattices"], want, err_msg="block %d" % blk)
            del pick, acc
        np.testing.assert_array_equal(st["jumps"], jumps_want)
        assert (jumps_want > 100).all()


def test_angle_topology_kmc_outputs_vs_reference(golden):
    """The flagship config of the reference (tests/integration/mdlmc_run.py:37-70): KMCLattice on
    AngleTopology + FermiAngle.  _determine_groups leaves trajectory frame 0 in the frame cache
    (topology.py:142-146), so continuous_output numbers it 0 and flushes it with the first event,
    xyz_output starts with it, the MSD / autocorrelation start there, and the reset / print
    phases count from it.  Event trace, frame numbers, flush times, the lattice every yielded
    frame is seen with, and the observable rows against the reference's own seeded run
    (tests/golden/angle_kmc.npz, oracle/make_golden.py gen_angle_kmc)."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import KMCLattice
    from cmdlmc_b200.topology import AngleTopology
    from cmdlmc_b200.trajectory import ArrayTrajectory
    g = golden("angle_kmc")
    w = synth.workload("C1")
    nfr = int(g["nframes"])
    frames = synth.trajectory(w, nfr, with_extra=True)
    names = np.array(["O"] * w.n_oxygen + ["P"] * w.n_extra)
    box = make_box(w.cell)

    def make(chunk):
        np.random.seed(int(g["seed"]))
        top = AngleTopology(ArrayTrajectory(frames, names, time_step=w.time_step), box, donor_atoms="O",
                            extra_atoms="P", group_size=w.group_size, cutoff=w.cutoff, buffer=w.buffer)
        return KMCLattice(top, atom_box=box, jumprate_function=cm.FermiAngle(*w.rate_params, np.pi / 2),
                          lattice_size=w.n_oxygen, proton_number=w.n_protons, donor_atoms="O",
                          time_step=w.time_step, rng="replay", chunk_size=chunk)

    for chunk in (1024, 41):
        # continuous output + the lattice xyz_output appends
        kmc = make(chunk)
        np.testing.assert_array_equal(kmc.lattice, g["cont_lattice0"])
        got = list(kmc._frames_with_lattice())
        ev = kmc.event_log
        ne = len(g["cont_ev_start"])
        n = min(ne, len(ev["start"]))
        assert n >= ne - 1 and n > 5
        np.testing.assert_array_equal(ev["frame"][:n], g["cont_ev_frame"][:n])
        np.testing.assert_array_equal(ev["start"][:n], g["cont_ev_start"][:n])
        np.testing.assert_array_equal(ev["dest"][:n], g["cont_ev_dest"][:n])
        np.testing.assert_allclose(ev["time"][:n], g["cont_ev_time"][:n], rtol=1e-12)
        ft, marks = g["cont_frame_times"], g["cont_lattice_marks"]
        m = min(len(ft), len(got))
        assert m >= len(ft) - 1 and m > 100
        assert got[0][0] == 0 and got[1][0] == 1
        np.testing.assert_array_equal([x[0] for x in got[:m]], ft[:m, 0].astype(int))
        np.testing.assert_allclose([x[1] for x in got[:m]], ft[:m, 1], rtol=1e-12)
        # frame number 0 is trajectory frame 0, which the topology never walked
        np.testing.assert_array_equal(got[0][2]["O"].atom_positions, frames[0, :w.n_oxygen])
        idx = np.arange(1, w.n_oxygen + 1)
        for (num, t, frame, lat), (mark_lat, mark_pos) in zip(got[:m], marks[:m]):
            assert float(np.dot(lat, idx)) == mark_lat, num
            assert frame["O"].atom_positions[np.where(lat > 0)[0]].sum() == pytest.approx(mark_pos, rel=1e-12)
        # observables
        rows = list(make(chunk).observables_output(100, 10))
        want = g["obs_obs"]
        assert len(want) > 10 and len(rows) >= len(want) - 1
        for (f, t, msd, auto), wr in zip(rows, want):
            assert f == int(wr[0])
            assert t == pytest.approx(wr[1], rel=1e-12)
            np.testing.assert_allclose(msd, wr[2:5], rtol=1e-9, atol=1e-12)
            assert auto == int(wr[5])


def test_event_log_overflow_is_an_error():
    """Philox mode has no stream to run out of: a block with more events than the log holds must
    not hand back outputs rebuilt from a truncated log."""
    w = synth.workload("C2")
    frames = synth.trajectory(w, 64)
    kmc = make_kmc(w, frames, rng="philox", seed=3)
    kmc.events_per_frame_bound = 1          # C2 makes ~3.4 events per frame
    with pytest.raises(RuntimeError, match="event log"):
        list(kmc)
    kmc = make_kmc(w, frames, rng="philox", seed=3)
    assert len(list(kmc)) > 40
