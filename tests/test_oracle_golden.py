"""Pins the CPU oracle (oracle/cmdlmc_oracle.c) against golden vectors produced by the REAL
reference (oracle/make_golden.py) -- the -m "not gpu" half of the parity story.

Tolerances: the reference's own build uses -O3 -ffast-math (setup.py:59) so it is itself only
defined up to reassociation; lengths/distances/angles are compared at 1e-12 relative (the
north-star bound), integer outputs (index sets, event traces, lattices) bit-exactly."""
import numpy as np
import pytest

from cmdlmc_b200 import synth

RTOL = 1e-12
CELLS = ["ortho", "cubic10", "diag9", "mono", "tri", "tri2"]


@pytest.mark.parametrize("name", CELLS)
def test_geometry_vs_reference(golden, orc, name):
    g = golden("geometry")
    box = orc.OracleBox(g[name + "_cell"])
    a, b, c = g[name + "_a"], g[name + "_b"], g[name + "_c"]
    np.testing.assert_allclose(box.length(a, b), g[name + "_length"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(box.distance(a, b), g[name + "_distance"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(box.angle(a, b, c), g[name + "_angle"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(box.length_all_to_all(a[:40], b[:50]), g[name + "_all"],
                               rtol=RTOL, atol=0)
    for i in range(20):
        idx, dist = box.next_neighbor(a[i], b[:60])
        assert idx == g[name + "_nn_idx"][i]
        assert dist == pytest.approx(g[name + "_nn_dist"][i], rel=RTOL)


def test_tie_behaviour(golden, orc):
    """Q10: +-L/2 stays in the ortho path (strict compare), C round() flips it in the general
    path (numpyatom.pyx:39-42 vs :72)."""
    g = golden("geometry")
    z = np.zeros((2, 3))
    t = np.array([[5.0, 5.0, 5.0], [-5.0, -5.0, -5.0]])
    np.testing.assert_array_equal(orc.OracleBox(g["cubic10_cell"]).distance(z, t), g["tie_ortho"])
    np.testing.assert_array_equal(orc.OracleBox(g["diag9_cell"]).distance(z, t), g["tie_general"])
    np.testing.assert_array_equal(g["tie_ortho"], t)
    np.testing.assert_array_equal(g["tie_general"], -t)


def test_extended_box(golden, orc):
    g = golden("geometry")
    box = orc.OracleBox(np.array([10.0, 10, 10]), box_multiplier=(2, 3, 4))
    np.testing.assert_array_equal(box.periodic_boundaries_extended, g["ext_pbc"])
    pos = np.array([box.position_extended_box(i, g["ext_frame"]) for i in range(5 * 24)])
    np.testing.assert_allclose(pos, g["ext_pos"], rtol=0, atol=1e-13)


def test_reference_known_answers(orc):
    """tests/cython_exts/LMC/test_AtomBox.py:19-64,161-174 restated on the oracle."""
    box = orc.OracleBox([10.0, 10, 10])
    a1, a2 = np.zeros(3), np.array([6.0, 6, 6])
    for i in range(-5, 5):
        assert box.length(a1, a2 + 10 * i)[0] == pytest.approx(np.sqrt(48))
    np.testing.assert_allclose(box.distance(a1, a2), [-4, -4, -4])
    assert box.angle(np.zeros(3), np.array([3.0, 0, 0]), np.array([3.0, 34, 0])) == \
        pytest.approx(np.pi / 2)
    atoms = np.array([[0.0, 0, 0], [1, 1, 1], [5, 5, 5], [10, 10, 10]])
    s3 = np.sqrt(3)
    want = np.array([[0, s3, 5 * s3, 0], [s3, 0, 4 * s3, s3], [5 * s3, 4 * s3, 0, 5 * s3],
                     [0, s3, 5 * s3, 0]])
    np.testing.assert_allclose(box.length_all_to_all(atoms, atoms), want)


def test_water_conversions(orc):
    """tests/cython_exts/LMC/test_AtomBox.py:177-226: exact equality of the ramp formula."""
    a, b, d0, lb, rb = 0.5, 2.3, 2.45, 2.3, 3.33
    box = orc.OracleBox([10.0, 10, 10])
    ramp = orc.OracleBox([10.0, 10, 10], conversion=dict(a=a, b=b, d0=d0, left_bound=lb,
                                                         right_bound=rb))
    z = np.zeros((1, 3))
    len1 = float(box.length(z, np.array([2.7, 0, 0]))[0])
    assert a * (len1 - d0) + b == float(ramp.length(z, np.array([2.7, 0, 0]))[0])
    assert float(ramp.length(z, np.array([2.3, 0, 0]))[0]) == 2.3  # on the bound: unchanged
    assert float(ramp.length(z, np.array([2.4, 0, 0]))[0]) == b
    lin = orc.OracleBox([10.0, 10, 10], conversion=dict(a=0.5, b=1.1, left_bound=2.2,
                                                        right_bound=3.3))
    assert float(lin.length(z, np.array([2.5, 0, 0]))[0]) == pytest.approx(0.5 * 2.5 + 1.1)


def test_topology_known_answer(golden, orc):
    """tests/topo/test_topology.py:32-65."""
    g = golden("topology")
    row, col, dist = orc.topology_bruteforce(orc.OracleBox([10.0, 10, 10]), g["kat_pos"], 2.0, 0)
    np.testing.assert_array_equal(row, [0, 0, 1, 1, 2, 4])
    np.testing.assert_array_equal(col, [1, 4, 0, 2, 1, 0])
    np.testing.assert_array_equal(dist, [1.5, 1.0, 1.5, 1.5, 1.5, 1.0])
    np.testing.assert_array_equal(row, g["kat_row"])
    np.testing.assert_array_equal(col, g["kat_col"])
    assert row.dtype == np.int32 and col.dtype == np.int32 and dist.dtype == np.float64


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_topology_and_verlet_vs_reference(golden, orc, cfg):
    g = golden("topology")
    w = synth.workload(cfg)
    nfr = int(g[cfg + "_nframes"])
    frames = synth.trajectory(w, nfr)
    box = orc.OracleBox(w.cell)
    row, col, dist = orc.topology_bruteforce(box, frames[0], w.cutoff, w.buffer)
    np.testing.assert_array_equal(row, g[cfg + "_bf_row"])
    np.testing.assert_array_equal(col, g[cfg + "_bf_col"])
    np.testing.assert_allclose(dist, g[cfg + "_bf_dist"], rtol=RTOL, atol=0)
    prev = None
    for k, (row, col, dist, rebuilt) in enumerate(orc.verlet_generator(box, frames, w.cutoff,
                                                                       w.buffer)):
        assert len(row) == g[cfg + "_verlet_counts"][k]
        assert np.sum(dist) == pytest.approx(g[cfg + "_verlet_dsum"][k], rel=1e-12)
        changed = prev is None or len(row) != len(prev[0]) or not (
            np.array_equal(row, prev[0]) and np.array_equal(col, prev[1]))
        assert changed == bool(g[cfg + "_verlet_changed"][k])
        if changed:
            assert rebuilt
        prev = (row, col)
        key = "%s_verlet_f%d_row" % (cfg, k)
        if key in g.files:
            np.testing.assert_array_equal(row, g[key])
            np.testing.assert_array_equal(col, g["%s_verlet_f%d_col" % (cfg, k)])
            np.testing.assert_allclose(dist, g["%s_verlet_f%d_dist" % (cfg, k)], rtol=RTOL)


def test_np_sum_order(orc):
    rng = np.random.RandomState(3)
    for n in list(range(0, 40)) + [127, 128, 129, 255, 256, 257, 1000, 1777, 8156, 65537]:
        x = rng.uniform(0, 1, n) * 10.0 ** rng.randint(-8, 8, n)
        assert orc.np_sum(x) == np.sum(x), n


def _oracle_rates_per_frame(orc, w, frames):
    box = orc.OracleBox(w.cell)
    fptr, starts, dests, omegas = [0], [], [], []
    for row, col, dist, _ in orc.verlet_generator(box, frames, w.cutoff, w.buffer):
        starts.append(row)
        dests.append(col)
        omegas.append(orc.rates(w.rate_kind, w.rate_params, dist))
        fptr.append(fptr[-1] + len(row))
    return np.array(fptr), np.concatenate(starts), np.concatenate(dests), np.concatenate(omegas)


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_kmc_replay_vs_reference(golden, orc, cfg):
    """Bit-exact event trace (frame, start, dest, proton) and final lattice against the
    reference's KMCLattice run with np.random.seed(seed); times to 1e-12."""
    g = golden("kmc")
    w = synth.workload(cfg)
    frames = synth.trajectory(w, int(g[cfg + "_nframes"]))
    fptr, start, dest, omega = _oracle_rates_per_frame(orc, w, frames)
    lattice = g[cfg + "_trace_lattice0"].copy()
    lat2, _ = synth.initial_lattice(w.n_oxygen, w.n_protons, int(g[cfg + "_seed"]))
    np.testing.assert_array_equal(lattice, lat2)
    u = g[cfg + "_trace_u"]
    ne = len(g[cfg + "_trace_ev_time"])
    res = orc.kmc_replay(fptr, start, dest, omega, lattice, w.time_step, u, ne)
    assert res["n_events"] >= ne - 1
    n = min(ne, res["n_events"])
    assert n > 100
    np.testing.assert_array_equal(res["frame"][:n], g[cfg + "_trace_ev_frame"][:n])
    np.testing.assert_array_equal(res["dframe"][:n], g[cfg + "_trace_ev_dframe"][:n])
    np.testing.assert_array_equal(res["start"][:n], g[cfg + "_trace_ev_start"][:n])
    np.testing.assert_array_equal(res["dest"][:n], g[cfg + "_trace_ev_dest"][:n])
    np.testing.assert_array_equal(res["proton"][:n], g[cfg + "_trace_ev_proton"][:n])
    np.testing.assert_allclose(res["time"][:n], g[cfg + "_trace_ev_time"][:n], rtol=1e-12)
    if n == ne:
        np.testing.assert_array_equal(lattice, g[cfg + "_trace_lattice_final"])
    # frame time stamps as yielded by iter(kmc) (MDMC.py:94-96)
    ft = g[cfg + "_trace_frame_times"]
    fe = res["frame_event"]
    for fn, t in ft:
        e = fe[int(fn)]
        if 0 <= e < n:
            assert res["time"][e] == pytest.approx(t, rel=1e-12)


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_observables_vs_reference(golden, orc, cfg):
    g = golden("kmc")
    w = synth.workload(cfg)
    frames = synth.trajectory(w, int(g[cfg + "_nframes"]))
    fptr, start, dest, omega = _oracle_rates_per_frame(orc, w, frames)
    lattice0 = g[cfg + "_obs_lattice0"].copy()
    lattice = lattice0.copy()
    u = g[cfg + "_obs_u"]
    ne = len(g[cfg + "_obs_ev_time"])
    res = orc.kmc_replay(fptr, start, dest, omega, lattice, w.time_step, u, ne)
    obs = orc.observables(orc.OracleBox(w.cell), frames, lattice0, res, 100, 10)
    want = g[cfg + "_obs_obs"]
    assert len(want) > 0
    assert len(obs) >= len(want) - 1
    for (f, t, msd, auto), wrow in zip(obs, want):
        assert f == int(wrow[0])
        assert t == pytest.approx(wrow[1], rel=1e-12)
        np.testing.assert_allclose(msd, wrow[2:5], rtol=1e-9, atol=1e-12)
        assert auto == int(wrow[5])


def test_fastforward_vs_reference(golden, orc):
    """tests/LMC/test_MDMC.py:10-73 rate streams through the restated time stepper."""
    g = golden("fastforward")
    k = 0
    for dt in (0.1, 0.5, 1.3):
        for omega in (0.03, 0.06, 0.13):
            rows = g["const_%d" % k]
            assert tuple(g["const_%d_par" % k]) == (dt, omega)
            k += 1
            got = orc.fastforward([omega], dt, g["const_u"], len(rows))
            np.testing.assert_array_equal(got[:, :2], rows[:, :2])
            np.testing.assert_allclose(got[:, 2], rows[:, 2], rtol=1e-14)
            # the reference's own assertions (test_MDMC.py:44-51)
            t_fixed = np.cumsum(-np.log(1 - g["const_u"][:len(rows)]) / omega)
            np.testing.assert_allclose(got[:, 2], t_fixed, atol=1e-7)
            assert ((got[:, 2] // dt).astype(int) == got[:, 0].astype(int)).all()
    rows = g["sin_rows"]
    got = orc.fastforward(g["sin_rates"], 0.5, g["sin_u"], len(rows))
    np.testing.assert_array_equal(got[:, :2], rows[:, :2])
    np.testing.assert_allclose(got[:, 2], rows[:, 2], rtol=1e-14)
    # test_variable_rates_index (test_MDMC.py:76-93): the event lands on the non-zero frame
    rates = np.zeros(117)
    rates[73] = 0.17
    u = np.random.RandomState(1).random_sample(101)
    got = orc.fastforward(rates, 0.22, u, 101)
    assert (got[1:, 0].astype(int) % 117 == 73).all()


def test_legacy_rates_and_mt(orc):
    """A9' / GSL-flavoured stream: self-consistency only (PARITY UNPINNED upstream)."""
    x = np.linspace(2.0, 4.0, 50)
    fermi = orc.rates("Fermi", (0.06, 2.3, 0.1), x)
    np.testing.assert_allclose(fermi, 0.06 / (1 + np.exp((x - 2.3) / 0.1)), rtol=1e-14)
    fa = orc.rates("FermiAngle", (0.06, 2.3, 0.1, np.pi / 2), x, np.linspace(0, np.pi, 50))
    assert (fa[:25] == 0).all() and (fa[25:] == fermi[25:]).all()
    ae = orc.rates("ActivationEnergy", (0.06, 1.2, 30.0, 2.2, 510.0), x)
    assert ae[0] == 0.06 and (np.diff(ae) <= 0).all()
    ex = orc.rates("Exponential", (2.0, -1.5), x)
    np.testing.assert_allclose(ex, 2.0 * np.exp(-1.5 * x), rtol=1e-14)
    # MT19937 == NumPy's legacy RandomState core generator
    mt = orc.MT19937(12345)
    rs = np.random.RandomState(12345)
    want = rs.random_sample(20)
    got = np.array([mt.double53() for _ in range(20)])
    np.testing.assert_array_equal(got, want)
    mt = orc.MT19937(7)
    ks = [mt.gsl_uniform_int(1000) for _ in range(1000)]
    assert 0 <= min(ks) and max(ks) < 1000


def test_gsl_stream_helper_matches_oracle_mt(orc):
    """cmdlmc_b200.lmc.gsl_streams (host helper feeding the LMC replay mode) == the oracle's
    MT19937 with GSL's uniform_int / uniform conventions (PARITY UNPINNED upstream)."""
    from cmdlmc_b200 import lmc
    counts = [7, 0, 1000, 3, 4093]
    pick, acc = lmc.gsl_streams(99, counts, sweeps_per_frame=2)
    mt = orc.MT19937(99)
    k = 0
    for p in np.repeat(counts, 2):
        for _ in range(int(p)):
            assert pick[k] == mt.gsl_uniform_int(int(p))
            assert acc[k] == mt.gsl_uniform()
            k += 1
    assert k == len(pick)
    raw = lmc.mt19937_u32(5, 10)
    mt = orc.MT19937(5)
    assert [int(x) for x in raw] == [mt.u32() for _ in range(10)]


def test_angle_topology_vs_reference(golden, orc):
    """AngleTopology colvars + FermiAngle (reference run in oracle/make_golden.py gen_angle) vs the
    oracle restatement: same pair list, angles and masked rates."""
    g = golden("angle")
    w = synth.workload("C1")
    two = synth.trajectory(w, 2, with_extra=True)
    # groups come from trajectory frame 0, the first yielded topology from frame 1 (the reference
    # consumes frame 0 in AngleTopology._determine_groups, topology.py:43,145)
    frames = two[1]
    oxy, pho = frames[:w.n_oxygen], frames[w.n_oxygen:]
    obox = orc.OracleBox(w.cell)
    start, dest, dist = orc.topology_bruteforce(obox, oxy, w.cutoff, w.buffer)
    np.testing.assert_array_equal(start, g["start0"])
    np.testing.assert_array_equal(dest, g["dest0"])
    np.testing.assert_allclose(dist, g["dist0"], rtol=1e-12)
    d_po = obox.length_all_to_all(two[0][w.n_oxygen:], two[0][:w.n_oxygen])
    group = np.full(w.n_oxygen, -1)
    for p_index, os_ in enumerate(np.argsort(d_po, axis=1)[:, :w.group_size]):
        group[os_] = p_index
    np.testing.assert_array_equal(group, g["group"])
    ang = obox.angle(pho[group[start]], oxy[start], oxy[dest])
    np.testing.assert_allclose(ang, g["angle0"], rtol=1e-12)
    rate = orc.rates("FermiAngle", tuple(w.rate_params) + (np.pi / 2,), dist, ang)
    np.testing.assert_allclose(rate, g["rate0"], rtol=1e-10)
    assert ((rate == 0) == (g["rate0"] == 0)).all()
