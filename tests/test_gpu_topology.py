"""GPU parity, rows A7-A9: neighbour lists (index sets BIT-EXACT, distances bit-exact vs the
oracle / 1e-12 vs the reference's golden vectors), Verlet schedule, fused rates."""
import numpy as np
import pytest

from cmdlmc_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def make_box(cell):
    import cmdlmc_b200 as cm
    cell = np.asarray(cell, dtype=float)
    return cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)


def make_traj(frames, time_step=0.5, name="O"):
    from cmdlmc_b200.trajectory import ArrayTrajectory
    return ArrayTrajectory(frames, np.array([name] * frames.shape[1]), time_step=time_step)


def test_known_answer_reference_test(golden):
    """tests/topo/test_topology.py:32-65."""
    from cmdlmc_b200.topology import NeighborTopology
    g = golden("topology")
    pos = g["kat_pos"]
    top = NeighborTopology(make_traj(pos[None]), make_box([10.0, 10, 10]), cutoff=2.0, buffer=0,
                           donor_atoms="O")
    row, col, dist = top.get_topology_bruteforce(pos)
    np.testing.assert_array_equal(row, [0, 0, 1, 1, 2, 4])
    np.testing.assert_array_equal(col, [1, 4, 0, 2, 1, 0])
    np.testing.assert_array_equal(dist, [1.5, 1.0, 1.5, 1.5, 1.5, 1.0])
    assert row.dtype == np.int32 and col.dtype == np.int32 and dist.dtype == np.float64


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_bruteforce_and_verlet_vs_reference_golden(golden, cfg):
    from cmdlmc_b200.topology import NeighborTopology
    g = golden("topology")
    w = synth.workload(cfg)
    nfr = int(g[cfg + "_nframes"])
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    top = NeighborTopology(make_traj(frames, w.time_step), box, donor_atoms="O", cutoff=w.cutoff,
                           buffer=w.buffer)
    row, col, dist = top.get_topology_bruteforce(frames[0])
    np.testing.assert_array_equal(row, g[cfg + "_bf_row"])
    np.testing.assert_array_equal(col, g[cfg + "_bf_col"])
    np.testing.assert_allclose(dist, g[cfg + "_bf_dist"], rtol=RTOL, atol=0)
    top.chunk_size = 37   # exercise the carried Verlet state across blocks
    prev = None
    k = -1
    for k, (row, col, dist, _) in enumerate(top.topology_verlet_list_generator()):
        assert len(row) == g[cfg + "_verlet_counts"][k]
        assert np.sum(dist) == pytest.approx(g[cfg + "_verlet_dsum"][k], rel=1e-12)
        changed = prev is None or len(row) != len(prev[0]) or not (
            np.array_equal(row, prev[0]) and np.array_equal(col, prev[1]))
        assert changed == bool(g[cfg + "_verlet_changed"][k])
        prev = (row, col)
        key = "%s_verlet_f%d_row" % (cfg, k)
        if key in g.files:
            np.testing.assert_array_equal(row, g[key])
            np.testing.assert_array_equal(col, g["%s_verlet_f%d_col" % (cfg, k)])
            np.testing.assert_allclose(dist, g["%s_verlet_f%d_dist" % (cfg, k)], rtol=RTOL)
    assert k == nfr - 1


@pytest.mark.parametrize("cfg,nfr", [("C1", 300), ("C2", 200), ("C4", 64)])
def test_bit_exact_vs_oracle(orc, cfg, nfr):
    """Every frame: identical (row, col) arrays and bit-identical distances; rates to 1e-10."""
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    rate = cm.Fermi(*w.rate_params)
    for mode in (0, 1):
        topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                           mode, rate, cap), frames)
        counts, rebuilt, rate_sum = topo.frame_info()
        oracle_iter = orc.verlet_generator(obox, frames, w.cutoff, w.buffer) if mode == 1 else (
            orc.topology_bruteforce(obox, fr, w.cutoff, w.buffer) + (True,) for fr in frames)
        n_reb = 0
        for f, (orow, ocol, odist, oreb) in enumerate(oracle_iter):
            start, dest, dist, omega = topo.get_frame(f, int(counts[f]))
            np.testing.assert_array_equal(start, orow)
            np.testing.assert_array_equal(dest, ocol)
            np.testing.assert_array_equal(dist, odist)
            assert bool(rebuilt[f]) == bool(oreb), f
            n_reb += bool(oreb)
            want = orc.rates("Fermi", w.rate_params, odist)
            np.testing.assert_allclose(omega, want, rtol=1e-10, atol=0)
            assert rate_sum[f] == pytest.approx(want.sum(), rel=1e-9)
        assert f == nfr - 1
        if mode == 1:
            assert 1 < n_reb < nfr // 4    # the schedule really alternates


def test_reference_verlet_equals_bruteforce():
    """tests/topo/test_topology.py:68-101: Verlet list == brute force for a random walk, arrays
    compared with assert_array_equal."""
    from cmdlmc_b200.topology import NeighborTopology
    rng = np.random.RandomState(0)
    pos = rng.uniform(0, 10, size=(5, 3))
    frames = []
    for _ in range(51):
        pos = pos + rng.normal(size=(5, 3), scale=1)
        frames.append(pos.copy())
    frames = np.array(frames)
    box = make_box([10.0, 10, 10])
    top1 = NeighborTopology(make_traj(frames, name="H"), box, cutoff=3, buffer=10, donor_atoms="H")
    top2 = NeighborTopology(make_traj(frames, name="H"), box, cutoff=3, buffer=10, donor_atoms="H")
    n = 0
    for n1, n2 in zip(top1.topology_verlet_list_generator(), top2.topology_bruteforce_generator()):
        np.testing.assert_array_equal(n1[0], n2[0])
        np.testing.assert_array_equal(n1[1], n2[1])
        np.testing.assert_array_equal(n1[2], n2[2])
        n += 1
    assert n == 51


def test_edge_cases(orc):
    from cmdlmc_b200.topology import NeighborTopology
    box, obox = make_box([10.0, 10, 10]), orc.OracleBox([10.0, 10, 10])
    # coincident atoms: a pair at exactly 0.0 is dropped (sparse zero, topology.py:68-69)
    pos = np.array([[1.0, 1, 1], [1.0, 1, 1], [2.0, 1, 1], [11.0, 11, 11]])
    top = NeighborTopology(make_traj(pos[None]), box, cutoff=3.0, buffer=0.0, donor_atoms="O")
    got = top.get_topology_bruteforce(pos)
    want = orc.topology_bruteforce(obox, pos, 3.0, 0.0)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert (0, 1) not in set(zip(got[0], got[1])) and (0, 3) not in set(zip(got[0], got[1]))
    # exact-threshold pair: dist <= cutoff + buffer keeps it, one ulp beyond drops it
    p2 = np.array([[0.0, 0, 0], [3.0, 0, 0], [0, np.nextafter(3.0, 4.0), 0]])
    got = NeighborTopology(make_traj(p2[None]), box, cutoff=3.0, buffer=0.0,
                           donor_atoms="O").get_topology_bruteforce(p2)
    np.testing.assert_array_equal(got[0], [0, 1])
    np.testing.assert_array_equal(got[1], [1, 0])
    # no neighbours at all, single atom, odd/even atom counts
    far = np.array([[0.0, 0, 0], [5.0, 5, 5]])
    assert len(NeighborTopology(make_traj(far[None]), box, cutoff=1.0, buffer=0.0,
                                donor_atoms="O").get_topology_bruteforce(far)[0]) == 0
    rng = np.random.RandomState(5)
    for n in (1, 2, 3, 32, 33, 64, 65, 257, 700, 1024):
        p = rng.uniform(0, 10, size=(n, 3))
        cut = 2.0 if n < 200 else 0.8
        got = NeighborTopology(make_traj(p[None]), box, cutoff=cut, buffer=0.5,
                               donor_atoms="O").get_topology_bruteforce(p)
        want = orc.topology_bruteforce(obox, p, cut, 0.5)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)


def test_float32_frames_and_full_size_properties():
    """HDF5-style float32 storage is up-cast on the device; at full C2 frame counts check
    size-independent properties: symmetry, sortedness, count parity, rate range."""
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload("C2")
    frames = synth.trajectory(w, 2048)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    t64 = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate,
                                                      cap), frames.astype(np.float32).astype(np.float64))
    t32 = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate,
                                                      cap), frames.astype(np.float32))
    c64, _, s64 = t64.frame_info()
    c32, _, s32 = t32.frame_info()
    np.testing.assert_array_equal(c64, c32)
    np.testing.assert_array_equal(s64, s32) if False else np.testing.assert_allclose(s64, s32, rtol=1e-12)
    assert (c64 % 2 == 0).all() and (c64 > 0).all()
    for f in (0, 1000, 2047):
        start, dest, dist, omega = t32.get_frame(f, int(c32[f]))
        key = start.astype(np.int64) * w.n_oxygen + dest
        assert (np.diff(key) > 0).all()                       # row-major, columns ascending
        fwd = dict(zip(zip(start, dest), dist))
        assert all(fwd[(d, s)] == v for (s, d), v in fwd.items())   # symmetric, bitwise
        assert (dist <= w.cutoff + w.buffer).all() and (dist > 0).all()
        assert (omega > 0).all() and (omega <= w.rate_params[0]).all()


@pytest.mark.parametrize("mode,dtype", [(0, np.float64), (0, np.float32), (1, np.float32)])
def test_pageable_blocks_go_through_the_staging_ring(mode, dtype):
    """A plain NumPy block (pageable) is staged through the library's page-locked ring, a block in
    runtime.pinned_empty memory is copied from directly; both give the same lists.  The block is
    larger than the ring (3 x 8 MiB), so slots are recycled while their DMA is in flight."""
    from cmdlmc_b200 import runtime
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload("C2")
    nfr = 6144 if dtype == np.float32 else 3072            # 29.5 MB either way
    frames = np.ascontiguousarray(synth.trajectory(w, nfr).astype(dtype))
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    make = lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, mode, rate, cap)
    s0 = runtime.staging_stats()
    tp = build_with_retry(make, frames)
    s1 = runtime.staging_stats()
    assert s1[0] - s0[0] >= frames.nbytes and s1[2] >= 1      # went through the ring
    pinned = runtime.pinned_empty(frames.shape, dtype)
    pinned[...] = frames
    tq = build_with_retry(make, pinned)
    s2 = runtime.staging_stats()
    assert s2[0] == s1[0] and s2[1] - s1[1] >= frames.nbytes  # copied from directly
    cp, rp_, sp = tp.frame_info()
    cq, rq, sq = tq.frame_info()
    np.testing.assert_array_equal(cp, cq)
    np.testing.assert_array_equal(rp_, rq)
    np.testing.assert_array_equal(sp, sq)
    for f in (0, nfr // 2, nfr - 1):
        for a, b in zip(tp.get_frame(f, int(cp[f])), tq.get_frame(f, int(cq[f]))):
            np.testing.assert_array_equal(a, b)


def _lists(box, w, frames, rate):
    """One launch over the whole resident block (host blocks are cut into upload chunks, and a
    chunk of a few hundred frames leaves a persistent CTA too few frames for a list)."""
    import torch
    from cmdlmc_b200.topology import DeviceTopology
    d = torch.from_numpy(frames).cuda()
    t = DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate, 0, path=0)
    t.build_dev(d.data_ptr(), frames.shape[0])
    counts, _, rsum = t.frame_info()
    assert (counts > 0).all()
    start, dest, dist, om = t.get_block(0, len(counts), counts, omega=True)
    pad = np.arange(start.shape[1])[None, :] >= counts[:, None]     # beyond a frame's count: not written
    for a in (start, dest, dist, om):
        a[pad] = 0
    return t, counts, rsum, start, dest, dist, om


@pytest.mark.parametrize("cfg", ["C1", "C2", "T333"])
def test_skin_list_of_the_dense_kernel_changes_no_bit(orc, monkeypatch, cfg):
    """The dense kernel keeps the filter's candidate set (radius + skin) across consecutive frames
    and re-filters it while the two largest displacements stay below the skin
    (pairs_dense.cuh).  Whatever the skin -- none, the default, a huge one, a list too short to
    hold the candidates -- the lists, distances and rates are the same bits, and they are the
    oracle's.  The trajectory holds what the displacement test has to survive: atoms re-imaged by
    whole cell vectors between frames, one atom jumping 3 A in a single step, a frame repeated."""
    import cmdlmc_b200 as cm
    if cfg == "T333":   # a full triclinic cell and an odd atom count (no TMA: plain frame loads)
        w = synth.Workload("T333", synth._triclinic_cell(24.0, 23.0, 25.0, 0.15, -0.2, 0.1), 333, 0, 100,
                           4096, 0.5, 3.0, 2.0, "Fermi", (0.06, 2.3, 0.1), 21, group_size=0)
    else:
        w = synth.workload(cfg)
    nfr = 4096                                   # several frames for every resident CTA
    frames = synth.trajectory(w, nfr)
    rng = np.random.RandomState(11)
    cellm = w.cell_matrix
    for f in range(40, nfr, 97):                 # re-imaging: whole cell vectors, from here on
        a = rng.randint(w.n_oxygen)
        frames[f:, a] += rng.randint(-1, 2, size=3) @ cellm
    frames[700:, 5] += np.array([1.9, -1.7, 1.6])          # a 3 A jump inside one time step
    frames[901] = frames[900]                               # no displacement at all
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    rate = cm.Fermi(*w.rate_params)
    runs = {}
    for name, env in (("direct", {"CMDLMC_B200_DENSE_SKIN": "0"}), ("default", {}),
                      ("wide", {"CMDLMC_B200_DENSE_SKIN": "1.5"}),
                      ("short list", {"CMDLMC_B200_DENSE_LIST_CAP": "3000"})):
        for k in ("CMDLMC_B200_DENSE_SKIN", "CMDLMC_B200_DENSE_LIST_CAP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        runs[name] = _lists(box, w, frames, rate)
    fr_d, reb_d, _ = runs["direct"][0].skin_stats()
    fr_s, reb_s, _ = runs["default"][0].skin_stats()
    assert fr_d == 0 and fr_s >= nfr - 16 and 0 < reb_s < fr_s // 2      # the list really is reused
    fr_l, reb_l, _ = runs["short list"][0].skin_stats()
    if cfg != "C1":
        assert fr_l < fr_s                        # frames whose list did not fit went the direct way
    ref = runs["direct"]
    for name, got in runs.items():
        for a, b in zip(ref[1:], got[1:]):
            np.testing.assert_array_equal(a, b, err_msg=name)
    _, counts, _, start, dest, dist, _ = ref
    for f in (0, 39, 40, 41, 699, 700, 701, 900, 901, 2500, nfr - 1):
        orow, ocol, odist = orc.topology_bruteforce(obox, frames[f], w.cutoff, w.buffer)
        c = int(counts[f])
        np.testing.assert_array_equal(start[f, :c], orow)
        np.testing.assert_array_equal(dest[f, :c], ocol)
        np.testing.assert_array_equal(dist[f, :c], odist)


@pytest.mark.parametrize("mode,dtype,nfr", [(0, np.float32, 2304), (1, np.float32, 300), (1, np.float64, 300)])
def test_donor_selection_on_the_device_equals_the_host_gather(mode, dtype, nfr):
    """Whole frames (oxygens and the heavy atoms they are bonded to, in shuffled order) go up as they
    lie in memory and a gather kernel behind the copy picks the donor rows
    (cmd_topo_set_selection); the lists are those of a block gathered on the host first -- through
    the chunked brute-force upload and through the one-piece Verlet upload."""
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload("C2")
    allpos = synth.trajectory(w, nfr, with_extra=True)                 # [F, 400 O + 100 P, 3]
    order = np.random.RandomState(2).permutation(allpos.shape[1])
    allpos = np.ascontiguousarray(allpos[:, order].astype(dtype))
    rows = np.flatnonzero(order < w.n_oxygen).astype(np.int32)         # where the oxygens ended up
    rows = rows[np.argsort(order[rows])]                               # donor i = oxygen i
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)

    def on_device(cap):
        t = DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, mode, rate, cap)
        t.set_selection(allpos.shape[1], rows)
        return t
    td = build_with_retry(on_device, allpos)
    th = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, mode, rate, cap),
                          np.ascontiguousarray(allpos[:, rows]))
    cd, rd, sd = td.frame_info()
    ch, rh, sh = th.frame_info()
    np.testing.assert_array_equal(cd, ch)
    np.testing.assert_array_equal(rd, rh)
    np.testing.assert_array_equal(sd, sh)
    for a, b in zip(td.get_block(0, nfr, cd, omega=True), th.get_block(0, nfr, ch, omega=True)):
        pad = np.arange(a.shape[1])[None, :] >= cd[:, None]
        a[pad] = 0
        b[pad] = 0
        np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        td.build(allpos[:, rows])          # donor-only block while the selection is on


@pytest.mark.parametrize("cfg,nfr", [("C1", 6), ("C2", 6), ("C4", 3)])
def test_cell_list_path_equals_dense(orc, cfg, nfr):
    """The cell-list search (large boxes) and the dense search give the same arrays, bit for bit."""
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    res = []
    for path in (0, 1):
        t = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate,
                                                        cap, path=path), frames)
        assert t.path == path
        counts = t.frame_info()[0]
        res.append([t.get_frame(f, int(counts[f])) for f in range(nfr)])
    for fa, fb in zip(*res):
        for a, b in zip(fa, fb):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("cfg,path,mode", [("C2", 0, 0), ("C2", 0, 1), ("C3", 1, 0), ("C4", 1, 0)])
def test_structural_zeros_of_the_cell_matrix_change_no_bit(monkeypatch, cfg, path, mode):
    """C2/C4 (monoclinic, four non-zero entries) and C3 (upper-triangular h) take the shortened
    matrix products of pbc.cuh matvec3_norm_sp; CMDLMC_B200_NO_SPARSE=1 forces the full products of
    math_helper.pyx:50-60.  Lists, distances and rates have to agree bit for bit."""
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload(cfg)
    nfr = 12
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params) if w.rate_kind == "Fermi" else cm.ActivationEnergy(*w.rate_params)
    res = []
    for full in (False, True):
        if full:
            monkeypatch.setenv("CMDLMC_B200_NO_SPARSE", "1")
        else:
            monkeypatch.delenv("CMDLMC_B200_NO_SPARSE", raising=False)
        t = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, mode, rate,
                                                        cap, path=path), frames)
        counts = t.frame_info()[0]
        res.append([t.get_frame(f, int(counts[f])) for f in range(nfr)])
    for fa, fb in zip(*res):
        for a, b in zip(fa, fb):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("cfg,nfr", [("C3", 3)])
def test_large_box_vs_oracle(orc, cfg, nfr):
    """C3: 2048 O in a triclinic cell, activation-energy rate -- cell-list path vs the oracle's
    all-pairs loop (index sets and distances bit-exact, rates 1e-10)."""
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    rate = cm.ActivationEnergy(*w.rate_params)
    t = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate, cap),
                         frames)
    assert t.path == 1
    counts = t.frame_info()[0]
    for f in range(nfr):
        start, dest, dist, omega = t.get_frame(f, int(counts[f]))
        want = orc.topology_bruteforce(obox, frames[f], w.cutoff, w.buffer)
        np.testing.assert_array_equal(start, want[0])
        np.testing.assert_array_equal(dest, want[1])
        np.testing.assert_array_equal(dist, want[2])
        np.testing.assert_allclose(omega, orc.rates("ActivationEnergy", w.rate_params, want[2]),
                                   rtol=1e-10)


def test_cell_list_random_boxes(orc):
    """Random orthorhombic and triclinic boxes through the cell list, incl. axes with fewer than
    three cells (collapsed) and a Verlet run across blocks."""
    from cmdlmc_b200.topology import NeighborTopology, DeviceTopology, build_with_retry
    rng = np.random.RandomState(11)
    cells = [np.array([31.0, 12.0, 47.0]),                              # y: 2 cells -> collapsed
             np.array([40.0, 0, 0, 7.0, 36.0, 0, -5.0, 9.0, 33.0]),
             np.array([14.0, 0, 0, 3.0, 13.0, 0, 2.0, 1.0, 60.0])]      # x, y collapsed
    for cell in cells:
        box, obox = make_box(cell), orc.OracleBox(cell)
        hm = np.diag(cell) if cell.size == 3 else cell.reshape(3, 3)
        for n in (1500, 2100):
            p = rng.uniform(-0.3, 1.3, size=(n, 3)) @ hm      # also outside the cell
            top = NeighborTopology(make_traj(p[None]), box, cutoff=3.2, buffer=1.0, donor_atoms="O")
            got = top.get_topology_bruteforce(p)
            want = orc.topology_bruteforce(obox, p, 3.2, 1.0)
            for a, b in zip(got, want):
                np.testing.assert_array_equal(a, b)
    # Verlet schedule + refresh on the cell path, blocks of 7 frames
    w = synth.workload("C3")
    frames = synth.trajectory(w, 20)
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    top = NeighborTopology(make_traj(frames, w.time_step), box, donor_atoms="O", cutoff=w.cutoff,
                           buffer=w.buffer)
    top.chunk_size = 7
    gen = orc.verlet_generator(obox, frames, w.cutoff, w.buffer)
    k = -1
    for k, ((row, col, dist, _), want) in enumerate(zip(top.topology_verlet_list_generator(), gen)):
        np.testing.assert_array_equal(row, want[0])
        np.testing.assert_array_equal(col, want[1])
        np.testing.assert_array_equal(dist, want[2])
    assert k == 19


def test_c5_water_box_cell_list_vs_oracle(orc):
    """C5: 32 768 O in a 99.4 A orthorhombic box (cell-list path, 19^3 cells): one frame against
    the oracle's all-pairs loop (5.4e8 pairs), plus size-independent properties on more frames."""
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload("C5")
    frames = synth.trajectory(w, 3)
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    rate = cm.Fermi(*w.rate_params)
    t = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate, cap),
                         frames)
    assert t.path == 1
    counts = t.frame_info()[0]
    start, dest, dist, omega = t.get_frame(0, int(counts[0]))
    want = orc.topology_bruteforce(obox, frames[0], w.cutoff, w.buffer)
    np.testing.assert_array_equal(start, want[0])
    np.testing.assert_array_equal(dest, want[1])
    np.testing.assert_array_equal(dist, want[2])
    for f in (1, 2):
        s, d, di, om = t.get_frame(f, int(counts[f]))
        key = s.astype(np.int64) * w.n_oxygen + d
        assert (np.diff(key) > 0).all() and (di <= w.cutoff + w.buffer).all() and (di > 0).all()
        rev = np.argsort(d.astype(np.int64) * w.n_oxygen + s, kind="stable")
        np.testing.assert_array_equal(di[rev], di)          # (j, i) carries the same distance
    hist = t.distance_histogram(0.0, 5.0, 500)
    assert hist.sum() == counts.sum()


def test_angle_topology_and_fermi_angle_vs_reference(golden):
    """F1: AngleTopology (P-O...O angle colvar) + FermiAngle against the reference's own run on
    the C1 integration config with its P atoms (tests/golden/angle.npz)."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import AngleTopology
    from cmdlmc_b200.trajectory import ArrayTrajectory
    g = golden("angle")
    w = synth.workload("C1")
    nfr = int(g["nframes"])
    # like upstream, building the groups consumes trajectory frame 0: nfr topologies need nfr + 1
    frames = synth.trajectory(w, nfr + 1, with_extra=True)
    names = np.array(["O"] * w.n_oxygen + ["P"] * w.n_extra)
    box = make_box(w.cell)
    rate = cm.FermiAngle(*w.rate_params, np.pi / 2)

    def make():
        top = AngleTopology(ArrayTrajectory(frames, names, time_step=w.time_step), box,
                            donor_atoms="O", extra_atoms="P", group_size=w.group_size,
                            cutoff=w.cutoff, buffer=w.buffer)
        top.chunk_size = 5
        return top
    top = make()
    np.testing.assert_array_equal(top._group, g["group"])
    assert len(list(top.get_cached_frames())) == int(g["cached_frames_after_init"])
    # 1. the reference's iteration protocol: (start, dest, dist, angle) per frame
    k = -1
    for k, (start, dest, dist, angle) in enumerate(top):
        assert len(start) == g["counts"][k]
        assert angle.sum() == pytest.approx(g["angle_sum"][k], rel=1e-12)
        r = rate(dist, angle)
        assert r.sum() == pytest.approx(g["rate_sum"][k], rel=1e-10)
        if k == 0:
            np.testing.assert_array_equal(start, g["start0"])
            np.testing.assert_array_equal(dest, g["dest0"])
            np.testing.assert_allclose(dist, g["dist0"], rtol=1e-12)
            np.testing.assert_allclose(angle, g["angle0"], rtol=1e-12)
            np.testing.assert_allclose(r, g["rate0"], rtol=1e-10)
            assert ((r == 0) == (g["rate0"] == 0)).all() and (r == 0).sum() > 100
            # the host-level _determine_colvars (one batched angle call) gives the same angles
            full = next(iter(ArrayTrajectory(frames[1:2], names, time_step=w.time_step)))
            np.testing.assert_array_equal(top._determine_colvars(start, dest, dist, full)[3], angle)
    assert k == nfr - 1
    np.testing.assert_allclose(angle, g["angle_last"], rtol=1e-12)
    # 2. device pipeline: the fused rates carry the FermiAngle mask
    top = make()
    top.attach_jumprate(rate)
    f = 0
    for topo, full_frames, _ in top.device_blocks():
        counts, _, rsum = topo.frame_info()
        for j in range(len(full_frames)):
            om = topo.get_frame(j, int(counts[j]))[3]
            assert om.sum() == pytest.approx(g["rate_sum"][f], rel=1e-10)
            assert rsum[j] == pytest.approx(g["rate_sum"][f], rel=1e-10)
            if f == 0:
                np.testing.assert_allclose(om, g["rate0"], rtol=1e-10)
            f += 1
    assert f == nfr
    # 3. KMC on top of it runs and only uses unmasked transitions
    from cmdlmc_b200.kmc import KMCLattice
    np.random.seed(5)
    frames = synth.trajectory(w, 600, with_extra=True)      # long enough for a few events
    kmc = KMCLattice(make(), atom_box=box, jumprate_function=rate, lattice_size=w.n_oxygen,
                     proton_number=w.n_protons, donor_atoms="O", time_step=w.time_step,
                     extra_atoms="P", chunk_size=128)
    n = sum(1 for _ in kmc)
    ev = kmc.event_log
    assert n > 0 and len(ev["frame"]) > 0


def test_full_size_c2_trajectory_properties():
    """BASELINE.json config 2 at its full size (100 000 frames, float32 storage like the HDF5
    path) through the Verlet pipeline in blocks: size-independent properties of the reference
    (tests/topo/test_topology.py:68-101: a Verlet list restricted to cutoff+buffer equals the
    brute-force list) on sampled frames, list symmetry / order on every sampled frame, and
    conservation laws over the whole run."""
    from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE, MODE_VERLET, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload("C2")
    total, block = 100000, 20000
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    verlet = bf = None
    rng = np.random.RandomState(0)
    n_rebuilds = n_pairs = 0
    for b0 in range(0, total, block):
        fr32 = synth.trajectory(w, block, start=b0, dtype=np.float32)
        if verlet is None:
            verlet = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                                 MODE_VERLET, rate, cap), fr32)
        else:
            verlet.build(fr32)
        counts, rebuilt, rsum = verlet.frame_info()
        assert (counts > 0).all() and (counts % 2 == 0).all() and np.isfinite(rsum).all()
        n_rebuilds += int(rebuilt.sum())
        n_pairs += int(counts.sum())
        sample = np.sort(rng.choice(block, 6, replace=False))
        sub = np.ascontiguousarray(fr32[sample])
        if bf is None:
            bf = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                             MODE_BRUTEFORCE, rate, cap), sub)
        else:
            bf.build(sub)
        bcounts = bf.frame_info()[0]
        for j, f in enumerate(sample):
            s, d, dist, om = verlet.get_frame(int(f), int(counts[f]))
            key = s.astype(np.int64) * w.n_oxygen + d
            assert (np.diff(key) > 0).all()                          # row-major, columns ascending
            order = np.argsort(d.astype(np.int64) * w.n_oxygen + s, kind="stable")
            np.testing.assert_array_equal(dist[order], dist)         # (j, i) mirrors (i, j) bitwise
            bs, bd, bdist, bom = bf.get_frame(j, int(bcounts[j]))
            # the buffer guarantees that every pair within `cutoff` is on the Verlet list
            keep, bkeep = dist <= w.cutoff, bdist <= w.cutoff
            assert keep.sum() > 100
            np.testing.assert_array_equal(s[keep], bs[bkeep])
            np.testing.assert_array_equal(d[keep], bd[bkeep])
            np.testing.assert_array_equal(dist[keep], bdist[bkeep])  # refresh == rebuild, bit for bit
            np.testing.assert_array_equal(om[keep], bom[bkeep])
    assert 0.01 * total < n_rebuilds < 0.1 * total
    assert n_pairs > 6000 * total


def test_full_size_c2_bruteforce_host_blocks(orc):
    """BASELINE.json config 2 at its full size in brute-force mode, the way the bench's end-to-end
    number runs: float32 host blocks of 16 384 frames through cmd_topo_build (chunked upload, chunk
    kernels on two streams, skin list per persistent CTA).  Size-independent properties on every
    frame (even counts, finite positive rate sums, counts equal to a second topology that takes the
    direct filter path), order / symmetry and the oracle's lists on sampled frames -- among them the
    frames right at the chunk boundaries."""
    from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE, build_with_retry
    import cmdlmc_b200 as cm
    import os
    w = synth.workload("C2")
    total, block = 98304, 16384
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    rate = cm.Fermi(*w.rate_params)
    topo = direct = None
    rng = np.random.RandomState(1)
    for b0 in range(0, total, block):
        fr32 = synth.trajectory(w, block, start=b0, dtype=np.float32)
        if topo is None:
            topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                               MODE_BRUTEFORCE, rate, cap), fr32)
            os.environ["CMDLMC_B200_DENSE_SKIN"] = "0"
            try:
                direct = DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate,
                                        topo.stride)
            finally:
                del os.environ["CMDLMC_B200_DENSE_SKIN"]
        else:
            topo.build(fr32)
        direct.build(fr32)
        counts, _, rsum = topo.frame_info()
        dcounts, _, drsum = direct.frame_info()
        assert (counts > 0).all() and (counts % 2 == 0).all()
        np.testing.assert_array_equal(counts, dcounts)
        np.testing.assert_array_equal(rsum, drsum)
        edges = [0, 2047, 2048, 2049, 4095, 4096, block - 1]
        for f in sorted(set(edges) | set(rng.choice(block, 3, replace=False).tolist())):
            s, d, dist, om = topo.get_frame(int(f), int(counts[f]))
            key = s.astype(np.int64) * w.n_oxygen + d
            assert (np.diff(key) > 0).all()
            mirror = np.argsort(d.astype(np.int64) * w.n_oxygen + s, kind="stable")
            np.testing.assert_array_equal(dist[mirror], dist)
            orow, ocol, odist = orc.topology_bruteforce(obox, fr32[f].astype(np.float64), w.cutoff, w.buffer)
            np.testing.assert_array_equal(s, orow)
            np.testing.assert_array_equal(d, ocol)
            np.testing.assert_array_equal(dist, odist)
    fr, reb, _ = topo.skin_stats()
    assert fr > 0.9 * total and reb < fr // 4        # the skin list carried most frames
    assert direct.skin_stats()[0] == 0


def test_skin_list_on_unrelated_frames(orc, monkeypatch):
    """Brute-force callers may hand over frames that have nothing to do with each other: the
    displacement test then fails on (nearly) every frame and the list is rebuilt -- same lists as
    the direct path, and the oracle's on sampled frames.  Uniformly random positions, so pairs
    come arbitrarily close and the counts fluctuate."""
    import torch
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import DeviceTopology
    w = synth.workload("C2")
    nfr = 2048
    rng = np.random.RandomState(8)
    frames = rng.uniform(0, 1, size=(nfr, w.n_oxygen, 3)) @ w.cell_matrix
    frames[100] = frames[99]                       # and one exact repeat in between
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    rate = cm.Fermi(*w.rate_params)
    d = torch.from_numpy(frames).cuda()
    res = []
    for skin in ("0", None):
        if skin is None:
            monkeypatch.delenv("CMDLMC_B200_DENSE_SKIN", raising=False)
        else:
            monkeypatch.setenv("CMDLMC_B200_DENSE_SKIN", skin)
        t = DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate, 16384, path=0)
        t.build_dev(d.data_ptr(), nfr)
        counts, _, rsum = t.frame_info()
        assert (counts > 0).all()
        blk = t.get_block(0, nfr, counts, omega=True)
        pad = np.arange(blk[0].shape[1])[None, :] >= counts[:, None]
        for a in blk:
            a[pad] = 0
        res.append((t, counts, rsum) + blk)
    fr, reb, _ = res[1][0].skin_stats()
    assert fr >= nfr - 16 and reb >= fr - 8         # rebuilt on (nearly) every frame
    for a, b in zip(res[0][1:], res[1][1:]):
        np.testing.assert_array_equal(a, b)
    _, counts, _, start, dest, dist, _ = res[1]
    for f in (0, 99, 100, 101, 1500, nfr - 1):
        orow, ocol, odist = orc.topology_bruteforce(obox, frames[f], w.cutoff, w.buffer)
        c = int(counts[f])
        np.testing.assert_array_equal(start[f, :c], orow)
        np.testing.assert_array_equal(dest[f, :c], ocol)
        np.testing.assert_array_equal(dist[f, :c], odist)


def test_streaming_blocks_on_two_topologies_equal_blocking_builds():
    """cmd_topo_build_async on two topologies used alternately (the upload of block k+1 overlaps the
    kernels of block k; results are read one block late) against blocking builds of the same
    blocks: counts, rate sums and sampled lists are the same bits.  Blocks come from page-locked
    and from plain memory."""
    from cmdlmc_b200 import runtime
    from cmdlmc_b200.topology import DeviceTopology, MODE_BRUTEFORCE, build_with_retry
    import cmdlmc_b200 as cm
    w = synth.workload("C2")
    block, nblocks = 4096, 6
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    blocks = [np.ascontiguousarray(synth.trajectory(w, block, start=k * block, dtype=np.float32))
              for k in range(nblocks)]
    ref = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                      MODE_BRUTEFORCE, rate, cap), blocks[0])
    want = []
    for b in blocks:
        ref.build(b)
        c, _, r = ref.frame_info()
        want.append((c, r, ref.get_frame(block - 1, int(c[-1])), ref.get_frame(7, int(c[7]))))
    two = [DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, MODE_BRUTEFORCE, rate, ref.stride)
           for _ in range(2)]
    pinned = [runtime.pinned_empty(blocks[0].shape, np.float32) for _ in range(2)]
    for t in two:
        t.build(blocks[0])                       # the first block of a topology is a blocking one
        with pytest.raises(ValueError):
            t.build_async(blocks[0][:, ::2])     # not contiguous / wrong shape
    got = [None] * nblocks

    def collect(k):
        t = two[k % 2]
        t.wait()
        c, _, r = t.frame_info()
        got[k] = (c, r, t.get_frame(block - 1, int(c[-1])), t.get_frame(7, int(c[7])))
    for k in range(nblocks):
        src = blocks[k]
        if k % 3 != 2:                           # two of three blocks through page-locked memory
            pinned[k % 2][...] = blocks[k]
            src = pinned[k % 2]
        two[k % 2].build_async(src)
        if k >= 1:
            collect(k - 1)                       # the previous block, while this one is in flight
    collect(nblocks - 1)
    for k in range(nblocks):
        np.testing.assert_array_equal(got[k][0], want[k][0])
        np.testing.assert_array_equal(got[k][1], want[k][1])
        for fa, fb in zip(got[k][2:], want[k][2:]):
            for a, b in zip(fa, fb):
                np.testing.assert_array_equal(a, b)


def test_randomised_sweep_vs_oracle(orc):
    """Seeded sweep over cell shapes (incl. strongly skewed cells that keep periodic images in the
    filter), atom counts (odd / even / tiny / beyond one warp), radii (up to half the smallest
    height) and both search paths: every list must equal the oracle's, bit for bit; then Verlet
    runs with random block sizes against the oracle's generator."""
    from cmdlmc_b200.topology import DeviceTopology, NeighborTopology, build_with_retry
    import cmdlmc_b200 as cm
    rng = np.random.RandomState(2024)
    n_images_seen = set()
    for case in range(40):
        if case % 3 == 0:
            cell = rng.uniform(7.0, 16.0, size=3)
            hm = np.diag(cell)
        else:
            hm = np.diag(rng.uniform(8.0, 15.0, size=3))
            hm[1, 0], hm[2, 0], hm[2, 1] = rng.uniform(-0.45, 0.45, size=3) * np.array([hm[0, 0], hm[0, 0], hm[1, 1]])
            if case % 3 == 2:
                hm[0, 1], hm[0, 2], hm[1, 2] = rng.uniform(-1.5, 1.5, size=3)
            cell = hm.ravel()
        box, obox = make_box(cell), orc.OracleBox(cell)
        heights = 1.0 / np.linalg.norm(np.linalg.inv(hm.T), axis=1)
        rc = float(rng.uniform(0.15, 0.5) * heights.min())
        if case % 4 == 1:       # radius beyond half the smallest height: periodic images matter
            rc = float(rng.uniform(0.55, 0.95) * heights.min())
        n = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 64, 97, 200, 333]))
        if case % 4 == 1:
            n = min(n, 97)      # nearly every pair is listed at such radii: keep the lists small
        p = rng.uniform(-0.6, 1.6, size=(n, 3)) @ hm
        if n > 3 and case % 5 == 0:
            p[1] = p[0]                                   # coincident pair: dropped (sparse zero)
        cutoff, buffer = 0.7 * rc, 0.3 * rc
        want = orc.topology_bruteforce(obox, p, cutoff, buffer)
        for path in (0, 1):
            t = build_with_retry(lambda cap: DeviceTopology(box, n, cutoff, buffer, 0, None, cap,
                                                            path=path), p[None])
            n_images_seen.add(t.n_images)
            c = t.frame_info()[0]
            got = t.get_frame(0, int(c[0]))
            np.testing.assert_array_equal(got[0], want[0], err_msg="case %d path %d" % (case, path))
            np.testing.assert_array_equal(got[1], want[1])
            np.testing.assert_array_equal(got[2], want[2])
    assert max(n_images_seen) > 0          # some cells needed periodic images in the filter
    # Verlet with random chunking on a random walk in a skewed cell
    hm = np.array([[11.0, 0, 0], [3.0, 10.0, 0], [-2.0, 4.0, 12.0]])
    box, obox = make_box(hm.ravel()), orc.OracleBox(hm.ravel())
    n, nfr = 150, 90
    pos0 = rng.uniform(0, 1, size=(n, 3)) @ hm
    frames = pos0[None] + np.cumsum(rng.normal(scale=0.06, size=(nfr, n, 3)), axis=0)
    top = NeighborTopology(make_traj(frames), box, donor_atoms="O", cutoff=2.5, buffer=0.8)
    top.chunk_size = 13
    gen = orc.verlet_generator(obox, frames, 2.5, 0.8)
    n_reb = 0
    for (row, col, dist, _), want in zip(top.topology_verlet_list_generator(), gen):
        np.testing.assert_array_equal(row, want[0])
        np.testing.assert_array_equal(col, want[1])
        np.testing.assert_array_equal(dist, want[2])
        n_reb += bool(want[3])
    assert 2 < n_reb < nfr


def test_large_system_verlet_schedule_and_refresh(orc):
    """8192 oxygens: the rebuild schedule comes from the cluster kernel (eight CTAs sharing the
    displacement vector) and the refresh is split over several CTAs per frame.  The rebuild
    frames have to be the sequential walk's (oracle orc_verlet_step), the lists the brute-force
    lists of the same frames inside the cutoff."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    n, nfr = 8192, 48
    L = (n / 0.0334) ** (1.0 / 3.0)
    w = synth.Workload("big", np.array([L, L, L]), n, 0, 1, nfr, 0.5, 3.0, 2.0, "Fermi",
                       (0.06, 2.3, 0.1), 11, group_size=0)
    frames = synth.trajectory(w, nfr, amplitude=0.6, noise=0.05)
    box, obox = make_box(w.cell), orc.OracleBox(w.cell)
    rate = cm.Fermi(*w.rate_params)
    lib, P = orc.lib(), orc._p
    disp = np.zeros(n)
    want_rebuilt = []
    for f in range(nfr):
        cur = np.ascontiguousarray(frames[f])
        reb = lib.orc_verlet_step(obox.handle, P(np.ascontiguousarray(frames[f - 1])) if f else None,
                                  P(cur), n, float(w.buffer), P(disp))
        want_rebuilt.append(bool(reb) or f == 0)
    assert 2 < sum(want_rebuilt) < nfr // 2
    tv = build_with_retry(lambda cap: DeviceTopology(box, n, w.cutoff, w.buffer, 1, rate, cap), frames)
    tb = build_with_retry(lambda cap: DeviceTopology(box, n, w.cutoff, w.buffer, 0, rate, cap), frames)
    cv, rebuilt, rsum_v = tv.frame_info()
    cb, _, rsum_b = tb.frame_info()
    np.testing.assert_array_equal(rebuilt.astype(bool), np.array(want_rebuilt))
    refreshed = [f for f in range(nfr) if not want_rebuilt[f]]
    for f in (0, refreshed[0], refreshed[-1], int(np.nonzero(want_rebuilt)[0][-1])):
        sv, dv, xv, ov = tv.get_frame(f, int(cv[f]))
        sb, db, xb, ob = tb.get_frame(f, int(cb[f]))
        if want_rebuilt[f]:
            np.testing.assert_array_equal(sv, sb)
            np.testing.assert_array_equal(dv, db)
            np.testing.assert_array_equal(xv, xb)
            assert rsum_v[f] == pytest.approx(rsum_b[f], rel=1e-12)
        else:
            kv, kb = xv <= w.cutoff, xb <= w.cutoff
            np.testing.assert_array_equal(sv[kv], sb[kb])
            np.testing.assert_array_equal(dv[kv], db[kb])
            np.testing.assert_array_equal(xv[kv], xb[kb])
            np.testing.assert_allclose(ov[kv], ob[kb], rtol=1e-12)
            assert rsum_v[f] == pytest.approx(ov.sum(), rel=1e-10)
