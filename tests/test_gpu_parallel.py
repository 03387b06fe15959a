"""GPU side of the sharding rows and the histogram kernels (K5): ranks emulated
one after the other on one GPU must reproduce the sequential run exactly."""
import numpy as np
import pytest

from cmdlmc_b200 import synth

pytestmark = pytest.mark.gpu


def make_box(cell):
    import cmdlmc_b200 as cm
    cell = np.asarray(cell, dtype=float)
    return cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)


@pytest.mark.parametrize("cfg,nfr,world", [("C1", 240, 2), ("C2", 150, 4)])
def test_frame_block_sharding_is_exact(cfg, nfr, world):
    """Verlet mode across frame blocks: skip pass + own block == sequential run, frame by frame,
    bit for bit (index arrays, distances, rates, rebuilt flags except at the block head)."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.parallel import ShardedTopology
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET, build_with_retry
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    seq = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                      MODE_VERLET, rate, cap), frames)
    counts, rebuilt, rsum = seq.frame_info()
    want = [seq.get_frame(f, int(counts[f])) for f in range(nfr)]
    seen = 0
    for rank in range(world):
        sh = ShardedTopology(box, w.n_oxygen, w.cutoff, w.buffer, MODE_VERLET, rate,
                             lambda a, b: frames[a:b], nfr, rank=rank, world=world, chunk=29)
        for first, topo in sh.blocks():
            c, r, s = topo.frame_info()
            for k in range(len(c)):
                f = first + k
                got = topo.get_frame(k, int(c[k]))
                for a, b in zip(got, want[f]):
                    np.testing.assert_array_equal(a, b)
                assert s[k] == pytest.approx(rsum[f], rel=1e-12)
                if not (rank > 0 and f == sh.start):
                    assert bool(r[k]) == bool(rebuilt[f])
                seen += 1
    assert seen == nfr


def test_pair_distance_histogram(orc):
    """K5: histogram of all listed distances of a block == numpy on the same lists; float32
    (HDF5-style) blocks and the cell-list path give the same counts."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    w = synth.workload("C2")
    frames = synth.trajectory(w, 40)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate,
                                                       cap), frames)
    counts = topo.frame_info()[0]
    alld = np.concatenate([topo.get_frame(f, int(counts[f]))[2] for f in range(40)])
    nb = 500
    want = np.floor((alld - 0.0) * (nb / 5.0)).astype(np.int64)
    want = np.bincount(want[(want >= 0) & (want < nb)], minlength=nb)
    got = topo.distance_histogram(0.0, 5.0, nb)
    np.testing.assert_array_equal(got, want)
    assert got.sum() == counts.sum()
    got2 = topo.distance_histogram(0.0, 5.0, nb, out=got.copy())      # accumulates
    np.testing.assert_array_equal(got2, 2 * want)
    cell = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, 0, rate,
                                                       cap, path=1), frames)
    np.testing.assert_array_equal(cell.distance_histogram(0.0, 5.0, nb), want)


def test_jump_histogram_and_event_distances():
    """jumpstat numerator: distances of the jump pairs, logged per event, histogrammed on the
    device over all replicas."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import DeviceKMC, RNG_PHILOX
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET, build_with_retry
    w = synth.workload("C1")
    frames = synth.trajectory(w, 200)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                       MODE_VERLET, rate, cap), frames)
    counts = topo.frame_info()[0]
    R = 6
    lat = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 50 + r)[0] for r in range(R)])
    kmc = DeviceKMC(box, lat, w.time_step, RNG_PHILOX, seed=3)
    kmc.set_event_log(4000)
    kmc.advance(topo)
    nb = 100
    want = np.zeros(nb, np.int64)
    total = 0
    for r in range(R):
        ev = kmc.events(r)
        assert len(ev["dist"]) == len(ev["frame"]) > 0
        # the logged distance is the listed distance of (start, dest) in the event's frame
        for e in range(0, len(ev["frame"]), 37):
            f = int(ev["frame"][e])
            s, d, dist, _ = topo.get_frame(f, int(counts[f]))
            k = np.flatnonzero((s == ev["start"][e]) & (d == ev["dest"][e]))
            assert len(k) == 1 and dist[k[0]] == ev["dist"][e]
        b = np.floor(ev["dist"] * (nb / 5.0)).astype(np.int64)
        want += np.bincount(b[(b >= 0) & (b < nb)], minlength=nb)
        total += len(ev["dist"])
    got = kmc.jump_histogram(0.0, 5.0, nb)
    np.testing.assert_array_equal(got, want)
    assert got.sum() == total


def test_replica_sharded_ensemble_is_world_size_invariant():
    """C4-style ensemble (Philox KMC replicas on one lattice): sharding the replicas over 2 or 3
    ranks (emulated one after the other) gives the same per-replica trajectories as one rank --
    the Philox counter carries the global replica id -- and the same reduced statistics."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200 import parallel
    from cmdlmc_b200.ensemble import run_kmc_ensemble
    w = synth.workload("C4")
    nfr, R = 160, 12
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    kw = dict(n_sites=w.n_oxygen, n_protons=w.n_protons, cutoff=w.cutoff, buffer=w.buffer,
              jumprate=rate, time_step=w.time_step, n_replicas=R, seed=21, reset_frequency=80,
              print_frequency=20, chunk=64, histogram=(0.0, 5.0, 50))
    one = run_kmc_ensemble(box, lambda a, b: frames[a:b], nfr, rank=0, world=1, **kw)
    assert one["n_replicas"] == R and one["events"] > R
    assert one["jump_hist"].sum() == one["events"]
    assert one["pair_hist"].sum() > 0 and one["observables"]["n"] == R
    assert one["replica_frames"] == R * nfr
    assert one["occupancy_counts"].sum() == w.n_protons * R * nfr        # protons are conserved
    for world in (2, 3):
        parts = [run_kmc_ensemble(box, lambda a, b: frames[a:b], nfr, rank=q, world=world,
                                  reduce=False, **kw) for q in range(world)]
        lat = np.zeros_like(one["local"]["lattices"])
        nev = np.zeros(R, np.int64)
        rows = [None] * R
        for p in parts:
            ids = p["local"]["replica_ids"]
            lat[ids] = p["local"]["lattices"]
            nev[ids] = p["local"]["n_events"]
            for k, r in enumerate(ids):
                rows[r] = p["local"]["observables"][k]
        np.testing.assert_array_equal(lat, one["local"]["lattices"])
        np.testing.assert_array_equal(nev, one["local"]["n_events"])
        merged = parallel.merge_observables(rows)
        np.testing.assert_allclose(merged["mean"], one["observables"]["mean"], rtol=1e-13)
        assert sum(p["jump_hist"].sum() for p in parts) == one["events"]
        np.testing.assert_array_equal(sum(p["jump_hist"] for p in parts), one["jump_hist"])
        np.testing.assert_array_equal(sum(p["occupancy_counts"] for p in parts), one["occupancy_counts"])


def test_jumpstat_probability_follows_the_rate():
    """jumpstat: jumps per listed pair-frame per distance bin.  A pair can only carry a jump when its
    start site is occupied and its destination empty; with protons spread uniformly that factor is
    the same in every bin, so p(d) / Fermi(d) must be flat over the bins with enough jumps."""
    import cmdlmc_b200 as cm
    from cmdlmc_b200.jumpstat import jump_statistics
    w = synth.workload("C4")
    nfr = 400
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    rate = cm.Fermi(*w.rate_params)
    st = jump_statistics(box, lambda a, b: frames[a:b], nfr, n_sites=w.n_oxygen, n_protons=w.n_protons,
                         cutoff=w.cutoff, buffer=w.buffer, jumprate=rate, time_step=w.time_step,
                         n_replicas=96, seed=2, nbins=50, chunk=200)
    assert st["jumps"].sum() == st["events"] > 20000
    ok = st["jumps"] > 400
    assert ok.sum() >= 4
    ratio = st["probability"][ok] / rate(st["centers"][ok])
    # statistical error per bin < 5 %, bin-centre approximation of the steep Fermi function ~10 %
    assert ratio.max() / ratio.min() < 1.5, ratio
    assert (st["probability"][st["centers"] > 3.5] < 1e-6).all()
