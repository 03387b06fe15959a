"""Legacy LMC sweep (row A14, PARITY UNPINNED upstream): the CUDA kernel against the oracle's
restatement of the same specification text -- occupancy trajectories bit-exact in replay mode,
jump statistics within error bars in Philox mode."""
import numpy as np
import pytest

from cmdlmc_b200 import synth

pytestmark = pytest.mark.gpu


def make_box(cell):
    import cmdlmc_b200 as cm
    cell = np.asarray(cell, dtype=float)
    return cm.AtomBoxCubic(cell) if cell.size == 3 else cm.AtomBoxMonoclinic(cell)


def build(cfg, nfr, mode=1):
    import cmdlmc_b200 as cm
    from cmdlmc_b200.topology import DeviceTopology, build_with_retry
    w = synth.workload(cfg)
    frames = synth.trajectory(w, nfr)
    box = make_box(w.cell)
    # a larger amplitude than the MD rate so that hops (and hop-hop dependencies inside one
    # 32-attempt chunk) are frequent enough to exercise the commit / re-evaluate logic
    rate = cm.Fermi(8.0, w.rate_params[1] + 0.4, w.rate_params[2] * 3)
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer, mode,
                                                       rate, cap), frames)
    return w, topo


@pytest.mark.parametrize("cfg,nfr,nrep,spf", [("C1", 40, 5, 1), ("C4", 12, 3, 2)])
def test_replay_occupancy_bit_exact(orc, cfg, nfr, nrep, spf):
    from cmdlmc_b200.lmc import DeviceLMC, RNG_REPLAY, gsl_streams
    w, topo = build(cfg, nfr)
    counts = topo.frame_info()[0]
    lists = [topo.get_frame(f, int(counts[f])) for f in range(nfr)]
    lat0 = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 900 + r)[0] for r in range(nrep)])
    streams = [gsl_streams(1000 + r, counts, spf) for r in range(nrep)]
    pick = np.stack([s[0] for s in streams])
    acc = np.stack([s[1] for s in streams])
    dev = DeviceLMC(lat0, RNG_REPLAY)
    dev.set_replay_stream(pick, acc)
    dev.enable_jump_matrix()
    dev.advance(topo, w.time_step, spf)
    st = dev.state()
    jm_want = np.zeros((w.n_oxygen, w.n_oxygen), np.int64)
    total_jumps = 0
    for r in range(nrep):
        lat = lat0[r].copy()
        pos = jumps = 0
        for f in range(nfr):
            s, d, _, om = lists[f]
            prob = om * w.time_step
            for _ in range(spf):
                p = len(s)
                jumps += orc.lmc_sweep(s, d, prob, lat, pick[r, pos:pos + p], acc[r, pos:pos + p],
                                       jm_want)
                pos += p
        np.testing.assert_array_equal(st["lattices"][r], lat)
        assert st["jumps"][r] == jumps and st["attempts"][r] == pos
        assert st["sweeps"][r] == nfr * spf and not st["halted"][r]
        total_jumps += jumps
    assert total_jumps > 20 * nrep            # the test really moves protons
    np.testing.assert_array_equal(dev.jump_matrix(), jm_want)
    # a stream that is too short halts the replica instead of reading past its end
    dev2 = DeviceLMC(lat0[:1], RNG_REPLAY)
    dev2.set_replay_stream(pick[:1, :int(counts[0]) + 3], acc[:1, :int(counts[0]) + 3])
    dev2.advance(topo, w.time_step, spf)
    st2 = dev2.state()
    assert st2["halted"][0] and st2["sweeps"][0] == 1


def test_philox_statistics(orc):
    from cmdlmc_b200.lmc import DeviceLMC, RNG_PHILOX
    nfr, R = 30, 256
    w, topo = build("C1", nfr)
    counts = topo.frame_info()[0]
    lists = [topo.get_frame(f, int(counts[f])) for f in range(nfr)]
    lat0 = np.stack([synth.initial_lattice(w.n_oxygen, w.n_protons, 40 + r)[0] for r in range(R)])
    dev = DeviceLMC(lat0, RNG_PHILOX, seed=12)
    dev.advance(topo, w.time_step, 1)
    st = dev.state()
    assert (st["attempts"] == counts.sum()).all()
    assert ((st["lattices"] > 0).sum(axis=1) == w.n_protons).all()      # protons are conserved
    for r in range(R):                                                   # labels only move
        assert sorted(st["lattices"][r][st["lattices"][r] > 0]) == list(range(1, w.n_protons + 1))
    # same seed -> same trajectories; other seed -> different ones
    dev_b = DeviceLMC(lat0, RNG_PHILOX, seed=12)
    dev_b.advance(topo, w.time_step, 1)
    np.testing.assert_array_equal(dev_b.state()["lattices"], st["lattices"])
    dev_c = DeviceLMC(lat0, RNG_PHILOX, seed=13)
    dev_c.advance(topo, w.time_step, 1)
    assert (dev_c.state()["lattices"] != st["lattices"]).any()
    # jump counts vs the oracle driven by NumPy streams: means agree within 4 standard errors
    rng = np.random.RandomState(3)
    ref = []
    for r in range(48):
        lat = lat0[r].copy()
        j = 0
        for f in range(nfr):
            s, d, _, om = lists[f]
            p = len(s)
            j += orc.lmc_sweep(s, d, om * w.time_step, lat, rng.randint(0, p, size=p),
                               rng.random_sample(p))
        ref.append(j)
    g, c = st["jumps"].astype(float), np.array(ref, float)
    se = np.sqrt(g.var(ddof=1) / len(g) + c.var(ddof=1) / len(c))
    assert abs(g.mean() - c.mean()) < 4 * se, (g.mean(), c.mean(), se)
