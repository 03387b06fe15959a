"""F2: HydroniumTopology + ReLUTransformation / InterpolatedTransformation / DistanceInterpolator
(mdlmc/topo/topology.py:170-353).  Golden traces come from the reference's own classes driven by
its KMCLattice (oracle/make_golden.py gen_hydronium, tests/golden/hydronium.npz): the rates depend
on every proton's residence time, so a bit-identical event trace pins the per-replica rescaling
inside the KMC kernel, the nearest-neighbour kernel and the jump-time bookkeeping at once."""
import copy

import numpy as np
import pytest

from cmdlmc_b200 import synth

pytestmark = pytest.mark.gpu

HYD_RELU = dict(a=0.9, b=2.35, d0=2.5, left_bound=2.2, right_bound=3.2)
HYD_RELAX = 6.0
HYD_PROTONS = 24
HYD_FERMI = (0.3, 2.45, 0.12)


def setup(g, variant, rng="replay", seed=0, nfr=150, chunk=64):
    import cmdlmc_b200 as cm
    from cmdlmc_b200.kmc import KMCLattice
    from cmdlmc_b200.topology import (DistanceInterpolator, HydroniumTopology,
                                      InterpolatedTransformation, ReLUTransformation)
    from cmdlmc_b200.trajectory import ArrayTrajectory
    w = copy.deepcopy(synth.workload("C2"))
    w.n_protons, w.rate_params = HYD_PROTONS, HYD_FERMI
    frames = synth.trajectory(w, nfr)
    box = cm.AtomBoxMonoclinic(w.cell)
    traj = ArrayTrajectory(frames, np.array(["O"] * w.n_oxygen), time_step=w.time_step)
    if variant == "relu":
        tf, ip = ReLUTransformation(**HYD_RELU), DistanceInterpolator(HYD_RELAX)
    else:
        tf, ip = InterpolatedTransformation(g["interp_x"], g["interp_y"]), None
    top = HydroniumTopology(traj, box, donor_atoms="O", cutoff=w.cutoff, buffer=w.buffer,
                            distance_transformation_function=tf, distance_interpolator=ip)
    kmc = KMCLattice(top, atom_box=box, jumprate_function=cm.Fermi(*w.rate_params),
                     lattice_size=w.n_oxygen, proton_number=w.n_protons, donor_atoms="O",
                     time_step=w.time_step, rng=rng, seed=seed, chunk_size=chunk)
    return w, frames, box, top, kmc


@pytest.mark.parametrize("variant", ["relu", "interp"])
def test_event_trace_vs_reference_golden(golden, variant):
    g = golden("hydronium")
    np.random.seed(31)
    w, frames, box, top, kmc = setup(g, variant)
    np.testing.assert_array_equal(kmc.lattice, g[variant + "_lattice0"])
    for _ in kmc:
        pass
    ev = kmc.event_log
    want_n = len(g[variant + "_ev_time"])
    n = min(want_n, len(ev["time"]))
    assert n >= want_n - 1 and n > 300
    np.testing.assert_array_equal(ev["frame"][:n], g[variant + "_ev_frame"][:n])
    np.testing.assert_array_equal(ev["start"][:n], g[variant + "_ev_start"][:n])
    np.testing.assert_array_equal(ev["dest"][:n], g[variant + "_ev_dest"][:n])
    np.testing.assert_array_equal(ev["proton"][:n], g[variant + "_ev_proton"][:n])
    np.testing.assert_allclose(ev["time"][:n], g[variant + "_ev_time"][:n], rtol=1e-12)
    if len(ev["time"]) == want_n:
        np.testing.assert_array_equal(kmc.lattice, g[variant + "_lattice_final"])


def test_host_colvars_and_nearest_kernel_vs_reference(golden):
    """_determine_colvars (host mirror) and cmd_topo_nearest (device) against the reference's
    colvars for a fixed lattice / jump-time state."""
    from cmdlmc_b200.topology import DeviceTopology, MODE_VERLET, build_with_retry
    g = golden("hydronium")
    w, frames, box, top, kmc = setup(g, "relu", nfr=3)
    lattice = g["colvar_lattice"].copy()
    top.take_lattice_reference(lattice)
    top._time_of_last_jump_vec[:] = g["colvar_tlast"]
    got = list(top)
    assert len(got) == 3
    for k, (s, d, dist) in enumerate(got):
        np.testing.assert_array_equal(s, g["colvar%d_start" % k])
        np.testing.assert_array_equal(d, g["colvar%d_dest" % k])
        np.testing.assert_allclose(dist, g["colvar%d_dist" % k], rtol=1e-12)
    # the device's nearest-neighbour arrays hold the same destinations and the unscaled distances
    import ctypes as C
    from cmdlmc_b200 import _abi
    topo = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, w.cutoff, w.buffer,
                                                       MODE_VERLET, None, cap), frames)
    _abi.check(_abi.lib().cmd_topo_nearest(topo.handle, 4))
    dest = np.zeros(w.n_oxygen * 4, np.int32)
    dist = np.zeros(w.n_oxygen * 4)
    _abi.check(_abi.lib().cmd_topo_get_frame_nearest(topo.handle, 1, _abi.ptr(dest, C.c_int),
                                                     _abi.ptr(dist)))
    np.testing.assert_array_equal(dest, g["colvar1_dest"])
    counts = topo.frame_info()[0]
    s, d, di, _ = topo.get_frame(1, int(counts[1]))
    for site in (0, 7, 399):
        row = np.sort(di[s == site])[:4]
        np.testing.assert_array_equal(dist[4 * site:4 * site + 4], row)
    # a site with fewer than four listed neighbours is an error, like upstream (topology.py:250)
    sparse = build_with_retry(lambda cap: DeviceTopology(box, w.n_oxygen, 1.0, 0.5, MODE_VERLET,
                                                         None, cap), frames)
    assert _abi.lib().cmd_topo_nearest(sparse.handle, 4) != 0


def test_philox_mode_runs_and_tracks_jump_times(golden):
    g = golden("hydronium")
    w, frames, box, top, kmc = setup(g, "relu", rng="philox", seed=9)
    n = sum(1 for _ in kmc)
    ev = kmc.event_log
    assert n > 0 and len(ev["time"]) > 100
    tl = kmc._device.last_jump_times()[0]
    # protons that jumped carry the time of their last jump, the others still -1
    last = {}
    for p, t in zip(ev["proton"], ev["time"]):
        last[int(p)] = t
    for p in range(1, w.n_protons + 1):
        assert tl[p - 1] == last.get(p, -1.0)
    assert (np.sort(kmc.lattice[kmc.lattice > 0]) == np.arange(1, w.n_protons + 1)).all()
