"""Replica ensembles across the GPUs of one box (BASELINE.json config 4: many independent KMC
replicas on one lattice, replica-sharded, statistics reduced over NCCL).

Every rank builds the (cheap) Verlet topology of all frames itself and runs the replicas it owns
(`parallel.replica_ids`: r mod G) through the streaming KMC kernel; the Philox counter carries the
GLOBAL replica id, so the ensemble is the same set of trajectories for any number of GPUs.  Only the
statistics cross the NVLink fabric: one all-reduce of a few KB at the end."""
import numpy as np

from . import parallel
from .kmc import DeviceKMC, RNG_PHILOX
from .topology import DeviceTopology, MODE_VERLET, build_with_retry


def initial_lattices(n_sites, n_protons, replica_ids, seed):
    """Lattices like KMCLattice._initialize_lattice (MDMC.py:68-72), one legacy RandomState per
    GLOBAL replica id so that the start configurations do not depend on the sharding either."""
    out = np.zeros((len(replica_ids), n_sites), dtype=np.int32)
    for k, r in enumerate(replica_ids):
        rng = np.random.RandomState((int(seed) * 1000003 + int(r)) % (2 ** 31 - 1))
        lat = np.zeros(n_sites, dtype=np.int32)
        lat[:n_protons] = np.arange(1, n_protons + 1)
        rng.shuffle(lat)
        out[k] = lat
    return out


def run_kmc_ensemble(atom_box, frames_source, n_frames, *, n_sites, n_protons, cutoff, buffer,
                     jumprate, time_step, n_replicas, seed=0, reset_frequency=0, print_frequency=0,
                     chunk=1024, histogram=None, rank=None, world=None, reduce=True):
    """frames_source(start, stop) -> donor positions [stop - start, n_sites, 3].
    histogram = (lo, hi, nbins) adds the jump-distance and pair-distance histograms (jumpstat).
    Returns a dict: per-replica arrays of this rank (`local`) and the ensemble statistics summed
    over all ranks (`n_replicas`, `events`, `site_updates`, `observables` mean / sem per printed
    row, `jump_hist`, `pair_hist`)."""
    r0, w0 = parallel.rank_world()
    rank = r0 if rank is None else rank
    world = w0 if world is None else world
    ids = parallel.replica_ids(n_replicas, rank, world)
    observe = print_frequency > 0
    first = frames_source(0, min(chunk, n_frames))
    topo = build_with_retry(lambda cap: DeviceTopology(atom_box, n_sites, cutoff, buffer, MODE_VERLET,
                                                       jumprate, cap), first)
    kmc = None
    if len(ids):
        kmc = DeviceKMC(atom_box, initial_lattices(n_sites, n_protons, ids, seed), time_step,
                        RNG_PHILOX, seed)
        kmc.set_replica_ids(rank, world)
        if observe:
            kmc.set_observables(reset_frequency, print_frequency)
        kmc.enable_occupancy()
        if histogram:
            kmc.set_event_log(64 * n_frames + 64)
    pair_hist = np.zeros(histogram[2], np.int64) if histogram else None
    pos = 0
    while pos < n_frames:
        hi = min(n_frames, pos + chunk)
        if pos:
            topo.build(frames_source(pos, hi))
        if kmc is not None:
            kmc.advance(topo, topo.positions_ptr() if observe else None)
        if histogram and rank == 0:
            topo.distance_histogram(histogram[0], histogram[1], histogram[2], out=pair_hist)
        pos = hi
    local = {"replica_ids": ids}
    stats = {"events": np.zeros(1, np.int64), "site_updates": np.zeros(1, np.int64),
             "replicas": np.array([len(ids)], np.int64), "occupancy": np.zeros(n_sites, np.int64),
             "replica_frames": np.zeros(1, np.int64)}
    rows = []
    if kmc is not None:
        st = kmc.state()
        local.update(lattices=st["lattices"], n_events=st["n_events"], time=st["time"])
        stats["events"][0] = st["n_events"].sum()
        stats["site_updates"][0] = st["site_updates"].sum()
        occ, rf = kmc.occupancy()
        stats["occupancy"][:] = occ
        stats["replica_frames"][0] = rf
        if observe:
            rows = [kmc.observables(k) for k in range(len(ids))]
            local["observables"] = rows
    if histogram:
        jh = np.zeros(histogram[2], np.int64)
        if kmc is not None:
            dropped = kmc.events_dropped()
            if dropped:
                raise RuntimeError("%d jumps did not fit the event log (64 per frame and replica): the "
                                   "jump-distance histogram would be incomplete" % dropped)
            kmc.jump_histogram(histogram[0], histogram[1], histogram[2], out=jh)
        stats["jump_hist"] = jh
        stats["pair_hist"] = pair_hist if rank == 0 else np.zeros_like(pair_hist)
    tot = parallel.allreduce_sum(stats) if reduce else stats
    out = {"local": local, "n_replicas": int(tot["replicas"][0]), "events": int(tot["events"][0]),
           "site_updates": int(tot["site_updates"][0]),
           # fraction of (replica, frame) pairs in which a site carried a proton
           "occupancy": tot["occupancy"] / max(int(tot["replica_frames"][0]), 1),
           "occupancy_counts": tot["occupancy"], "replica_frames": int(tot["replica_frames"][0])}
    if histogram:
        out["jump_hist"], out["pair_hist"] = tot["jump_hist"], tot["pair_hist"]
    if observe:
        out["observables"] = parallel.merge_observables(rows) if (reduce or rows) else None
    return out
