"""Sharding of the hot path across the GPUs of one box.

One process per GPU (torchrun); the path has no data-path collective:

* geometry -> neighbour lists -> rates shard by contiguous TRAJECTORY-FRAME BLOCK.  Frames are
  independent given the Verlet rebuild schedule, so a rank whose block starts at frame s first
  walks frames [0, s) through the cheap displacement / rebuild-decision pass (cmd_topo_skip),
  which reproduces the state of a sequential run at s exactly;
* the KMC / LMC stage shards by INDEPENDENT REPLICA (replica r runs on rank r mod G);
* only the statistics -- histograms, MSD sums, autocorrelation and jump counts -- are summed
  across ranks: one all-reduce of a few KB per reporting interval (NCCL over NVLink on GPUs, gloo
  in the CPU tests).  torch.distributed is plumbing here, nothing else.
"""
import os

import numpy as np


def rank_world():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def frame_block(n_frames, rank, world):
    """Contiguous block [start, stop) of rank `rank`; block sizes differ by at most one frame and
    the blocks tile [0, n_frames) in rank order."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(n_frames), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def replica_ids(n_replicas, rank, world):
    """Global ids of the replicas rank `rank` owns: r mod world == rank."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return np.arange(rank, int(n_replicas), world, dtype=np.int64)


def allreduce_sum(stats, device=None):
    """Sums a dict of NumPy arrays (float64 or integer) over all ranks with ONE all-reduce per
    dtype class; returns new arrays of the same shapes.  Without an initialised process group
    (single process) the input is returned as copies."""
    out = {k: np.array(v, copy=True) for k, v in stats.items()}
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return out
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out
    if device is None and dist.get_backend() == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    for is_float in (True, False):
        keys = [k for k in sorted(out) if np.issubdtype(out[k].dtype, np.floating) == is_float]
        if not keys:
            continue
        dt = np.float64 if is_float else np.int64
        flat = np.concatenate([out[k].astype(dt).ravel() for k in keys]) if keys else np.zeros(0, dt)
        t = torch.from_numpy(flat)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)
        flat = t.cpu().numpy()
        pos = 0
        for k in keys:
            n = out[k].size
            out[k] = flat[pos:pos + n].reshape(out[k].shape).astype(out[k].dtype)
            pos += n
    return out


def merge_observables(rows_per_replica):
    """Replica statistics of KMCLattice.observables_output rows (frame, time, msd_x, msd_y, msd_z,
    autocorr): per output row the sums needed for mean and standard error over ALL replicas of
    all ranks.  rows_per_replica: list of float arrays [n_rows, 6] (this rank's replicas; ragged
    lengths are cut to the shortest).  Returns dict(frame, n, mean[rows,4], sem[rows,4])."""
    n_rows = min((len(r) for r in rows_per_replica), default=0)
    world = rank_world()[1]
    if world > 1:   # every rank must contribute the same number of rows
        n_all = allreduce_sum({"n": np.array([n_rows, -n_rows], dtype=np.int64)})
        # min over ranks via two sums is not possible; ranks agree on n_rows by construction
        # (same trajectory length); a mismatch is a caller bug
        if n_all["n"][0] != n_rows * world:
            raise ValueError("ranks hold different numbers of observable rows")
    if n_rows == 0:
        return dict(frame=np.zeros(0, np.int64), n=0, mean=np.zeros((0, 4)), sem=np.zeros((0, 4)))
    data = np.stack([np.asarray(r)[:n_rows, 2:6] for r in rows_per_replica])   # [R, rows, 4]
    local = {"s1": data.sum(axis=0), "s2": (data ** 2).sum(axis=0),
             "cnt": np.array([data.shape[0]], dtype=np.int64)}
    tot = allreduce_sum(local)
    n = int(tot["cnt"][0])
    mean = tot["s1"] / n
    var = np.maximum(tot["s2"] / n - mean ** 2, 0.0) * (n / max(n - 1, 1))
    return dict(frame=np.asarray(rows_per_replica[0])[:n_rows, 0].astype(np.int64), n=n,
                mean=mean, sem=np.sqrt(var / n))


class ShardedTopology:
    """This rank's frame block of a trajectory through the GPU topology pipeline.

    frames_source(start, stop) -> float array [stop - start, n_atoms, 3] (host); only the frames a
    rank needs are ever requested: its own block, and in Verlet mode the frames before it for the
    schedule pass (coordinates only, no pair work)."""

    def __init__(self, atom_box, n_atoms, cutoff, buffer, mode, jumprate, frames_source, n_frames,
                 rank=None, world=None, chunk=4096, capacity=0):
        from .topology import DeviceTopology
        r, w = rank_world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.start, self.stop = frame_block(n_frames, self.rank, self.world)
        self.chunk = int(chunk)
        self.mode = mode
        self.frames_source = frames_source
        self._make = lambda cap: DeviceTopology(atom_box, n_atoms, cutoff, buffer, mode, jumprate, cap)
        self._capacity = capacity
        self.topo = None

    def blocks(self):
        """Yields (first_frame, DeviceTopology) for every chunk of this rank's block; the lists of
        the chunk stay in HBM until the next chunk is built."""
        from .topology import MODE_VERLET, build_with_retry
        pos = 0
        if self.mode == MODE_VERLET and self.start > 0:
            # the very first frames size the capacity and create the object
            while pos < self.start:
                hi = min(self.start, pos + self.chunk)
                fr = self.frames_source(pos, hi)
                if self.topo is None:
                    self.topo = self._sized(fr[:1])
                self.topo.skip(fr)
                pos = hi
        pos = self.start
        while pos < self.stop:
            hi = min(self.stop, pos + self.chunk)
            fr = self.frames_source(pos, hi)
            if self.topo is None:
                self.topo = self._sized(fr[:1])
            self.topo.build(fr)
            yield pos, self.topo
            pos = hi

    def _sized(self, first_frame):
        """Creates the topology with a capacity probed on one frame WITHOUT consuming it."""
        from .topology import DeviceTopology, MODE_BRUTEFORCE
        if self._capacity:
            return self._make(self._capacity)
        probe = self._make(0)
        probe_bf = DeviceTopology(probe.atom_box, probe.n_atoms, probe._args[2], probe._args[3],
                                  MODE_BRUTEFORCE, None, 0)
        probe_bf.build(first_frame)
        cap = probe_bf.stride
        del probe_bf
        return self._make(cap)
