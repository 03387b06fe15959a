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


_comm = {"up": False, "rank": 0, "world": 1}


def comm_init(rank=None, world=None, unique_id=None):
    """Builds the LIBRARY's communicator (cmd_comm_init: NCCL bound inside libcmdlmc_b200, over
    NVLink / NVSwitch) for this process's GPU.  Rank 0 creates the 128-byte unique id; it travels
    to the other ranks over torch.distributed when a process group exists (any backend), or is
    passed in by a host that has its own transport.  Returns a small dict for logging."""
    import ctypes as C
    from . import _abi, runtime
    if _comm["up"]:
        return dict(_comm)
    r0, w0 = rank_world()
    rank = r0 if rank is None else int(rank)
    world = w0 if world is None else int(world)
    runtime.ensure_init()
    lib = _abi.lib()
    ident = np.zeros(128, np.uint8)
    if world > 1:
        if unique_id is not None:
            ident[:] = np.frombuffer(bytes(unique_id), np.uint8)[:128]
        else:
            import torch
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("comm_init needs a unique_id or an initialised torch.distributed group")
            if rank == 0:
                _abi.check(lib.cmd_comm_unique_id(ident.ctypes.data_as(_abi.u8p)))
            t = torch.from_numpy(ident)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, 0)
            ident[:] = t.cpu().numpy()
    _abi.check(lib.cmd_comm_init(rank, world, ident.ctypes.data_as(_abi.u8p)))
    _comm.update(up=True, rank=rank, world=world,
                 nccl_version=int(lib.cmd_comm_nccl_version()) if world > 1 else None)
    return dict(_comm)


def comm_destroy():
    from . import _abi
    if _comm["up"]:
        _abi.lib().cmd_comm_destroy()
    _comm.update(up=False, rank=0, world=1)


def allreduce_sum(stats, device=None):
    """Sums a dict of NumPy arrays (float64 or integer) over all ranks with ONE all-reduce per
    dtype class; returns new arrays of the same shapes.  With the library communicator up
    (comm_init) the reduction is cmd_stats_allreduce -- NCCL inside the C ABI, no torch involved;
    otherwise torch.distributed (gloo in the CPU tests).  Single process: copies of the input."""
    out = {k: np.array(v, copy=True) for k, v in stats.items()}
    if _comm["up"] and _comm["world"] > 1:
        import ctypes as C
        from . import _abi
        fk = [k for k in sorted(out) if np.issubdtype(out[k].dtype, np.floating)]
        ik = [k for k in sorted(out) if not np.issubdtype(out[k].dtype, np.floating)]
        f = np.concatenate([out[k].astype(np.float64).ravel() for k in fk]) if fk else np.zeros(0)
        i = np.concatenate([out[k].astype(np.int64).ravel() for k in ik]) if ik else np.zeros(0, np.int64)
        _abi.check(_abi.lib().cmd_stats_allreduce(_abi.ptr(f) if f.size else None, f.size,
                                                  _abi.ptr(i, C.c_int64) if i.size else None, i.size))
        for keys, flat in ((fk, f), (ik, i)):
            pos = 0
            for k in keys:
                m = out[k].size
                out[k] = flat[pos:pos + m].reshape(out[k].shape).astype(out[k].dtype)
                pos += m
        return out
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
        self.n_frames = int(n_frames)
        self.topo_n_atoms = int(n_atoms)
        self.chunk = int(chunk)
        self.mode = mode
        self.frames_source = frames_source
        self._make = lambda cap: DeviceTopology(atom_box, n_atoms, cutoff, buffer, mode, jumprate, cap)
        self._capacity = capacity
        self.topo = None

    def blocks(self):
        """Yields (first_frame, DeviceTopology) for every chunk of this rank's block; the lists of
        the chunk stay in HBM until the next chunk is built.

        Verlet mode over several ranks: with the library communicator up (comm_init) every rank
        uploads ONLY its own block, the ranks all-gather the per-frame step lengths
        (N x 8 bytes per frame over NVLink) and each replays the rebuild schedule of the frames
        before its block from those (`_blocks_gathered`).  Without a communicator (a single
        process playing one rank of several) the frames before the block are walked through the
        displacement pass from their coordinates (cmd_topo_skip)."""
        from .topology import MODE_VERLET, build_with_retry
        if (self.mode == MODE_VERLET and self.world > 1 and _comm["up"]
                and _comm["world"] == self.world and _comm["rank"] == self.rank):
            yield from self._blocks_gathered()
            return
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

    def _blocks_gathered(self):
        import ctypes as C
        import torch
        from . import _abi, runtime
        dev = torch.device("cuda", runtime.ensure_init())
        n = self.topo_n_atoms
        own = self.stop - self.start
        halo = 1 if self.start > 0 else 0
        base, extra = divmod(int(self.n_frames), self.world)
        maxlen = base + (1 if extra else 0)
        # this rank's frames (+ the frame before them), resident in HBM for the whole pass
        lo = self.start - halo
        fr = np.ascontiguousarray(self.frames_source(lo, max(self.stop, lo + 1)), dtype=np.float64)
        if self.topo is None:
            self.topo = self._sized(fr[halo:halo + 1] if own else fr[:1])
        d_all = torch.from_numpy(fr).to(dev)
        d_own = d_all[halo:]
        send = torch.zeros((max(maxlen, 1), n), dtype=torch.float64, device=dev)
        recv = torch.empty((self.world, max(maxlen, 1), n), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        if own:
            self.topo.dr_dev(d_own.data_ptr(), own, d_all.data_ptr() if halo else None, send.data_ptr())
        _abi.check(_abi.lib().cmd_allgather_dev(C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()),
                                                send.numel() * 8))
        runtime.sync()
        if self.start > 0 and own:
            lens = [frame_block(self.n_frames, r, self.world) for r in range(self.rank)]
            before = torch.cat([recv[r, :b - a] for r, (a, b) in enumerate(lens)]).contiguous()
            torch.cuda.synchronize()
            last, pos = -1, 0
            while pos < self.start:
                hi = min(self.start, pos + self.chunk)
                k = self.topo.skip_dr_dev(before[pos:].data_ptr(), hi - pos)
                if k >= 0:
                    last = pos + k
                pos = hi
            reb = torch.from_numpy(np.ascontiguousarray(self.frames_source(last, last + 1)[0],
                                                        dtype=np.float64)).to(dev)
            torch.cuda.synchronize()
            self.topo.seed_dev(reb.data_ptr(), d_all.data_ptr())
            self.last_rebuild_before_block = last
        pos = 0
        while pos < own:
            cn = min(self.chunk, own - pos)
            self.topo.build_dev(d_own[pos:].data_ptr(), cn)
            yield self.start + pos, self.topo
            pos += cn
        runtime.sync()   # d_all must outlive the kernels that read it

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
