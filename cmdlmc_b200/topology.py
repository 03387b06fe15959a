"""Neighbour topology -- host mirror of mdlmc/topo/topology.py:18-121 on top of the C ABI.

`NeighborTopology` keeps the reference's constructor, generators and output contract
(`(start int32[P], destination int32[P], distance float64[P])` in LIL->COO order); the work is
done by the CUDA kernels of csrc/pairs.cu on blocks of frames.
"""
import ctypes as C
import logging
from collections import deque

import numpy as np

from . import _abi, runtime
from ._abi import as_f64, check, ptr
from .jumprate import NPAR

logger = logging.getLogger(__name__)

MODE_BRUTEFORCE = 0
MODE_VERLET = 1


class _ArrayFrames:
    """Frame stream of an array-backed trajectory with a visible cursor, so that whole blocks can be
    taken by slicing instead of frame by frame (same one-shot semantics as a generator)."""

    def __init__(self, traj):
        self.traj = traj
        self.pos = 0

    def __iter__(self):
        return self

    def frame(self, k):
        from .trajectory import Frame
        t = self.traj
        return Frame(t.atom_names, np.asarray(t.positions[k % len(t)], dtype=float), time=k * t.time_step)

    def __next__(self):
        if self.pos >= len(self.traj) and not self.traj.repeat:
            raise StopIteration
        k = self.pos
        self.pos += 1
        return self.frame(k)


class _LazyFrames:
    """Frames [k0, k1) of an array trajectory, materialised only when somebody looks at them."""

    def __init__(self, stream, k0, k1):
        self.stream, self.k0, self.k1 = stream, k0, k1

    def __len__(self):
        return self.k1 - self.k0

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self.stream.frame(self.k0 + i)

    def __iter__(self):
        for k in range(self.k0, self.k1):
            yield self.stream.frame(k)


class DeviceTopology:
    """Thin owner of a `cmd_topo` handle: neighbour lists + rates of a block of frames in HBM."""

    def __init__(self, atom_box, n_atoms, cutoff, buffer, mode, jumprate=None, capacity=0,
                 path=-1):
        if jumprate is not None and (getattr(jumprate, "kind", None) is None or not hasattr(jumprate, "_par")):
            # the reference accepts any callable JumpRate(*colvars); the rate is evaluated inside
            # the list kernels here, so it has to be one the device knows
            raise TypeError("jump rate %r cannot be evaluated on the device: use Fermi, FermiAngle, "
                            "ActivationEnergy or Exponential from cmdlmc_b200.jumprate (a JumpRate "
                            "subclass needs `kind` and `params`)" % (jumprate,))
        runtime.ensure_init()
        self.atom_box = atom_box          # keeps the box handle alive
        self.n_atoms = int(n_atoms)
        self.mode = mode
        kind = jumprate.kind if jumprate is not None else 0
        par = jumprate._par() if jumprate is not None else np.array([0.0, 0.0, 1.0] + [0.0] * 5)
        self._args = (atom_box, n_atoms, cutoff, buffer, mode, kind, par)
        self._handle = C.c_void_p()
        check(_abi.lib().cmd_topo_create(atom_box.handle, int(n_atoms), float(cutoff),
                                         float(buffer), int(mode), int(kind), ptr(par),
                                         int(capacity), C.byref(self._handle)))
        if path != -1:   # -1 auto, 0 dense (one CTA per frame), 1 cell list
            check(_abi.lib().cmd_topo_set_path(self._handle, int(path)))

    @property
    def path(self):
        """Search path the library chose (0 dense, 1 cell list); final after the first build."""
        return int(_abi.lib().cmd_topo_path(self._handle))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _abi.lib().cmd_topo_destroy(h)
            except Exception:
                pass
            self._handle = None

    @property
    def handle(self):
        return self._handle

    @property
    def stride(self):
        return int(_abi.lib().cmd_topo_stride(self._handle))

    @property
    def capacity_needed(self):
        """Directed pairs of the largest frame of a block that overflowed the per-frame capacity
        (0 when no build has returned CMD_ECAPACITY)."""
        return int(_abi.lib().cmd_topo_capacity_needed(self._handle))

    def skin_stats(self):
        """(frames, rebuilds, list entries filtered) of the dense kernel's skin list so far."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(_abi.lib().cmd_topo_skin_stats(self._handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    @property
    def n_images(self):
        """Periodic images the pair filter evaluates besides the wrapped vector."""
        return int(_abi.lib().cmd_topo_n_images(self._handle))

    @property
    def nframes(self):
        return int(_abi.lib().cmd_topo_nframes(self._handle))

    def set_selection(self, n_total, index):
        """Donor selection on the device: from now on build() / skip() take whole frames
        [F, n_total, 3] and the donors are rows `index` of each (n_total = 0: off again)."""
        index = np.ascontiguousarray(index, dtype=np.int32) if n_total else None
        if n_total and index.shape != (self.n_atoms,):
            raise ValueError("one selection index per donor atom expected")
        check(_abi.lib().cmd_topo_set_selection(self._handle, int(n_total),
                                                ptr(index, C.c_int) if n_total else None))
        self._n_rows = int(n_total) if n_total else self.n_atoms

    def build(self, frames):
        """frames: host ndarray [F, n, 3] float64 or float32 ([F, n_total, 3] after set_selection)."""
        frames = np.ascontiguousarray(frames)
        if frames.dtype not in (np.float32, np.float64):
            frames = frames.astype(np.float64)
        rows = getattr(self, "_n_rows", self.n_atoms)
        if frames.ndim != 3 or frames.shape[1:] != (rows, 3):
            raise ValueError("frames must have shape [F, %d, 3]" % rows)
        check(_abi.lib().cmd_topo_build(self._handle, frames.ctypes.data_as(C.c_void_p),
                                        frames.dtype.itemsize, frames.shape[0]))

    def build_async(self, frames):
        """Queues the upload and the kernels of a brute-force block and returns at once
        (cmd_topo_build_async); `frames` (page-locked for a real overlap) must stay untouched until
        wait().  Used alternately on two topologies, the upload of one block overlaps the kernels
        of the other."""
        if frames.dtype not in (np.float32, np.float64) or not frames.flags.c_contiguous:
            raise ValueError("build_async takes a C-contiguous float32 / float64 block as it is")
        rows = getattr(self, "_n_rows", self.n_atoms)
        if frames.ndim != 3 or frames.shape[1:] != (rows, 3):
            raise ValueError("frames must have shape [F, %d, 3]" % rows)
        self._async_block = frames           # keeps the host block alive
        check(_abi.lib().cmd_topo_build_async(self._handle, frames.ctypes.data_as(C.c_void_p),
                                              frames.dtype.itemsize, frames.shape[0]))

    def wait(self):
        """Completion of the last build_async block; raises on a capacity overflow."""
        try:
            check(_abi.lib().cmd_topo_wait(self._handle))
        finally:
            self._async_block = None

    def skip(self, frames):
        """Walks frames that precede this rank's block through the Verlet schedule pass only
        (frame-block sharding, cmd_topo_skip); produces no per-frame results."""
        frames = np.ascontiguousarray(frames)
        if frames.dtype not in (np.float32, np.float64):
            frames = frames.astype(np.float64)
        rows = getattr(self, "_n_rows", self.n_atoms)
        if frames.ndim != 3 or frames.shape[1:] != (rows, 3):
            raise ValueError("frames must have shape [F, %d, 3]" % rows)
        check(_abi.lib().cmd_topo_skip(self._handle, frames.ctypes.data_as(C.c_void_p),
                                       frames.dtype.itemsize, frames.shape[0]))

    def distance_histogram(self, lo, hi, nbins, out=None):
        """Adds the histogram of the listed pair distances of the last block to `out` (int64
        [nbins]; created when None).  Every unordered pair counts twice (both directions)."""
        if out is None:
            out = np.zeros(int(nbins), np.int64)
        check(_abi.lib().cmd_topo_distance_histogram(self._handle, float(lo), float(hi), int(nbins),
                                                     ptr(out, C.c_int64)))
        return out

    def build_dev(self, data_ptr, nframes):
        """Frames already in HBM: raw device pointer to float64 [F, n, 3]."""
        check(_abi.lib().cmd_topo_build_dev(self._handle, C.c_void_p(int(data_ptr)), int(nframes)))

    def dr_dev(self, frames_ptr, nframes, prev_ptr, out_ptr):
        """Step lengths dr[F, n] of a device block (topology.py:98); prev_ptr: the frame before the
        block, or None for the first frame of a trajectory (zeros)."""
        check(_abi.lib().cmd_topo_dr_dev(self._handle, C.c_void_p(int(frames_ptr)), int(nframes),
                                         C.c_void_p(int(prev_ptr)) if prev_ptr else None,
                                         C.c_void_p(int(out_ptr))))

    def skip_dr_dev(self, dr_ptr, nframes):
        """Walks `nframes` frames of the rebuild schedule from their step lengths (device [F, n]);
        returns the chunk-relative index of the last rebuild frame, -1 if there is none."""
        last = C.c_int64(-1)
        check(_abi.lib().cmd_topo_skip_dr_dev(self._handle, C.c_void_p(int(dr_ptr)), int(nframes),
                                              C.byref(last)))
        return int(last.value)

    def seed_dev(self, rebuild_frame_ptr, prev_frame_ptr):
        """After skip_dr_dev: the list of the last rebuild frame from its coordinates, and the
        frame right before the block (both device [n, 3])."""
        check(_abi.lib().cmd_topo_seed_dev(self._handle, C.c_void_p(int(rebuild_frame_ptr)),
                                           C.c_void_p(int(prev_frame_ptr))))

    def frame_info(self):
        n = self.nframes
        counts = np.zeros(n, np.int64)
        rebuilt = np.zeros(n, np.uint8)
        rate_sum = np.zeros(n)
        check(_abi.lib().cmd_topo_frame_info(self._handle, ptr(counts, C.c_int64),
                                             ptr(rebuilt, C.c_uint8), ptr(rate_sum)))
        return counts, rebuilt.astype(bool), rate_sum

    def get_frame(self, f, count=None):
        if count is None:
            count = int(self.frame_info()[0][f])
        start, dest = np.empty(count, np.int32), np.empty(count, np.int32)
        dist, omega = np.empty(count), np.empty(count)
        check(_abi.lib().cmd_topo_get_frame(self._handle, int(f), ptr(start, C.c_int),
                                            ptr(dest, C.c_int), ptr(dist), ptr(omega)))
        return start, dest, dist, omega

    def get_block(self, f0, nf, counts, omega=False):
        """(start, dest, dist[, omega]) of frames [f0, f0+nf) as host arrays [nf, width], width = the
        largest count among them: one strided copy per array instead of one round trip per frame."""
        width = max(int(np.max(counts[f0:f0 + nf])), 1)
        start, dest = np.empty((nf, width), np.int32), np.empty((nf, width), np.int32)
        dist = np.empty((nf, width))
        om = np.empty((nf, width)) if omega else None
        check(_abi.lib().cmd_topo_get_block(self._handle, int(f0), int(nf), width, ptr(start, C.c_int),
                                            ptr(dest, C.c_int), ptr(dist), ptr(om) if omega else None))
        return (start, dest, dist, om) if omega else (start, dest, dist)

    def tie_count(self):
        return int(_abi.lib().cmd_topo_tie_count(self._handle))

    def set_groups(self, group, n_extra):
        """donor index -> index of the extra atom it is bonded to (AngleTopology)."""
        group = np.ascontiguousarray(group, dtype=np.int32)
        if group.shape != (self.n_atoms,):
            raise ValueError("one group entry per donor atom expected")
        check(_abi.lib().cmd_topo_set_groups(self._handle, ptr(group, C.c_int), int(n_extra)))

    def apply_angles(self, extra_frames):
        """Angles of the last block's pairs from the extra-atom positions [F, n_extra, 3] of the
        same frames; masks the rates when the topology was created with a FermiAngle rate."""
        extra_frames = np.ascontiguousarray(extra_frames)
        if extra_frames.dtype not in (np.float32, np.float64):
            extra_frames = extra_frames.astype(np.float64)
        check(_abi.lib().cmd_topo_apply_angles(self._handle, extra_frames.ctypes.data_as(C.c_void_p),
                                               extra_frames.dtype.itemsize))

    def get_frame_angles(self, f, count=None):
        if count is None:
            count = int(self.frame_info()[0][f])
        theta = np.empty(count)
        check(_abi.lib().cmd_topo_get_frame_angles(self._handle, int(f), ptr(theta)))
        return theta

    def positions_ptr(self):
        """Device pointer of the frames the last block was built from."""
        p = C.c_void_p()
        check(_abi.lib().cmd_topo_positions(self._handle, C.byref(p)))
        return p


def build_with_retry(make, frames):
    """Runs make(capacity).build(frames); on a capacity overflow re-creates the object with the
    capacity the library asked for (plus head-room) and builds again."""
    capacity = 0
    for _ in range(6):
        topo = make(capacity)
        try:
            topo.build(frames)
            return topo
        except _abi.CmdError as e:
            if e.code != -5:
                raise
            need = topo.capacity_needed or max(64, topo.stride * 2)
            capacity = need + need // 4 + 64
    raise RuntimeError("could not size the per-frame pair capacity")


class NeighborTopology:
    """Keeps track of the connections between donor/acceptor atoms.
    Given a cutoff distance, for each atom the atoms within this
    distance will be determined.  (mdlmc/topo/topology.py:18-121)"""
    __show_in_config__ = True
    __no_config_parameter__ = ["trajectory", "atom_box"]

    #: frames pulled from the trajectory and sent to the GPU per launch
    chunk_size = 256

    def __init__(self, trajectory, atom_box, *, donor_atoms: str, cutoff: float = 3.0,
                 buffer: float = 2.0) -> None:
        self._raw_trajectory = trajectory
        self._cache = deque()
        self._frame_iter = None
        self.trajectory = trajectory
        self.trajectory_time_step = trajectory.time_step
        self.cutoff = cutoff
        self.buffer = buffer
        self.atombox = atom_box
        self.donor_atoms = donor_atoms
        self._jumprate = None

    # -- frame cache with the semantics of misc/tools.py:249-261 as installed at topology.py:43:
    # every frame handed downstream is remembered until get_cached_frames() drains it
    def get_cached_frames(self):
        while self._cache:
            yield self._cache.popleft()

    def attach_jumprate(self, jumprate):
        """Lets the device pipeline evaluate the jump rate inside the topology kernels."""
        self._jumprate = jumprate

    def _determine_colvars(self, start_indices, destination_indices, distances, frame):
        """Per convention, the first collective variable is the distance (topology.py:50-53)."""
        return start_indices, destination_indices, distances

    # -- topology.py:55-72
    def get_topology_bruteforce(self, frame):
        """Determine the distance for each atom pair.  If it is below cutoff + buffer, add it to
        the list of connections."""
        frame = as_f64(frame)
        topo = build_with_retry(
            lambda cap: DeviceTopology(self.atombox, frame.shape[0], self.cutoff, self.buffer,
                                       MODE_BRUTEFORCE, self._jumprate, cap), frame[None])
        start, dest, dist, _ = topo.get_frame(0)
        return start, dest, dist

    def _donor_positions(self, full_frame):
        return np.asarray(full_frame[self.donor_atoms].atom_positions, dtype=float)

    def _frames(self):
        """The ONE frame stream of this topology: like the reference's
        cache_last_elements(trajectory) generator (topology.py:43) it is shared by everything that
        pulls frames, so a frame taken once is not delivered again."""
        if self._frame_iter is None:
            from .trajectory import ArrayTrajectory
            if isinstance(self.trajectory, ArrayTrajectory):
                self._frame_iter = _ArrayFrames(self.trajectory)
            else:
                self._frame_iter = iter(self.trajectory)
        return self._frame_iter

    def _chunks(self):
        it = self._frames()
        if isinstance(it, _ArrayFrames) and not it.traj.repeat:
            while it.pos < len(it.traj):
                k0, k1 = it.pos, min(len(it.traj), it.pos + self.chunk_size)
                it.pos = k1
                yield _LazyFrames(it, k0, k1)
            return
        while True:
            frames = []
            for full_frame in it:
                frames.append(full_frame)
                if len(frames) == self.chunk_size:
                    break
            if not frames:
                return
            yield frames
            if len(frames) < self.chunk_size:
                return

    def device_blocks(self, mode=MODE_VERLET, chunk_size=None):
        """Yields (DeviceTopology, full_frames, host block) per block of frames: the block's neighbour
        lists and rates stay in HBM for the KMC kernel (no per-frame host traffic).  The host block is
        what was uploaded: the donor positions, or whole frames when the selection ran on the device."""
        if chunk_size is not None:
            self.chunk_size = int(chunk_size)
        topo = None
        select = None     # (n_total, donor rows): the selection runs on the device
        for full_frames in self._chunks():
            if isinstance(full_frames, _LazyFrames):   # array trajectory: one slice, no Frame objects
                traj = self.trajectory
                rows = traj.selection(self.donor_atoms)
                n_total = traj.positions.shape[1]
                if rows.size < n_total and 2 * rows.size >= n_total:
                    # most atoms are donors: the chunk goes up as it lies in memory (no pass over it
                    # on the host, the library stages pageable memory itself) and a gather kernel
                    # picks the donor rows; with few donors the host gather moves fewer bytes
                    select = (n_total, rows)
                    pos = traj.positions[full_frames.k0:full_frames.k1]
                    if not isinstance(pos, np.ndarray) or pos.dtype not in (np.float32, np.float64):
                        pos = np.asarray(pos, dtype=np.float32 if pos.dtype == np.float32 else np.float64)
                else:
                    pos = traj.block(self.donor_atoms, full_frames.k0, full_frames.k1)
            else:
                pos = np.stack([self._donor_positions(f) for f in full_frames])
            if topo is None:
                n_donors = select[1].size if select else pos.shape[1]

                def make(cap):
                    t = DeviceTopology(self.atombox, n_donors, self.cutoff, self.buffer, mode,
                                       self._jumprate, cap)
                    if select:
                        t.set_selection(*select)
                    return t
                topo = build_with_retry(make, pos)
            else:
                topo.build(pos)   # a capacity overflow mid-trajectory raises CmdError(-5)
            yield topo, full_frames, pos

    def _generate(self, mode):
        for topo, full_frames, _ in self.device_blocks(mode):
            counts, _, _ = topo.frame_info()
            if (counts < 0).any():
                raise _abi.CmdError(-5, "a frame overflowed its pair capacity")
            # the block's lists leave the GPU in three strided copies; the per-frame arrays the
            # protocol yields are fresh copies of their rows
            start, dest, dist = topo.get_block(0, len(counts), counts)
            for k, full_frame in enumerate(full_frames):
                c = int(counts[k])
                self._cache.append(full_frame)
                yield start[k, :c].copy(), dest[k, :c].copy(), dist[k, :c].copy(), full_frame

    # -- topology.py:74-78
    def topology_bruteforce_generator(self):
        yield from self._generate(MODE_BRUTEFORCE)

    # -- topology.py:80-114
    def topology_verlet_list_generator(self):
        """Keep track of the two maximum atom displacements.  As soon as their sum is larger
        than the buffer region, update the neighbor topology."""
        yield from self._generate(MODE_VERLET)

    def __iter__(self):
        for topo in self.topology_verlet_list_generator():
            yield self._determine_colvars(*topo)

    def update_time_of_last_jump(self, proton_idx, new_time):
        pass


class AngleTopology(NeighborTopology):
    """This topology class is used to calculate the POO angle as an additional collective variable.
    Of course, other atom types are possible as well.  In that case, the parameters for donor_atoms
    and extra_atoms just need to be changed accordingly.  (mdlmc/topo/topology.py:124-167)

                             O -- O
                            /
                           P
    """

    def __init__(self, trajectory, atom_box, *, donor_atoms: str, extra_atoms: str, group_size: int,
                 cutoff: float = 3.0, buffer: float = 2.0) -> None:
        super().__init__(trajectory, atom_box, donor_atoms=donor_atoms, cutoff=cutoff, buffer=buffer)
        self.extra_atoms = extra_atoms
        self.group_size = group_size
        self._determine_groups()

    def _determine_groups(self):
        """Find for each phosphorus atom the closest oxygen atoms (topology.py:142-156): the donor
        atoms belonging to one group.  Later groups overwrite earlier ones, like the dict upstream."""
        # upstream this is next(iter(self.trajectory)) on the cached one-shot generator
        # (topology.py:43,145): the first frame is consumed here, stays in the frame cache, and
        # the iteration over the topology starts at the second frame
        first_frame = next(self._frames())
        self._cache.append(first_frame)
        centres = first_frame[self.extra_atoms].atom_positions
        donors = first_frame[self.donor_atoms].atom_positions
        # one all-to-all distance matrix on the device, rows = extra atoms
        nearest = np.argsort(self.atombox.length_all_to_all(centres, donors), axis=1)[:, :self.group_size]
        self.n_extra = nearest.shape[0]
        # donor -> extra atom; a donor claimed by several extra atoms goes to the LAST of them, as
        # in the dict of the reference (later entries overwrite earlier ones) = the largest index
        self._group = np.full(donors.shape[0], -1, dtype=np.int32)
        np.maximum.at(self._group, nearest.ravel(),
                      np.repeat(np.arange(self.n_extra, dtype=np.int32), nearest.shape[1]))
        self.map_O_to_P = {int(o): int(p) for o, p in enumerate(self._group) if p >= 0}

    def _extra_positions(self, full_frame):
        return np.asarray(full_frame[self.extra_atoms].atom_positions, dtype=float)

    def _determine_colvars(self, start_indices, destination_indices, distances, frame):
        """Determine here the POO angles (topology.py:158-167), one batched CUDA call."""
        if len(start_indices) and (self._group[start_indices] < 0).any():
            raise KeyError(int(start_indices[self._group[start_indices] < 0][0]))
        p_atoms = self._extra_positions(frame)
        o_atoms = self._donor_positions(frame)
        angles = self.atombox.angle(p_atoms[self._group[start_indices]], o_atoms[start_indices],
                                    o_atoms[destination_indices])
        return start_indices, destination_indices, distances, np.atleast_1d(angles)

    def device_blocks(self, mode=MODE_VERLET, chunk_size=None):
        """Block pipeline with the angle colvar (and the FermiAngle mask) applied on the device."""
        for topo, full_frames, pos in super().device_blocks(mode, chunk_size):
            if not getattr(topo, "_groups_set", False):
                topo.set_groups(self._group, self.n_extra)
                topo._groups_set = True
            if isinstance(full_frames, _LazyFrames):
                extra = self.trajectory.block(self.extra_atoms, full_frames.k0, full_frames.k1)
            else:
                extra = np.stack([self._extra_positions(f) for f in full_frames])
            topo.apply_angles(extra)
            yield topo, full_frames, pos

    def _generate(self, mode):
        for topo, full_frames, _ in self.device_blocks(mode):
            counts, _, _ = topo.frame_info()
            for k, full_frame in enumerate(full_frames):
                start, dest, dist, _ = topo.get_frame(k, int(counts[k]))
                self._cache.append(full_frame)
                yield start, dest, dist, full_frame

    def __iter__(self):
        for topo, full_frames, _ in self.device_blocks(MODE_VERLET):
            counts, _, _ = topo.frame_info()
            for k, full_frame in enumerate(full_frames):
                start, dest, dist, _ = topo.get_frame(k, int(counts[k]))
                self._cache.append(full_frame)
                yield start, dest, dist, topo.get_frame_angles(k, int(counts[k]))


class DistanceTransformation:
    """If a topology which supports a transformation of the donor-acceptor distances is chosen
    (for example HydroniumTopology), this class specifies how the donor-acceptor distances are
    rescaled.  (mdlmc/topo/topology.py:260-270)"""
    __show_in_config__ = True
    transform_kind = 0

    def __call__(self, distances):
        return distances

    def device_parameters(self):
        """(kind, tpar[5], table_x, table_y) for cmd_kmc_set_hydronium."""
        return 0, np.zeros(5), None, None


class ReLUTransformation(DistanceTransformation):
    """Rectified Linear Unit transformation (topology.py:273-291): a constant b below d0, the line
    a (d - d0) + b above it, identity outside (left_bound, right_bound)."""
    transform_kind = 1

    def __init__(self, a: float, b: float, d0: float, left_bound: float, right_bound: float) -> None:
        self.a, self.b, self.d0 = a, b, d0
        self.left_bound, self.right_bound = left_bound, right_bound

    def __call__(self, distances):
        distances = np.asarray(distances, dtype=float)
        # same value per element as topology.py:288-292: inside (left_bound, right_bound) the
        # constant b below d0 and the line a (d - d0) + b above it, the distance itself elsewhere
        ramp = np.maximum(distances - self.d0, 0.0)
        inside = (self.left_bound < distances) & (distances < self.right_bound)
        return np.where(inside, np.where(ramp > 0.0, self.a * (distances - self.d0) + self.b, self.b),
                        distances)

    def device_parameters(self):
        return 1, np.array([self.a, self.b, self.d0, self.left_bound, self.right_bound], float), None, None


class InterpolatedTransformation(DistanceTransformation):
    """Transform O-O distances by linear interpolation in a table (topology.py:294-334; scipy's
    interp1d(kind="linear") upstream: slope * (x - x_lo) + y_lo)."""
    __show_signature_of__ = "from_file"
    transform_kind = 2

    def __init__(self, dist_array, conversion_array):
        self.x = np.ascontiguousarray(dist_array, dtype=float)
        self.y = np.ascontiguousarray(conversion_array, dtype=float)
        if self.x.ndim != 1 or self.x.shape != self.y.shape or self.x.size < 2:
            raise ValueError("dist_array and conversion_array must be 1-d arrays of equal length >= 2")
        self.x_min, self.x_max = self.x[0], self.x[-1]
        self.y_min = self.y[0]

    @classmethod
    def from_file(cls, dist_array_filename: str, conversion_array_filename: str):
        return cls(np.load(dist_array_filename), np.load(conversion_array_filename))

    def _interp(self, d):
        idx = np.clip(np.searchsorted(self.x, d), 1, self.x.size - 1)
        lo, hi = idx - 1, idx
        slope = (self.y[hi] - self.y[lo]) / (self.x[hi] - self.x[lo])
        return slope * (d - self.x[lo]) + self.y[lo]

    def __call__(self, distances):
        # per element the values of topology.py:328-334: the table inside [x_min, x_max], its first
        # value below it, the distance itself above it
        d = np.asarray(distances, dtype=float)
        table = self._interp(np.clip(d, self.x_min, self.x_max))
        return np.where(d < self.x_min, self.y_min, np.where(d <= self.x_max, table, d))

    def device_parameters(self):
        return 2, np.zeros(5), self.x, self.y


class DistanceInterpolator:
    """Interpolates linearly in time between neutral and relaxed donor-acceptor distances
    (topology.py:337-353)."""
    __show_in_config__ = True

    def __init__(self, relaxation_time: float):
        self.relaxation_time = relaxation_time

    def __call__(self, residence_time, distance_neutral, distance_relaxed):
        ratio = np.minimum(np.asarray(residence_time) / self.relaxation_time, 1)[:, None]
        return (1 - ratio) * distance_neutral + ratio * distance_relaxed


class HydroniumTopology(NeighborTopology):
    """Mimics the neighbor topology of a H3O+ ion in water by only defining connections to the
    closest oxygen neighbors (mdlmc/topo/topology.py:170-257): per site the four nearest listed
    neighbours, distances rescaled by a DistanceTransformation and relaxed with the residence
    time of the proton sitting on the site.

    In the device pipeline the lattice-independent part (nearest neighbours per site and frame)
    is computed with the lists (cmd_topo_nearest) and the per-replica rescaling + rate inside the
    KMC kernel (cmd_kmc_set_hydronium); `_determine_colvars` is the host-level mirror."""
    __no_config_parameter__ = ["trajectory", "atom_box", "distance_transformation_function",
                               "distance_interpolator"]
    n_nearest = 4

    def __init__(self, trajectory, atom_box, *, donor_atoms: str, cutoff: float, buffer: float = 0.0,
                 distance_transformation_function: "DistanceTransformation" = None,
                 distance_interpolator: "DistanceInterpolator" = None) -> None:
        super().__init__(trajectory, atom_box, donor_atoms=donor_atoms, cutoff=cutoff, buffer=buffer)
        self._time_of_last_jump_vec = None
        self._distance_transformation_function = distance_transformation_function or DistanceTransformation()
        self._distance_interpolator = distance_interpolator
        self._lattice = None

    def take_lattice_reference(self, lattice):
        """Stores a read-only view of KMCLattice's lattice (topology.py:201-211)."""
        self._lattice = lattice.view()
        self._lattice.flags.writeable = False
        self._proton_number = int((lattice != 0).sum())
        self._time_of_last_jump_vec = -np.ones(self._proton_number)

    def hydronium_parameters(self):
        """What cmd_kmc_set_hydronium needs besides the rate: transformation and relaxation time."""
        kind, tpar, tx, ty = self._distance_transformation_function.device_parameters()
        relax = self._distance_interpolator.relaxation_time if self._distance_interpolator else 0.0
        return dict(kind=kind, tpar=tpar, table_x=tx, table_y=ty, relaxation_time=float(relax),
                    frame_time_step=float(self.trajectory_time_step))

    def transform_distances(self, occupied_indices, distances, time):
        """Host mirror of topology.py:213-229 (the device does this per replica inside the KMC
        kernel): rescaled distances of the occupied sites, relaxed with the time the proton on
        each has spent there (infinite for a proton that has not jumped yet)."""
        rescaled = self._distance_transformation_function(distances)
        if self._distance_interpolator is None:
            return rescaled
        sites = np.unique(occupied_indices)
        since = self._time_of_last_jump_vec[self._lattice[sites] - 1]       # per proton label
        residence = np.full(since.shape, np.inf)
        jumped = since >= 0
        residence[jumped] = time - since[jumped]
        return self._distance_interpolator(residence, distances, rescaled)

    def _colvars_from_nearest(self, near_dest, near_dist, frame):
        """(start, dest, rescaled distance) of one frame from the device's nearest-neighbour arrays
        (topology.py:234-253): the selection ran in k_nearest, the user's transformation /
        interpolator objects are applied here like upstream."""
        donor_nr = len(self._lattice)
        n_atoms = self.n_nearest
        new_start = np.repeat(np.arange(donor_nr), n_atoms).reshape(donor_nr, n_atoms)
        new_dist = self.transform_distances(new_start, near_dist.reshape(donor_nr, n_atoms), frame.time)
        return new_start.flatten(), near_dest.astype(int), new_dist.flatten()

    def device_blocks(self, mode=MODE_VERLET, chunk_size=None):
        for topo, full_frames, pos in super().device_blocks(mode, chunk_size):
            check(_abi.lib().cmd_topo_nearest(topo.handle, self.n_nearest))
            yield topo, full_frames, pos

    def __iter__(self):
        if self._lattice is None:
            raise RuntimeError("take_lattice_reference has not been called (KMCLattice does it)")
        for topo, full_frames, _ in self.device_blocks(MODE_VERLET):
            for k, full_frame in enumerate(full_frames):
                dest = np.zeros(self.n_nearest * len(self._lattice), np.int32)
                dist = np.zeros(self.n_nearest * len(self._lattice))
                check(_abi.lib().cmd_topo_get_frame_nearest(topo.handle, k, ptr(dest, C.c_int), ptr(dist)))
                self._cache.append(full_frame)
                yield self._colvars_from_nearest(dest, dist, full_frame)

    def update_time_of_last_jump(self, proton_idx, new_time):
        self._time_of_last_jump_vec[proton_idx - 1] = new_time
