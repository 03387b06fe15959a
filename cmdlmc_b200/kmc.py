"""KMCLattice -- host mirror of mdlmc/LMC/MDMC.py:28-277 on top of the CUDA KMC kernel.

Same constructor, iteration protocol and outputs as the reference:
    iter(kmc)                      -> (frame_number, kmc_time, Frame)            MDMC.py:74-99
    kmc.xyz_output(particle_type)  -> Frame with the proton positions appended   MDMC.py:173-177
    kmc.observables_output(r, p)   -> (frame_number, kmc_time, msd[3], autocorr) MDMC.py:179-208
The time stepping, transition selection, proton moves and observables run on the GPU for whole
blocks of frames; this class only moves blocks and turns the event log back into the
reference's per-frame tuples.
"""
import ctypes as C
import logging
from abc import ABCMeta
from typing import Iterator

import numpy as np

from . import _abi, runtime
from ._abi import as_f64, check, ptr
from .topology import MODE_VERLET

logger = logging.getLogger(__name__)

RNG_REPLAY = 0
RNG_PHILOX = 1


class DeviceKMC:
    """Owner of a `cmd_kmc` handle: n_replicas independent KMC replicas on one topology stream."""

    def __init__(self, atom_box, lattices, time_step, rng_mode=RNG_PHILOX, seed=0):
        runtime.ensure_init()
        lattices = np.ascontiguousarray(np.atleast_2d(lattices), dtype=np.int32)
        self.atom_box = atom_box
        self.n_replicas, self.n_sites = lattices.shape
        self.rng_mode = rng_mode
        self._handle = C.c_void_p()
        check(_abi.lib().cmd_kmc_create(atom_box.handle, self.n_sites, self.n_replicas,
                                        ptr(lattices, C.c_int), float(time_step), int(rng_mode),
                                        int(seed) & (2 ** 64 - 1), C.byref(self._handle)))
        self._event_cap = 0

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _abi.lib().cmd_kmc_destroy(h)
            except Exception:
                pass
            self._handle = None

    def set_replica_ids(self, first, step):
        """Global ids of the local replicas (first + r * step) for the Philox counter: a replica's
        random stream is the same whatever the number of GPUs sharing the ensemble."""
        check(_abi.lib().cmd_kmc_set_replica_ids(self._handle, int(first), int(step)))

    def set_hydronium(self, jumprate, *, kind, tpar, table_x, table_y, relaxation_time,
                      frame_time_step):
        """HydroniumTopology mode: per-replica distance rescaling + rate inside the KMC kernel."""
        tx = np.ascontiguousarray(table_x, dtype=float) if table_x is not None else None
        ty = np.ascontiguousarray(table_y, dtype=float) if table_y is not None else None
        check(_abi.lib().cmd_kmc_set_hydronium(
            self._handle, int(jumprate.kind), ptr(jumprate._par()), int(kind),
            ptr(np.ascontiguousarray(tpar, dtype=float)), ptr(tx) if tx is not None else None,
            ptr(ty) if ty is not None else None, 0 if tx is None else tx.size,
            float(relaxation_time), float(frame_time_step)))

    def last_jump_times(self):
        out = np.zeros((self.n_replicas, self.n_sites))
        check(_abi.lib().cmd_kmc_get_last_jump_times(self._handle, ptr(out)))
        return out

    def enable_occupancy(self):
        check(_abi.lib().cmd_kmc_enable_occupancy(self._handle))

    def occupancy(self):
        """(counts int64[n_sites], replica_frames): how many (replica, consumed frame) pairs saw
        each site occupied, and the number of such pairs in total."""
        counts = np.zeros(self.n_sites, np.int64)
        frames = C.c_int64(0)
        check(_abi.lib().cmd_kmc_get_occupancy(self._handle, ptr(counts, C.c_int64), C.byref(frames)))
        return counts, frames.value

    def set_replay_stream(self, u):
        """u: per replica the uniforms the reference would draw from the legacy RandomState after
        its shuffle (random(), uniform-u, random(), ...).  The even entries are turned into the
        reference's time selectors -np.log(1 - u) here, with NumPy's log like upstream
        (MDMC.py:148); the device consumes them as they are."""
        u = np.array(np.atleast_2d(u), dtype=np.float64, order="C")
        if u.shape[0] != self.n_replicas:
            raise ValueError("one replay stream per replica expected")
        u[:, 0::2] = -np.log(1 - u[:, 0::2])
        check(_abi.lib().cmd_kmc_set_replay_stream(self._handle, ptr(u), u.shape[1]))

    def set_event_log(self, max_events_per_replica):
        check(_abi.lib().cmd_kmc_set_event_log(self._handle, int(max_events_per_replica)))
        self._event_cap = int(max_events_per_replica)

    def set_observables(self, reset_frequency, print_frequency):
        check(_abi.lib().cmd_kmc_set_observables(self._handle, int(reset_frequency),
                                                 int(print_frequency)))

    def advance(self, topo, positions_ptr=None):
        p = positions_ptr if positions_ptr is not None else C.c_void_p(0)
        check(_abi.lib().cmd_kmc_advance(self._handle, topo.handle, p))

    def state(self):
        r = self.n_replicas
        lattices = np.zeros((r, self.n_sites), np.int32)
        time, frame = np.zeros(r), np.zeros(r, np.int64)
        n_events, site_updates = np.zeros(r, np.int64), np.zeros(r, np.int64)
        check(_abi.lib().cmd_kmc_get_state(self._handle, ptr(lattices, C.c_int), ptr(time),
                                           ptr(frame, C.c_int64), ptr(n_events, C.c_int64),
                                           ptr(site_updates, C.c_int64)))
        return dict(lattices=lattices, time=time, frame=frame, n_events=n_events,
                    site_updates=site_updates)

    def status(self):
        r = self.n_replicas
        phase, reason = np.zeros(r, np.int32), np.zeros(r, np.int32)
        cursor = np.zeros(r, np.int64)
        check(_abi.lib().cmd_kmc_get_status(self._handle, ptr(phase, C.c_int),
                                            ptr(reason, C.c_int), ptr(cursor, C.c_int64)))
        return phase, reason, cursor

    def events(self, replica=0):
        cap = self._event_cap
        n = C.c_int64(0)
        frame, time = np.zeros(cap, np.int64), np.zeros(cap)
        start, dest, proton = (np.zeros(cap, np.int32) for _ in range(3))
        check(_abi.lib().cmd_kmc_get_events(self._handle, int(replica), cap, C.byref(n),
                                            ptr(frame, C.c_int64), ptr(time), ptr(start, C.c_int),
                                            ptr(dest, C.c_int), ptr(proton, C.c_int)))
        k = n.value
        dist = np.zeros(cap)
        check(_abi.lib().cmd_kmc_get_event_distances(self._handle, int(replica), cap, C.byref(n),
                                                     ptr(dist)))
        return dict(frame=frame[:k], time=time[:k], start=start[:k], dest=dest[:k],
                    proton=proton[:k], dist=dist[:k])

    def jump_histogram(self, lo, hi, nbins, out=None):
        """Adds the histogram of the O-O distances of all logged jumps (every replica) to `out`."""
        if out is None:
            out = np.zeros(int(nbins), np.int64)
        check(_abi.lib().cmd_kmc_jump_histogram(self._handle, float(lo), float(hi), int(nbins),
                                                ptr(out, C.c_int64)))
        return out

    def seed_observables(self, positions):
        """Donor positions [n_sites, 3] of a frame that becomes frame number 0 of the observables
        without being walked by the KMC (cmd_kmc_seed_observables)."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.shape != (self.n_sites, 3):
            raise ValueError("positions must have shape [%d, 3]" % self.n_sites)
        check(_abi.lib().cmd_kmc_seed_observables(self._handle, ptr(pos)))

    def observables(self, replica=0, first=0):
        """Time-stamped observable rows [first:] of a replica: (frame, time, msd_x, msd_y, msd_z,
        autocorrelation) each."""
        n = C.c_int64(0)
        check(_abi.lib().cmd_kmc_get_observables(self._handle, int(replica), 0, C.byref(n), None))
        count = max(int(n.value) - int(first), 0)
        rows = np.zeros((count, 6))
        if count:
            check(_abi.lib().cmd_kmc_get_observable_rows(self._handle, int(replica), int(first), count,
                                                         ptr(rows)))
        return rows

    def tie_count(self):
        return int(_abi.lib().cmd_kmc_tie_count(self._handle))

    def events_dropped(self):
        """Events that did not fit the event log since set_event_log (all replicas)."""
        return int(_abi.lib().cmd_kmc_events_dropped(self._handle))

    def selection_fallbacks(self):
        """Events the one-CTA-per-replica kernel re-decided with the sequential np.cumsum."""
        return int(_abi.lib().cmd_kmc_selection_fallbacks(self._handle))


class KMCLattice:
    """Implementation of the time-dependent Kinetic Monte Carlo Scheme (MDMC.py:28-226).

    rng="replay" (default) follows the reference's use of the global legacy np.random state (one
    shuffle at construction, then random() / uniform(0, S) alternating per event): after
    np.random.seed(s) the event trace (frame, start, destination, proton) is the one the oracle's
    restatement of MDMC.py produces from the same stream, bit for bit, and the reference's own
    seeded runs are reproduced by the golden tests (decisions closer than 1e-9 to a boundary are
    counted in `tie_count`: the rates come from CUDA's exp, a few ulp from NumPy's).  The uniforms
    are drawn from np.random one frame block AHEAD of their use (2 * (events_per_frame_bound *
    frames + 64) per block), so the global generator state differs from upstream during and after
    the run: other np.random consumers interleaved with the iteration see different numbers.
    rng="philox" uses the counter-based device generator (key = seed, counter = replica, event).
    """

    __show_in_config__ = True
    __no_config_parameter__ = ["topology", "atom_box", "jumprate_function"]

    #: upper bound of KMC events per trajectory frame used to size the replay stream of a block
    events_per_frame_bound = 16

    def __init__(self, topology: "NeighborTopology", *,
                 atom_box: "AtomBox",
                 jumprate_function: "JumpRate",
                 lattice_size: int,
                 proton_number: int,
                 donor_atoms: str,
                 time_step: float,
                 extra_atoms: str = None,
                 rng: str = "replay",
                 seed: int = 0,
                 chunk_size: int = 1024):
        self.topology = topology
        self._lattice = self._initialize_lattice(lattice_size, proton_number)
        if hasattr(self.topology, "take_lattice_reference"):
            self.topology.take_lattice_reference(self._lattice)
        self._atom_box = atom_box
        self._jumprate_function = jumprate_function
        self._donor_atoms = donor_atoms
        self._time_step = time_step
        self._extra_atoms = extra_atoms
        self._rng = rng
        self._seed = seed
        self._chunk_size = chunk_size
        if hasattr(self.topology, "attach_jumprate"):
            self.topology.attach_jumprate(jumprate_function)
        self._device = None
        self._pending_u = np.zeros(0)
        self._event_blocks = []
        self.tie_count = 0

    def _initialize_lattice(self, lattice_size, proton_number):
        lattice = np.zeros(lattice_size, dtype=np.int32)
        lattice[:proton_number] = range(1, proton_number + 1)
        np.random.shuffle(lattice)          # MDMC.py:71, global legacy state
        return lattice

    # ------------------------------------------------------------------ device pipeline ----
    def _blocks(self, observables=None):
        """Runs topology + KMC block by block on the GPU.  Yields
        (first_frame_number, full_frames, events_of_block)."""
        mode = RNG_REPLAY if self._rng == "replay" else RNG_PHILOX
        dev = DeviceKMC(self._atom_box, self._lattice, self._time_step, mode, self._seed)
        self._device = dev
        # Frames the topology consumed before the iteration starts (AngleTopology._determine_groups
        # pulls the first trajectory frame, topology.py:145) sit in the reference's frame cache and
        # come out first at the first event (MDMC.py:92-94): they take the frame numbers 0.., the
        # frames the KMC walks follow.
        cache = getattr(self.topology, "_cache", None)
        self._pre_frames = list(cache) if cache else []
        if cache:
            cache.clear()
        if observables:
            dev.set_observables(*observables)
            if len(self._pre_frames) > 1:
                raise NotImplementedError("more than one frame consumed before the KMC iteration")
            if self._pre_frames:
                f0 = self._pre_frames[0]
                dev.seed_observables(np.asarray(f0[self.donor_atoms].atom_positions, dtype=float))
        if hasattr(self.topology, "hydronium_parameters"):
            dev.set_hydronium(self._jumprate_function, **self.topology.hydronium_parameters())
        first = 0
        for topo, full_frames, _ in self.topology.device_blocks(MODE_VERLET, self._chunk_size):
            nfr = len(full_frames)
            dev.set_event_log(self.events_per_frame_bound * nfr + 64)   # restarts the log
            if mode == RNG_REPLAY:
                need = 2 * (self.events_per_frame_bound * nfr + 64)
                if self._pending_u.size < need:
                    fresh = np.random.random_sample(need - self._pending_u.size)
                    self._pending_u = np.concatenate([self._pending_u, fresh])
                dev.set_replay_stream(self._pending_u[None])
            dev.advance(topo, topo.positions_ptr() if observables else None)
            phase, reason, cursor = dev.status()
            if reason[0] == 1:
                raise RuntimeError("replay stream exhausted inside a block: raise "
                                   "KMCLattice.events_per_frame_bound")
            dropped = dev.events_dropped()
            if dropped:   # the lattices / flush times of the outputs are rebuilt from this log
                raise RuntimeError("%d events did not fit the event log of a block of %d frames: raise "
                                   "KMCLattice.events_per_frame_bound (now %d)"
                                   % (dropped, nfr, self.events_per_frame_bound))
            if mode == RNG_REPLAY:
                self._pending_u = self._pending_u[int(cursor[0]):]
            self._lattice[:] = dev.state()["lattices"][0]
            self.tie_count = dev.tie_count() + topo.tie_count()
            ev = dev.events(0)
            self._event_blocks.append(ev)
            yield first, full_frames, ev
            first += nfr
            if reason[0] == 2:   # the reference raises IndexError on an empty cumsum
                raise IndexError("no allowed proton transition left (MDMC.py:110)")

    def __iter__(self) -> Iterator:
        yield from self.continuous_output()

    def continuous_output(self):
        """(frame_number, kmc_time, Frame) for every frame flushed by an event (MDMC.py:77-99):
        frames consumed since the previous event carry the time of the event that flushed them."""
        for item in self._frames_with_lattice():
            yield item[0], item[1], item[2]

    def _frames_with_lattice(self):
        carry = None    # frames consumed but not yet flushed by an event (kept across blocks)
        lattice = self._lattice.copy()
        for first, full_frames, ev in self._blocks():
            if carry is None:   # (frame number, sweep it was consumed in, frame); see _blocks
                carry = [(k, -1, f) for k, f in enumerate(self._pre_frames)]
            m = len(self._pre_frames)
            frames = carry + [(first + m + k, first + k, f) for k, f in enumerate(full_frames)]
            e = 0
            out_upto = 0
            n_ev = len(ev["frame"])
            for idx, (number, sweep, frame) in enumerate(frames):
                # a frame is flushed by the first event whose sweep >= the sweep that consumed it;
                # events before it have already moved their protons (pre-jump lattice of the
                # flushing event)
                while e < n_ev and ev["frame"][e] < sweep:
                    lattice[ev["dest"][e]] = lattice[ev["start"][e]]
                    lattice[ev["start"][e]] = 0
                    e += 1
                if e >= n_ev:
                    break
                yield number, ev["time"][e], frame, lattice.copy()
                out_upto = idx + 1
            carry = frames[out_upto:]
            while e < n_ev:
                lattice[ev["dest"][e]] = lattice[ev["start"][e]]
                lattice[ev["start"][e]] = 0
                e += 1

    def xyz_output(self, particle_type: str = "H"):
        for _, _, frame, lattice in self._frames_with_lattice():
            occupied = np.where(lattice > 0)[0]
            particle_positions = frame[self.donor_atoms][occupied]
            particle_positions.atom_names = particle_type
            yield frame.append(particle_positions)

    def observables_output(self, reset_frequency: int, print_frequency: int):
        """(frame_number, kmc_time, msd[3], autocorrelation) rows computed on the device."""
        done = 0
        for _ in self._blocks(observables=(reset_frequency, print_frequency)):
            rows = self._device.observables(0, first=done)
            for row in rows:
                yield int(row[0]), row[1], row[2:5].copy(), int(row[5])
            done += len(rows)

    @property
    def event_log(self):
        """All events so far: dict of arrays frame (sweep), time, start, dest, proton."""
        keys = ("frame", "time", "start", "dest", "proton", "dist")
        if not self._event_blocks:
            return {k: np.zeros(0) for k in keys}
        return {k: np.concatenate([b[k] for b in self._event_blocks]) for k in keys}

    @property
    def lattice(self):
        return self._lattice

    @property
    def donor_atoms(self):
        return self._donor_atoms

    @property
    def extra_atoms(self):
        return self._extra_atoms

    @property
    def occupied_sites(self):
        return np.where(self._lattice > 0)[0]


class Output(metaclass=ABCMeta):
    __show_in_config__ = True
    __no_config_parameter__ = ["kmc"]


class XYZOutput(Output):
    def __init__(self, kmc: KMCLattice, particle_type: str) -> None:
        self.kmc = kmc
        self.particle_type = particle_type

    def __iter__(self):
        yield from self.kmc.xyz_output(self.particle_type)


class ObservablesOutput(Output):
    def __init__(self, kmc: KMCLattice, reset_frequency: int, print_frequency: int) -> None:
        self.kmc = kmc
        self.reset_frequency = reset_frequency
        self.print_frequency = print_frequency

    def __iter__(self):
        yield from self.kmc.observables_output(self.reset_frequency, self.print_frequency)
