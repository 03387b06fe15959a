"""Frame and the Trajectory protocol -- the input boundary of the hot path
(mdlmc/IO/trajectory_parser.py:43-135).  Parsing xyz / HDF5 files is host I/O and out of scope
(SURVEY.md section 8(f) F3); array-backed trajectories feed the GPU pipeline directly."""
from abc import ABCMeta, abstractmethod

import numpy as np


class Frame:
    """Names + positions (+ time) of all atoms of one MD step; selectable by atom name or by
    index (trajectory_parser.py:43-113)."""

    def __init__(self, names, positions, *, time=None):
        self._names = names
        self._positions = positions
        self._time = time

    @classmethod
    def from_recarray(cls, array, *, time=None):
        return cls(array["name"], array["pos"], time=time)

    def __getitem__(self, item):
        if isinstance(item, str):
            sel = self._names == item
        elif isinstance(item, (list, np.ndarray)):
            sel = item
        else:
            raise ValueError(f"Selection {item} not understood")
        return Frame(self._names[sel], self._positions[sel], time=self._time)

    def __repr__(self):
        body = "\n".join(f"{n}    {p[0]:20.10f} {p[1]:20.10f} {p[2]:20.10f}"
                         for n, p in zip(self.atom_names, self.atom_positions))
        return f"{self.atom_number}\n\n{body}"

    def append(self, f2):
        return Frame(np.hstack([self.atom_names, f2.atom_names]),
                     np.vstack([self.atom_positions, f2.atom_positions]))

    @property
    def atom_names(self):
        return self._names

    @atom_names.setter
    def atom_names(self, name):
        self._names[:] = name

    @property
    def atom_positions(self):
        return self._positions

    @property
    def atom_number(self):
        return self._names.size

    @property
    def time(self):
        return self._time


class Trajectory(metaclass=ABCMeta):
    """Iterable of Frame objects with a `time_step` attribute (trajectory_parser.py:116-135)."""
    __show_in_config__ = True

    @abstractmethod
    def __iter__(self):
        pass

    @property
    @abstractmethod
    def current_frame_number(self):
        pass

    @abstractmethod
    def __len__(self):
        pass


class ArrayTrajectory(Trajectory):
    """In-memory trajectory: positions [frames, atoms, 3] (float32 as HDF5Trajectory stores it,
    trajectory_parser.py:324, or float64) + atom names [atoms].  The GPU pipeline uploads whole
    frame blocks of `positions` instead of iterating Frame objects."""

    #: page-locked chunk buffers block() cycles through (a buffer is reused three blocks later)
    _ring_slots = 3

    def __init__(self, positions, atom_names, *, time_step: float, repeat: bool = False):
        # anything that slices like an array is taken as it is (an open HDF5 dataset stays on disk
        # and is read chunk by chunk, trajectory_parser.py:296,322)
        lazy = all(hasattr(positions, a) for a in ("shape", "ndim", "dtype", "__getitem__"))
        self.positions = positions if lazy else np.asarray(positions)
        self._ring = []
        self._ring_next = 0
        if self.positions.ndim != 3 or self.positions.shape[2] != 3:
            raise ValueError("positions must have shape [frames, atoms, 3]")
        self.atom_names = np.asarray(atom_names)
        if self.atom_names.shape[0] != self.positions.shape[1]:
            raise ValueError("atom_names must have one entry per atom")
        self.time_step = time_step
        self.repeat = repeat
        self._current_frame_number = 0

    def __iter__(self):
        step = 0
        while True:
            for pos in self.positions:
                self._current_frame_number = step
                yield Frame(self.atom_names, np.asarray(pos, dtype=float), time=step * self.time_step)
                step += 1
            if not self.repeat:
                break

    def __len__(self):
        return self.positions.shape[0]

    @property
    def current_frame_number(self):
        return self._current_frame_number

    def selection(self, name):
        """Indices of the atoms called `name`."""
        return np.where(self.atom_names == name)[0]

    def _chunk_buffer(self, shape, dtype):
        """The next buffer of the page-locked ring (runtime.pinned_empty: the library copies from it
        directly, the copy overlaps the kernels); plain memory when no CUDA device is bound."""
        need = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if len(self._ring) < self._ring_slots:
            self._ring.append(None)
        k = self._ring_next % len(self._ring)
        self._ring_next += 1
        buf = self._ring[k]
        if buf is None or buf.nbytes < need:
            try:
                from . import runtime
                buf = runtime.pinned_empty((max(need, 1),), np.uint8)
            except Exception:          # no device / no library: host-side tools only
                buf = np.empty(max(need, 1), np.uint8)
            self._ring[k] = buf
        else:
            try:                       # its last upload may still be in flight
                from . import runtime
                runtime.sync()
            except Exception:
                pass
        return buf[:need].view(dtype).reshape(shape)

    def block(self, name, start, stop):
        """[stop-start, n_selected, 3] positions of the atoms called `name` in a frame block, in the
        stored precision (float32 blocks are up-cast on the device, exactly) and in page-locked
        memory.  The array is valid until block() has been called `_ring_slots` more times."""
        sel = self.selection(name)
        blk = self.positions[start:stop]          # an HDF5 dataset reads the chunk here
        dtype = np.float32 if blk.dtype == np.float32 else np.float64
        out = self._chunk_buffer((blk.shape[0], sel.size, 3), dtype)
        if sel.size == blk.shape[1]:
            out[...] = blk
        else:
            np.take(blk, sel, axis=1, out=out) if blk.dtype == dtype else out.__setitem__(Ellipsis, blk[:, sel])
        return out


class XYZTrajectory(ArrayTrajectory):
    """xyz text trajectory with the constructor of the reference's XYZTrajectory
    (trajectory_parser.py:176-215): the file is read once on the host into a float64 block and
    then served like any array trajectory (the reference re-parses with np.genfromtxt per frame).
    `selection` keeps only atoms with that name / those names / those indices."""

    def __init__(self, filename, *, time_step: float, number_of_atoms: int = None,
                 selection=None, repeat: bool = False) -> None:
        names, frames = _read_xyz(filename, number_of_atoms)
        if selection is not None and selection != "None":
            if isinstance(selection, str):
                keep = names == selection
            elif len(selection) and isinstance(selection[0], str):
                keep = np.isin(names, list(selection))
            else:
                keep = np.zeros(names.shape[0], bool)
                keep[np.asarray(selection, dtype=int)] = True
            names, frames = names[keep], frames[:, keep]
        super().__init__(frames, names, time_step=time_step, repeat=repeat)
        self.filename = filename
        self.selection_spec = selection


def _read_xyz(filename, number_of_atoms=None):
    """(names [n], positions [frames, n, 3]) of a multi-frame xyz file or open text stream."""
    if hasattr(filename, "read"):
        lines = filename.read().splitlines()
    else:
        with open(filename, "r") as f:
            lines = f.read().splitlines()
    while lines and not lines[-1].strip():
        lines.pop()
    n = int(number_of_atoms) if number_of_atoms else int(lines[0].split()[0])
    per = n + 2
    if len(lines) < per or len(lines) % per:
        raise ValueError("xyz file does not hold whole frames of %d atoms" % n)
    n_frames = len(lines) // per
    names = np.array([lines[2 + i].split()[0] for i in range(n)])
    body = [ln for k, ln in enumerate(lines) if k % per >= 2]
    flat = np.array([ln.split()[1:4] for ln in body], dtype=float)
    return names, flat.reshape(n_frames, n, 3)


class NpzTrajectory(ArrayTrajectory):
    """Array trajectory from an .npz with `trajectory` [frames, atoms, 3] (float32 or float64, the
    HDF5 layout of the reference, IO/converters.py:38-43) and `atom_names` [atoms]."""

    def __init__(self, filename: str, *, time_step: float, repeat: bool = False) -> None:
        with np.load(filename) as z:
            super().__init__(z["trajectory"], z["atom_names"].astype(str), time_step=time_step,
                             repeat=repeat)
        self.filename = filename


class HDF5Trajectory(ArrayTrajectory):
    """HDF5 trajectory with the reference's layout (IO/converters.py:38-43: dataset `trajectory`
    float32 [frames, atoms, 3], `atom_names`) and constructor (trajectory_parser.py:290-311).  Needs
    h5py, which this image does not ship: the class exists so that reference drivers import, and
    works wherever h5py does.  The float32 block goes to the GPU as it is (up-cast on the device)."""

    def __init__(self, filename: str, time_step: float, selection=None, repeat: bool = False,
                 chunk_size: int = 1000) -> None:
        try:
            import h5py
        except ImportError as e:   # pragma: no cover - h5py is absent from the build image
            raise ImportError("HDF5Trajectory needs h5py; use NpzTrajectory / XYZTrajectory") from e
        self._file = h5py.File(filename, "r")
        names = np.asarray(self._file["atom_names"][:]).astype("<U2")
        positions = self._file["trajectory"]      # stays on disk: block() reads chunk by chunk
        if selection is not None and selection != "None":   # trajectory_parser.py:298-300 ignores it too
            import warnings
            warnings.warn("Selection is not implemented yet!")
        self.selection_spec = selection
        super().__init__(positions, names, time_step=time_step, repeat=repeat)
        self.filename = filename
        self._chunk_size = chunk_size

    def __iter__(self):
        step = 0
        while True:
            for c0 in range(0, len(self), self._chunk_size):     # trajectory_parser.py:313-337
                chunk = np.asarray(self.positions[c0:c0 + self._chunk_size], dtype=float)
                for pos in chunk:
                    self._current_frame_number = step
                    yield Frame(self.atom_names, pos, time=step * self.time_step)
                    step += 1
            if not self.repeat:
                break
