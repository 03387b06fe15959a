"""ctypes front-end of the CPU oracle (oracle/cmdlmc_oracle.c)  --  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the checker, never the product path.

The classes mirror the reference's Python-visible seams so that parity tests read like the
reference's own tests:

    OracleBox            ~ AtomBoxCubic / AtomBoxMonoclinic / AtomBoxWater*  (PBCHelper.pyx:25-351)
    topology_bruteforce  ~ NeighborTopology.get_topology_bruteforce          (topology.py:55-72)
    verlet_generator     ~ NeighborTopology.topology_verlet_list_generator   (topology.py:80-114)
    rates                ~ Fermi / FermiAngle (+ legacy AE / Exponential)    (jumprate_generators.py)
    kmc_replay           ~ KMCLattice.continuous_output in replay mode       (MDMC.py:77-171)
    observables          ~ KMCLattice.observables_output                     (MDMC.py:179-208)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cmdlmc_oracle.c")
LIB = os.path.join(HERE, "libcmdlmc_oracle.so")

RATE_KINDS = {"Fermi": 0, "FermiAngle": 1, "ActivationEnergy": 2, "Exponential": 3}

_lib = None


def build(force=False):
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
               "-fvisibility=hidden", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
            build()
        L = C.CDLL(LIB)
        dp, ip, lp, vp = (C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long),
                          C.c_void_p)
        L.orc_box_size.restype = C.c_int
        L.orc_box_init.argtypes = [vp, C.c_int, dp, dp, dp, C.c_int, dp]
        L.orc_length.argtypes = [vp, dp, dp, C.c_long, dp]
        L.orc_distance.argtypes = [vp, dp, dp, C.c_long, dp]
        L.orc_length_all_to_all.argtypes = [vp, dp, C.c_long, dp, C.c_long, dp]
        L.orc_angle.argtypes = [vp, dp, dp, dp, C.c_long, dp]
        L.orc_next_neighbor.argtypes = [vp, dp, dp, C.c_long, ip, dp]
        L.orc_topology_bruteforce.argtypes = [vp, dp, C.c_long, C.c_double, ip, ip, dp, C.c_long]
        L.orc_topology_bruteforce.restype = C.c_long
        L.orc_pairs_refresh.argtypes = [vp, dp, ip, ip, C.c_long, dp]
        L.orc_verlet_step.argtypes = [vp, dp, dp, C.c_long, C.c_double, dp]
        L.orc_verlet_step.restype = C.c_int
        L.orc_rates.argtypes = [C.c_int, dp, dp, dp, C.c_long, dp]
        L.orc_np_sum.argtypes = [dp, C.c_long]
        L.orc_np_sum.restype = C.c_double
        L.orc_kmc_replay.argtypes = [lp, ip, ip, dp, C.c_long, ip, C.c_long, C.c_double, dp,
                                     C.c_long, lp, lp, dp, ip, ip, ip, lp, ip]
        L.orc_kmc_replay.restype = C.c_long
        L.orc_fastforward.argtypes = [dp, C.c_long, C.c_int, C.c_double, dp, C.c_long, dp]
        L.orc_fastforward.restype = C.c_long
        L.orc_proton_positions.argtypes = [dp, ip, C.c_long, dp]
        L.orc_msd_update.argtypes = [vp, dp, dp, dp, ip, C.c_long, C.c_long]
        L.orc_autocorr.argtypes = [ip, ip, C.c_long]
        L.orc_autocorr.restype = C.c_long
        L.orc_mt_size.restype = C.c_int
        L.orc_mt_seed.argtypes = [vp, C.c_uint32]
        L.orc_mt_u32.argtypes = [vp]
        L.orc_mt_u32.restype = C.c_uint32
        L.orc_mt_double53.argtypes = [vp]
        L.orc_mt_double53.restype = C.c_double
        L.orc_gsl_uniform.argtypes = [vp]
        L.orc_gsl_uniform.restype = C.c_double
        L.orc_gsl_uniform_int.argtypes = [vp, C.c_uint32]
        L.orc_gsl_uniform_int.restype = C.c_uint32
        L.orc_lmc_sweep.argtypes = [ip, ip, dp, C.c_long, ip, ip, dp, C.c_long, lp]
        L.orc_lmc_sweep.restype = C.c_long
        L.orc_bench_frames.argtypes = [vp, dp, C.c_long, C.c_long, C.c_double, C.c_int, dp, dp]
        L.orc_bench_frames.restype = C.c_long
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, typ=C.c_double):
    return a.ctypes.data_as(C.POINTER(typ)) if a is not None else None


class OracleBox:
    """AtomBox restatement.  `periodic_boundaries`: 3 values -> AtomBoxCubic semantics,
    9 values -> AtomBoxMonoclinic semantics (PBCHelper.pyx:216-226, 248-260)."""

    def __init__(self, periodic_boundaries, box_multiplier=(1, 1, 1), conversion=None):
        pb = np.array(periodic_boundaries, dtype=float).ravel()
        self.periodic_boundaries = pb
        self.box_multiplier = np.array(box_multiplier, dtype=np.int32)
        self.kind = 0 if pb.size == 3 else 1
        if self.kind == 0:
            self.pbc_matrix = np.diag(pb)
            self.periodic_boundaries_extended = pb * self.box_multiplier
            self.h = self.h_inv = None
        else:
            ext = pb.copy()
            for i in range(3):
                for j in range(3):
                    ext[3 * i + j] *= self.box_multiplier[i]
            self.periodic_boundaries_extended = ext
            self.h = np.ascontiguousarray(ext.reshape(3, 3).T)
            self.h_inv = np.array(np.linalg.inv(self.h), order="C")
            self.pbc_matrix = pb.reshape(3, 3)
        conv, conv5 = 0, None
        if conversion is not None:
            if "d0" in conversion:
                conv = 2
                conv5 = _d([conversion["a"], conversion["b"], conversion["d0"],
                            conversion["left_bound"], conversion["right_bound"]])
            else:
                conv = 1
                conv5 = _d([conversion["a"], conversion["b"], 0.0,
                            conversion["left_bound"], conversion["right_bound"]])
        self._buf = C.create_string_buffer(lib().orc_box_size())
        lib().orc_box_init(self._buf, self.kind,
                           _p(_d(self.periodic_boundaries_extended[:3])),
                           _p(self.h) if self.h is not None else None,
                           _p(self.h_inv) if self.h_inv is not None else None, conv,
                           _p(conv5) if conv5 is not None else None)

    @property
    def handle(self):
        return self._buf

    def length(self, arr1, arr2):
        a = _d(np.asarray(arr1, dtype=float).reshape(-1, 3))
        b = _d(np.asarray(arr2, dtype=float).reshape(-1, 3))
        out = np.zeros(a.shape[0])
        lib().orc_length(self._buf, _p(a), _p(b), a.shape[0], _p(out))
        return out

    def distance(self, arr1, arr2):
        a = _d(np.asarray(arr1, dtype=float).reshape(-1, 3))
        b = _d(np.asarray(arr2, dtype=float).reshape(-1, 3))
        out = np.zeros(a.shape)
        lib().orc_distance(self._buf, _p(a), _p(b), a.shape[0], _p(out))
        return np.squeeze(out)

    def length_all_to_all(self, arr1, arr2):
        a, b = _d(arr1), _d(arr2)
        out = np.zeros((a.shape[0], b.shape[0]))
        lib().orc_length_all_to_all(self._buf, _p(a), a.shape[0], _p(b), b.shape[0], _p(out))
        return out

    def angle(self, a1, a2, a3):
        a1 = _d(np.asarray(a1, dtype=float).reshape(-1, 3))
        a2 = _d(np.asarray(a2, dtype=float).reshape(-1, 3))
        a3 = _d(np.asarray(a3, dtype=float).reshape(-1, 3))
        out = np.zeros(a1.shape[0])
        lib().orc_angle(self._buf, _p(a1), _p(a2), _p(a3), a1.shape[0], _p(out))
        return out if out.size > 1 else float(out[0])

    def next_neighbor(self, pos, frame):
        pos, frame = _d(pos), _d(frame)
        idx, dist = C.c_int(), C.c_double()
        lib().orc_next_neighbor(self._buf, _p(pos), _p(frame), frame.shape[0], C.byref(idx),
                                C.byref(dist))
        return idx.value, dist.value

    def position_extended_box(self, index, frame):
        """PBCHelper.pyx:34-53: atom = index % n, image = index // n, (i, j, k) with k fastest."""
        frame = _d(frame)
        n = frame.shape[0]
        atom, box = index % n, index // n
        m = self.box_multiplier
        i, j, k = box // (m[1] * m[2]), (box // m[2]) % m[1], box % m[2]
        pm = self.pbc_matrix
        return frame[atom] + i * pm[0] + j * pm[1] + k * pm[2]


def topology_bruteforce(box, frame, cutoff, buffer=0.0):
    """topology.py:55-72.  Returns (row i32[P], col i32[P], dist f64[P])."""
    frame = _d(frame)
    n = frame.shape[0]
    rc = cutoff + buffer
    cap = max(64, 64 * n)
    while True:
        row, col = np.empty(cap, np.int32), np.empty(cap, np.int32)
        dist = np.empty(cap)
        p = lib().orc_topology_bruteforce(box.handle, _p(frame), n, rc, _p(row, C.c_int),
                                          _p(col, C.c_int), _p(dist), cap)
        if p >= 0:
            return row[:p].copy(), col[:p].copy(), dist[:p].copy()
        cap = -p


def pairs_refresh(box, frame, row, col):
    """topology.py:110."""
    frame, row, col = _d(frame), _i(row), _i(col)
    out = np.empty(row.shape[0])
    lib().orc_pairs_refresh(box.handle, _p(frame), _p(row, C.c_int), _p(col, C.c_int),
                            row.shape[0], _p(out))
    return out


def verlet_generator(box, frames, cutoff, buffer):
    """topology.py:80-114.  `frames`: iterable of f64[N,3].  Yields (row, col, dist, rebuilt)."""
    last = None
    displacement = None
    topo = None
    for frame in frames:
        frame = _d(frame)
        n = frame.shape[0]
        first = last is None
        if first:
            topo = topology_bruteforce(box, frame, cutoff, buffer)
            displacement = np.zeros(n)
        rebuild = lib().orc_verlet_step(box.handle, _p(last) if not first else None, _p(frame), n,
                                        float(buffer), _p(displacement))
        if rebuild:
            topo = topology_bruteforce(box, frame, cutoff, buffer)
        else:
            topo = (topo[0], topo[1], pairs_refresh(box, frame, topo[0], topo[1]))
        yield topo[0], topo[1], topo[2], bool(rebuild) or first
        last = frame


def rates(kind, params, x, theta=None):
    """jumprate_generators.py:33-34, 42-43 (+ legacy AE / Exponential, parity unpinned)."""
    x = _d(x)
    par = _d(list(params) + [0.0] * (8 - len(params)))
    th = _d(theta) if theta is not None else None
    out = np.empty(x.shape[0])
    lib().orc_rates(RATE_KINDS[kind] if isinstance(kind, str) else kind, _p(par), _p(x),
                    _p(th) if th is not None else None, x.shape[0], _p(out))
    return out


def np_sum(a):
    a = _d(a)
    return lib().orc_np_sum(_p(a), a.shape[0])


def kmc_replay(fptr, start, dest, omega, lattice, dt, u, max_events, trace_lattice=False):
    """MDMC.py:77-171 in replay mode.  lattice (i32[nsites]) is modified in place.
    Returns dict of event arrays (+ frame_event, lattice_trace)."""
    fptr = np.ascontiguousarray(fptr, dtype=np.int64)
    start, dest, omega = _i(start), _i(dest), _d(omega)
    u = time_selectors(u)
    assert lattice.dtype == np.int32 and lattice.flags.c_contiguous
    nframes = fptr.shape[0] - 1
    nsites = lattice.shape[0]
    max_events = int(min(max_events, u.shape[0] // 2))
    ev_frame = np.zeros(max_events, np.int64)
    ev_dframe = np.zeros(max_events, np.int64)
    ev_time = np.zeros(max_events)
    ev_start = np.zeros(max_events, np.int32)
    ev_dest = np.zeros(max_events, np.int32)
    ev_proton = np.zeros(max_events, np.int32)
    frame_event = np.zeros(nframes, np.int64)
    trace = np.zeros((max_events, nsites), np.int32) if trace_lattice else None
    nev = lib().orc_kmc_replay(_p(fptr, C.c_long), _p(start, C.c_int), _p(dest, C.c_int),
                               _p(omega), nframes, _p(lattice, C.c_int), nsites, float(dt),
                               _p(u), max_events, _p(ev_frame, C.c_long),
                               _p(ev_dframe, C.c_long), _p(ev_time), _p(ev_start, C.c_int),
                               _p(ev_dest, C.c_int), _p(ev_proton, C.c_int),
                               _p(frame_event, C.c_long),
                               _p(trace, C.c_int) if trace is not None else None)
    out = dict(n_events=nev, frame=ev_frame[:nev], dframe=ev_dframe[:nev], time=ev_time[:nev],
               start=ev_start[:nev], dest=ev_dest[:nev], proton=ev_proton[:nev],
               frame_event=frame_event)
    if trace is not None:
        out["lattice_trace"] = trace[:nev]
    return out


def time_selectors(u):
    """Uniform stream (random(), uniform-u, random(), ...) -> stream whose even entries are the
    reference's time selectors -np.log(1 - u) (MDMC.py:148), evaluated by NumPy like upstream."""
    u = np.array(u, dtype=np.float64)
    u[0::2] = -np.log(1 - u[0::2])
    return u


def fastforward(rates, dt, u, n_events, cycle=True):
    """MDMC.py:121-171 on a stream of per-frame total rates.  Returns f64[n,3] rows of
    (sweep, delta_frame, kmc_time)."""
    rates, u = _d(rates), -np.log(1 - _d(u))
    rows = np.zeros((n_events, 3))
    n = lib().orc_fastforward(_p(rates), rates.shape[0], int(cycle), float(dt), _p(u), n_events,
                              _p(rows))
    return rows[:n]


def init_lattice(lattice_size, proton_number, rng):
    """MDMC.py:68-72 with an explicit legacy RandomState instead of the global one."""
    lattice = np.zeros(lattice_size, dtype=np.int32)
    lattice[:proton_number] = range(1, proton_number + 1)
    rng.shuffle(lattice)
    return lattice


def observables(box, positions, lattice0, events, reset_frequency, print_frequency):
    """MDMC.py:179-208 + output.py on a finished replay.  `positions` f64[F,N,3] donor-site
    coordinates; `events` = kmc_replay() result; lattice0 = lattice before the first event.
    Returns list of (frame_number, time, msd[3], autocorr)."""
    positions = _d(positions)
    nsites = positions.shape[1]
    lattice = np.ascontiguousarray(lattice0, dtype=np.int32).copy()
    nprot = int((lattice > 0).sum())
    fe = events["frame_event"]
    out = []
    snapshot = np.zeros((nprot, 3))
    displacement = np.zeros((nprot, 3))
    auto0 = None
    applied = 0  # number of events already applied to `lattice`
    for f in range(positions.shape[0]):
        e = fe[f]
        if e < 0:
            break
        while applied < e:  # frames of event e see the lattice after events 0..e-1
            s, d = events["start"][applied], events["dest"][applied]
            lattice[d] = lattice[s]
            lattice[s] = 0
            applied += 1
        t = events["time"][e]
        if f == 0:
            auto0 = lattice.copy()
            lib().orc_proton_positions(_p(positions[0]), _p(lattice, C.c_int), nsites,
                                       _p(snapshot))
            continue
        if f % reset_frequency == 0:
            auto0 = lattice.copy()
            displacement[:] = 0
        lib().orc_msd_update(box.handle, _p(snapshot), _p(displacement), _p(positions[f]),
                             _p(lattice, C.c_int), nsites, nprot)
        if f % print_frequency == 0:
            auto = lib().orc_autocorr(_p(lattice, C.c_int), _p(auto0, C.c_int), nsites)
            msd = np.sum(displacement ** 2, axis=0) / displacement.shape[0]
            out.append((f, t, msd, int(auto)))
    return out


def lmc_sweep(start, dest, prob, lattice, pick, acc, jumpmatrix=None):
    """Legacy LMC sweep restatement (PARITY UNPINNED, see cmdlmc_oracle.c)."""
    start, dest, prob, pick, acc = _i(start), _i(dest), _d(prob), _i(pick), _d(acc)
    assert lattice.dtype == np.int32
    jm = None
    if jumpmatrix is not None:
        assert jumpmatrix.dtype == np.int64
        jm = _p(jumpmatrix, C.c_long)
    return lib().orc_lmc_sweep(_p(start, C.c_int), _p(dest, C.c_int), _p(prob), pick.shape[0],
                               _p(lattice, C.c_int), _p(pick, C.c_int), _p(acc),
                               lattice.shape[0], jm)


class MT19937:
    def __init__(self, seed):
        self._buf = C.create_string_buffer(lib().orc_mt_size())
        lib().orc_mt_seed(self._buf, seed)

    def u32(self):
        return lib().orc_mt_u32(self._buf)

    def double53(self):
        return lib().orc_mt_double53(self._buf)

    def gsl_uniform(self):
        return lib().orc_gsl_uniform(self._buf)

    def gsl_uniform_int(self, n):
        return lib().orc_gsl_uniform_int(self._buf, n)


def bench_frames(box, frames, rc, rate_kind, params):
    frames = _d(frames)
    par = _d(list(params) + [0.0] * (8 - len(params)))
    rs = C.c_double()
    tot = lib().orc_bench_frames(box.handle, _p(frames), frames.shape[0], frames.shape[1],
                                 float(rc), RATE_KINDS[rate_kind], _p(par), C.byref(rs))
    return tot, rs.value
