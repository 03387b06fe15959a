"""Seeded synthetic oxygen lattices and trajectories for the five BASELINE.json configs.

The reference ships no data for its integration config (tests/integration/trajectory.xyz is a
missing blob, tests/integration/mdlmc_run.py:37-70 only fixes box, atom counts and parameters),
so every workload is synthesised here.  Host-side NumPy only.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Workload:
    name: str
    cell: np.ndarray            # f64[3] (orthorhombic) or f64[9] (rows = cell vectors)
    n_oxygen: int
    n_extra: int                # heavy atoms (P) the oxygens are bonded to
    n_protons: int
    n_frames: int
    time_step: float
    cutoff: float
    buffer: float
    rate_kind: str
    rate_params: tuple
    seed: int
    group_size: int = 3
    extra: dict = field(default_factory=dict)

    @property
    def is_ortho(self):
        return np.asarray(self.cell).size == 3

    @property
    def cell_matrix(self):
        c = np.asarray(self.cell, dtype=float)
        return np.diag(c) if c.size == 3 else c.reshape(3, 3)


def _monoclinic_cell(a, b, c, beta_deg):
    beta = np.deg2rad(beta_deg)
    return np.array([a, 0, 0, 0, b, 0, c * np.cos(beta), 0, c * np.sin(beta)], dtype=float)


def _triclinic_cell(lx, ly, lz, xy, xz, yz):
    return np.array([lx, 0, 0, xy * ly, ly, 0, xz * lz, yz * lz, lz], dtype=float)


def workload(name, n_frames=None):
    """C1..C5 of SURVEY.md 8(d).  n_frames overrides the nominal trajectory length."""
    fermi = (0.06, 2.3, 0.1)
    if name == "C1":   # reference integration config: tests/integration/mdlmc_run.py:37-70
        w = Workload("C1", np.array([29.122, 25.354, 12.363]), 144, 48, 96, 2000, 0.4, 3.0, 2.0,
                     "Fermi", fermi, 0, group_size=3)
    elif name == "C2":  # CsH2PO4-like monoclinic, ~400 O (100 PO4 groups), 100k frames
        w = Workload("C2", _monoclinic_cell(23.4, 19.2, 26.6, 107.7), 400, 100, 200, 100000, 0.5,
                     3.0, 2.0, "Fermi", fermi, 1, group_size=4)
    elif name == "C3":  # triclinic phosphonic-acid-like, 2048 O, activation-energy rate
        w = Workload("C3", _triclinic_cell(40.0, 38.0, 39.0, 0.2, -0.3, 0.15), 2048, 683, 1024,
                     20000, 0.5, 3.0, 2.0, "ActivationEnergy", (0.06, 1.2, 30.0, 2.2, 510.0), 2,
                     group_size=3)
    elif name == "C4":  # 384-O lattice (C2-like cell), 1024 replicas
        w = Workload("C4", _monoclinic_cell(23.4, 19.2, 25.5, 107.7), 384, 96, 256, 10000, 0.5,
                     3.0, 2.0, "Fermi", fermi, 3, group_size=4, extra={"replicas": 1024})
    elif name == "C5":  # 32k-O water-like orthorhombic box, cell list + histogram
        w = Workload("C5", np.array([99.4, 99.4, 99.4]), 32768, 0, 1, 200, 0.5, 3.0, 2.0,
                     "Fermi", fermi, 4, group_size=0)
    else:
        raise ValueError(name)
    if n_frames is not None:
        w.n_frames = int(n_frames)
    return w


def _grid_counts(n, lengths):
    """Integer grid (nx, ny, nz) with nx*ny*nz >= n and near-cubic cells."""
    lengths = np.asarray(lengths, dtype=float)
    vol = np.prod(lengths)
    s = (vol / n) ** (1.0 / 3.0)
    counts = np.maximum(1, np.round(lengths / s).astype(int))
    while np.prod(counts) < n:
        k = int(np.argmax(lengths / counts))
        counts[k] += 1
    return counts


def initial_positions(w):
    """Returns (oxygen f64[N,3], extra f64[M,3]) Cartesian positions inside the cell."""
    rng = np.random.RandomState(w.seed)
    hm = w.cell_matrix                       # rows = cell vectors
    lengths = np.linalg.norm(hm, axis=1)
    if w.n_extra > 0:
        n_centres = w.n_extra
    else:
        n_centres = w.n_oxygen
    counts = _grid_counts(n_centres, lengths)
    gi = np.stack(np.meshgrid(*[np.arange(c) for c in counts], indexing="ij"), -1).reshape(-1, 3)
    pick = rng.permutation(gi.shape[0])[:n_centres]
    frac = (gi[np.sort(pick)] + 0.5) / counts
    centres = frac @ hm
    spacing = lengths / counts
    if w.n_extra > 0:
        centres = centres + rng.uniform(-0.12, 0.12, size=centres.shape) * spacing
        # group_size oxygens per heavy atom on a randomly rotated tripod / tetrahedron, r = 1.52
        gs = w.group_size
        base = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=float) / np.sqrt(3)
        oxy = []
        for c in centres:
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            a, b, cc, d = q
            rot = np.array([[a*a+b*b-cc*cc-d*d, 2*(b*cc-a*d), 2*(b*d+a*cc)],
                            [2*(b*cc+a*d), a*a-b*b+cc*cc-d*d, 2*(cc*d-a*b)],
                            [2*(b*d-a*cc), 2*(cc*d+a*b), a*a-b*b-cc*cc+d*d]])
            oxy.append(c + 1.52 * base[:gs] @ rot.T)
        oxy = np.concatenate(oxy)[:w.n_oxygen]
        extra = centres
    else:
        oxy = centres + rng.uniform(-0.08, 0.08, size=centres.shape) * spacing
        extra = np.zeros((0, 3))
    return np.ascontiguousarray(oxy), np.ascontiguousarray(extra)


def trajectory(w, n_frames=None, start=0, dtype=np.float64, amplitude=0.25, noise=0.01,
               with_extra=False):
    """Frames [start, start+n_frames) of the synthetic trajectory: every atom oscillates around
    its site (per-component sinusoid, amplitude 0.25 A, period 60-240 frames, random phase) plus
    white noise N(0, 0.01 A).  Deterministic in (seed, frame index): any block can be generated
    independently, which is what frame-block sharding across GPUs needs.  f64[F,N,3]."""
    n_frames = w.n_frames if n_frames is None else int(n_frames)
    oxy, extra = initial_positions(w)
    base = np.concatenate([oxy, extra]) if with_extra else oxy
    rng = np.random.RandomState(w.seed + 7919)
    period = rng.uniform(60.0, 240.0, size=base.shape)
    phase = rng.uniform(0, 2 * np.pi, size=base.shape)
    out = np.empty((n_frames,) + base.shape, dtype=dtype)
    block = 4096
    stop = start + n_frames
    for blk in range(start // block, (stop + block - 1) // block if n_frames else 0):
        g0 = blk * block                                   # absolute first frame of the noise block
        lo, hi = max(start, g0), min(stop, g0 + block)
        t = np.arange(lo, hi, dtype=float)[:, None, None]
        x = base[None] + amplitude * np.sin(2 * np.pi * t / period[None] + phase[None])
        # counter-style noise: one seeded stream per aligned block of 4096 frames, so any frame
        # range reproduces the same values as the whole trajectory sliced
        nrng = np.random.RandomState((w.seed * 1000003 + g0) % (2**31 - 1))
        x += nrng.normal(scale=noise, size=(hi - g0,) + base.shape)[lo - g0:]
        out[lo - start:hi - start] = x
    return out


def initial_lattice(n_sites, n_protons, seed):
    """Lattice as KMCLattice._initialize_lattice builds it (MDMC.py:68-72), from an explicit
    legacy RandomState(seed) instead of the global one."""
    rng = np.random.RandomState(seed)
    lattice = np.zeros(n_sites, dtype=np.int32)
    lattice[:n_protons] = np.arange(1, n_protons + 1)
    rng.shuffle(lattice)
    return lattice, rng
