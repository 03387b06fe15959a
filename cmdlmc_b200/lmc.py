"""Legacy LMC sweep engine -- host side of csrc/lmc.cu.

PARITY UNPINNED upstream: the Cython LMCHelper / LMCRoutine of the reference is not in the tree;
the semantics follow mdlmc/IO/config_parser.py:182-189 and the .bak tests.  `DeviceLMC` runs
n_replicas independent lattices over the frames of a topology block (one sweep = P attempts per
frame); `gsl_streams` produces the replay numbers a GSL-driven run would consume
(gsl_rng_mt19937: gsl_rng_uniform_int(P) for the pair, gsl_rng_uniform for the acceptance)."""
import ctypes as C

import numpy as np

from . import _abi, runtime
from ._abi import check, ptr

RNG_REPLAY = 0
RNG_PHILOX = 1


class DeviceLMC:
    def __init__(self, lattices, rng_mode=RNG_PHILOX, seed=0):
        runtime.ensure_init()
        lattices = np.ascontiguousarray(np.atleast_2d(lattices), dtype=np.int32)
        self.n_replicas, self.n_sites = lattices.shape
        self._handle = C.c_void_p()
        check(_abi.lib().cmd_lmc_create(self.n_sites, self.n_replicas, ptr(lattices, C.c_int),
                                        int(rng_mode), int(seed) & (2 ** 64 - 1),
                                        C.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _abi.lib().cmd_lmc_destroy(h)
            except Exception:
                pass
            self._handle = None

    def set_replay_stream(self, pick, acc):
        pick = np.ascontiguousarray(np.atleast_2d(pick), dtype=np.int32)
        acc = np.ascontiguousarray(np.atleast_2d(acc), dtype=np.float64)
        if pick.shape != acc.shape or pick.shape[0] != self.n_replicas:
            raise ValueError("pick and acc must both be [n_replicas, n_attempts]")
        check(_abi.lib().cmd_lmc_set_replay_stream(self._handle, ptr(pick, C.c_int), ptr(acc),
                                                   pick.shape[1]))

    def enable_jump_matrix(self, enable=True):
        check(_abi.lib().cmd_lmc_enable_jump_matrix(self._handle, int(bool(enable))))

    def advance(self, topo, prob_scale, sweeps_per_frame=1):
        """sweeps_per_frame sweeps on every frame of topo's current block; a hop is accepted with
        probability omega * prob_scale (prob_scale = the MD time step)."""
        check(_abi.lib().cmd_lmc_advance(self._handle, topo.handle, float(prob_scale),
                                         int(sweeps_per_frame)))

    def state(self):
        r = self.n_replicas
        lattices = np.zeros((r, self.n_sites), np.int32)
        jumps, attempts, sweeps = (np.zeros(r, np.int64) for _ in range(3))
        halted = np.zeros(r, np.int32)
        check(_abi.lib().cmd_lmc_get_state(self._handle, ptr(lattices, C.c_int),
                                           ptr(jumps, C.c_int64), ptr(attempts, C.c_int64),
                                           ptr(sweeps, C.c_int64), ptr(halted, C.c_int)))
        return dict(lattices=lattices, jumps=jumps, attempts=attempts, sweeps=sweeps,
                    halted=halted.astype(bool))

    def jump_matrix(self):
        m = np.zeros((self.n_sites, self.n_sites), np.int64)
        check(_abi.lib().cmd_lmc_get_jump_matrix(self._handle, ptr(m, C.c_int64)))
        return m


def mt19937_u32(seed, n):
    """n raw 32-bit outputs of MT19937 seeded with init_genrand(seed) -- the core generator of
    both GSL's gsl_rng_mt19937 and NumPy's legacy RandomState (host-side, NumPy only)."""
    rs = np.random.RandomState()
    key = np.empty(624, dtype=np.uint32)
    key[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        key[i] = (1812433253 * (int(key[i - 1]) ^ (int(key[i - 1]) >> 30)) + i) & 0xFFFFFFFF
    rs.set_state(("MT19937", key, 624))
    return rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)


def gsl_streams(seed, counts, sweeps_per_frame=1):
    """Replay streams of one replica for frames with `counts[f]` listed pairs: per attempt
    gsl_rng_uniform_int(P) (rejection on u32 / (0xffffffff // P)) then gsl_rng_uniform
    (u32 / 2^32), drawn from one MT19937 sequence like a GSL-driven sweep would."""
    counts = np.asarray(counts, dtype=np.int64)
    total = int(counts.sum()) * int(sweeps_per_frame)
    raw = mt19937_u32(seed, 2 * total + total // 64 + 4096).astype(np.int64)
    pick = np.empty(total, np.int32)
    acc = np.empty(total, np.float64)
    pos = out = 0
    for p in np.repeat(counts, sweeps_per_frame):
        p = int(p)
        if p == 0:
            continue
        scale = 0xFFFFFFFF // p
        seg = raw[pos:pos + 2 * p]
        k = seg[0::2] // scale
        if seg.size == 2 * p and (k < p).all():      # no rejection in this sweep (the usual case)
            pick[out:out + p] = k
            acc[out:out + p] = seg[1::2] / 4294967296.0
            pos += 2 * p
            out += p
            continue
        for _ in range(p):                            # a rejected draw shifts the sequence
            while True:
                kk = int(raw[pos]) // scale
                pos += 1
                if kk < p:
                    break
            pick[out] = kk
            acc[out] = int(raw[pos]) / 4294967296.0
            pos += 1
            out += 1
    return pick, acc
