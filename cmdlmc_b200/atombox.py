"""AtomBox family -- host-side mirror of the reference's Cython classes
(mdlmc/cython_exts/LMC/PBCHelper.pyx:25-351) on top of the C ABI.

Same constructor signatures, attribute names, argument conventions (array-likes of any shape
reshaped to [-1, 3] float64, PBCHelper.pyx:60-61,78-79) and return shapes.  Every method runs a
CUDA kernel through libcmdlmc_b200; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _abi, runtime
from ._abi import as_f64, check, ptr


class AtomBox:
    """Base class (PBCHelper.pyx:25-211).  Use AtomBoxCubic or AtomBoxMonoclinic."""

    _n_values = None

    def __init__(self, periodic_boundaries, *args, box_multiplier=(1, 1, 1), **kwargs):
        runtime.ensure_init()
        pb = np.array(periodic_boundaries, dtype=float).ravel()
        if self._n_values is not None and pb.size != self._n_values:
            raise ValueError("%s needs %d periodic boundary values, got %d"
                             % (type(self).__name__, self._n_values, pb.size))
        self.periodic_boundaries = pb
        self.box_multiplier = np.array(box_multiplier, dtype=np.int32)
        handle = C.c_void_p()
        check(_abi.lib().cmd_box_create(ptr(pb), pb.size, ptr(self.box_multiplier, C.c_int),
                                        C.byref(handle)))
        self._handle = handle
        ext = np.zeros(pb.size)
        pbcm, h = np.zeros(9), np.zeros(9)
        check(_abi.lib().cmd_box_query(handle, ptr(ext), ptr(pbcm), ptr(h), None))
        self.periodic_boundaries_extended = ext
        self.pbc_matrix = pbcm.reshape(3, 3)
        self._h = h.reshape(3, 3)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _abi.lib().cmd_box_destroy(h)
            except Exception:
                pass
            self._handle = None

    @property
    def handle(self):
        return self._handle

    # -- PBCHelper.pyx:34-53
    def position_extended_box(self, index, frame):
        frame = as_f64(frame)
        out = np.zeros(3)
        check(_abi.lib().cmd_position_extended_box(self._handle, int(index), ptr(frame),
                                                   frame.shape[0], ptr(out)))
        return out

    # -- PBCHelper.pyx:56-70
    def distance(self, arr1, arr2):
        """Calculates for two arrays of positions an array of vector distances"""
        a = as_f64(np.asarray(arr1, dtype=float).reshape((-1, 3)))
        b = as_f64(np.asarray(arr2, dtype=float).reshape((-1, 3)))
        if a.shape != b.shape:
            raise ValueError("arr1 and arr2 must hold the same number of positions")
        out = np.zeros(a.shape)
        check(_abi.lib().cmd_distance(self._handle, ptr(a), ptr(b), a.shape[0], ptr(out)))
        return np.squeeze(out)

    # -- PBCHelper.pyx:74-85
    def length(self, arr1, arr2):
        """Calculates for two arrays of positions an array of scalar distances"""
        a = as_f64(np.asarray(arr1, dtype=float).reshape((-1, 3)))
        b = as_f64(np.asarray(arr2, dtype=float).reshape((-1, 3)))
        if a.shape != b.shape:
            raise ValueError("arr1 and arr2 must hold the same number of positions")
        out = np.zeros(a.shape[0])
        check(_abi.lib().cmd_length(self._handle, ptr(a), ptr(b), a.shape[0], ptr(out)))
        return out

    # -- PBCHelper.pyx:88-95
    def length_all_to_all(self, arr1, arr2):
        a, b = as_f64(arr1), as_f64(arr2)
        if a.ndim != 2 or b.ndim != 2 or a.shape[1] != 3 or b.shape[1] != 3:
            raise ValueError("length_all_to_all expects [n, 3] and [m, 3] arrays")
        out = np.zeros((a.shape[0], b.shape[0]))
        check(_abi.lib().cmd_length_all_to_all(self._handle, ptr(a), a.shape[0], ptr(b),
                                               b.shape[0], ptr(out)))
        return out

    # -- PBCHelper.pyx:133-134 (vertex is atompos_2); batched inputs are an extension
    def angle(self, atompos_1, atompos_2, atompos_3):
        a1 = as_f64(np.asarray(atompos_1, dtype=float).reshape((-1, 3)))
        a2 = as_f64(np.asarray(atompos_2, dtype=float).reshape((-1, 3)))
        a3 = as_f64(np.asarray(atompos_3, dtype=float).reshape((-1, 3)))
        out = np.zeros(a1.shape[0])
        check(_abi.lib().cmd_angle(self._handle, ptr(a1), ptr(a2), ptr(a3), a1.shape[0], ptr(out)))
        return float(out[0]) if np.ndim(atompos_1) == 1 else out

    # -- PBCHelper.pyx:153-167
    def next_neighbor(self, pos, frame_2):
        pos, frame = as_f64(pos), as_f64(frame_2)
        idx, dist = C.c_int(-1), C.c_double(0)
        check(_abi.lib().cmd_next_neighbor(self._handle, ptr(pos), ptr(frame), frame.shape[0],
                                           C.byref(idx), C.byref(dist)))
        return idx.value, dist.value

    # -- PBCHelper.pyx:169-185
    def next_neighbor_extended_box(self, index_1, frame_1, frame_2):
        f1, f2 = as_f64(frame_1), as_f64(frame_2)
        idx, dist = C.c_int(-1), C.c_double(0)
        check(_abi.lib().cmd_next_neighbor_extended_box(self._handle, int(index_1), ptr(f1),
                                                        f1.shape[0], ptr(f2), f2.shape[0],
                                                        C.byref(idx), C.byref(dist)))
        return idx.value, dist.value

    def determine_phosphorus_oxygen_pairs(self, oxygen_atoms, phosphorus_atoms):
        """int32 [oxygens of the extended box]: the phosphorus of the extended box each of them is
        closest to (same result as PBCHelper.pyx:187-196)."""
        oxygens = as_f64(oxygen_atoms)
        count = oxygens.shape[0] * int(np.prod(self.box_multiplier))
        nearest = (self.next_neighbor_extended_box(k, oxygens, phosphorus_atoms)[0] for k in range(count))
        return np.fromiter(nearest, dtype=np.int32, count=count)

    def get_acidic_proton_indices(self, atoms, verbose=False):
        """Indices (into `atoms`, a record array with 'name' and 'pos') of the hydrogens whose
        nearest heavy atom is an oxygen (same result as PBCHelper.pyx:198-211), from one
        all-to-all distance matrix on the device."""
        is_h = atoms["name"] == "H"
        heavy = atoms[~is_h]
        if not is_h.any() or heavy.size == 0:
            return []
        dist = self.length_all_to_all(as_f64(atoms["pos"][is_h]), as_f64(heavy["pos"]))
        bonded_to_oxygen = heavy["name"][np.argmin(dist, axis=1)] == "O"
        acidic = [int(i) for i in np.flatnonzero(is_h)[bonded_to_oxygen]]
        if verbose:
            print("# %d acidic protons: %s" % (len(acidic), acidic))
        return acidic


class AtomBoxCubic(AtomBox):
    """Subclass of AtomBox for orthogonal periodic MD boxes (PBCHelper.pyx:213-239)"""
    _n_values = 3


class AtomBoxMonoclinic(AtomBox):
    """Subclass of AtomBox for nonorthogonal periodic MD boxes (PBCHelper.pyx:242-275)"""
    _n_values = 9

    def __init__(self, periodic_boundaries, *args, box_multiplier=(1, 1, 1), **kwargs):
        super().__init__(periodic_boundaries, *args, box_multiplier=box_multiplier, **kwargs)
        self.h = np.ascontiguousarray(self._h)
        # the reference inverts with np.linalg.inv (PBCHelper.pyx:259): hand the library the very
        # same matrix so device results are bit-identical with the reference's arithmetic
        self.h_inv = np.array(np.linalg.inv(self.h), order="C")
        check(_abi.lib().cmd_box_set_hinv(self._handle, ptr(self.h_inv)))
        self.pbc_matrix = self.periodic_boundaries.reshape((3, 3))


class AtomBoxWater(AtomBoxCubic):
    """Converts oxygen-oxygen distances to typical hydronium-oxygen distances
    (PBCHelper.pyx:278-303); the base class converts nothing."""


class AtomBoxWaterLinearConversion(AtomBoxWater):
    """PBCHelper.pyx:306-324"""

    def __init__(self, periodic_boundaries, *args, box_multiplier=(1, 1, 1), **kwargs):
        super().__init__(periodic_boundaries, box_multiplier=box_multiplier)
        p = args[0]
        par = as_f64([p["a"], p["b"], 0.0, p["left_bound"], p["right_bound"]])
        check(_abi.lib().cmd_box_set_conversion(self._handle, 1, ptr(par)))


class AtomBoxWaterRampConversion(AtomBoxWater):
    """PBCHelper.pyx:327-351"""

    def __init__(self, periodic_boundaries, *args, box_multiplier=(1, 1, 1), **kwargs):
        super().__init__(periodic_boundaries, box_multiplier=box_multiplier)
        p = args[0]
        par = as_f64([p["a"], p["b"], p["d0"], p["left_bound"], p["right_bound"]])
        check(_abi.lib().cmd_box_set_conversion(self._handle, 2, ptr(par)))
