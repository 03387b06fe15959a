"""ctypes binding of libcmdlmc_b200.so (the C ABI declared in include/cmdlmc_b200.h).

The product path is CUDA-only: if the shared library has not been built, or no CUDA device can
be bound, every entry point raises -- there is no CPU fallback and nothing here imports the
oracle.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# CMDLMC_B200_LIB lets a developer A/B-test another build of the same library
LIB_PATH = os.environ.get("CMDLMC_B200_LIB") or os.path.join(HERE, "libcmdlmc_b200.so")


class CmdError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libcmdlmc_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/cmdlmc_b200.h declares
SIGNATURES = {
    "cmd_abi_version": (C.c_int, []),
    "cmd_last_error": (C.c_char_p, []),
    "cmd_init": (C.c_int, [C.c_int]),
    "cmd_shutdown": (C.c_int, []),
    "cmd_device_count": (C.c_int, [ip]),
    "cmd_set_stream": (C.c_int, [vp]),
    "cmd_sync": (C.c_int, []),
    "cmd_launch_count": (C.c_int64, []),
    "cmd_fp64_peak": (C.c_int, [C.c_int, dp]),
    "cmd_smem_peak": (C.c_int, [C.c_int, dp]),
    "cmd_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "cmd_host_free": (C.c_int, [vp]),
    "cmd_staging_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), ip]),
    "cmd_box_create": (C.c_int, [dp, C.c_int, ip, C.POINTER(vp)]),
    "cmd_box_set_hinv": (C.c_int, [vp, dp]),
    "cmd_box_n_images": (C.c_int, [vp]),
    "cmd_box_set_conversion": (C.c_int, [vp, C.c_int, dp]),
    "cmd_box_query": (C.c_int, [vp, dp, dp, dp, dp]),
    "cmd_box_destroy": (None, [vp]),
    "cmd_length": (C.c_int, [vp, dp, dp, C.c_int64, dp]),
    "cmd_distance": (C.c_int, [vp, dp, dp, C.c_int64, dp]),
    "cmd_length_all_to_all": (C.c_int, [vp, dp, C.c_int64, dp, C.c_int64, dp]),
    "cmd_angle": (C.c_int, [vp, dp, dp, dp, C.c_int64, dp]),
    "cmd_next_neighbor": (C.c_int, [vp, dp, dp, C.c_int64, ip, dp]),
    "cmd_position_extended_box": (C.c_int, [vp, C.c_int, dp, C.c_int, dp]),
    "cmd_next_neighbor_extended_box": (C.c_int, [vp, C.c_int, dp, C.c_int, dp, C.c_int, ip, dp]),
    "cmd_length_dev": (C.c_int, [vp, vp, vp, C.c_int64, vp]),
    "cmd_distance_dev": (C.c_int, [vp, vp, vp, C.c_int64, vp]),
    "cmd_length_all_to_all_dev": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int64, vp]),
    "cmd_angle_dev": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp]),
    "cmd_rates": (C.c_int, [C.c_int, dp, dp, dp, C.c_int64, dp]),
    "cmd_rates_dev": (C.c_int, [C.c_int, dp, vp, vp, C.c_int64, vp]),
    "cmd_topo_create": (C.c_int, [vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp,
                                  C.c_int64, C.POINTER(vp)]),
    "cmd_topo_destroy": (None, [vp]),
    "cmd_topo_set_path": (C.c_int, [vp, C.c_int]),
    "cmd_topo_path": (C.c_int, [vp]),
    "cmd_topo_build_dev": (C.c_int, [vp, vp, C.c_int64]),
    "cmd_topo_build": (C.c_int, [vp, vp, C.c_int, C.c_int64]),
    "cmd_topo_skip_dev": (C.c_int, [vp, vp, C.c_int64]),
    "cmd_topo_skip": (C.c_int, [vp, vp, C.c_int, C.c_int64]),
    "cmd_topo_frame_info": (C.c_int, [vp, lp, u8p, dp]),
    "cmd_topo_stride": (C.c_int64, [vp]),
    "cmd_topo_capacity_needed": (C.c_int64, [vp]),
    "cmd_topo_skin_stats": (C.c_int, [vp, lp, lp, lp]),
    "cmd_topo_n_images": (C.c_int, [vp]),
    "cmd_topo_nframes": (C.c_int64, [vp]),
    "cmd_topo_get_frame": (C.c_int, [vp, C.c_int64, ip, ip, dp, dp]),
    "cmd_topo_set_selection": (C.c_int, [vp, C.c_int, ip]),
    "cmd_topo_build_async": (C.c_int, [vp, vp, C.c_int, C.c_int64]),
    "cmd_topo_wait": (C.c_int, [vp]),
    "cmd_topo_get_block": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int64, ip, ip, dp, dp]),
    "cmd_topo_device_arrays": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                         C.POINTER(vp), C.POINTER(vp)]),
    "cmd_topo_tie_count": (C.c_int64, [vp]),
    "cmd_topo_nearest": (C.c_int, [vp, C.c_int]),
    "cmd_topo_near_arrays": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), lp]),
    "cmd_topo_get_frame_nearest": (C.c_int, [vp, C.c_int64, ip, dp]),
    "cmd_topo_set_groups": (C.c_int, [vp, ip, C.c_int]),
    "cmd_topo_apply_angles": (C.c_int, [vp, vp, C.c_int]),
    "cmd_topo_apply_angles_dev": (C.c_int, [vp, vp]),
    "cmd_topo_get_frame_angles": (C.c_int, [vp, C.c_int64, dp]),
    "cmd_topo_distance_histogram": (C.c_int, [vp, C.c_double, C.c_double, C.c_int, lp]),
    "cmd_topo_distance_histogram_dev": (C.c_int, [vp, C.c_double, C.c_double, C.c_int, vp]),
    "cmd_kmc_get_event_distances": (C.c_int, [vp, C.c_int, C.c_int64, lp, dp]),
    "cmd_kmc_jump_histogram": (C.c_int, [vp, C.c_double, C.c_double, C.c_int, lp]),
    "cmd_kmc_jump_histogram_dev": (C.c_int, [vp, C.c_double, C.c_double, C.c_int, vp]),
    "cmd_topo_row_offsets": (C.c_int, [vp, C.POINTER(vp)]),
    "cmd_topo_n_atoms": (C.c_int, [vp]),
    "cmd_topo_positions": (C.c_int, [vp, C.POINTER(vp)]),
    "cmd_kmc_create": (C.c_int, [vp, C.c_int, C.c_int, ip, C.c_double, C.c_int, C.c_uint64,
                                 C.POINTER(vp)]),
    "cmd_kmc_destroy": (None, [vp]),
    "cmd_kmc_set_replica_ids": (C.c_int, [vp, C.c_int, C.c_int]),
    "cmd_kmc_set_hydronium": (C.c_int, [vp, C.c_int, dp, C.c_int, dp, dp, dp, C.c_int, C.c_double,
                                        C.c_double]),
    "cmd_kmc_get_last_jump_times": (C.c_int, [vp, dp]),
    "cmd_kmc_enable_occupancy": (C.c_int, [vp]),
    "cmd_kmc_get_occupancy": (C.c_int, [vp, lp, lp]),
    "cmd_kmc_set_replay_stream": (C.c_int, [vp, dp, C.c_int64]),
    "cmd_kmc_set_event_log": (C.c_int, [vp, C.c_int64]),
    "cmd_kmc_set_observables": (C.c_int, [vp, C.c_int, C.c_int]),
    "cmd_kmc_seed_observables": (C.c_int, [vp, dp]),
    "cmd_kmc_advance": (C.c_int, [vp, vp, vp]),
    "cmd_kmc_get_state": (C.c_int, [vp, ip, dp, lp, lp, lp]),
    "cmd_kmc_get_status": (C.c_int, [vp, ip, ip, lp]),
    "cmd_kmc_get_events": (C.c_int, [vp, C.c_int, C.c_int64, lp, lp, dp, ip, ip, ip]),
    "cmd_kmc_get_observables": (C.c_int, [vp, C.c_int, C.c_int64, lp, dp]),
    "cmd_kmc_get_observable_rows": (C.c_int, [vp, C.c_int, C.c_int64, C.c_int64, dp]),
    "cmd_kmc_tie_count": (C.c_int64, [vp]),
    "cmd_kmc_events_dropped": (C.c_int64, [vp]),
    "cmd_kmc_selection_fallbacks": (C.c_int64, [vp]),
    "cmd_kmc_debug_counter": (C.c_int64, [vp, C.c_int]),
    "cmd_lmc_create": (C.c_int, [C.c_int, C.c_int, ip, C.c_int, C.c_uint64, C.POINTER(vp)]),
    "cmd_lmc_destroy": (None, [vp]),
    "cmd_lmc_set_replay_stream": (C.c_int, [vp, ip, dp, C.c_int64]),
    "cmd_lmc_enable_jump_matrix": (C.c_int, [vp, C.c_int]),
    "cmd_lmc_advance": (C.c_int, [vp, vp, C.c_double, C.c_int]),
    "cmd_lmc_get_state": (C.c_int, [vp, ip, lp, lp, lp, ip]),
    "cmd_lmc_get_jump_matrix": (C.c_int, [vp, lp]),
    "cmd_comm_unique_id": (C.c_int, [u8p]),
    "cmd_comm_init": (C.c_int, [C.c_int, C.c_int, u8p]),
    "cmd_comm_destroy": (C.c_int, []),
    "cmd_comm_rank": (C.c_int, []),
    "cmd_comm_world": (C.c_int, []),
    "cmd_comm_nccl_version": (C.c_int, []),
    "cmd_stats_allreduce": (C.c_int, [dp, C.c_int64, lp, C.c_int64]),
    "cmd_stats_allreduce_dev": (C.c_int, [vp, C.c_int64, vp, C.c_int64]),
    "cmd_allgather_dev": (C.c_int, [vp, vp, C.c_int64]),
    "cmd_topo_block_stats_dev": (C.c_int, [vp, vp]),
    "cmd_topo_dr_dev": (C.c_int, [vp, vp, C.c_int64, vp, vp]),
    "cmd_topo_skip_dr_dev": (C.c_int, [vp, vp, C.c_int64, lp]),
    "cmd_topo_seed_dev": (C.c_int, [vp, vp, vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m cmdlmc_b200.build` (nvcc, sm_100a). "
                "cmdlmc_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if "CMDLMC_B200_LIB" in os.environ and not hasattr(L, name):
                continue   # A/B test against an older build: newer entry points are simply absent
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise CmdError(code, lib().cmd_last_error().decode("utf-8", "replace"))
    return code


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a, typ=C.c_double):
    return a.ctypes.data_as(C.POINTER(typ)) if a is not None else None
