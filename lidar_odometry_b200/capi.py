"""ctypes binding of ``libb2lo.so`` — the C ABI declared in ``include/b2lo.h``.

This is plumbing only: every call lands in hand-written sm_100a kernels (``lidar_odometry_b200/csrc``).
There is no CPU fallback; :func:`lib` raises if the shared library is missing, and ``b2lo_ctx_create``
fails with ``B2LO_E_CUDA`` when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2LO_LIB") or os.path.join(HERE, "libb2lo.so")   # B2LO_LIB: an alternative build of the same library (kernel A/B experiments)
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(os.path.dirname(HERE), "include", "b2lo.h")

B2LO_OK, B2LO_S_INSUFFICIENT, B2LO_S_EMPTY = 0, 1, 2
B2LO_E_CUDA, B2LO_E_ARG, B2LO_E_RANGE, B2LO_E_CAPACITY, B2LO_E_NOMEM = -1, -2, -3, -4, -5
B2LO_MAX_ITERS = 16


class B2loError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b2lo error {code}: {msg}")
        self.code = code


class IcpCfg(C.Structure):  # b2lo_icp_cfg
    _fields_ = [
        ("max_iterations", C.c_int), ("translation_tolerance", C.c_double), ("rotation_tolerance", C.c_double),
        ("max_correspondence_distance", C.c_double), ("min_correspondence_points", C.c_int), ("use_robust_loss", C.c_int),
        ("robust_loss_delta", C.c_double), ("use_surfel_correspondence", C.c_int), ("use_adaptive_m_estimator", C.c_int),
        ("loss_type", C.c_int), ("min_scale_factor", C.c_double), ("max_scale_factor", C.c_double),
        ("num_alpha_segments", C.c_int), ("truncated_threshold", C.c_double), ("gmm_components", C.c_int),
        ("gmm_sample_size", C.c_int), ("pko_kernel_type", C.c_int),
    ]


class IterTrace(C.Structure):  # b2lo_iter_trace
    _fields_ = [
        ("n_corr", C.c_int), ("scale", C.c_double), ("delta", C.c_double),
        ("H", C.c_double * 36), ("g", C.c_double * 6), ("cost", C.c_double),
        ("dx", C.c_float * 6), ("T_in", C.c_float * 16), ("T_out", C.c_float * 16),
        ("em_iters", C.c_int), ("kmeans_iters", C.c_int),
    ]


class IcpStats(C.Structure):  # b2lo_icp_stats
    _fields_ = [
        ("status", C.c_int), ("num_iterations", C.c_int), ("num_correspondences", C.c_int), ("converged", C.c_int),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("device_ms", C.c_float),
        ("it", IterTrace * B2LO_MAX_ITERS),
    ]


class OdomCfg(C.Structure):  # b2lo_odom_cfg
    _fields_ = [
        ("voxel_size", C.c_float), ("point_stride", C.c_int), ("map_voxel_size", C.c_float), ("max_range", C.c_double),
        ("surfel_planarity_threshold", C.c_float), ("keyframe_distance_threshold", C.c_double),
        ("keyframe_rotation_threshold", C.c_double), ("icp", IcpCfg),
    ]


class OdomResult(C.Structure):  # b2lo_odom_result
    _fields_ = [
        ("pose", C.c_float * 16), ("keyframe", C.c_int), ("icp_status", C.c_int), ("n_features", C.c_int),
        ("n_corr", C.c_int), ("n_iters", C.c_int), ("device_ms", C.c_float), ("l0", C.c_size_t), ("l1", C.c_size_t),
    ]


class RecordFmt(C.Structure):  # b2lo_record_fmt
    _fields_ = [("record_bytes", C.c_uint32), ("off_x", C.c_uint32), ("off_y", C.c_uint32), ("off_z", C.c_uint32)]


def build(force=False, verbose=False):
    """Compile libb2lo.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-s", "-C", CSRC, "clean"], check=True)
    r = subprocess.run(["make", "-s", "-j8", "-C", CSRC], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libb2lo.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None
_vp, _sz, _i, _f, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_double

# name -> (restype, argtypes); every function declared in include/b2lo.h
SIGNATURES = {
    "b2lo_default_icp_cfg": (None, [C.POINTER(IcpCfg)]),
    "b2lo_version": (C.c_char_p, []),
    "b2lo_last_error": (C.c_char_p, []),
    "b2lo_struct_sizes": (None, [_vp]),
    "b2lo_process_env_for_batches": (_i, [_i]),
    "b2lo_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "b2lo_ctx_destroy": (_i, [_vp]),
    "b2lo_ctx_sync": (_i, [_vp]),
    "b2lo_ctx_stream": (_vp, [_vp]),
    "b2lo_ctx_launch_count": (C.c_longlong, [_vp]),
    "b2lo_ctx_io_bytes": (_i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "b2lo_ctx_debug_clocks": (_i, [_vp, _vp]),
    "b2lo_ctx_host_us": (_i, [_vp, _vp, _i]),
    "b2lo_ctx_profile": (_i, [_vp, _i]),
    "b2lo_ctx_profile_read": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(C.c_longlong)]),
    "b2lo_filter": (_i, [_vp, _vp, _sz, _sz, _i, _f, _vp, _vp, C.POINTER(_sz)]),
    "b2lo_filter_dev": (_i, [_vp, _vp, _sz, _sz, _i, _f]),
    "b2lo_ctx_features": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "b2lo_kitti_record_fmt": (None, [C.POINTER(RecordFmt)]),
    "b2lo_ply_parse_header": (_i, [_vp, _sz, C.POINTER(RecordFmt), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_i), C.POINTER(_sz)]),
    "b2lo_ply_read_ascii": (_i, [_vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "b2lo_filter_records": (_i, [_vp, _vp, _sz, C.POINTER(RecordFmt), _i, _f, _vp, _vp, C.POINTER(_sz)]),
    "b2lo_filter_records_dev": (_i, [_vp, _vp, _sz, C.POINTER(RecordFmt), _i, _f]),
    "b2lo_voxel_grid_filter": (_i, [_vp, _vp, _sz, _sz, _f, _vp, _sz, C.POINTER(_sz)]),
    "b2lo_map_create": (_i, [_vp, _f, _i, _f, _i, _sz, C.POINTER(_vp)]),
    "b2lo_map_destroy": (_i, [_vp]),
    "b2lo_map_clear": (_i, [_vp]),
    "b2lo_map_set_planarity_threshold": (_i, [_vp, _f]),
    "b2lo_map_set_compute_surfels": (_i, [_vp, _i]),
    "b2lo_map_update": (_i, [_vp, _vp, _sz, _sz, _vp, _d]),
    "b2lo_map_counts": (_i, [_vp, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "b2lo_map_lookup": (_i, [_vp, _vp, _vp, _vp]),
    "b2lo_map_export_l0": (_i, [_vp, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "b2lo_map_export_surfels": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "b2lo_map_export_l1": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "b2lo_map_rebuild_knn": (_i, [_vp]),
    "b2lo_map_has_knn": (_i, [_vp]),
    "b2lo_map_transform_rehash": (_i, [_vp, _vp]),
    "b2lo_icp_correspondences": (_i, [_vp, _vp, _sz, _sz, _vp, _d, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_sz)]),
    "b2lo_icp_correspondences_knn": (_i, [_vp, _vp, _sz, _sz, _vp, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "b2lo_icp_optimize": (_i, [_vp, _vp, _sz, _sz, _vp, C.POINTER(IcpCfg), _vp, C.POINTER(IcpStats)]),
    "b2lo_icp_iterate": (_i, [_vp, _vp, _sz, _sz, _vp, _d, C.POINTER(IcpCfg), _vp, C.POINTER(IcpStats)]),
    "b2lo_icp_optimize_loop": (_i, [_vp, _vp, _sz, _sz, _vp, _vp, _sz, _sz, _vp, C.POINTER(IcpCfg), _vp, C.POINTER(C.c_float), C.POINTER(IcpStats)]),
    "b2lo_icp_optimize_features": (_i, [_vp, _vp, C.POINTER(IcpCfg), _vp, C.POINTER(IcpStats)]),
    "b2lo_icp_shard_begin": (_i, [_vp, _vp, _sz, _sz, _vp, C.POINTER(IcpCfg)]),
    "b2lo_icp_shard_corr": (_i, [_vp, C.POINTER(IcpCfg), _vp]),
    "b2lo_icp_shard_sample": (_i, [_vp, C.POINTER(IcpCfg), C.c_longlong, C.c_longlong, _d, _vp]),
    "b2lo_icp_shard_accumulate": (_i, [_vp, C.POINTER(IcpCfg), C.c_longlong, _d, _vp, _vp]),
    "b2lo_icp_shard_finish": (_i, [_vp, C.POINTER(IcpCfg), _vp, _vp, C.POINTER(_i), C.POINTER(IcpStats)]),
    "b2lo_shard_unique_id": (_i, [_vp, _sz]),
    "b2lo_shard_comm_create": (_i, [_vp, _i, _i, _vp, _sz, C.POINTER(_vp)]),
    "b2lo_shard_comm_destroy": (_i, [_vp]),
    "b2lo_shard_comm_ipc_handle": (_i, [_vp, _vp, _sz]),
    "b2lo_shard_comm_open_peers": (_i, [_vp, _vp, _sz]),
    "b2lo_icp_shard_optimize": (_i, [_vp, _vp, _vp, _sz, _sz, _vp, C.POINTER(IcpCfg), _vp, C.POINTER(IcpStats), C.POINTER(C.c_float)]),
    "b2lo_se3_mul": (None, [_vp, _vp, _vp]),
    "b2lo_se3_inv": (None, [_vp, _vp]),
    "b2lo_se3_from_rt": (None, [_vp, _vp]),
    "b2lo_so3_log": (None, [_vp, _vp]),
    "b2lo_so3_exp": (None, [_vp, _vp]),
    "b2lo_svd3": (None, [_vp, _vp, _vp, _vp]),
    "b2lo_ldlt6_solve": (None, [_vp, _vp, _vp]),
    "b2lo_fit_plane": (None, [_vp, _i, _vp, _vp, _vp]),
    "b2lo_voxel_key_hash": (C.c_uint64, [_i, _i, _i]),
    "b2lo_default_odom_cfg": (None, [C.POINTER(OdomCfg), _i]),
    "b2lo_odom_create": (_i, [_vp, C.POINTER(OdomCfg), C.POINTER(_vp)]),
    "b2lo_odom_destroy": (_i, [_vp]),
    "b2lo_odom_reset": (_i, [_vp]),
    "b2lo_odom_graph_stats": (_i, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "b2lo_odom_map": (_vp, [_vp]),
    "b2lo_odom_process": (_i, [_vp, _vp, _sz, _sz, C.POINTER(OdomResult)]),
    "b2lo_odom_process_dev": (_i, [_vp, _vp, _sz, _sz, C.POINTER(OdomResult)]),
    "b2lo_odom_lookahead": (_i, [_vp, _vp, _sz, _sz, C.c_int]),
    "b2lo_odom_process_la": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, C.POINTER(OdomResult)]),
    "b2lo_odom_set_record_fmt": (_i, [_vp, C.POINTER(RecordFmt)]),
    "b2lo_odom_process_batch_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "b2lo_lockstep_create": (_i, [_vp, _i, C.POINTER(_vp)]),
    "b2lo_lockstep_destroy": (_i, [_vp]),
    "b2lo_lockstep_process_dev": (_i, [_vp, _vp, _vp, _sz, _vp, C.POINTER(C.c_float)]),
    "b2lo_lockstep_stats": (_i, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
}


def lib():
    """Load libb2lo.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built (run __graft_entry__.build() or "
                f"`make -C {CSRC}`).  lidar_odometry_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("B2LO_LIB") and not hasattr(L, name):
                continue        # A/B runs against an older build of the library: entry points it does not have yet stay unbound
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().b2lo_last_error().decode()


def check(rc):
    """Negative codes are errors; non-negative codes (soft outcomes) are returned."""
    if rc < 0:
        raise B2loError(rc, last_error())
    return rc
