"""ctypes binding of libcubeslam_b200.so (the C ABI declared in include/cube_slam_b200.h).

The library is the product; there is no Python or CPU fallback.  If the shared object is missing or
no CUDA device is usable, loading / cs_create fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcubeslam_b200.so")

CS_OK = 0
STATUS_NAMES = {0: "CS_OK", -1: "CS_ERR_INVALID_ARG", -2: "CS_ERR_CUDA", -3: "CS_ERR_CAPACITY",
                -4: "CS_ERR_NOT_PREPARED", -5: "CS_ERR_NCCL", -6: "CS_ERR_UNSUPPORTED"}


class CuboidParams(C.Structure):
    """cs_cuboid_params: detect_3d_cuboid's mode members + detect_cuboid's hard-coded locals."""
    _fields_ = [
        ("consider_config_1", C.c_int32), ("consider_config_2", C.c_int32),
        ("whether_sample_cam_roll_pitch", C.c_int32), ("whether_sample_bbox_height", C.c_int32),
        ("max_cuboid_num", C.c_int32), ("reweight_edge_distance", C.c_int32),
        ("whether_normalize_two_errors", C.c_int32), ("top_sample_count_override", C.c_int32),
        ("nominal_skew_ratio", C.c_double), ("max_cut_skew", C.c_double),
        ("vp12_edge_angle_thre", C.c_double), ("vp3_edge_angle_thre", C.c_double),
        ("shorted_edge_thre", C.c_double), ("weight_vp_angle", C.c_double), ("weight_skew_error", C.c_double),
        ("pre_merge_dist_thre", C.c_double), ("pre_merge_angle_thre", C.c_double),
        ("edge_length_threshold", C.c_double), ("canny_low", C.c_double), ("canny_high", C.c_double),
        ("yaw_half_range_deg", C.c_double), ("yaw_step_deg", C.c_double),
    ]


class LineParams(C.Structure):
    _fields_ = [("use_LSD", C.c_int32), ("numoctaves", C.c_int32), ("octaveratio", C.c_float),
                ("line_length_thres", C.c_float)]


class BatchStats(C.Structure):
    _fields_ = [("n_frames", C.c_int64), ("n_objects", C.c_int64), ("n_roi_jobs", C.c_int64),
                ("n_candidates", C.c_int64), ("n_valid", C.c_int64), ("n_kernel_launches", C.c_int64),
                ("roi_pixels", C.c_int64), ("n_lines_in", C.c_int64)]


# numpy view of cs_cuboid_rec (512 bytes)
CUBOID_DTYPE = np.dtype([
    ("pos", "f8", 3), ("scale", "f8", 3), ("rotY", "f8"), ("box_config_type", "f8", 2),
    ("box_corners_2d", "i4", (2, 8)), ("box_corners_3d_world", "f8", (3, 8)),
    ("rect_detect_2d", "f8", 4), ("edge_distance_error", "f8"), ("edge_angle_error", "f8"),
    ("normalized_error", "f8"), ("skew_ratio", "f8"), ("down_expand_height", "f8"),
    ("camera_roll_delta", "f8"), ("camera_pitch_delta", "f8"), ("combined_score", "f8"),
    ("proposal_index", "i4"), ("height_sample_id", "i4"), ("valid", "i4"), ("pad_", "i4"),
])

# numpy views of cs_keyline (40 bytes) and cs_dmatch (16 bytes)
KEYLINE_DTYPE = np.dtype([("start_x", "f4"), ("start_y", "f4"), ("end_x", "f4"), ("end_y", "f4"), ("angle", "f4"), ("line_length", "f4"),
                          ("response", "f4"), ("size", "f4"), ("num_pixels", "i4"), ("class_id", "i4")])
DMATCH_DTYPE = np.dtype([("query_idx", "i4"), ("train_idx", "i4"), ("img_idx", "i4"), ("distance", "f4")])

_lib = None

EXPORTS = [
    "cs_abi_version", "cs_create", "cs_destroy", "cs_last_error", "cs_default_cuboid_params",
    "cs_default_line_params", "cs_set_calibration", "cs_cam_pose", "cs_cuboid_measurement", "cs_cuboid_measurement_orb", "cs_detect_cuboids", "cs_detect_cuboids_batch",
    "cs_batch_upload", "cs_batch_upload_online", "cs_detect_frames_batch", "cs_batch_run", "cs_batch_run_async", "cs_batch_fetch", "cs_batch_stats_get",
    "cs_batch_device_records", "cs_stream", "cs_stage_ms", "cs_set_profiling", "cs_debug_roi",
    "cs_debug_candidates", "cs_detect_lines", "cs_detect_lines_batch", "cs_debug_lsd", "cs_debug_lsd_stats", "cs_debug_lsd_prof", "cs_debug_atan2", "cs_atan2_host", "cs_cuboid_draw_edges", "cs_debug_edlines", "cs_debug_stage_offsets", "cs_comm_unique_id", "cs_comm_init",
    "cs_allgather_topk", "cs_allgather_wait", "cs_fetch_gathered",
    "cs_keylines_from_lines", "cs_lbd_compute", "cs_lbd_compute_batch", "cs_detect_descrip_lines", "cs_detect_descrip_lines_batch",
    "cs_match_line_descrip", "cs_match_line_descrip_batch", "cs_lbd_debug_prepare", "cs_lbd_debug_keylines_edl", "cs_debug_last_set_pose",
]


def load():
    """dlopen the product library; raises if it has not been built (python -m cube_slam_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("%s is missing: build it with `python -m cube_slam_b200.build` (needs nvcc); "
                      "cube_slam_b200 has no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i, d_p = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    u8_p, i32_p, f_p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float)
    L.cs_abi_version.restype = i
    L.cs_create.restype = vp
    L.cs_create.argtypes = [i] * 6
    L.cs_destroy.argtypes = [vp]
    L.cs_destroy.restype = None
    L.cs_last_error.restype = C.c_char_p
    L.cs_last_error.argtypes = [vp]
    L.cs_default_cuboid_params.argtypes = [C.POINTER(CuboidParams)]
    L.cs_default_line_params.argtypes = [C.POINTER(LineParams)]
    L.cs_set_calibration.argtypes = [vp, d_p]
    L.cs_cam_pose.argtypes = [d_p, d_p, d_p, d_p]
    L.cs_cuboid_measurement.argtypes = [vp, d_p, d_p, d_p, d_p, d_p, d_p, d_p]
    L.cs_cuboid_measurement_orb.argtypes = [vp, d_p, C.c_double, d_p, d_p, d_p, d_p]
    L.cs_detect_cuboids.argtypes = [vp, u8_p, i, i, i, i, d_p, d_p, i, d_p, i, C.POINTER(CuboidParams), vp, i32_p]
    L.cs_detect_cuboids_batch.argtypes = [vp, vp, i, i, i, i, i, d_p, d_p, i32_p, d_p, i32_p, C.POINTER(CuboidParams), vp, i32_p]
    L.cs_batch_upload.argtypes = [vp, vp, i, i, i, i, i, d_p, d_p, i32_p, d_p, i32_p, C.POINTER(CuboidParams)]
    L.cs_batch_upload_online.argtypes = [vp, vp, i, i, i, i, i, d_p, d_p, i32_p, C.POINTER(LineParams), C.POINTER(CuboidParams)]
    L.cs_detect_frames_batch.argtypes = [vp, vp, i, i, i, i, i, d_p, d_p, i32_p, C.POINTER(LineParams), C.POINTER(CuboidParams), vp, i32_p]
    L.cs_batch_run.argtypes = [vp]
    L.cs_batch_run_async.argtypes = [vp]
    L.cs_batch_fetch.argtypes = [vp, vp, i32_p]
    L.cs_batch_stats_get.argtypes = [vp, C.POINTER(BatchStats)]
    L.cs_batch_device_records.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cs_stream.restype = vp
    L.cs_stream.argtypes = [vp]
    L.cs_stage_ms.argtypes = [vp, C.c_char_p, C.POINTER(C.c_float)]
    L.cs_set_profiling.argtypes = [vp, i]
    L.cs_debug_stage_offsets.argtypes = [vp, vp, f_p]
    L.cs_debug_roi.argtypes = [vp, i, i32_p, u8_p, f_p, i, d_p, i, i32_p, i32_p]
    L.cs_debug_candidates.argtypes = [vp, i, i32_p, u8_p, d_p, d_p, i]
    L.cs_detect_lines.argtypes = [vp, vp, i, i, i, i, C.POINTER(LineParams), f_p, i32_p]
    L.cs_detect_lines_batch.argtypes = [vp, vp, i, i, i, i, i, C.POINTER(LineParams), f_p, C.c_int32, i32_p]
    L.cs_debug_lsd.argtypes = [vp, i, i32_p, d_p, d_p, d_p, i32_p, i32_p, f_p, i32_p, i]
    L.cs_debug_lsd_stats.argtypes = [vp, i32_p, i32_p, i]
    L.cs_debug_lsd_prof.argtypes = [vp, C.POINTER(C.c_uint64), i]
    L.cs_debug_atan2.argtypes = [vp, d_p, d_p, d_p, i]
    L.cs_atan2_host.argtypes = [C.c_double, C.c_double]
    L.cs_atan2_host.restype = C.c_double
    L.cs_cuboid_draw_edges.argtypes = [vp, i32_p]
    L.cs_debug_edlines.argtypes = [vp, i, u8_p, C.POINTER(C.c_int16), C.POINTER(C.c_int16), C.POINTER(C.c_int16), u8_p, i32_p, i32_p, u8_p, f_p, i32_p, i]
    L.cs_comm_unique_id.argtypes = [vp, C.c_char_p, u8_p]
    L.cs_comm_init.argtypes = [vp, C.c_char_p, u8_p, i, i]
    L.cs_allgather_topk.argtypes = [vp, i, C.POINTER(vp)]
    L.cs_allgather_wait.argtypes = [vp]
    L.cs_fetch_gathered.argtypes = [vp, vp, i]
    L.cs_keylines_from_lines.argtypes = [f_p, i, i, i, vp]
    L.cs_lbd_compute.argtypes = [vp, vp, i, i, i, i, vp, i, u8_p, f_p]
    L.cs_lbd_compute_batch.argtypes = [vp, vp, i, i, i, i, i, vp, i32_p, u8_p, f_p]
    L.cs_detect_descrip_lines.argtypes = [vp, vp, i, i, i, i, C.POINTER(LineParams), vp, u8_p, i32_p]
    L.cs_detect_descrip_lines_batch.argtypes = [vp, vp, i, i, i, i, i, C.POINTER(LineParams), vp, u8_p, C.c_int32, i32_p]
    L.cs_match_line_descrip.argtypes = [vp, u8_p, i, u8_p, i, C.c_float, vp, i32_p]
    L.cs_match_line_descrip_batch.argtypes = [vp, u8_p, i32_p, u8_p, i32_p, i, C.c_float, vp, i32_p]
    L.cs_lbd_debug_prepare.argtypes = [vp, i, vp, f_p, f_p]
    L.cs_lbd_debug_keylines_edl.argtypes = [f_p, f_p, i, i, i, vp]
    L.cs_debug_last_set_pose.argtypes = [u8_p, d_p, d_p, i, i, i32_p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("cs_abi_version",):
            fn.restype = C.c_int
    _lib = L
    return L


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))
