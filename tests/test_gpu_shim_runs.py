"""The maintainer-side C++ shim, RUN: shim/line_lbd_b200.cpp compiled against the reference's own class header
(line_lbd/include/line_lbd/line_lbd_allclass.h) and called like object_slam/src/main_obj.cpp:363-366,428 calls the reference class --
construct line_lbd_detect, set use_LSD / line_length_thres, detect_filter_lines(img, lines_mat) -- through the harness
shim/test/line_shim_driver.cpp (oracle/_ref/libshim_line.so, built by oracle/Makefile where the reference checkout exists; cv::Mat is the
container stand-in oracle/ref/minicv.hpp because this image has no OpenCV C++ headers).  The n x 4 CV_32F matrix that comes out must be
the oracle's segments bit for bit, for both detectors."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "libshim_line.so")


@pytest.mark.gpu
@pytest.mark.parametrize("use_lsd", [1, 0])
def test_line_shim_returns_the_reference_segments(oracle, fixture_a, fixture_b, use_lsd):
    if not os.path.exists(SHIM):
        pytest.skip("oracle/_ref/libshim_line.so not built (needs the reference's headers at build time)")
    import cube_slam_b200  # noqa: F401  (fails loudly if the product library is missing)
    L = C.CDLL(SHIM)
    L.shim_line_detect_filter.restype = C.c_int
    for img, thres in ((fixture_a["img"], 15.0), (fixture_b["frames"][3][0], 15.0), (fixture_b["frames"][20][0], 30.0)):
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        out = np.zeros((8192, 4), np.float32)
        n = L.shim_line_detect_filter(img.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, ch, use_lsd, C.c_float(thres),
                                      out.ctypes.data_as(C.POINTER(C.c_float)), 8192)
        assert n >= 0
        want = (oracle.lsd_detect(img, thres) if use_lsd else oracle.edl_detect(img, thres))["lines"]
        assert n == len(want)
        np.testing.assert_array_equal(out[:n], want)


SHIM_CUBOID = os.path.join(ROOT, "oracle", "_ref", "libshim_cuboid.so")


@pytest.mark.gpu
@pytest.mark.parametrize("mode,kw", [("default", {}), ("top5", dict(max_cuboid_num=5)), ("roll_pitch", dict(whether_sample_cam_roll_pitch=1))])
def test_cuboid_shim_returns_the_reference_cuboids(oracle, fixture_a, mode, kw):
    """shim/detect_3d_cuboid_b200.cpp built in place of box_proposal_detail.cpp, against the reference's class header and next to the
    reference's own object_3d_util.cpp / matrix_utils.cpp (oracle/_ref/libshim_cuboid.so, shim/test/cuboid_shim_driver.cpp), called as
    main_obj.cpp:354-361,449 calls the class: the std::vector<ObjectSet> it fills must hold the oracle's cuboids -- same count, same
    order, discrete fields equal, continuous ones to 1e-9 (the tolerance of the direct parity tests) -- and cam_pose_raw.euler_angle the
    oracle's angles."""
    if not os.path.exists(SHIM_CUBOID):
        pytest.skip("oracle/_ref/libshim_cuboid.so not built (needs the reference's headers at build time)")
    import cube_slam_b200  # noqa: F401
    L = C.CDLL(SHIM_CUBOID)
    L.shim_detect_cuboid.restype = C.c_int
    p = oracle.default_params(**kw)
    k = max(int(p.max_cuboid_num), 1)
    img = np.ascontiguousarray(fixture_a["img"], np.uint8)
    h, w, ch = img.shape
    K = np.ascontiguousarray(fixture_a["K"], np.float64)
    T = np.ascontiguousarray(fixture_a["T"], np.float64)
    boxes = np.ascontiguousarray(fixture_a["boxes"], np.float64).reshape(-1, 5)
    lines = np.ascontiguousarray(fixture_a["lines"], np.float64).reshape(-1, 4)
    flags = np.array([p.consider_config_1, p.consider_config_2, p.whether_sample_cam_roll_pitch, p.whether_sample_bbox_height], np.int32)
    out = np.zeros((len(boxes), k), oracle.REF_CUBOID_DTYPE)
    counts = np.zeros(len(boxes), np.int32)
    euler = np.zeros(3)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
    rc = L.shim_detect_cuboid(img.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, ch, dp(K), dp(T), dp(boxes), len(boxes), dp(lines), len(lines),
                              flags.ctypes.data_as(C.POINTER(C.c_int32)), C.c_double(p.nominal_skew_ratio), int(p.max_cuboid_num),
                              out.ctypes.data_as(C.POINTER(C.c_double)), k, counts.ctypes.data_as(C.POINTER(C.c_int32)), dp(euler))
    assert rc == 0
    want = oracle.detect_cuboid(img, K, T, boxes, lines, p, topk_cap=k)["cuboids"]
    np.testing.assert_allclose(euler, oracle.cam_pose(K, T)["euler"], rtol=0, atol=1e-12)
    for b in range(len(boxes)):
        assert counts[b] == len(want[b])
        for j in range(counts[b]):
            g, o = out[b, j], want[b][j]
            np.testing.assert_array_equal(g["box_corners_2d"], o["box_corners_2d"])
            np.testing.assert_array_equal(g["box_config_type"], o["box_config_type"])
            np.testing.assert_array_equal(g["rect_detect_2d"], o["rect_detect_2d"])
            for f in ("pos", "rotY", "scale", "box_corners_3d_world", "edge_distance_error", "edge_angle_error", "normalized_error", "skew_ratio",
                      "down_expand_height", "camera_roll_delta", "camera_pitch_delta"):
                np.testing.assert_allclose(np.asarray(g[f], np.float64), np.asarray(o[f], np.float64), rtol=1e-9, atol=1e-9, err_msg=f)
