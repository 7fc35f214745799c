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
