"""The distance transform kernels evaluate the 3x3 chamfer metric in "cone form" (two vertical sweeps without a dependency inside a
row, then two 1-D min-plus scans per row) instead of OpenCV's two raster passes.  This test restates that
evaluation order in numpy and checks it against the oracle's raster scan (itself pinned bit for bit to cv2): same integers on
random, sparse, single-pixel and empty edge maps."""
import numpy as np

A, B, BIG = 62587, 89738, 1 << 30   # DT_HV, DT_DG, DT_BIG of cs_kernels_image.cu


def cone_dt(edges):
    h, w = edges.shape
    S = np.empty((h, w), np.int64)
    up = np.full(w + 2, BIG, np.int64)
    for y in range(h):                                   # the down sweep of k_dt_bi
        nv = np.minimum(np.minimum(np.minimum(up[:-2], up[2:]) + B, up[1:-1] + A), BIG)
        nv[edges[y] > 0] = 0
        S[y] = nv
        up[1:-1] = nv
    N = np.empty_like(S)
    dn = np.full(w + 2, BIG, np.int64)
    for y in range(h - 1, -1, -1):                       # the up sweep, the mirror image
        nv = np.minimum(np.minimum(np.minimum(dn[:-2], dn[2:]) + B, dn[1:-1] + A), BIG)
        nv[edges[y] > 0] = 0
        N[y] = nv
        dn[1:-1] = nv
    V = np.minimum(S, N)                                  # taken by the scan warps
    cols = np.arange(w)                                   # the scan warps
    F = np.minimum.accumulate(V - A * cols, axis=1) + A * cols
    D = np.minimum.accumulate((F + A * cols)[:, ::-1], axis=1)[:, ::-1] - A * cols
    dist_max = np.float32(0xffffffff - B) * np.float32(1 / 65536.0)
    return np.where(D >= BIG, dist_max, D.astype(np.uint32).astype(np.float32) * np.float32(1 / 65536.0)).astype(np.float32)


def test_cone_form_equals_raster_scan(oracle):
    rng = np.random.default_rng(1)
    for trial in range(200):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        dens = float(rng.choice([0.0, 0.002, 0.01, 0.05, 0.3]))
        e = (rng.random((h, w)) < dens).astype(np.uint8) * 255
        if trial % 7 == 0:
            e[:] = 0
            e[rng.integers(0, h), rng.integers(0, w)] = 255
        np.testing.assert_array_equal(cone_dt(e), oracle.chamfer_dt(e), err_msg="trial %d (%dx%d, density %g)" % (trial, w, h, dens))


def test_cone_form_on_the_fixture_roi(oracle):
    import os
    from conftest import GOLD
    pins = np.load(os.path.join(GOLD, "cv_pins.npz"))
    np.testing.assert_array_equal(cone_dt(pins["roi_canny"]), pins["roi_dt"])       # cv2's own output on the fixture-A ROI
    np.testing.assert_array_equal(cone_dt(np.zeros((9, 11), np.uint8)), pins["empty_dt"])
