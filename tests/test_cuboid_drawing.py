"""Result drawing: the 12 edges plot_image_with_cuboid draws (object_3d_util.cpp:54-131) through the C ABI,
host-only (no GPU): corner indices / colours / thickness against the reference's tables restated here, and the rasterisation against
the same cv::line call the reference makes."""
import numpy as np
import pytest


# object_3d_util.cpp:60-69 (visible / hidden edge corner ids, final_universal_object) and :86-92 (marker types), 1-based as written there
REF_EDGES = {
    (1, 1): ([3, 4, 4, 1, 4, 8, 1, 2, 2, 3, 2, 6, 1, 5, 3, 7, 5, 6, 6, 7, 7, 8, 8, 5], [4, 2, 6, 3, 1, 5, 5, 5, 3, 1, 3, 1]),
    (1, 2): ([2, 3, 3, 4, 3, 7, 1, 2, 1, 4, 2, 6, 1, 5, 4, 8, 5, 6, 6, 7, 7, 8, 8, 5], [2, 4, 6, 3, 1, 5, 5, 5, 3, 1, 3, 1]),
    (2, 1): ([2, 3, 3, 4, 4, 1, 3, 7, 4, 8, 1, 2, 2, 6, 1, 5, 5, 6, 6, 7, 7, 8, 8, 5], [2, 4, 2, 6, 6, 3, 5, 5, 3, 1, 3, 1]),
    (2, 2): ([2, 3, 3, 4, 4, 1, 3, 7, 4, 8, 1, 2, 2, 6, 1, 5, 5, 6, 6, 7, 7, 8, 8, 5], [2, 4, 2, 6, 6, 3, 5, 5, 3, 1, 3, 1]),
}
LINE_MARKERS = [[0, 0, 255, 2], [0, 0, 255, 1], [0, 255, 0, 2], [0, 255, 0, 1], [255, 0, 0, 2], [255, 0, 0, 1]]  # :111-112


@pytest.mark.parametrize("cfg", sorted(REF_EDGES))
def test_draw_edges_follow_the_reference_tables(cfg):
    import cv2
    import cube_slam_b200 as cs
    rng = np.random.default_rng(sum(cfg))
    rec = np.zeros(1, cs.CUBOID_DTYPE)
    rec["box_config_type"][0] = cfg
    corners = rng.integers(20, 300, (2, 8)).astype(np.int32)
    rec["box_corners_2d"][0] = corners
    edges = cs.cuboid_draw_edges(rec)
    pts, marks = REF_EDGES[cfg]
    want = []
    for e in range(12):
        a, b, m = pts[2 * e] - 1, pts[2 * e + 1] - 1, marks[e] - 1
        want.append([corners[0, a], corners[1, a], corners[0, b], corners[1, b]] + LINE_MARKERS[m])
    np.testing.assert_array_equal(edges, np.array(want, np.int32))
    # rasterisation = the reference's own cv::line(..., CV_AA) calls in the same order
    img = np.full((320, 320, 3), 90, np.uint8)
    ref = img.copy()
    for x1, y1, x2, y2, bb, gg, rr, th in want:
        cv2.line(ref, (int(x1), int(y1)), (int(x2), int(y2)), (int(bb), int(gg), int(rr)), int(th), cv2.LINE_AA, 0)
    got = cs.plot_image_with_cuboid(img.copy(), rec)
    np.testing.assert_array_equal(got, ref)
    assert (got != img).any()
