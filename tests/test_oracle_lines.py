"""The line-detector oracles (oracle/lsd_oracle.cpp, oracle/edl_oracle.cpp) against third-party pins.

The reference ships no tests for line_lbd.  What it does ship is ONE output of its own LSD node for the demo frame
(detect_3d_cuboid/data/edge_detection/LSD/0000_edge.txt, written by line_lbd/src/detect_lines.cpp:94-104 with
use_LSD = true, length > 15, six significant digits).  The oracle reproduces that file whole: 271 of 271 segments, in the file's
order, to the printed precision -- a hard pin of the LSD flavour of detect_filter_lines end to end (cvtColor, blur, resize,
gradient, the raster seed order of the vendored lsd.cpp, region growing, refinement, NFA, the length filter).  cv2's own
(rewritten) LSD reproduces almost none of them: the vendored lsd.cpp quirks matter.  tests/test_oracle_ref_lsd.py adds the
reference's lsd.cpp itself, compiled, on more images.  The OpenCV calls inside both detectors are pinned against cv2 where it
is importable."""
import os

import numpy as np
import pytest


def _nearest(gold, lines):
    lines = lines.astype(np.float64)
    sw = lines[:, [2, 3, 0, 1]]
    out = np.empty(len(gold))
    for i, g in enumerate(gold):
        out[i] = min(np.abs(lines - g).max(axis=1).min(), np.abs(sw - g).max(axis=1).min())
    return out


def test_lsd_reproduces_shipped_segments(oracle, fixture_a):
    gold = fixture_a["lines"]
    assert gold.shape == (271, 4)
    res = oracle.lsd_detect(fixture_a["img"], 15.0)
    lines = res["lines"].astype(np.float64)
    assert lines.shape == gold.shape                      # every segment, none extra
    # same order as the file (the order the seed loop found them in); six significant digits of coordinates < 1000
    np.testing.assert_allclose(lines, gold, rtol=2e-5, atol=2e-5)
    assert _nearest(gold, res["lines"]).max() < 1e-3
    lens = np.hypot(lines[:, 0] - lines[:, 2], lines[:, 1] - lines[:, 3])
    assert lens.min() > 15.0


def test_lsd_fast_atan2_and_blur_resize_against_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    cv2.setUseOptimized(False)
    cv2.setNumThreads(1)
    rng = np.random.default_rng(11)
    ys = np.concatenate([rng.normal(0, 300, 4000), [0, 0, 1, -1, 5, -5, 0.0]]).astype(np.float32)
    xs = np.concatenate([rng.normal(0, 300, 4000), [0, 1, 0, 0, 5, 5, -3.0]]).astype(np.float32)
    for y, x in zip(ys, xs):
        assert oracle.lsd_fast_atan2(float(y), float(x)) == cv2.fastAtan2(float(y), float(x))
    for shape in [(120, 160), (97, 211), (480, 640)]:
        src = rng.integers(0, 256, shape).astype(np.float64)
        blur, scaled = oracle.lsd_blur_resize(src)
        ref_blur = cv2.GaussianBlur(src, (7, 7), 0.6 / 0.8)   # lsd.cpp:452-456: sigma = sigma_scale / scale, ksize 7
        np.testing.assert_array_equal(blur, ref_blur)
        ref_scaled = cv2.resize(ref_blur, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)   # lsd.cpp:459: Size(), SCALE, SCALE
        np.testing.assert_array_equal(scaled, ref_scaled)


def test_edlines_maps_against_cv2(oracle, fixture_a, fixture_b):
    cv2 = pytest.importorskip("cv2")
    cv2.setUseOptimized(False)
    cv2.setNumThreads(1)
    rng = np.random.default_rng(13)
    noise = rng.integers(0, 256, (101, 143, 3), dtype=np.uint8)
    for img in (fixture_a["img"], fixture_b["frames"][3][0], noise):
        st = oracle.edl_detect(img, 50.0, want_stages=True)["stages"]
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        blur = cv2.GaussianBlur(gray, (5, 5), 1.0)              # binary_descriptor.cpp:811-812
        np.testing.assert_array_equal(st["blur"], blur)
        dx = cv2.Sobel(blur, cv2.CV_16S, 1, 0, ksize=3)           # :1617-1618
        dy = cv2.Sobel(blur, cv2.CV_16S, 0, 1, ksize=3)
        np.testing.assert_array_equal(st["dx"], dx)
        np.testing.assert_array_equal(st["dy"], dy)
        s = np.abs(dx.astype(np.int32)) + np.abs(dy.astype(np.int32))
        s16 = s.astype(np.int16)
        _, thr = cv2.threshold(s16, 81, 255, cv2.THRESH_TOZERO)   # :1624 (gradienThreshold_ + 1)
        g = cv2.multiply(thr, np.ones_like(thr), scale=0.25)      # `gImg_ / 4` is a scaled convert: cvRound, half to even
        np.testing.assert_array_equal(st["g"], g)
        np.testing.assert_array_equal(st["dir"], np.where(np.abs(dx.astype(np.int32)) < np.abs(dy.astype(np.int32)), 255, 0))
        # anchors: local maxima of g across the edge direction, on the odd grid, scanned column by column (:1640-1666)
        h, w = gray.shape
        anc = []
        gi = st["g"].astype(np.int32)
        for x in range(1, w - 1, 2):
            for y in range(1, h - 1, 2):
                if st["dir"][y, x] == 255:
                    ok = gi[y, x] >= gi[y - 1, x] + 8 and gi[y, x] >= gi[y + 1, x] + 8
                else:
                    ok = gi[y, x] >= gi[y, x - 1] + 8 and gi[y, x] >= gi[y, x + 1] + 8
                if ok:
                    anc.append(y * w + x)
        np.testing.assert_array_equal(st["anchors"], np.array(anc, np.int32))
        # routing only ever marks pixels above the gradient threshold, and every surviving anchor is on an edge
        assert not (st["edge"][st["g"] == 0]).any()


def test_edlines_segments_are_supported_by_edges(oracle, fixture_a):
    """Sanity of the sequential half: each segment lies on routed edge pixels and is longer than the length filter."""
    res = oracle.edl_detect(fixture_a["img"], 50.0, want_stages=True)
    edge = res["stages"]["edge"] > 0
    L = res["lines"]
    assert len(L) == len(oracle.edl_detect(fixture_a["img"], 50.0)["lines"]) and len(L) >= 15
    h, w = edge.shape
    for x1, y1, x2, y2 in L:
        assert np.hypot(x2 - x1, y2 - y1) > 50.0
        t = np.linspace(0.05, 0.95, 40)
        xs, ys = x1 + t * (x2 - x1), y1 + t * (y2 - y1)
        hit = 0
        for x, y in zip(xs, ys):
            xi, yi = int(round(float(x))), int(round(float(y)))
            hit += edge[max(0, yi - 2):yi + 3, max(0, xi - 2):xi + 3].any()
        assert hit >= 36
    raw = res["raw_lines"]
    assert len(raw) >= len(L)
    # the length filter is the only difference between the raw and the filtered list
    keep = np.hypot(raw[:, 0] - raw[:, 2], raw[:, 1] - raw[:, 3]) > 50.0
    np.testing.assert_array_equal(raw[keep], L)


def test_edlines_segments_lie_on_cv2_canny_edges(oracle, fixture_a, fixture_b):
    """Third-party sanity for the sequential half of EDLines (no reference output exists): every validated segment follows an edge that
    cv2's Canny finds in the same blurred image."""
    cv2 = pytest.importorskip("cv2")
    n_seg = 0
    for img in [fixture_a["img"]] + [fixture_b["frames"][i][0] for i in (0, 15, 30, 45)]:
        L = oracle.edl_detect(img, 30.0)["lines"]
        g = cv2.GaussianBlur(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), (5, 5), 1.0)
        e = cv2.dilate(cv2.Canny(g, 20, 60), np.ones((5, 5), np.uint8)) > 0
        h, w = e.shape
        for x1, y1, x2, y2 in L:
            t = np.linspace(0.03, 0.97, 50)
            xs = np.clip(np.rint(x1 + t * (x2 - x1)).astype(int), 0, w - 1)
            ys = np.clip(np.rint(y1 + t * (y2 - y1)).astype(int), 0, h - 1)
            assert e[ys, xs].mean() >= 0.9
            n_seg += 1
    assert n_seg >= 60
