"""GPU parity of the EDLines line detector (cs_detect_lines, use_LSD = 0: the class default of line_lbd_detect,
line_lbd_allclass.cpp:121) against the CPU oracle's restatement of BinaryDescriptor / EDLineDetector.

Integer stages (blur, Sobel maps, gradient / direction maps, anchors in scan order, edge map after smart routing) must be
bit-exact; the fitted segments are float32 and must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det():
    import cube_slam_b200 as cs
    d = cs.line_lbd_detect()      # use_LSD = False, line_length_thres = 50: the constructor defaults
    assert d.use_LSD is False
    return d


def _check_frame(det, oracle, img, frame, thres):
    ref = oracle.edl_detect(img, thres, want_stages=True)
    h, w = img.shape[:2]
    dbg = det.debug_frame_edlines(w, h, frame)
    for k in ("blur", "dx", "dy", "g", "dir"):
        np.testing.assert_array_equal(dbg[k], ref["stages"][k], err_msg=k)
    np.testing.assert_array_equal(dbg["anchors"], ref["stages"]["anchors"])
    np.testing.assert_array_equal(dbg["edge"], ref["stages"]["edge"])
    assert len(dbg["raw_lines"]) == len(ref["raw_lines"])
    np.testing.assert_array_equal(dbg["raw_lines"], ref["raw_lines"])
    return ref


def test_fixture_frames(det, oracle, fixture_a, fixture_b):
    imgs = [fixture_b["frames"][i][0] for i in (0, 17, 40)]
    lines = det.detect_filter_lines_batch(np.stack(imgs))
    for f, img in enumerate(imgs):
        ref = _check_frame(det, oracle, img, f, 50.0)
        np.testing.assert_array_equal(lines[f], ref["lines"])
    one = det.detect_filter_lines(fixture_a["img"])
    ref = _check_frame(det, oracle, fixture_a["img"], 0, 50.0)
    np.testing.assert_array_equal(one, ref["lines"])
    assert len(one) > 5


def test_short_threshold_synthetic_gray_and_flat(det, oracle):
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    d = cs.line_lbd_detect(context=det._ctx)
    d.line_length_thres = 15
    imgs, Ts, boxes, lines, K = S.make_batch(43, 4, 640, 480, 3)
    out = d.detect_filter_lines_batch(imgs)
    total = 0
    for f in range(4):
        ref = _check_frame(d, oracle, imgs[f], f, 15.0)
        np.testing.assert_array_equal(out[f], ref["lines"])
        total += len(out[f])
    assert total > 10
    gray = np.ascontiguousarray(imgs[:, :, :, 1])
    out = d.detect_filter_lines_batch(gray)
    for f in range(4):
        np.testing.assert_array_equal(out[f], oracle.edl_detect(gray[f], 15.0)["lines"])
    flat = np.full((1, 240, 320, 3), 90, np.uint8)
    assert len(d.detect_filter_lines_batch(flat)[0]) == 0
    # odd sizes exercise the borders of the scan grid and of the routing walk
    rng = np.random.RandomState(5)
    odd = np.kron(rng.randint(0, 255, (9, 13)).astype(np.uint8), np.ones((23, 19), np.uint8))[:203, :241]
    got = d.detect_filter_lines(odd)
    ref = _check_frame(d, oracle, odd, 0, 15.0)
    np.testing.assert_array_equal(got, ref["lines"])
    assert len(got) > 5


def test_sequence_frames(det, oracle, fixture_b):
    """Every frame of the shipped object_slam sequence (object_slam/data/raw_imgs), one batch."""
    imgs = np.stack([fr[0] for fr in fixture_b["frames"]])
    out = det.detect_filter_lines_batch(imgs)
    for f in range(len(imgs)):
        ref = oracle.edl_detect(imgs[f], 50.0)
        np.testing.assert_array_equal(out[f], ref["lines"], err_msg="frame %d" % f)


def test_online_mode_with_edlines(det, oracle):
    """cs_detect_frames_batch with the EDLines flavour == cs_detect_lines_batch followed by cs_detect_cuboids_batch == oracle."""
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    F = 4
    imgs, Ts, boxes, _, K = S.make_batch(53, F, 640, 480, 3, poisson=True)
    ctx = cs.Context(0, 640, 480, F, 16, 2048)
    ctx.set_calibration(K)
    p = cs.default_params(max_cuboid_num=2)
    d = cs.line_lbd_detect(context=det._ctx)
    d.line_length_thres = 15
    lp = d.params()
    assert lp.use_LSD == 0
    out1, cnt1 = ctx.detect_frames_host(imgs, Ts, boxes, lp, p)
    out1, cnt1 = out1.copy(), cnt1.copy()
    lines = d.detect_filter_lines_batch(imgs)
    out2, cnt2 = ctx.detect_batch_host(imgs, Ts, boxes, [l.astype(np.float64) for l in lines], p)
    np.testing.assert_array_equal(cnt1, cnt2)
    assert out1.tobytes() == out2.tobytes()
    o = 0
    for f in range(F):
        rl = oracle.edl_detect(imgs[f], 15.0)["lines"].astype(np.float64)
        ref = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], rl, oracle.default_params(max_cuboid_num=2))
        for b in range(len(boxes[f])):
            assert cnt1[o] == len(ref["cuboids"][b])
            for k in range(cnt1[o]):
                assert int(out1[o, k]["proposal_index"]) == int(ref["cuboids"][b][k]["proposal_index"])
                assert abs(float(out1[o, k]["normalized_error"]) - float(ref["cuboids"][b][k]["normalized_error"])) < 1e-9
            o += 1
    ctx.close()
