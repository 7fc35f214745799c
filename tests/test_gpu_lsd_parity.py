"""GPU parity of the LSD line detector (cs_detect_lines) against the CPU oracle's restatement of line_lbd's LSD path.

Streaming stages (blur, resize, gradient modulus / angle, pseudo-ordering) must be bit-exact; the detected segments are
float32 and must be identical (the only non-IEEE operations are double cos/sin/log of CUDA vs glibc, <= 2 ulp before the
float rounding)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det():
    import cube_slam_b200 as cs
    d = cs.line_lbd_detect()
    d.use_LSD = True              # object_slam/src/main_obj.cpp:365
    d.line_length_thres = 15      # :366
    return d


def _check_frame(det, oracle, img, frame=0):
    ref = oracle.lsd_detect(img, 15.0, want_stages=True)
    dbg = det.debug_frame(frame)
    np.testing.assert_array_equal(dbg["scaled"], ref["stages"]["scaled"])
    np.testing.assert_array_equal(dbg["modgrad"][:-1, :-1], ref["stages"]["modgrad"][:-1, :-1])
    np.testing.assert_array_equal(dbg["angles"], ref["stages"]["angles"])
    # the GPU list keeps only pixels with a defined angle (the seed loop skips the others): same order otherwise
    rl = ref["stages"]["list"]
    np.testing.assert_array_equal(dbg["list"], rl[ref["stages"]["angles"].ravel()[rl] != -1024.0])
    assert len(dbg["raw_lines"]) == len(ref["raw_lines"])
    np.testing.assert_array_equal(dbg["raw_lines"], ref["raw_lines"])
    return ref


def test_fixture_frames(det, oracle, fixture_a, fixture_b):
    imgs = [fixture_b["frames"][i][0] for i in (0, 17, 40)]
    lines = det.detect_filter_lines_batch(np.stack(imgs))
    for f, img in enumerate(imgs):
        ref = _check_frame(det, oracle, img, f)
        np.testing.assert_array_equal(lines[f], ref["lines"])
        assert len(lines[f]) > 5
    one = det.detect_filter_lines(fixture_a["img"])
    ref = _check_frame(det, oracle, fixture_a["img"], 0)
    np.testing.assert_array_equal(one, ref["lines"])


def test_synthetic_and_gray_input(det, oracle):
    from cube_slam_b200 import synthetic as S
    imgs, Ts, boxes, lines, K = S.make_batch(41, 4, 640, 480, 3)
    out = det.detect_filter_lines_batch(imgs)
    for f in range(4):
        ref = oracle.lsd_detect(imgs[f], 15.0)
        np.testing.assert_array_equal(out[f], ref["lines"])
    gray = np.ascontiguousarray(imgs[:, :, :, 1])
    out = det.detect_filter_lines_batch(gray)
    for f in range(4):
        np.testing.assert_array_equal(out[f], oracle.lsd_detect(gray[f], 15.0)["lines"])
    # a flat image has no gradient above the threshold: no lines, no crash
    flat = np.full((1, 240, 320, 3), 90, np.uint8)
    assert len(det.detect_filter_lines_batch(flat)[0]) == 0


def test_more_octaves_give_the_same_lines_and_bad_counts_fail_loudly(det, oracle, fixture_a):
    """filter_lines keeps octave 0 only (line_lbd_allclass.cpp:200-207) and octave 0 does not depend on the higher octaves: a detector built
    with 2 or 3 octaves returns the one-octave matrix (the compiled reference does: tests/test_oracle_ref_octaves.py)."""
    import cube_slam_b200 as cs
    d = cs.line_lbd_detect(context=det._ctx)
    d.use_LSD = True
    d.line_length_thres = 15
    want = oracle.lsd_detect(fixture_a["img"], 15.0)["lines"]
    for n in (2, 3):
        d.numoctaves_ = n
        np.testing.assert_array_equal(d.detect_filter_lines(fixture_a["img"]), want)
    d.numoctaves_ = 0
    with pytest.raises(cs.CubeSlamError, match="INVALID_ARG"):
        d.detect_filter_lines(np.zeros((64, 64), np.uint8))


def test_lines_feed_detect_cuboid(det, oracle, fixture_b):
    """The online path of object_slam: line_lbd lines -> detect_cuboid (main_obj.cpp:424-450), frame 0 (no sampling)."""
    import cube_slam_b200 as cs
    img, boxes = fixture_b["frames"][0]
    lines = det.detect_filter_lines(img)
    d3 = cs.detect_3d_cuboid()
    d3.set_calibration(fixture_b["K"])
    d3.nominal_skew_ratio = 2  # main_obj.cpp:360
    got = d3.detect_cuboid(img, fixture_b["T"], boxes, lines.astype(np.float64))
    ref_lines = oracle.lsd_detect(img, 15.0)["lines"].astype(np.float64)
    ref = oracle.detect_cuboid(img, fixture_b["K"], fixture_b["T"], boxes, ref_lines, oracle.default_params(nominal_skew_ratio=2))
    assert len(got) == len(boxes) == 1
    assert got[0][0].proposal_index == int(ref["cuboids"][0][0]["proposal_index"])
    assert abs(got[0][0].normalized_error - float(ref["cuboids"][0][0]["normalized_error"])) < 1e-9


def test_object_slam_sequence_parity(det, oracle, fixture_b):
    """north_star: best-proposal-index parity on the shipped object_slam/data sequence, online mode
    (object_slam/src/main_obj.cpp:392-450): LSD lines (length > 15) -> detect_cuboid with nominal_skew_ratio 2,
    roll/pitch sampling on every frame but the first, all frames at the first frame's pose."""
    import cube_slam_b200 as cs
    frames = fixture_b["frames"]
    K, T = fixture_b["K"], fixture_b["T"]
    imgs = np.stack([f[0] for f in frames])
    gpu_lines = det.detect_filter_lines_batch(imgs)
    ctx = cs.Context(0, 640, 480, len(frames), 8, 4096)
    ctx.set_calibration(K)
    n_checked = 0
    for sampling, ids in ((0, [0]), (1, list(range(1, len(frames))))):
        p = cs.default_params(whether_sample_cam_roll_pitch=sampling, nominal_skew_ratio=2.0)
        out, counts = ctx.detect_batch_host(imgs[ids], np.stack([T] * len(ids)), [frames[i][1] for i in ids],
                                            [gpu_lines[i].astype(np.float64) for i in ids], p)
        o = 0
        for i in ids:
            ref_lines = oracle.lsd_detect(frames[i][0], 15.0)["lines"]
            np.testing.assert_array_equal(gpu_lines[i], ref_lines)
            ref = oracle.detect_cuboid(frames[i][0], K, T, frames[i][1], ref_lines.astype(np.float64),
                                       oracle.default_params(whether_sample_cam_roll_pitch=sampling, nominal_skew_ratio=2.0))
            for b in range(len(frames[i][1])):
                assert counts[o] == len(ref["cuboids"][b]), (i, b)
                if counts[o]:
                    g, r = out[o, 0], ref["cuboids"][b][0]
                    assert int(g["proposal_index"]) == int(r["proposal_index"]), i
                    assert abs(float(g["normalized_error"]) - float(r["normalized_error"])) < 1e-4
                    np.testing.assert_allclose(g["pos"], r["pos"], rtol=1e-9, atol=1e-9)
                    np.testing.assert_array_equal(g["box_corners_2d"], r["box_corners_2d"])
                    n_checked += 1
                o += 1
    assert n_checked >= 45  # 51 of the 58 frames carry a box
    ctx.close()


def test_online_batch_equals_two_calls(det, oracle):
    """cs_detect_frames_batch (lines stay on the device) == cs_detect_lines_batch followed by cs_detect_cuboids_batch."""
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    F = 5
    imgs, Ts, boxes, _, K = S.make_batch(51, F, 640, 480, 3, poisson=True)
    ctx = cs.Context(0, 640, 480, F, 16, 2048)
    ctx.set_calibration(K)
    p = cs.default_params(max_cuboid_num=2)
    lp = det.params()
    out1, cnt1 = ctx.detect_frames_host(imgs, Ts, boxes, lp, p)
    out1, cnt1 = out1.copy(), cnt1.copy()
    lines = det.detect_filter_lines_batch(imgs)
    out2, cnt2 = ctx.detect_batch_host(imgs, Ts, boxes, [l.astype(np.float64) for l in lines], p)
    np.testing.assert_array_equal(cnt1, cnt2)
    assert out1.tobytes() == out2.tobytes()
    # and both equal the oracle's two-stage result
    from test_gpu_cuboid_parity import _compare_box
    o = 0
    n_knife = 0
    for f in range(F):
        rl = oracle.lsd_detect(imgs[f], float(lp.line_length_thres))["lines"].astype(np.float64)

        def redo(cut_flip=-1, f=f, rl=rl):
            return oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], rl, oracle.default_params(max_cuboid_num=2), cut_flip=cut_flip)
        ref = redo()
        for b in range(len(boxes[f])):
            n_knife += _compare_box(oracle, out1[o], cnt1[o], ref, b, redo)
            o += 1
    assert n_knife == 0
    ctx.close()


def test_sequence_against_the_shipped_matlab_cuboids(det, fixture_b):
    """The CUDA path end to end (cs_detect_lines -> cs_detect_cuboids_batch) against the cuboids the reference's authors ship for this
    sequence (object_slam/data/detect_cuboids_saved.txt, MATLAB, local ground frame) at the per-frame poses of pop_cam_poses_saved.txt.
    A soft check (another Canny / DT in MATLAB): same cuboid up to the sampling grid; tests/test_oracle_matlab_crosscheck.py is the CPU twin."""
    import os
    import cube_slam_b200 as cs
    from scipy.spatial.transform import Rotation
    from conftest import GOLD
    fb = os.path.join(GOLD, "fixture_b")
    pop = np.loadtxt(os.path.join(fb, "pop_cam_poses_saved.txt"))
    sav = np.loadtxt(os.path.join(fb, "detect_cuboids_saved.txt"))
    ids = [int(r[0]) for r in sav]
    frames = fixture_b["frames"]
    imgs = np.stack([frames[i][0] for i in ids])
    Ts = []
    for i in ids:
        T = np.eye(4)
        T[:3, :3] = Rotation.from_quat(pop[i][4:8]).as_matrix()
        T[:3, 3] = pop[i][1:4]
        Ts.append(T)
    lines = det.detect_filter_lines_batch(imgs)
    ctx = cs.Context(0, 640, 480, len(ids), 8, 4096)
    ctx.set_calibration(fixture_b["K"])
    out, counts = ctx.detect_batch_host(imgs, np.stack(Ts), [frames[i][1] for i in ids], [l.astype(np.float64) for l in lines],
                                        cs.default_params(nominal_skew_ratio=2.0))
    assert list(counts) == [1] * len(ids)
    rows = []
    for k, row in enumerate(sav):
        c = out[k, 0]
        d_pos = float(np.linalg.norm(np.array(c["pos"]) - row[1:4]))
        d_yaw = (float(c["rotY"]) - row[4] + np.pi / 4) % (np.pi / 2) - np.pi / 4
        swapped = abs(((float(c["rotY"]) - row[4] + np.pi / 2) % np.pi) - np.pi / 2) > np.pi / 4
        sc = np.array(c["scale"])[[1, 0, 2]] if swapped else np.array(c["scale"])
        rows.append((d_pos, abs(d_yaw), float((np.abs(sc - row[5:8]) / row[5:8]).max())))
    rows = np.array(rows)
    step = 6.0 / 180 * np.pi
    assert np.median(rows[:, 0]) < 0.05 and np.median(rows[:, 1]) < 0.02 and np.median(rows[:, 2]) < 0.15
    good = (rows[:, 0] < 0.15) & (rows[:, 1] < 2.1 * step) & (rows[:, 2] < 0.3)
    assert good.sum() >= 42, int(good.sum())     # 43 of 51, the oracle's own count (tests/test_oracle_matlab_crosscheck.py)
    ctx.close()


def test_seed_loop_on_a_giant_region(det, oracle):
    """A smooth ramp is one line-support region far larger than the shared-memory part of the region list (it spills to HBM):
    result unchanged."""
    ramp = np.tile((np.arange(640) * 0.35).astype(np.uint8)[None, :], (480, 1))
    ramp = np.ascontiguousarray(ramp)
    got = det.detect_filter_lines_batch(ramp[None])
    _, redo = det.seed_loop_stats(1)
    ref = oracle.lsd_detect(ramp, 15.0)
    np.testing.assert_array_equal(got[0], ref["lines"])
    np.testing.assert_array_equal(det.debug_frame(0)["raw_lines"], ref["raw_lines"])


def test_tma_staged_tiles_equal_byte_staged_tiles(det, oracle):
    """k_lsd_blur<true> / k_ed_front<true> (BGR tiles fetched by the copy engine, the default on 640 / 1280 wide BGR frames) against the
    byte-load instantiations (cs_set_profiling bit 8): identical segments for both detectors, and equal to the oracle."""
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    imgs = S.make_batch(71, 3, 640, 480, 3)[0]
    for use_lsd, ref_fn in ((True, oracle.lsd_detect), (False, oracle.edl_detect)):
        d = cs.line_lbd_detect(context=det._ctx)
        d.use_LSD = use_lsd
        d.line_length_thres = 15
        got = {}
        for flags in (0, 256):
            det._ctx.set_profiling(flags)
            got[flags] = d.detect_filter_lines_batch(imgs)
        det._ctx.set_profiling(0)
        for f in range(3):
            np.testing.assert_array_equal(got[0][f], got[256][f])
            np.testing.assert_array_equal(got[0][f], ref_fn(imgs[f], 15.0)["lines"])
