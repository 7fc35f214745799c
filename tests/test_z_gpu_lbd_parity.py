"""GPU parity of the descriptor / matcher half of line_lbd_detect: cs_detect_descrip_lines, cs_lbd_compute,
cs_match_line_descrip through the Python mirror of the class, against the CPU oracle (oracle/lbd_oracle.cpp, itself equal to the compiled
reference: tests/test_oracle_ref_lbd.py) and against the committed goldens.

Key lines (end points, angle, lineLength, response, size, numOfPixels), the 32-byte descriptors, the 72-float descriptors and the matches
(which of several equally near codes included) must be identical -- nothing here is compared with a tolerance.

(The file sorts last on purpose: the kernels behind it were written after this round's GPU budget was spent, so the round-end run is their
first launch; the suite's earlier files do not depend on them.)"""
import json
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = (("start_x", "sx"), ("start_y", "sy"), ("end_x", "ex"), ("end_y", "ey"), ("angle", "angle"), ("line_length", "line_length"),
          ("response", "response"), ("size", "size"), ("num_pixels", "num_pixels"))


@pytest.fixture(scope="module")
def det():
    import cube_slam_b200 as cs
    d = cs.line_lbd_detect()
    d.line_length_thres = 15
    return d


def _same_keylines(got, want):
    assert len(got) == len(want)
    for a, b in FIELDS:
        np.testing.assert_array_equal(got[a], want[b], err_msg=a)
    np.testing.assert_array_equal(got["class_id"], np.arange(len(got)))


@pytest.mark.parametrize("use_lsd,n", [(True, 271), (False, 88)])
def test_detect_descrip_lines_demo_frame(det, oracle, fixture_a, use_lsd, n):
    det.use_LSD = use_lsd
    img = fixture_a["img"]
    kl, desc = det.detect_descrip_lines(img)
    want = oracle.lbd_detect_keylines(img, use_lsd, 15.0)
    assert len(kl) == n
    _same_keylines(kl, want)
    np.testing.assert_array_equal(desc, oracle.lbd_compute(img, want))
    # the float descriptor behind the bytes, through BinaryDescriptor::compute on the same key lines
    d2, f2 = det.compute_descriptors(img, kl, want_float=True)
    wd, wf = oracle.lbd_compute(img, want, want_float=True)
    np.testing.assert_array_equal(d2, wd)
    np.testing.assert_array_equal(f2, wf)


@pytest.mark.parametrize("use_lsd", [True, False])
def test_batches_of_sequence_and_synthetic_frames(det, oracle, fixture_b, use_lsd):
    from cube_slam_b200 import synthetic as S
    det.use_LSD = use_lsd
    for imgs in (np.stack([fixture_b["frames"][i][0] for i in (0, 9, 17, 33, 40)]), S.make_batch(31, 4, 640, 480, 3)[0],
                 S.make_batch(32, 2, 1242, 375, 3, kind="kitti")[0]):
        out = det.detect_descrip_lines_batch(imgs)
        assert len(out) == len(imgs)
        for f, (kl, desc) in enumerate(out):
            want = oracle.lbd_detect_keylines(imgs[f], use_lsd, 15.0)
            _same_keylines(kl, want)
            np.testing.assert_array_equal(desc, oracle.lbd_compute(imgs[f], want))


@pytest.mark.parametrize("use_lsd", [True, False])
def test_mat_overload_octaves_variant_and_gray_input(det, oracle, fixture_a, use_lsd):
    import cv2
    img = fixture_a["img"]
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    for use_lsd in (use_lsd,):
        det.use_LSD = use_lsd
        # detect_descrip_lines(gray, lines_mat, descrips): every octave-0 line, no length filter (line_lbd_allclass.cpp:224-250)
        lines, desc = det.detect_descrip_lines(img, as_mat=True)
        want = oracle.lbd_detect_keylines(img, use_lsd, -1.0)
        assert det.line_length_thres == 15 and len(lines) == len(want) == (445 if use_lsd else 88)   # 271 / 88 of them are longer than 15
        np.testing.assert_array_equal(lines, np.stack([want["sx"], want["sy"], want["ex"], want["ey"]], 1))
        np.testing.assert_array_equal(desc, oracle.lbd_compute(img, want))
        # detect_descrip_lines_octaves (:285-339): start x <= end x
        kls, descs = det.detect_descrip_lines_octaves(img)
        w15 = oracle.lbd_detect_keylines(img, use_lsd, 15.0)
        _same_keylines(kls[0], oracle.lbd_order_keylines(w15))
        np.testing.assert_array_equal(descs[0], oracle.lbd_compute(img, w15))
        assert (kls[0]["start_x"] <= kls[0]["end_x"]).all()
        # a one-channel input goes the same way (computeImpl :608-612 converts a colour one itself)
        kg, dg = det.detect_descrip_lines(gray)
        wg = oracle.lbd_detect_keylines(gray, use_lsd, 15.0)
        _same_keylines(kg, wg)
        np.testing.assert_array_equal(dg, oracle.lbd_compute(gray, wg))


def test_given_keylines_border_and_degenerate_lines(det, oracle):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    img[30:90, 40:120] = 220
    rows = np.array([[2, 2, 150, 3], [5, 110, 5, 4], [158, 1, 158, 118], [0, 0, 159, 119], [80, 60, 80.4, 60.2], [10, 117, 150, 118.5],
                     [40, 30, 120, 30], [40.5, 90.2, 120.3, 89.7]], np.float32)
    rows = np.concatenate([rows, rng.uniform(0, 1, (300, 4)).astype(np.float32) * np.array([159, 119, 159, 119], np.float32)])
    kl = det.keylines_from_lines(rows, 160, 120)
    want = oracle.lbd_keylines_from_lsd(rows, 160, 120)
    _same_keylines(kl, want)
    d, f = det.compute_descriptors(img, kl, want_float=True)
    wd, wf = oracle.lbd_compute(img, want, want_float=True)
    np.testing.assert_array_equal(d, wd)
    np.testing.assert_array_equal(f, wf)                       # NaNs of one-pixel lines in the same places
    np.testing.assert_array_equal(det.get_line_descriptors(img, rows), wd)
    assert det.compute_descriptors(img, kl[:0]).shape == (0, 32)   # "keypoint list is empty": nothing computed, no error


def test_match_line_descrip(det, oracle, fixture_b):
    rng = np.random.default_rng(11)

    def flip(c, bits):
        c = c.copy()
        for b in bits:
            c[b // 8] ^= np.uint8(1 << (b % 8))
        return c

    qs, ts = [], []
    for trial in range(20):
        nq, nt = int(rng.integers(1, 60)), int(rng.integers(6, 400))
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        q = np.stack([flip(t[int(rng.integers(0, nt))], rng.integers(0, 256, int(rng.integers(0, 40)))) for _ in range(nq)])
        k = int(rng.integers(1, 7))
        for j in range(min(6, nt)):                       # several train codes at the same distance from query 0
            t[(j * 7) % nt] = flip(q[0], [int(x) for x in rng.choice(256, k, replace=False)])
        t[nt - 1] = t[0]                                   # an exact duplicate: the lower train index wins
        qs.append(q)
        ts.append(t)
    qs.append(rng.integers(0, 256, (100, 32), dtype=np.uint8))     # unrelated codes: nearest neighbours around 100 .. 130 bits away
    ts.append(rng.integers(0, 256, (5, 32), dtype=np.uint8))
    for thres in (25.0, 300.0):
        batch = det.match_line_descrip_batch(qs, ts, thres)
        for m, q, t in zip(batch, qs, ts):
            wq, wt, wd = oracle.lbd_match(q, t, thres)
            np.testing.assert_array_equal(m["query_idx"], wq)
            np.testing.assert_array_equal(m["train_idx"], wt)
            np.testing.assert_array_equal(m["distance"], wd)
            assert (m["img_idx"] == 0).all()
        one = det.match_line_descrip(qs[3], ts[3], thres)
        np.testing.assert_array_equal(one, batch[3])
    assert len(det.match_line_descrip(qs[0], ts[0][:0])) == 0 and len(det.match_line_descrip(qs[0][:0], ts[0])) == 0
    # descriptors of two frames of the real sequence, matched on the device
    det.use_LSD = True
    (k0, d0), (k1, d1) = det.detect_descrip_lines_batch(np.stack([fixture_b["frames"][0][0], fixture_b["frames"][1][0]]))
    m = det.match_line_descrip(d0, d1, 40.0)
    wq, wt, wd = oracle.lbd_match(d0, d1, 40.0)
    np.testing.assert_array_equal(m["query_idx"], wq)
    np.testing.assert_array_equal(m["train_idx"], wt)
    assert len(m) >= 3


@pytest.mark.parametrize("flavour", ["lsd", "edlines"])
def test_against_the_committed_goldens(det, flavour):
    """No oracle in the loop: counts and checksums recorded by tools/make_golden_lbd.py after oracle == compiled reference held."""
    import cv2
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "expected_lbd.json")))
    for case in g["frames"]:
        img = cv2.imread(os.path.join(os.path.dirname(__file__), "golden", case["image"]), 1)
        for flav in (flavour,):
            det.use_LSD = flav == "lsd"
            kl, desc = det.detect_descrip_lines(img)
            assert len(kl) == case[flav]["n"]
            assert zlib.crc32(np.ascontiguousarray(desc).tobytes()) == case[flav]["desc_crc32"]
            assert zlib.crc32(np.ascontiguousarray(kl["angle"]).tobytes()) == case[flav]["angle_crc32"]
            assert int(kl["num_pixels"].sum()) == case[flav]["num_pixels_sum"]
    if flavour != "lsd":
        return
    det.use_LSD = True
    imgs = [cv2.imread(os.path.join(os.path.dirname(__file__), "golden", p), 1) for p in g["match"]["images"]]
    (k0, d0), (k1, d1) = det.detect_descrip_lines_batch(np.stack(imgs))
    m = det.match_line_descrip(d0, d1, g["match"]["thres"])
    assert [list(map(int, x)) for x in zip(m["query_idx"], m["train_idx"], m["distance"])] == g["match"]["triples"]


SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libshim_line.so")


@pytest.mark.parametrize("use_lsd", [1, 0])
def test_cpp_shim_descriptor_and_matcher_members(oracle, fixture_a, fixture_b, use_lsd):
    """shim/line_lbd_b200.cpp's descriptor / matcher members, compiled against the reference's own class header and called as a user of
    class line_lbd_detect calls them (shim/test/line_shim_driver.cpp -> oracle/_ref/libshim_line.so): detect_descrip_lines (both
    overloads), detect_descrip_lines_octaves, get_line_descriptors, match_line_descrip."""
    import ctypes as C
    if not os.path.exists(SHIM):
        pytest.skip("oracle/_ref/libshim_line.so not built (needs the reference's headers at build time)")
    import cube_slam_b200  # noqa: F401
    L = C.CDLL(SHIM)
    for name in ("shim_line_detect_descrip", "shim_line_descriptors_of", "shim_line_match"):
        getattr(L, name).restype = C.c_int
    img = np.ascontiguousarray(fixture_a["img"], np.uint8)
    h, w, ch = img.shape
    u8 = C.POINTER(C.c_uint8)
    cap = 8192
    res = {}
    for mode in (0, 1, 2):
        kl = np.zeros(cap, oracle.KEYLINE_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = L.shim_line_detect_descrip(img.ctypes.data_as(u8), w, h, ch, use_lsd, C.c_float(15.0), mode, kl.ctypes.data_as(C.c_void_p),
                                       desc.ctypes.data_as(u8), cap)
        assert n >= 0
        res[mode] = (kl[:n], desc[:n])
    want = oracle.lbd_detect_keylines(img, bool(use_lsd), 15.0)
    wdesc = oracle.lbd_compute(img, want)
    for mode, wk in ((0, want), (1, oracle.lbd_order_keylines(want))):
        kl, desc = res[mode]
        assert len(kl) == len(wk)
        for f in ("sx", "sy", "ex", "ey", "angle", "line_length", "response", "size", "num_pixels"):
            np.testing.assert_array_equal(kl[f], wk[f], err_msg="mode %d %s" % (mode, f))
        np.testing.assert_array_equal(kl["class_id"], np.arange(len(kl)))
        np.testing.assert_array_equal(desc, wdesc)
    wall = oracle.lbd_detect_keylines(img, bool(use_lsd), -1.0)
    kl, desc = res[2]
    assert len(kl) == len(wall)
    for f in ("sx", "sy", "ex", "ey"):
        np.testing.assert_array_equal(kl[f], wall[f])
    np.testing.assert_array_equal(desc, oracle.lbd_compute(img, wall))
    # get_line_descriptors on the rows detect_filter_lines returned
    rows = np.ascontiguousarray(np.stack([want["sx"], want["sy"], want["ex"], want["ey"]], 1), np.float32)
    d2 = np.zeros((len(rows), 32), np.uint8)
    assert L.shim_line_descriptors_of(img.ctypes.data_as(u8), w, h, ch, rows.ctypes.data_as(C.POINTER(C.c_float)), len(rows), d2.ctypes.data_as(u8)) == len(rows)
    np.testing.assert_array_equal(d2, oracle.lbd_compute(img, oracle.lbd_keylines_from_lsd(rows, w, h)))
    # match_line_descrip between two frames of the sequence
    f0, f1 = fixture_b["frames"][0][0], fixture_b["frames"][1][0]
    d0 = oracle.lbd_compute(f0, oracle.lbd_detect_keylines(f0, True, 15.0))
    d1 = oracle.lbd_compute(f1, oracle.lbd_detect_keylines(f1, True, 15.0))
    qi, ti, di = np.zeros(len(d0), np.int32), np.zeros(len(d0), np.int32), np.zeros(len(d0), np.float32)
    n = L.shim_line_match(d0.ctypes.data_as(u8), len(d0), d1.ctypes.data_as(u8), len(d1), C.c_float(40.0), qi.ctypes.data_as(C.POINTER(C.c_int32)),
                          ti.ctypes.data_as(C.POINTER(C.c_int32)), di.ctypes.data_as(C.POINTER(C.c_float)))
    wq, wt, wd = oracle.lbd_match(d0, d1, 40.0)
    assert n == len(wq) >= 3
    np.testing.assert_array_equal(qi[:n], wq)
    np.testing.assert_array_equal(ti[:n], wt)
    np.testing.assert_array_equal(di[:n], wd)


def test_detect_raw_lines_with_and_without_downsampling(det, oracle, fixture_a):  # verified kernels only (cs_detect_lines_batch), both flavours
    """detect_raw_lines(gray, lines_mat, downsample_img) (line_lbd_allclass.cpp:174-189) through the Python mirror."""
    import cv2
    img = fixture_a["img"]
    half = cv2.resize(img, None, fx=0.5, fy=0.5)
    for use_lsd, fn in ((True, oracle.lsd_detect), (False, oracle.edl_detect)):
        det.use_LSD = use_lsd
        np.testing.assert_array_equal(det.detect_raw_lines(img), fn(img, -1.0)["lines"])
        np.testing.assert_array_equal(det.detect_raw_lines(img, downsample_img=True), fn(half, -1.0)["lines"] * np.float32(2))
    assert det.line_length_thres == 15


# ---------------------------------------------------------------------------------------------------------------------------------------
# Not LBD, but new this session and not yet run on a GPU either, so it lives in the file that sorts last: several boxes in a roll / pitch-sampled
# frame (cs_set_profiling bit 10).
def test_sampled_frames_with_several_boxes_default_and_carried_pose(oracle):
    """Frames picked on the CPU (tests/test_sampling_deviation.py): in (seed 102, frame 10) and (seed 104, frame 1) the pose the reference
    carries from box to box gives the third box 16 yaw samples where the raw pose gives 15 (or the other way round), and a different best
    proposal.  Default mode: every box starts from the raw pose == the oracle with orc_set_independent_boxes(1).  Bit 10: the library carries
    the pose == the oracle as it is (== the compiled reference).  Both strictly, every box, every kept cuboid."""
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    from test_gpu_cuboid_parity import _compare_cuboid
    L = oracle.lib()
    for seed, pick in ((102, (10, 3)), (104, (1, 0))):
        imgs, Ts, boxes, lines, K = S.make_batch(seed, 12, 640, 480, 3, poisson=True)
        sel = list(pick)
        imgs, Ts = imgs[sel], Ts[sel]
        boxes, lines = [boxes[i] for i in sel], [lines[i] for i in sel]
        for kw in (dict(whether_sample_cam_roll_pitch=1, max_cuboid_num=3), dict(whether_sample_cam_roll_pitch=1, whether_sample_bbox_height=1, max_cuboid_num=2)):
            differs = False
            for carried in (0, 1):
                ctx = cs.Context(0, 640, 480, 2, 16, 4096)
                ctx.set_calibration(K)
                ctx.L.cs_set_profiling(ctx.h, 1024 if carried else 0)
                out, counts = ctx.detect_batch_host(imgs, Ts, boxes, lines, cs.default_params(**kw))
                ctx.close()
                o = 0
                try:
                    L.orc_set_independent_boxes(0 if carried else 1)
                    for f in range(len(sel)):
                        ref = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], np.asarray(lines[f], float), oracle.default_params(**kw))
                        for b in range(len(boxes[f])):
                            assert counts[o] == len(ref["cuboids"][b]), (seed, f, b, carried)
                            for k in range(counts[o]):
                                _compare_cuboid(out[o, k], ref["cuboids"][b][k])
                            o += 1
                finally:
                    L.orc_set_independent_boxes(0)
                if carried:
                    differs = differs or out.tobytes() != prev
                prev = out.tobytes()
            assert differs, "the two modes returned the same records on frames chosen because they differ"


def test_class_mirror_carries_the_pose_like_the_reference(oracle):
    """cs.detect_3d_cuboid (the Python mirror of the class, like the C++ shim) turns bit 10 on: a sampled frame with three boxes gives the
    oracle's -- the reference's -- cuboids for every box, on a frame where starting from the raw pose would not."""
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    from test_gpu_cuboid_parity import _compare_cuboid
    imgs, Ts, boxes, lines, K = S.make_batch(102, 12, 640, 480, 3, poisson=True)
    f = 10
    det = cs.detect_3d_cuboid()
    det.set_calibration(K)
    det.whether_sample_cam_roll_pitch = True
    det.max_cuboid_num = 3
    got = det.detect_cuboid(imgs[f], Ts[f], boxes[f], lines[f])
    ref = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], np.asarray(lines[f], float), oracle.default_params(whether_sample_cam_roll_pitch=1, max_cuboid_num=3))
    assert len(got) == len(boxes[f]) == 3
    for b in range(3):
        assert len(got[b]) == len(ref["cuboids"][b])
        for k in range(len(got[b])):
            _compare_cuboid({name: getattr(got[b][k], name) for name in got[b][k].__slots__}, ref["cuboids"][b][k])
