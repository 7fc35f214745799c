"""The oracle's OpenMP batch driver (oracle/batch_oracle.cpp, the CPU arm of bench.py) against the per-frame oracle calls."""
import numpy as np


def test_batch_driver_equals_per_frame_calls(oracle):
    from cube_slam_b200 import synthetic as S
    imgs, Ts, boxes, lines, K = S.make_batch(5, 5, 640, 480, 2)
    p = oracle.default_params()
    per = [oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines[f], p) for f in range(len(imgs))]
    for nt in (1, 3):
        got = oracle.detect_frames_batch(imgs, K, Ts, boxes, lines, p, line_mode=0, n_threads=nt, want_records=True)
        assert got["n_valid"].tolist() == [r["n_valid"] for r in per]
        assert got["n_cand"].tolist() == [r["n_candidates"] for r in per]
        o = 0
        for f, r in enumerate(per):
            for b in range(len(boxes[f])):
                assert got["counts"][o] == len(r["cuboids"][b])
                if got["counts"][o]:
                    assert got["records"][o, 0]["proposal_index"] == r["cuboids"][b][0]["proposal_index"]
                    assert got["records"][o, 0]["normalized_error"] == r["cuboids"][b][0]["normalized_error"]
                o += 1


def test_batch_driver_online_modes(oracle):
    from cube_slam_b200 import synthetic as S
    imgs, Ts, boxes, _, K = S.make_batch(6, 3, 640, 480, 2)
    p = oracle.default_params()
    for mode, detect in ((1, oracle.lsd_detect), (2, oracle.edl_detect)):
        got = oracle.detect_frames_batch(imgs, K, Ts, boxes, None, p, line_mode=mode, line_length_thres=15.0, n_threads=2)
        for f in range(len(imgs)):
            seg = detect(imgs[f], 15.0)["lines"].astype(np.float64)
            ref = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], seg, p)
            assert got["n_lines"][f] == len(seg)
            assert got["n_valid"][f] == ref["n_valid"]
