"""The dataset replay CLI (wire formats of the reference's data folders)."""
import os

import numpy as np
import pytest

from conftest import GOLD


def test_read_all_number_txt_semantics(tmp_path):
    from cube_slam_b200.replay import read_all_number_txt, write_edges_txt
    p = tmp_path / "a.txt"
    p.write_text("1 2 3 4\n\n5\t6\t7\t8\n9 10\n")
    m = read_all_number_txt(str(p), 4)
    np.testing.assert_array_equal(m, [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 0, 0]])
    e = tmp_path / "e.txt"
    lines = np.array([[1.5, 2.25, 300.125, 4.0]], np.float32)
    write_edges_txt(str(e), lines)
    np.testing.assert_array_equal(read_all_number_txt(str(e), 4), lines.astype(np.float64))
    # the reference's own shipped edge file parses to 271 x 4
    ref = read_all_number_txt(os.path.join(GOLD, "fixture_a", "0000_edge.txt"), 4)
    assert ref.shape == (271, 4)


@pytest.mark.gpu
def test_replay_object_slam_sequence(tmp_path, oracle, fixture_b):
    from cube_slam_b200 import replay
    out = tmp_path / "cubes.txt"
    rows = replay.main([os.path.join(GOLD, "fixture_b"), "--out", str(out), "--save-edges", str(tmp_path / "edges")])
    got = np.loadtxt(str(out), ndmin=2)
    assert got.shape[1] == 9 and len(got) == len(rows) >= 45
    # spot-check three frames against the oracle's two-stage result
    for fi in (0, 9, 33):
        img, boxes = fixture_b["frames"][fi]
        if len(boxes) == 0:
            continue
        lines = oracle.lsd_detect(img, 15.0)["lines"]
        saved = replay.read_all_number_txt(str(tmp_path / "edges" / ("%04d_edge.txt" % fi)), 4)
        np.testing.assert_array_equal(saved.astype(np.float32), lines)
        ref = oracle.detect_cuboid(img, fixture_b["K"], fixture_b["T"], boxes, lines.astype(np.float64),
                                   oracle.default_params(whether_sample_cam_roll_pitch=int(fi != 0), nominal_skew_ratio=2.0))
        row = got[got[:, 0] == fi][0]
        c = ref["cuboids"][0][0]
        np.testing.assert_allclose(row[1:4], c["pos"], atol=2e-6)
        np.testing.assert_allclose(row[4], c["rotY"], atol=2e-6)
        np.testing.assert_allclose(row[8], c["normalized_error"], atol=2e-6)
