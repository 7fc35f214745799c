"""Soft cross-check of the whole path (line detection -> cuboid proposals) against results the reference's authors ship.

`object_slam/data/detect_cuboids_saved.txt` holds the cuboid their MATLAB implementation detected in 51 of the 58 frames (`frame x y z
yaw l w h score`, in the local ground frame: camera above the origin, yaw 0) and `pop_cam_poses_saved.txt` the per-frame camera pose
(height, roll, pitch) those detections were made with (object_slam/src/main_obj.cpp:477-488,617-625 reads both in offline mode).  The
MATLAB code uses another Canny / distance transform (`detect_3d_cuboid/README.md:2`: "the final output might be slightly differently"),
so this is not equality: the oracle, run on the same frames, boxes and poses with LSD lines, must land on the same cuboid up to the
sampling grid -- most frames pick the SAME yaw sample, positions agree to a few centimetres."""
import os

import numpy as np
from scipy.spatial.transform import Rotation

from conftest import GOLD


def test_oracle_agrees_with_the_shipped_matlab_cuboids(oracle, fixture_b):
    fb = os.path.join(GOLD, "fixture_b")
    pop = np.loadtxt(os.path.join(fb, "pop_cam_poses_saved.txt"))      # time x y z qx qy qz qw
    sav = np.loadtxt(os.path.join(fb, "detect_cuboids_saved.txt"))     # frame x y z yaw l w h score
    assert pop.shape == (58, 8) and sav.shape == (51, 9)
    K = fixture_b["K"]
    rows = []
    for row in sav:
        i = int(row[0])
        img, boxes = fixture_b["frames"][i]
        assert len(boxes) == 1
        T = np.eye(4)
        T[:3, :3] = Rotation.from_quat(pop[i][4:8]).as_matrix()
        T[:3, 3] = pop[i][1:4]
        lines = oracle.lsd_detect(img, 15.0)["lines"].astype(np.float64)             # main_obj.cpp:364-366
        res = oracle.detect_cuboid(img, K, T, boxes, lines, oracle.default_params(nominal_skew_ratio=2.0))
        assert len(res["cuboids"][0]) == 1
        c = res["cuboids"][0][0]
        d_pos = float(np.linalg.norm(np.array(c["pos"]) - row[1:4]))
        d_yaw = (float(c["rotY"]) - row[4] + np.pi / 4) % (np.pi / 2) - np.pi / 4      # a cuboid has no front: compare modulo 90 degrees
        swapped = abs(((float(c["rotY"]) - row[4] + np.pi / 2) % np.pi) - np.pi / 2) > np.pi / 4
        sc = np.array(c["scale"])[[1, 0, 2]] if swapped else np.array(c["scale"])
        d_scale = float((np.abs(sc - row[5:8]) / row[5:8]).max())
        rows.append((d_pos, abs(d_yaw), d_scale))
    rows = np.array(rows)
    step = 6.0 / 180 * np.pi                                                           # the yaw sampling step
    assert np.median(rows[:, 0]) < 0.05 and np.percentile(rows[:, 0], 80) < 0.12      # metres (measured: 0.032 / 0.094)
    assert np.median(rows[:, 1]) < 0.02                                                # same yaw sample on most frames (0.003)
    assert (rows[:, 1] < 0.5 * step).sum() >= 25 and np.percentile(rows[:, 1], 80) < 1.1 * step
    assert np.median(rows[:, 2]) < 0.15                                                # relative size (0.10)
    good = (rows[:, 0] < 0.15) & (rows[:, 1] < 2.1 * step) & (rows[:, 2] < 0.3)
    print("matlab cross-check:", np.median(rows, axis=0), np.percentile(rows, 80, axis=0), int((rows[:, 1] < 0.5 * step).sum()), int(good.sum()))
    assert good.sum() >= 42, int(good.sum())                                           # 43 of 51
