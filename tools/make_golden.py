#!/usr/bin/env python
"""Generates tests/golden/ from the read-only reference checkout (run in the build container only).

    python tools/make_golden.py [/root/reference]

Writes
  fixture_a/   inputs of the reference's single-frame demo (detect_3d_cuboid/src/main.cpp:35-48):
               0000_rgb_raw.jpg, 0000_edge.txt, meta.json (K, pose, bbox)
  fixture_b/   the 58-frame object_slam/data sequence: raw_imgs/*.jpg, filter_2d_obj_txts/*.txt, meta.json
               (K from main_obj.cpp:346-349, pose = truth_cam_poses.txt row 0 -> SE3 matrix)
  cv_pins.npz  third-party pins: cv2 4.x outputs (cvtColor / Canny / distanceTransform with
               setUseOptimized(False)) for seeded inputs -- the vectors the oracle's OpenCV restatements
               are checked against even where cv2 is not importable
  expected_*.json  oracle outputs for the fixtures (regression pins of the oracle itself)
The reference ships no expected outputs of its own, so these are the goldens.
"""
import json
import os
import shutil
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def quat_pose_to_T(row):
    """g2o::SE3Quat(Vector7d [x y z qx qy qz qw]).to_homogeneous_matrix() (main_obj.cpp:384)."""
    t = np.asarray(row[:3], float)
    qx, qy, qz, qw = row[3:7]
    n = np.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    qx, qy, qz, qw = qx / n, qy / n, qz / n, qw / n
    R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                  [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                  [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def cub_summary(c):
    return dict(proposal_index=int(c["proposal_index"]), height_sample_id=int(c["height_sample_id"]),
                normalized_error=float(c["normalized_error"]), combined_score=float(c["combined_score"]),
                edge_distance_error=float(c["edge_distance_error"]), edge_angle_error=float(c["edge_angle_error"]),
                pos=c["pos"].tolist(), scale=c["scale"].tolist(), rotY=float(c["rotY"]),
                box_config_type=c["box_config_type"].tolist(), box_corners_2d=c["box_corners_2d"].tolist())


def main():
    os.makedirs(GOLD, exist_ok=True)
    # ---- fixture A
    fa = os.path.join(GOLD, "fixture_a")
    os.makedirs(fa, exist_ok=True)
    shutil.copy(os.path.join(REF, "detect_3d_cuboid/data/0000_rgb_raw.jpg"), fa)
    shutil.copy(os.path.join(REF, "detect_3d_cuboid/data/edge_detection/LSD/0000_edge.txt"), fa)
    K = [[529.5, 0, 365.0], [0, 529.5, 265.0], [0, 0, 1.0]]
    T = [[1, 0.0011, 0.0004, 0], [0, -0.3376, 0.9413, 0], [0.0011, -0.9413, -0.3376, 1.35], [0, 0, 0, 1]]
    bbox = [[188 - 1, 189 - 1, 201, 311, 0.88]]  # main.cpp:46-48: matlab -> c++ coordinates
    json.dump(dict(K=K, T=T, boxes=bbox, source="detect_3d_cuboid/src/main.cpp:35-48"), open(os.path.join(fa, "meta.json"), "w"), indent=1)
    img = cv2.imread(os.path.join(fa, "0000_rgb_raw.jpg"), 1)
    lines = np.loadtxt(os.path.join(fa, "0000_edge.txt"))
    exp = {}
    for name, kw in [("default", {}), ("sample_roll_pitch", dict(whether_sample_cam_roll_pitch=1)),
                     ("sample_height_top5", dict(whether_sample_bbox_height=1, max_cuboid_num=5)),
                     ("config1_only", dict(consider_config_2=0)), ("config2_only", dict(consider_config_1=0))]:
        r = O.detect_cuboid(img, K, T, bbox, lines, O.default_params(**kw), trace_object=0)
        tr = r["trace"]
        exp[name] = dict(n_candidates=r["n_candidates"], n_valid=r["n_valid"], roi=list(tr["roi"]), n_lines_roi=tr["n_lines_roi"],
                         n_lines_merged=tr["n_lines_merged"], n_kept=tr["n_kept"], canny_pixels=int((tr["canny"] > 0).sum()),
                         dist_sum=float(tr["dist"].astype(np.float64).sum()), cuboids=[cub_summary(c) for c in r["cuboids"][0]])
    json.dump(exp, open(os.path.join(GOLD, "expected_fixture_a.json"), "w"), indent=1)

    # ---- fixture B
    fb = os.path.join(GOLD, "fixture_b")
    os.makedirs(os.path.join(fb, "raw_imgs"), exist_ok=True)
    os.makedirs(os.path.join(fb, "filter_2d_obj_txts"), exist_ok=True)
    src = os.path.join(REF, "object_slam/data")
    n = 0
    for f in sorted(os.listdir(os.path.join(src, "raw_imgs"))):
        shutil.copy(os.path.join(src, "raw_imgs", f), os.path.join(fb, "raw_imgs"))
        n += 1
    for f in sorted(os.listdir(os.path.join(src, "filter_2d_obj_txts"))):
        shutil.copy(os.path.join(src, "filter_2d_obj_txts", f), os.path.join(fb, "filter_2d_obj_txts"))
    # the authors' own (MATLAB) cuboids for this sequence and the per-frame local-ground-frame camera poses they go with:
    # a soft cross-check of the whole path (tests/test_oracle_matlab_crosscheck.py)
    for f in ("detect_cuboids_saved.txt", "pop_cam_poses_saved.txt"):
        shutil.copy(os.path.join(src, f), fb)
    poses = np.loadtxt(os.path.join(src, "truth_cam_poses.txt"))
    Tb = quat_pose_to_T(poses[0, 1:8])
    Kb = [[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]]
    json.dump(dict(K=Kb, T=Tb.tolist(), n_frames=n, truth_row0=poses[0].tolist(),
                   source="object_slam/src/main_obj.cpp:346-349,384,436-447; object_slam/data"), open(os.path.join(fb, "meta.json"), "w"), indent=1)

    # ---- third-party pins from cv2 (OpenCV 4.x, non-IPP paths)
    cv2.setUseOptimized(False)
    cv2.setNumThreads(1)
    rng = np.random.default_rng(20260922)
    bgr = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    gray_pin = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (120, 160), dtype=np.uint8), (7, 7), 2.0)
    roi = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)[168:168 + 351, 167:167 + 241].copy()
    pins = dict(bgr=bgr, gray=gray_pin, smooth=smooth, smooth_canny=cv2.Canny(smooth, 30, 90), roi=roi, roi_canny=cv2.Canny(roi, 80, 200))
    pins["smooth_dt"] = cv2.distanceTransform(255 - pins["smooth_canny"], cv2.DIST_L2, 3)
    pins["roi_dt"] = cv2.distanceTransform(255 - pins["roi_canny"], cv2.DIST_L2, 3)
    pins["empty_dt"] = cv2.distanceTransform(np.full((9, 11), 255, np.uint8), cv2.DIST_L2, 3)
    np.savez_compressed(os.path.join(GOLD, "cv_pins.npz"), cv2_version=np.array(cv2.__version__), **pins)
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
