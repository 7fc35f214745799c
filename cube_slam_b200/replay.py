"""Replays a dataset laid out like the reference's data folders through the library.

    python -m cube_slam_b200.replay DATA_DIR [--offline-edges] [--out cuboids.txt] [--save-edges DIR] [--batch N]

DATA_DIR layout (object_slam/data and detect_3d_cuboid/data of the reference):
    raw_imgs/%04d_rgb_raw.jpg  (or %04d_rgb_raw.jpg in DATA_DIR itself)
    filter_2d_obj_txts/%04d_yolo2_0.15.txt     rows `x y w h prob`, 1-based x,y (main_obj.cpp:436-439)
    edge_detection/LSD/%04d_edge.txt           rows `x1 y1 x2 y2` (only with --offline-edges; Tracking.cc:1583-1590)
    meta.json                                  {"K": 3x3, "T": 4x4}  camera intrinsics and camera-to-world pose
Online mode runs the LSD line detector (use_LSD, length > 15) like object_slam (main_obj.cpp:364-366,428), samples
camera roll/pitch on every frame but the first (main_obj.cpp:442) and uses nominal_skew_ratio 2 (main_obj.cpp:360).
Output rows follow object_slam/data/detect_cuboids_saved.txt: `frame x y z yaw l w h score`.
Text parsing mirrors read_all_number_txt (matrix_utils.cpp:197-232): whitespace separated numbers, empty lines skipped.
"""
import argparse
import json
import os
import sys

import numpy as np


def read_all_number_txt(path, n_cols):
    """matrix_utils.cpp:197-232: every non-empty line is a row; missing trailing columns stay 0."""
    rows = []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            vals = []
            for tok in line.split():
                try:
                    vals.append(float(tok))
                except ValueError:
                    break
            row = np.zeros(n_cols)
            row[:min(len(vals), n_cols)] = vals[:n_cols]
            rows.append(row)
    return np.array(rows, np.float64).reshape(-1, n_cols)


def write_edges_txt(path, lines):
    """line_lbd/src/detect_lines.cpp:85-96: `x1\\ty1\\tx2\\ty2` per row."""
    with open(path, "w") as f:
        for l in lines:
            f.write("%s\t%s\t%s\t%s\n" % tuple(repr(float(v)) for v in l))


def find_frames(data_dir):
    img_dir = os.path.join(data_dir, "raw_imgs")
    if not os.path.isdir(img_dir):
        img_dir = data_dir
    ids = sorted(int(f[:4]) for f in os.listdir(img_dir) if f.endswith("_rgb_raw.jpg") and f[:4].isdigit())
    return img_dir, ids


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("data_dir")
    ap.add_argument("--offline-edges", action="store_true", help="read edge_detection/LSD/%%04d_edge.txt instead of detecting lines")
    ap.add_argument("--out", default=None, help="write `frame x y z yaw l w h score` rows here")
    ap.add_argument("--save-edges", default=None, help="directory for %%04d_edge.txt files of the detected lines")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-sampling", action="store_true", help="never sample camera roll/pitch")
    args = ap.parse_args(argv)

    import cv2
    import cube_slam_b200 as cs

    meta = json.load(open(os.path.join(args.data_dir, "meta.json")))
    K, T = np.array(meta["K"], float), np.array(meta["T"], float)
    img_dir, ids = find_frames(args.data_dir)
    if not ids:
        sys.exit("no %04d_rgb_raw.jpg frames under " + args.data_dir)
    frames, boxes = [], []
    for i in ids:
        img = cv2.imread(os.path.join(img_dir, "%04d_rgb_raw.jpg" % i), 1)
        frames.append(img)
        bpath = os.path.join(args.data_dir, "filter_2d_obj_txts", "%04d_yolo2_0.15.txt" % i)
        b = read_all_number_txt(bpath, 5) if os.path.exists(bpath) else np.zeros((0, 5))
        b[:, :2] -= 1  # matlab -> c++ coordinates
        boxes.append(b)
    h, w = frames[0].shape[:2]
    ctx = cs.Context(args.device, w, h, args.batch, 32, 8192)
    ctx.set_calibration(K)
    det = cs.line_lbd_detect(context=ctx)
    det.use_LSD = True
    det.line_length_thres = 15

    rows = []
    groups = ([[0]] if not args.no_sampling else []) + [list(range(0 if args.no_sampling else 1, len(ids)))]
    for grp_i, grp in enumerate(groups):
        sampling = 0 if (args.no_sampling or grp_i == 0) else 1
        p = cs.default_params(whether_sample_cam_roll_pitch=sampling, nominal_skew_ratio=2.0)
        for s in range(0, len(grp), args.batch):
            sel = grp[s:s + args.batch]
            imgs = np.stack([frames[k] for k in sel])
            Ts = np.stack([T] * len(sel))
            bl = [boxes[k] for k in sel]
            if args.offline_edges:
                ll = [read_all_number_txt(os.path.join(args.data_dir, "edge_detection", "LSD", "%04d_edge.txt" % ids[k]), 4) for k in sel]
                out, cnt = ctx.detect_batch_host(imgs, Ts, bl, ll, p)
            else:
                if args.save_edges:
                    os.makedirs(args.save_edges, exist_ok=True)
                    for k, l in zip(sel, det.detect_filter_lines_batch(imgs)):
                        write_edges_txt(os.path.join(args.save_edges, "%04d_edge.txt" % ids[k]), l)
                out, cnt = ctx.detect_frames_host(imgs, Ts, bl, det.params(), p)
            o = 0
            for k in sel:
                for _ in range(len(boxes[k])):
                    if cnt[o] > 0:
                        c = out[o, 0]
                        rows.append([ids[k], c["pos"][0], c["pos"][1], c["pos"][2], c["rotY"], c["scale"][0], c["scale"][1], c["scale"][2],
                                     c["normalized_error"]])
                    o += 1
    rows.sort(key=lambda r: r[0])
    text = "".join("%d\t" % r[0] + "\t".join("%.6f" % v for v in r[1:]) + "\n" for r in rows)
    if args.out:
        open(args.out, "w").write(text)
    else:
        sys.stdout.write(text)
    return rows


if __name__ == "__main__":
    main()
