"""Writes tests/golden/expected_lbd.json: counts and checksums of the descriptor / matcher half of line_lbd_detect
on committed fixture frames, from the CPU oracle -- and only after the oracle's key lines, descriptors and matches were found EQUAL to the
reference's own code (oracle/_ref/liblinelbd_ref.so, compiled from /root/reference) on every one of them; the script refuses to write
otherwise, and refuses to run where that library is missing.  The GPU test (tests/test_z_gpu_lbd_parity.py) and the CPU test
(tests/test_goldens_lbd.py) check against the file without the reference in the loop.

    python tools/make_golden_lbd.py
"""
import json
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
IMAGES = ["fixture_a/0000_rgb_raw.jpg"] + ["fixture_b/raw_imgs/%04d_rgb_raw.jpg" % i for i in (0, 1, 17, 40)]
MATCH = dict(images=["fixture_b/raw_imgs/0000_rgb_raw.jpg", "fixture_b/raw_imgs/0001_rgb_raw.jpg"], thres=40.0)
FIELDS = ("sx", "sy", "ex", "ey", "angle", "line_length", "response", "size", "num_pixels")


def main():
    if not O.ref_detect_filter_lines_available():
        raise SystemExit("oracle/_ref/liblinelbd_ref.so is missing: goldens are only written where the reference itself can be run")
    O.build()
    frames = []
    for rel in IMAGES:
        img = cv2.imread(os.path.join(GOLD, rel), 1)
        case = dict(image=rel)
        for flav, use_lsd in (("lsd", True), ("edlines", False)):
            kl = O.lbd_detect_keylines(img, use_lsd, 15.0)
            desc = O.lbd_compute(img, kl)
            kr, dr = O.ref_detect_descrip_lines(img, use_lsd, 15.0)
            if len(kr) != len(kl) or any((kr[f] != kl[f]).any() for f in FIELDS) or (dr != desc).any():
                raise SystemExit("oracle != reference on %s (%s): not writing" % (rel, flav))
            case[flav] = dict(n=int(len(kl)), desc_crc32=zlib.crc32(np.ascontiguousarray(desc).tobytes()),
                              angle_crc32=zlib.crc32(np.ascontiguousarray(kl["angle"]).tobytes()), num_pixels_sum=int(kl["num_pixels"].sum()))
        frames.append(case)
    imgs = [cv2.imread(os.path.join(GOLD, p), 1) for p in MATCH["images"]]
    d = [O.lbd_compute(im, O.lbd_detect_keylines(im, True, 15.0)) for im in imgs]
    got, want = O.lbd_match(d[0], d[1], MATCH["thres"]), O.ref_match_line_descrip(d[0], d[1], MATCH["thres"])
    if any(len(a) != len(b) or (a != b).any() for a, b in zip(got, want)):
        raise SystemExit("oracle != reference on the match: not writing")
    out = dict(generator="tools/make_golden_lbd.py", line_length_thres=15.0, frames=frames,
               match=dict(MATCH, triples=[[int(q), int(t), int(x)] for q, t, x in zip(*got)]))
    path = os.path.join(GOLD, "expected_lbd.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, [(c["lsd"]["n"], c["edlines"]["n"]) for c in frames], len(out["match"]["triples"]), "matches")


if __name__ == "__main__":
    main()
