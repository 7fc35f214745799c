"""The descriptor / matcher oracle against the committed goldens (tests/golden/expected_lbd.json, written by tools/make_golden_lbd.py only
after oracle == compiled reference held on every case): catches oracle drift on machines without the reference."""
import json
import os
import zlib

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_the_lbd_goldens(oracle):
    import cv2
    g = json.load(open(os.path.join(GOLD, "expected_lbd.json")))
    for case in g["frames"]:
        img = cv2.imread(os.path.join(GOLD, case["image"]), 1)
        for flav in ("lsd", "edlines"):
            kl = oracle.lbd_detect_keylines(img, flav == "lsd", g["line_length_thres"])
            desc = oracle.lbd_compute(img, kl)
            assert len(kl) == case[flav]["n"]
            assert zlib.crc32(np.ascontiguousarray(desc).tobytes()) == case[flav]["desc_crc32"]
            assert zlib.crc32(np.ascontiguousarray(kl["angle"]).tobytes()) == case[flav]["angle_crc32"]
            assert int(kl["num_pixels"].sum()) == case[flav]["num_pixels_sum"]
    imgs = [cv2.imread(os.path.join(GOLD, p), 1) for p in g["match"]["images"]]
    d = [oracle.lbd_compute(im, oracle.lbd_detect_keylines(im, True, g["line_length_thres"])) for im in imgs]
    q, t, x = oracle.lbd_match(d[0], d[1], g["match"]["thres"])
    assert [[int(a), int(b), int(c)] for a, b, c in zip(q, t, x)] == g["match"]["triples"]
    assert len(g["match"]["triples"]) >= 3
