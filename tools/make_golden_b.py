#!/usr/bin/env python
"""Oracle goldens for the 58-frame object_slam sequence (fixture B) and for seeded synthetic frames: regression pins of the oracle itself
(SURVEY.md section 7 step 1), also checked against the CUDA path on the GPU.

    python tools/make_golden_b.py

Needs oracle/_ref/*.so (the reference's line_lbd and detect_3d_cuboid sources compiled where /root/reference exists; every record written
here was first checked equal to the reference's own output for the same inputs):
the raw LSD segments (`n_raw_ref`, `raw_checksum_ref`) and raw EDLines key lines (`edl_n_raw_ref`, `edl_raw_checksum_ref`) recorded here
are the REFERENCE's output, so the committed file pins the oracle to the reference wherever the tests
run.  Per frame and mode: LSD segment count and a checksum of the segments, candidates, valid proposals, and per box the best cuboid
(proposal index, normalised error, position, yaw, scale).  Modes: default, and 5 x 5 camera roll / pitch sampling
(whether_sample_cam_roll_pitch).  object_slam's own settings: length threshold 15, nominal_skew_ratio 2 (main_obj.cpp:359-366)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def seg_checksum(lines):
    a = np.ascontiguousarray(lines, np.float32).view(np.uint32).astype(np.uint64)
    w = (np.arange(a.size, dtype=np.uint64).reshape(a.shape) * np.uint64(2654435761) + np.uint64(1)) & np.uint64(0xffffffff)
    return int(((a * w) & np.uint64(0xffffffffffff)).sum() & np.uint64(0xffffffffffff))


def best(c):
    return dict(proposal_index=int(c["proposal_index"]), normalized_error=float(c["normalized_error"]), pos=[float(v) for v in c["pos"]],
                rotY=float(c["rotY"]), scale=[float(v) for v in c["scale"]])


def frame_record(img, K, T, boxes, modes):
    res = O.lsd_detect(img, 15.0)
    lines = res["lines"]
    # the raw segments come from the REFERENCE's own lsd.cpp (oracle/_ref/liblsd_ref.so, built from /root/reference by oracle/Makefile);
    # the oracle restatement has to agree with it bit for bit before anything is written
    ref_raw = O.ref_lsd_detect(img)
    if not np.array_equal(ref_raw, res["raw_lines"]):
        raise SystemExit("oracle LSD differs from the reference's lsd.cpp on this frame: fix the oracle first")
    # the EDLines flavour (use_LSD = false) the same way: the reference's own binary_descriptor.cpp (oracle/_ref/libedl_ref.so)
    edl = O.edl_detect(img, 15.0)
    ref_edl = O.ref_edl_detect(img)
    if not np.array_equal(ref_edl, edl["raw_lines"]):
        raise SystemExit("oracle EDLines differs from the reference's binary_descriptor.cpp on this frame: fix the oracle first")
    # ... and the whole function, line_lbd_detect::detect_filter_lines as the reference's class computes it (oracle/_ref/liblinelbd_ref.so):
    # the recorded `lines_checksum` / `edl_lines_checksum` ARE the reference's output
    for use_lsd, got in ((True, lines), (False, edl["lines"])):
        if not np.array_equal(O.ref_detect_filter_lines(img, use_lsd, 15.0), got):
            raise SystemExit("oracle detect_filter_lines differs from the reference's on this frame: fix the oracle first")
    rec = dict(n_lines=int(len(lines)), lines_checksum=seg_checksum(lines), n_raw_ref=int(len(ref_raw)), raw_checksum_ref=seg_checksum(ref_raw),
               edl_n_lines=int(len(edl["lines"])), edl_lines_checksum=seg_checksum(edl["lines"]), edl_n_raw_ref=int(len(ref_edl)),
               edl_raw_checksum_ref=seg_checksum(ref_edl), modes={})
    for name, kw in modes:
        # stage (ii): the reference's own detect_cuboid (oracle/_ref/libcuboid_ref.so) must return the oracle's records, field for field,
        # when both use libm's atan2; the golden itself is then written with the arithmetic atan2 the CUDA path shares (tests/test_pmath.py)
        pr = O.default_params(nominal_skew_ratio=2.0, **kw)
        O.lib().orc_set_portable_atan2(0)
        got = O.detect_cuboid(img, K, T, boxes, lines.astype(np.float64), pr)["cuboids"]
        O.lib().orc_set_portable_atan2(1)
        want = O.ref_detect_cuboid(img, K, T, boxes, lines.astype(np.float64), pr)
        for b in range(len(want)):
            if len(got[b]) != len(want[b]) or any(not np.array_equal(np.asarray(got[b][j][f], np.float64), np.asarray(want[b][j][f], np.float64))
                                                  for j in range(len(want[b])) for f in ("pos", "rotY", "scale", "normalized_error", "box_corners_2d")):
                raise SystemExit("oracle detect_cuboid differs from the reference's on this frame: fix the oracle first")
        r = O.detect_cuboid(img, K, T, boxes, lines.astype(np.float64), O.default_params(nominal_skew_ratio=2.0, **kw))
        rec["modes"][name] = dict(n_candidates=int(r["n_candidates"]), n_valid=int(r["n_valid"]),
                                  boxes=[best(c[0]) if len(c) else None for c in r["cuboids"]])
    return rec


def main():
    import conftest
    fb = conftest.fixture_b.__wrapped__()
    modes = [("default", {}), ("sample_roll_pitch", dict(whether_sample_cam_roll_pitch=1))]
    out = dict(source="oracle (oracle/*.cpp) on tests/golden/fixture_b, raw LSD segments from the reference lsd.cpp (oracle/_ref); tools/make_golden_b.py", frames=[])
    for i, (img, boxes) in enumerate(fb["frames"]):
        out["frames"].append(frame_record(img, fb["K"], fb["T"], boxes, modes))
        print("fixture B frame", i, out["frames"][-1]["n_lines"], flush=True)
    json.dump(out, open(os.path.join(GOLD, "expected_fixture_b.json"), "w"), indent=0)
    from cube_slam_b200 import synthetic as S
    syn = dict(source="oracle on cube_slam_b200.synthetic.make_batch(seed, 4, w, h, nb, kind); tools/make_golden_b.py", cases=[])
    for seed, w, h, kind, nb in ((101, 640, 480, "indoor", 3), (102, 1242, 375, "kitti", 8), (103, 1280, 960, "indoor", 4)):
        imgs, Ts, boxes, _, K = S.make_batch(seed, 4, w, h, nb, kind=kind, poisson=(kind == "indoor"))
        case = dict(seed=seed, w=w, h=h, kind=kind, nb=nb, image_checksum=int(imgs.astype(np.uint64).sum()), frames=[])
        for f in range(4):
            case["frames"].append(frame_record(imgs[f], K, Ts[f], boxes[f], modes[:1]))
        syn["cases"].append(case)
        print("synthetic", seed, [fr["n_lines"] for fr in case["frames"]], flush=True)
    json.dump(syn, open(os.path.join(GOLD, "expected_synthetic.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
