"""Committed goldens of the whole path (tests/golden/expected_fixture_b.json, expected_synthetic.json; tools/make_golden_b.py): the
oracle must keep reproducing them (CPU), and the CUDA path must produce the same numbers without the oracle in the loop (GPU)."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLD

sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))


def _checksum(lines):
    a = np.ascontiguousarray(lines, np.float32).view(np.uint32).astype(np.uint64)
    w = (np.arange(a.size, dtype=np.uint64).reshape(a.shape) * np.uint64(2654435761) + np.uint64(1)) & np.uint64(0xffffffff)
    return int(((a * w) & np.uint64(0xffffffffffff)).sum() & np.uint64(0xffffffffffff))


def _check_box(got, want):
    if want is None:
        assert got is None
        return
    assert got is not None
    assert int(got["proposal_index"]) == want["proposal_index"]
    assert abs(float(got["normalized_error"]) - want["normalized_error"]) < 1e-9
    np.testing.assert_allclose(got["pos"], want["pos"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=1e-9, atol=1e-9)
    assert abs(float(got["rotY"]) - want["rotY"]) < 1e-12


MODES = {"default": {}, "sample_roll_pitch": dict(whether_sample_cam_roll_pitch=1)}


def test_oracle_reproduces_the_fixture_b_goldens(oracle, fixture_b):
    exp = json.load(open(os.path.join(GOLD, "expected_fixture_b.json")))
    assert len(exp["frames"]) == len(fixture_b["frames"])
    for i in range(0, len(exp["frames"]), 3):  # every third frame keeps the CPU suite short; the GPU test walks all of them
        img, boxes = fixture_b["frames"][i]
        e = exp["frames"][i]
        res = oracle.lsd_detect(img, 15.0)
        lines = res["lines"]
        assert len(lines) == e["n_lines"] and _checksum(lines) == e["lines_checksum"]
        # the raw segments of the reference's own lsd.cpp, recorded when the golden was made (oracle/_ref): pins the oracle to the reference
        assert len(res["raw_lines"]) == e["n_raw_ref"] and _checksum(res["raw_lines"]) == e["raw_checksum_ref"]
        edl = oracle.edl_detect(img, 15.0)   # the EDLines flavour against the reference's own binary_descriptor.cpp, the same way
        assert len(edl["raw_lines"]) == e["edl_n_raw_ref"] and _checksum(edl["raw_lines"]) == e["edl_raw_checksum_ref"]
        assert len(edl["lines"]) == e["edl_n_lines"] and _checksum(edl["lines"]) == e["edl_lines_checksum"]
        for mode, kw in MODES.items():
            if mode != "default" and i % 6:
                continue
            r = oracle.detect_cuboid(img, fixture_b["K"], fixture_b["T"], boxes, lines.astype(np.float64), oracle.default_params(nominal_skew_ratio=2.0, **kw))
            em = e["modes"][mode]
            assert (r["n_candidates"], r["n_valid"]) == (em["n_candidates"], em["n_valid"])
            for b, want in enumerate(em["boxes"]):
                _check_box(r["cuboids"][b][0] if len(r["cuboids"][b]) else None, want)


def test_oracle_reproduces_the_synthetic_goldens(oracle):
    from cube_slam_b200 import synthetic as S
    exp = json.load(open(os.path.join(GOLD, "expected_synthetic.json")))
    for case in exp["cases"][:2]:
        imgs, Ts, boxes, _, K = S.make_batch(case["seed"], 4, case["w"], case["h"], case["nb"], kind=case["kind"], poisson=(case["kind"] == "indoor"))
        assert int(imgs.astype(np.uint64).sum()) == case["image_checksum"]  # the generator itself is part of the golden
        for f in (0, 3):
            e = case["frames"][f]
            res = oracle.lsd_detect(imgs[f], 15.0)
            lines = res["lines"]
            assert len(lines) == e["n_lines"] and _checksum(lines) == e["lines_checksum"]
            assert len(res["raw_lines"]) == e["n_raw_ref"] and _checksum(res["raw_lines"]) == e["raw_checksum_ref"]
            edl = oracle.edl_detect(imgs[f], 15.0)
            assert len(edl["raw_lines"]) == e["edl_n_raw_ref"] and _checksum(edl["raw_lines"]) == e["edl_raw_checksum_ref"]
            r = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines.astype(np.float64), oracle.default_params(nominal_skew_ratio=2.0))
            em = e["modes"]["default"]
            assert (r["n_candidates"], r["n_valid"]) == (em["n_candidates"], em["n_valid"])
            for b, want in enumerate(em["boxes"]):
                _check_box(r["cuboids"][b][0] if len(r["cuboids"][b]) else None, want)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(MODES))
def test_cuda_path_reproduces_the_fixture_b_goldens(fixture_b, mode):
    """cs_detect_frames_batch (LSD on the device, then the cuboid stage) against the committed numbers: no oracle involved."""
    import cube_slam_b200 as cs
    exp = json.load(open(os.path.join(GOLD, "expected_fixture_b.json")))
    frames = fixture_b["frames"]
    F = len(frames)
    imgs = np.stack([fr[0] for fr in frames])
    boxes = [fr[1] for fr in frames]
    ctx = cs.Context(0, 640, 480, F, 8, 4096)
    ctx.set_calibration(fixture_b["K"])
    det = cs.line_lbd_detect(context=ctx)
    det.use_LSD = True
    det.line_length_thres = 15
    lines = det.detect_filter_lines_batch(imgs)
    for i in range(F):
        assert len(lines[i]) == exp["frames"][i]["n_lines"] and _checksum(lines[i]) == exp["frames"][i]["lines_checksum"], i
    p = cs.default_params(nominal_skew_ratio=2.0, **MODES[mode])
    out, counts = ctx.detect_frames_host(imgs, np.stack([fixture_b["T"]] * F), boxes, det.params(), p)
    if mode == "default":   # the EDLines flavour of the line stage against the same goldens
        det.use_LSD = False
        elines = det.detect_filter_lines_batch(imgs)
        for i in range(F):
            assert len(elines[i]) == exp["frames"][i]["edl_n_lines"] and _checksum(elines[i]) == exp["frames"][i]["edl_lines_checksum"], i
        det.use_LSD = True
    o = 0
    for i in range(F):
        for b, want in enumerate(exp["frames"][i]["modes"][mode]["boxes"]):
            if mode == "sample_roll_pitch" and b > 0:
                o += 1  # later boxes of a sampled frame: the documented cam_pose deviation, checked elsewhere
                continue
            _check_box(out[o, 0] if counts[o] else None, want)
            o += 1
    ctx.close()


@pytest.mark.gpu
def test_cuda_path_reproduces_the_synthetic_goldens():
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    exp = json.load(open(os.path.join(GOLD, "expected_synthetic.json")))
    for case in exp["cases"]:
        imgs, Ts, boxes, _, K = S.make_batch(case["seed"], 4, case["w"], case["h"], case["nb"], kind=case["kind"], poisson=(case["kind"] == "indoor"))
        assert int(imgs.astype(np.uint64).sum()) == case["image_checksum"]
        ctx = cs.Context(0, case["w"], case["h"], 4, 16, 4096)
        ctx.set_calibration(K)
        det = cs.line_lbd_detect(context=ctx)
        det.use_LSD = True
        det.line_length_thres = 15
        lines = det.detect_filter_lines_batch(imgs)
        for f in range(4):
            assert len(lines[f]) == case["frames"][f]["n_lines"] and _checksum(lines[f]) == case["frames"][f]["lines_checksum"], (case["seed"], f)
        out, counts = ctx.detect_frames_host(imgs, Ts, boxes, det.params(), cs.default_params(nominal_skew_ratio=2.0))
        det.use_LSD = False
        elines = det.detect_filter_lines_batch(imgs)
        for f in range(4):
            assert len(elines[f]) == case["frames"][f]["edl_n_lines"] and _checksum(elines[f]) == case["frames"][f]["edl_lines_checksum"], (case["seed"], f)
        o = 0
        for f in range(4):
            for want in case["frames"][f]["modes"]["default"]["boxes"]:
                _check_box(out[o, 0] if counts[o] else None, want)
                o += 1
        ctx.close()
