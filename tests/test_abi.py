"""The C-ABI library loads on a CPU-only host, exports every symbol the header declares, and fails loudly
(no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "cube_slam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if n not in ("cs_status",)))


def test_library_exports_every_declared_symbol():
    from cube_slam_b200 import _lib
    L = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 25
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(set(_lib.EXPORTS)) == declared
    assert L.cs_abi_version() == 1


def test_struct_layouts_match_header():
    from cube_slam_b200 import _lib
    assert _lib.CUBOID_DTYPE.itemsize == 440
    assert C.sizeof(_lib.CuboidParams) == 8 * 4 + 14 * 8
    assert C.sizeof(_lib.LineParams) == 16
    assert C.sizeof(_lib.BatchStats) == 64
    p = _lib.CuboidParams()
    _lib.load().cs_default_cuboid_params(C.byref(p))
    # the reference's literals (detect_3d_cuboid.h:72-79, box_proposal_detail.cpp:79-87,177-179,197,128)
    assert (p.consider_config_1, p.consider_config_2, p.whether_sample_cam_roll_pitch, p.whether_sample_bbox_height) == (1, 1, 0, 0)
    assert (p.max_cuboid_num, p.nominal_skew_ratio, p.max_cut_skew) == (1, 1.0, 3.0)
    assert (p.vp12_edge_angle_thre, p.vp3_edge_angle_thre, p.shorted_edge_thre) == (15.0, 10.0, 20.0)
    assert (p.weight_vp_angle, p.weight_skew_error) == (0.8, 1.5)
    assert (p.pre_merge_dist_thre, p.pre_merge_angle_thre, p.edge_length_threshold) == (20.0, 5.0, 30.0)
    assert (p.canny_low, p.canny_high, p.yaw_half_range_deg, p.yaw_step_deg) == (80.0, 200.0, 45.0, 6.0)
    lp = _lib.LineParams()
    _lib.load().cs_default_line_params(C.byref(lp))
    assert (lp.use_LSD, lp.numoctaves, lp.octaveratio, lp.line_length_thres) == (0, 1, 1.0, 50.0)


def test_cam_pose_host_function_matches_oracle(oracle, fixture_a):
    """cs_cam_pose is host-only (set_cam_pose, box_proposal_detail.cpp:42-54): bit-identical to the oracle's restatement."""
    from cube_slam_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(0)
    K = np.ascontiguousarray(fixture_a["K"]).reshape(9)
    for i in range(20):
        T = fixture_a["T"].copy()
        if i:
            a = rng.uniform(-3, 3, 3)
            from cube_slam_b200.synthetic import euler_zyx_to_rot
            T[:3, :3] = euler_zyx_to_rot(*a)
            T[:3, 3] = rng.uniform(-2, 2, 3)
        e = np.zeros(3)
        kr = np.zeros(9)
        assert L.cs_cam_pose(_lib.ptr(K, C.c_double), _lib.ptr(np.ascontiguousarray(T).reshape(16), C.c_double), _lib.ptr(e, C.c_double),
                             _lib.ptr(kr, C.c_double)) == 0
        ref = oracle.cam_pose(fixture_a["K"], T)
        np.testing.assert_array_equal(e, ref["euler"])
        np.testing.assert_array_equal(kr.reshape(3, 3), ref["KinvR"])


def test_no_cpu_fallback():
    """Without a CUDA device cs_create returns NULL and the Python mirror raises; with one, this test is skipped."""
    import subprocess
    import sys
    code = ("import cube_slam_b200 as cs\n"
            "try:\n    cs.Context(0)\n    print('HAS_GPU')\nexcept cs.CubeSlamError as e:\n    print('RAISED', e)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert "RAISED" in out.stdout and "no CPU path" in (out.stdout + out.stderr), out.stdout + out.stderr


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (a product path through the oracle voids parity claims)."""
    pkg = os.path.join(ROOT, "cube_slam_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "orc_" not in text, os.path.join(root, f)
