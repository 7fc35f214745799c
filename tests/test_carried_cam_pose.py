"""Host half of the reference-exact handling of several boxes in a roll / pitch-sampled frame (cs_set_profiling bit 10).

cs_debug_last_set_pose is the function detect_batch_carried (cs_context.cu) runs on the candidate records a pass brings back: which pose
hypothesis the reference's cam_pose holds after a height sample -- that of the LAST proposal fuse_normalize_scores_v2 keeps
(box_proposal_detail.cpp:479-487 over object_3d_util.cpp:495-527), or the last sampled one when nothing is kept.  Checked here against the
oracle's trace of the same height sample (valid rows, their candidate indices, the kept ids in the order the reference visits them), on
real jobs and on synthetic records that hit every branch (n <= 4, the saturated-angle branch, an empty intersection, NaNs, ties)."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    from cube_slam_b200 import _lib
    return _lib.load()


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _last_pose(L, valid, dist, angle, n_pose):
    out = C.c_int32(-1)
    valid, dist, angle = np.ascontiguousarray(valid, np.uint8), np.ascontiguousarray(dist, np.float64), np.ascontiguousarray(angle, np.float64)
    assert L.cs_debug_last_set_pose(_p(valid, C.c_uint8), _p(dist, C.c_double), _p(angle, C.c_double), len(valid), n_pose, C.byref(out)) == 0
    return out.value


def _linespace(start, end, step):
    """matrix_utils.cpp:350-363: for (i = start; i <= end; i += step)"""
    out, v = [], start
    while v <= end:
        out.append(v)
        v = v + step
    return np.array(out)


def test_against_the_oracles_trace_of_sampled_jobs(L, oracle):
    from cube_slam_b200 import synthetic as S
    p = oracle.default_params(whether_sample_cam_roll_pitch=1)
    seen = 0
    for seed in (101, 102):
        imgs, Ts, boxes, lines, K = S.make_batch(seed, 6, 640, 480, 3, poisson=True)
        for f in range(len(imgs)):
            for b in range(len(boxes[f])):
                tr = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], np.asarray(lines[f], float), p, trace_object=b, trace_caps=(4096, 1 << 21, 1 << 17))["trace"]
                n_cand, n_valid = tr["n_candidates"], tr["n_valid"]
                if n_valid < 200:
                    continue
                # the sampled roll / pitch grids (4 or 5 values each: linespace(-6 deg, +6 deg, 3 deg) sits on a rounding edge, like the yaw
                # one; box_proposal_detail.cpp:215-226); candidates are enumerated roll-major, then pitch (:230-232)
                euler = oracle.cam_pose(K, Ts[f])["euler"]
                d6, d3 = 6.0 / 180.0 * np.pi, 3.0 / 180.0 * np.pi          # as the reference writes them: (6 / 180) * pi, not 6 * (pi / 180)
                rolls, pitches = _linespace(euler[0] - d6, euler[0] + d6, d3), _linespace(euler[1] - d6, euler[1] + d6, d3)
                n_pose = len(rolls) * len(pitches)
                assert n_pose in (16, 20, 25) and n_cand % n_pose == 0
                assert set(np.unique(tr["rows"][:, 7])) <= set(rolls) and set(np.unique(tr["rows"][:, 8])) <= set(pitches)
                valid = np.zeros(n_cand, np.uint8)
                dist, angle = np.zeros(n_cand), np.zeros(n_cand)
                valid[tr["cand_index"]] = 1
                dist[tr["cand_index"]] = tr["rows"][:, 4]
                angle[tr["cand_index"]] = tr["rows"][:, 5]
                if tr["n_kept"] == 0:
                    want = n_pose - 1
                else:
                    last = tr["kept_ids"][-1]
                    want = int(tr["cand_index"][last]) // (n_cand // n_pose)
                    # ... and that index is the pose the reference sets last: the (roll, pitch) recorded in the kept row (:453,484)
                    assert want == int(np.searchsorted(rolls, tr["rows"][last, 7])) * len(pitches) + int(np.searchsorted(pitches, tr["rows"][last, 8]))
                assert _last_pose(L, valid, dist, angle, n_pose) == want
                seen += n_valid > 4
    assert seen >= 20


def _reference_last(dist, angle):
    """fuse_normalize_scores_v2's final_keep_inds[-1] (object_3d_util.cpp:495-527) with ties by index, NaN last; None when nothing is kept."""
    n = len(dist)
    if n <= 4:
        return n - 1 if n else None
    bn = int(round(float(np.float32(n)) / 3.0 * 2.0))
    key = lambda v: sorted(range(n), key=lambda i: (np.isnan(v[i]), v[i] if not np.isnan(v[i]) else 0.0, i))
    ds, as_ = key(dist), key(angle)
    dk = ds[:bn - 1]
    if angle[as_[bn - 1]] > angle[as_[bn - 2]]:
        inter = sorted(set(dk) & set(as_[:bn - 1]))
        return inter[-1] if inter else None
    return dk[-1]


def test_every_branch_on_synthetic_records(L):
    rng = np.random.default_rng(9)
    n_pose, per_pose = 5, 40
    hit = {"small": 0, "saturated": 0, "intersection": 0, "empty": 0, "nan": 0}
    for trial in range(400):
        n_cand = n_pose * per_pose
        valid = (rng.random(n_cand) < rng.choice([0.02, 0.2, 0.6])).astype(np.uint8)
        dist = np.round(rng.random(n_cand) * rng.choice([4, 1000]), 0 if trial % 3 else 6)      # coarse values: plenty of ties
        angle = np.round(rng.random(n_cand) * 8, 0 if trial % 2 else 6)
        mode = trial % 5
        if mode == 1:
            angle[:] = np.minimum(angle, 3.0)                                                    # saturates at the cut
        if mode == 2 and valid.sum() > 6:                                                        # best by distance == worst by angle: empty intersection
            v = np.flatnonzero(valid)
            dist[v] = np.arange(len(v), dtype=float)
            angle[v] = np.arange(len(v), 0, -1, dtype=float)
        if mode == 3:
            angle[rng.random(n_cand) < 0.3] = np.nan
            hit["nan"] += 1
        v = np.flatnonzero(valid)
        last = _reference_last(dist[v], angle[v])
        want = n_pose - 1 if last is None else int(v[last]) // per_pose
        assert _last_pose(L, valid, dist, angle, n_pose) == want, trial
        n = len(v)
        if n <= 4:
            hit["small"] += 1
        elif last is None:
            hit["empty"] += 1
        else:
            bn = int(round(float(np.float32(n)) / 3.0 * 2.0))
            srt = sorted(range(n), key=lambda i: (np.isnan(angle[v][i]), angle[v][i] if not np.isnan(angle[v][i]) else 0.0, i))
            hit["intersection" if angle[v][srt[bn - 1]] > angle[v][srt[bn - 2]] else "saturated"] += 1
    assert all(c > 5 for c in hit.values()), hit
    assert _last_pose(L, np.zeros(50, np.uint8), np.zeros(50), np.zeros(50), 25) == 24      # no valid proposal at all


def test_pass_structure_with_the_oracle_as_the_device(oracle):
    """cs_carried_passes (cube_slam_b200/csrc/cs_carried.h: what bit 10 runs) with the oracle standing in for the device
    (tests/host_core/carried_emu.cpp): one-box calls chained by the camera yaw alone reproduce the reference's loop over the boxes of a frame
    -- on frames where that chaining changes the number of yaw samples, on ordinary ones, on frames with one box or none, with height sampling."""
    import os
    import subprocess
    from cube_slam_b200 import _lib, synthetic as S
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "host_core", "carried_emu.cpp")
    out = os.path.join(here, "host_core", "_build", "libcarriedemu.so")
    orc = os.path.abspath(os.path.join(here, "..", "oracle", "_build"))
    oracle.build()
    deps = [src, os.path.join(here, "..", "cube_slam_b200", "csrc", "cs_carried.h"), os.path.join(orc, "liboracle.so")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src, "-L", orc, "-loracle", "-Wl,-rpath," + orc])
    E = C.CDLL(out)
    OL = oracle.lib()
    differed = 0
    for seed, pick in ((102, (10, 3, 0)), (104, (1, 0, 5))):
        imgs, Ts, boxes, lines, K = S.make_batch(seed, 12, 640, 480, 3, poisson=True)
        sel = list(pick)
        imgs = np.ascontiguousarray(imgs[sel])
        Ts = np.ascontiguousarray(np.asarray(Ts)[sel], np.float64)
        boxes = [np.asarray(boxes[i], np.float64).reshape(-1, 5) for i in sel] + [np.zeros((0, 5))]       # ... and a frame without boxes
        lines = [np.asarray(lines[i], np.float64).reshape(-1, 4) for i in sel]
        imgs = np.ascontiguousarray(np.concatenate([imgs, imgs[:1]]))
        Ts = np.ascontiguousarray(np.concatenate([Ts, Ts[:1]]))
        lines.append(lines[0])
        F, H, W = imgs.shape[:3]
        box_off = np.concatenate([[0], np.cumsum([len(b) for b in boxes])]).astype(np.int32)
        line_off = np.concatenate([[0], np.cumsum([len(l) for l in lines])]).astype(np.int32)
        all_boxes = np.ascontiguousarray(np.concatenate(boxes))
        all_lines = np.ascontiguousarray(np.concatenate(lines))
        for kw in (dict(whether_sample_cam_roll_pitch=1, max_cuboid_num=3), dict(whether_sample_cam_roll_pitch=1, whether_sample_bbox_height=1, max_cuboid_num=2)):
            p = oracle.default_params(**kw)
            topk = int(p.max_cuboid_num)
            out_recs = np.zeros((int(box_off[-1]), topk), _lib.CUBOID_DTYPE)
            counts = np.zeros(int(box_off[-1]), np.int32)
            n_passes = C.c_int32(0)
            Kc = np.ascontiguousarray(K, np.float64)
            rc = E.emu_carried(imgs.ctypes.data_as(C.POINTER(C.c_uint8)), F, W, H, W * 3, 3, _p(Kc, C.c_double), _p(Ts, C.c_double), _p(all_boxes, C.c_double),
                               _p(box_off, C.c_int32), _p(all_lines, C.c_double), _p(line_off, C.c_int32), C.byref(p), out_recs.ctypes.data_as(C.c_void_p),
                               _p(counts, C.c_int32), C.byref(n_passes))
            assert rc == 0 and n_passes.value == max(len(b) for b in boxes) == 3
            o = 0
            for f in range(F):
                if len(boxes[f]) == 0:
                    continue
                ref = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines[f], p)                    # the reference's loop over the boxes
                OL.orc_set_independent_boxes(1)
                ind = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines[f], p)
                OL.orc_set_independent_boxes(0)
                for b in range(len(boxes[f])):
                    assert counts[o] == len(ref["cuboids"][b])
                    got = out_recs[o, :counts[o]]
                    assert got.tobytes() == ref["cuboids"][b].view(_lib.CUBOID_DTYPE).tobytes() if got.dtype != ref["cuboids"][b].dtype else got.tobytes() == ref["cuboids"][b].tobytes()
                    differed += int(got.tobytes() != ind["cuboids"][b].tobytes())
                    o += 1
    assert differed >= 4        # the chosen frames are ones where starting every box from the raw pose gives other records


def _build_context_emu(oracle):
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.join(here, "..")
    srcs = [os.path.join(here, "host_core", "context_emu.cpp"), os.path.join(here, "host_core", "context_emu_absent.cpp")]
    out = os.path.join(here, "host_core", "_build", "libcontextemu.so")
    orc = os.path.abspath(os.path.join(root, "oracle", "_build"))
    csrc = os.path.join(root, "cube_slam_b200", "csrc")
    oracle.build()
    deps = srcs + [os.path.join(here, "host_core", "cuda_emu.h"), os.path.join(here, "host_core", "fake_cuda", "cuda_runtime.h"), os.path.join(orc, "liboracle.so")]
    deps += [os.path.join(csrc, f) for f in ("cs_context.cu", "cs_carried.h", "cs_internal.h", "cs_kernels.h", "cs_host_pose.cpp", "cs_host_pose.h", "cs_nccl_impl.inc")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-I", os.path.join(here, "host_core", "fake_cuda"),
                               "-x", "c++", "-o", out] + srcs + [os.path.join(csrc, "cs_host_pose.cpp"), "-L", orc, "-loracle", "-ldl", "-Wl,-rpath," + orc])
    return out


def test_the_librarys_own_host_code_with_the_oracle_as_the_device(oracle):
    """cs_context.cu itself compiled for the host (tests/host_core/context_emu.cpp: CUDA runtime calls inert, every kernel launcher a no-op
    except the last stage's, which fills records and candidate records from the oracle using the camera yaw the library's own yaw table
    was derived from).  cs_detect_cuboids_batch with bit 10 on sampled multi-box frames then returns the reference's cuboids byte for byte --
    and without bit 10 the independent-boxes ones: the real pass loop, the yaw override reaching build_tables, job / candidate offsets,
    last_set_pose, the pose-table lookup and the scatter, all exercised without a GPU."""
    from cube_slam_b200 import _lib, synthetic as S
    out = _build_context_emu(oracle)
    E = C.CDLL(out)
    E.cs_create.restype = C.c_void_p
    E.cs_last_error.restype = C.c_char_p
    OL = oracle.lib()
    for seed, pick in ((102, (10, 3, 0)), (104, (1, 0, 5))):
        imgs, Ts, boxes, lines, K = S.make_batch(seed, 12, 640, 480, 3, poisson=True)
        sel = list(pick)
        imgs = np.ascontiguousarray(imgs[sel])
        Ts = np.ascontiguousarray(np.asarray(Ts)[sel], np.float64)
        boxes = [np.asarray(boxes[i], np.float64).reshape(-1, 5) for i in sel]
        lines = [np.asarray(lines[i], np.float64).reshape(-1, 4) for i in sel]
        F, H, W = imgs.shape[:3]
        box_off = np.concatenate([[0], np.cumsum([len(b) for b in boxes])]).astype(np.int32)
        line_off = np.concatenate([[0], np.cumsum([len(l) for l in lines])]).astype(np.int32)
        all_boxes, all_lines = np.ascontiguousarray(np.concatenate(boxes)), np.ascontiguousarray(np.concatenate(lines))
        Kc = np.ascontiguousarray(K, np.float64)
        for kw in (dict(whether_sample_cam_roll_pitch=1, max_cuboid_num=3), dict(whether_sample_cam_roll_pitch=1, whether_sample_bbox_height=1, max_cuboid_num=2)):
            prm = _lib.CuboidParams()
            _lib.load().cs_default_cuboid_params(C.byref(prm))
            for k, v in kw.items():
                setattr(prm, k, v)
            topk = kw["max_cuboid_num"]
            outs = {}
            for carried in (1, 0):
                ctx = C.c_void_p(E.cs_create(0, 640, 480, F, 16, 4096))
                assert E.cs_set_calibration(ctx, _p(Kc, C.c_double)) == 0
                assert E.cs_set_profiling(ctx, 1024 if carried else 0) == 0
                rec = np.zeros((int(box_off[-1]), topk), _lib.CUBOID_DTYPE)
                cnt = np.zeros(int(box_off[-1]), np.int32)
                rc = E.emu_detect_cuboids_batch(ctx, imgs.ctypes.data_as(C.POINTER(C.c_uint8)), F, W, H, W * 3, 3, _p(Ts, C.c_double), _p(all_boxes, C.c_double), _p(box_off, C.c_int32),
                                                _p(all_lines, C.c_double), _p(line_off, C.c_int32), C.byref(prm), rec.ctypes.data_as(C.c_void_p), _p(cnt, C.c_int32))
                assert rc == 0, E.cs_last_error(ctx)
                assert E.emu_mismatch() == 0
                E.cs_destroy(ctx)
                outs[carried] = rec.tobytes()
                o = 0
                OL.orc_set_independent_boxes(0 if carried else 1)
                try:
                    for f in range(F):
                        ref = oracle.detect_cuboid(imgs[f], K, Ts[f], boxes[f], lines[f], oracle.default_params(**kw))
                        for b in range(len(boxes[f])):
                            assert cnt[o] == len(ref["cuboids"][b])
                            assert rec[o, :cnt[o]].tobytes() == ref["cuboids"][b].tobytes(), (seed, f, b, carried)
                            o += 1
                finally:
                    OL.orc_set_independent_boxes(0)
            assert outs[0] != outs[1]


def test_gpu_tests_of_the_carried_pose_rehearsed_through_the_python_package(oracle, monkeypatch):
    """The Python package bound to the emulated build instead of libcubeslam_b200.so (cube_slam_b200/_lib.py's loader pointed at it), and the two
    GPU tests of this subject run as they are: Context / detect_batch_host with and without bit 10, and the class mirror cs.detect_3d_cuboid."""
    import sys
    from cube_slam_b200 import _lib
    import test_z_gpu_lbd_parity as G
    monkeypatch.setattr(_lib, "LIB_PATH", _build_context_emu(oracle))
    monkeypatch.setattr(_lib, "_lib", None)
    try:
        G.test_sampled_frames_with_several_boxes_default_and_carried_pose(oracle)
        G.test_class_mirror_carries_the_pose_like_the_reference(oracle)
    finally:
        _lib._lib = None       # the next user gets the real library again
