"""Times the descriptor / matcher half of line_lbd_detect on one GPU and prints one JSON object.

    python tools/time_lbd.py [--frames 64] [--reps 5]

Synthetic VGA frames of the bench workload; every call goes through the C ABI with HOST buffers (frames in, key lines / descriptors /
matches out), wall clock around the synchronous calls after one warm-up call each.  bench.py runs this in a process of its own after its
timed region (key "lbd" of the bench line), so that nothing here can disturb the headline measurement."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import cube_slam_b200 as cs
    from cube_slam_b200 import synthetic as S
    imgs = np.ascontiguousarray(S.make_batch(20260922, args.frames, 640, 480, 3, poisson=True)[0])
    F = len(imgs)
    det = cs.line_lbd_detect()
    det.line_length_thres = 15
    out = {"frames": F, "size": "640x480", "line_length_thres": 15}
    for flav, use_lsd in (("lsd", True), ("edlines", False)):
        det.use_LSD = use_lsd
        res = det.detect_descrip_lines_batch(imgs)          # warm-up (allocations, weight upload)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            res = det.detect_descrip_lines_batch(imgs)
        dt = (time.perf_counter() - t0) / args.reps
        n = sum(len(k) for k, _ in res)
        out["detect_descrip_" + flav] = {"ms_per_batch": dt * 1e3, "frames_per_s": F / dt, "lines": n, "lines_per_s": n / dt}
        if use_lsd:
            kls = [k for k, _ in res]
            descs = [d for _, d in res]
    # descriptors alone, on the key lines just found (BinaryDescriptor::compute over a batch)
    L, h = det._ctx.L, det._ctx.h
    import ctypes as C
    from cube_slam_b200 import _lib
    off = np.concatenate([[0], np.cumsum([len(k) for k in kls])]).astype(np.int32)
    kl = np.ascontiguousarray(np.concatenate(kls))
    desc = np.zeros((len(kl), 32), np.uint8)

    def compute():
        det._ctx.check(L.cs_lbd_compute_batch(h, imgs.ctypes.data, F, 640, 480, 640 * 3, 3, kl.ctypes.data, _lib.ptr(off, C.c_int32), _lib.ptr(desc, C.c_uint8), None))

    compute()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        compute()
    dt = (time.perf_counter() - t0) / args.reps
    assert all((desc[off[f]:off[f + 1]] == descs[f]).all() for f in range(F)), "cs_lbd_compute_batch != cs_detect_descrip_lines_batch"
    px = int(np.concatenate([k["num_pixels"] for k in kls]).sum())
    out["compute"] = {"ms_per_batch": dt * 1e3, "lines": int(len(kl)), "lines_per_s": len(kl) / dt, "support_pixels": px * 63,
                      "gather_GBps": px * 63 * 4 / dt / 1e9, "note": "includes the H2D copy of the frames and the Sobel-map kernel"}
    # matching: every frame against the next one
    qs, ts = descs[:-1], descs[1:]
    det.match_line_descrip_batch(qs, ts, 40.0)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        m = det.match_line_descrip_batch(qs, ts, 40.0)
    dt = (time.perf_counter() - t0) / args.reps
    pairs = sum(len(q) * len(t) for q, t in zip(qs, ts))
    out["match"] = {"ms_per_batch": dt * 1e3, "frame_pairs": len(qs), "code_pairs": pairs, "code_pairs_per_s": pairs / dt, "good_matches": int(sum(len(x) for x in m))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
