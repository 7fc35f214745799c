"""The arithmetic atan2 of the angle-error chain (cube_slam_b200/csrc/cs_pmath.h, oracle/pmath.h): the library's host evaluation and the
oracle's copy agree bit for bit, stay within 1 ulp of the libm the reference calls, and the device evaluation equals both (GPU test).
Then the price of the substitution, measured on the oracle alone: libm atan2 vs the arithmetic one on whole detect_cuboid runs."""
import ctypes as C

import numpy as np
import pytest


def _inputs(n=200000, seed=3):
    rng = np.random.default_rng(seed)
    y = np.concatenate([rng.uniform(-700, 700, n), np.floor(rng.uniform(-700, 700, n)), rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-8, 8, n)])
    x = np.concatenate([rng.uniform(-700, 700, n), np.floor(rng.uniform(-700, 700, n)), rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-8, 8, n)])
    special = np.array([[0.0, 1], [0, -1], [-0.0, -1], [-0.0, 1], [1, 0], [-1, 0], [0, 0], [3, 3], [-3, 3], [5, -5], [-5, -5], [1e300, 1e-300],
                        [1e-300, -1e300], [np.inf, 1], [1, np.inf], [1, -np.inf], [np.inf, np.inf], [-np.inf, -np.inf]])
    return np.concatenate([y, special[:, 0]]), np.concatenate([x, special[:, 1]])


def _ulps(a, b):
    ia, ib = a.view(np.int64), b.view(np.int64)
    same_sign = (ia < 0) == (ib < 0)
    d = np.abs(ia - ib)
    return np.where(same_sign | (a == b), d, 1 << 40)


def test_oracle_and_library_host_copies_agree_and_stay_within_one_ulp_of_libm(oracle):
    import cube_slam_b200._lib as lib
    L = lib.load()
    O = oracle.lib()
    O.orc_atan2_portable.restype = C.c_double
    O.orc_atan2_portable.argtypes = [C.c_double, C.c_double]
    y, x = _inputs(20000)
    a = np.array([O.orc_atan2_portable(float(p), float(q)) for p, q in zip(y, x)])
    b = np.array([L.cs_atan2_host(float(p), float(q)) for p, q in zip(y, x)])
    assert a.tobytes() == b.tobytes()
    ref = np.arctan2(y, x)
    assert _ulps(a, ref).max() <= 1
    # signed zeros and the axes exactly as libm
    for yy, xx in ((0.0, 1.0), (-0.0, 1.0), (0.0, -1.0), (-0.0, -1.0), (1.0, 0.0), (-1.0, 0.0), (0.0, 0.0)):
        assert np.float64(O.orc_atan2_portable(yy, xx)).tobytes() == np.float64(np.arctan2(yy, xx)).tobytes()


def test_libm_vs_arithmetic_atan2_on_whole_frames(oracle, fixture_a):
    """What the substitution costs, on the oracle alone: with libm's atan2 (what the reference calls) instead of the arithmetic one the
    valid sets, kept counts and candidate errors agree to the last bits, and the selected cuboid is the same unless the box is one of
    the knife-edge cases (a mirror pair of proposals straddling fuse_normalize_scores_v2's cut), which the oracle flags."""
    from cube_slam_b200 import synthetic as S
    O = oracle.lib()
    cases = [(fixture_a["img"], fixture_a["K"], fixture_a["T"], fixture_a["boxes"], fixture_a["lines"])]
    imgs, Ts, boxes, lines, K = S.make_batch(12, 6, 640, 480, 5, poisson=True)
    cases += [(imgs[f], K, Ts[f], boxes[f], lines[f]) for f in range(6)]
    n_box = n_same = n_flagged = 0
    for img, K_, T, bx, ln in cases:
        try:
            O.orc_set_portable_atan2(0)
            r0 = oracle.detect_cuboid(img, K_, T, bx, ln, oracle.default_params(max_cuboid_num=1))
        finally:
            O.orc_set_portable_atan2(1)
        r1 = oracle.detect_cuboid(img, K_, T, bx, ln, oracle.default_params(max_cuboid_num=1))
        assert r0["n_valid"] == r1["n_valid"] and r0["n_candidates"] == r1["n_candidates"]
        for b in range(len(bx)):
            n_box += 1
            if not len(r0["cuboids"][b]):
                assert not len(r1["cuboids"][b])
                n_same += 1
                continue
            c0, c1 = r0["cuboids"][b][0], r1["cuboids"][b][0]
            same = int(c0["proposal_index"]) == int(c1["proposal_index"]) and abs(float(c0["normalized_error"]) - float(c1["normalized_error"])) < 1e-9
            n_same += same
            if not same:
                n_flagged += 1
            elif int(c0["proposal_index"]) == int(c1["proposal_index"]):
                np.testing.assert_allclose(c0["edge_angle_error"], c1["edge_angle_error"], rtol=1e-12)
    assert n_box >= 20 and n_same >= n_box - 3, (n_box, n_same, n_flagged)


@pytest.mark.gpu
def test_device_atan2_equals_the_host_evaluation_bit_for_bit(oracle):
    import cube_slam_b200 as cs
    from cube_slam_b200 import _lib
    O = oracle.lib()
    O.orc_atan2_portable.restype = C.c_double
    O.orc_atan2_portable.argtypes = [C.c_double, C.c_double]
    y, x = _inputs(200000)
    ctx = cs.Context(0, 64, 64, 1, 1, 16)
    out = np.zeros_like(y)
    ctx.check(ctx.L.cs_debug_atan2(ctx.h, _lib.ptr(y, C.c_double), _lib.ptr(x, C.c_double), _lib.ptr(out, C.c_double), len(y)))
    host = np.array([ctx.L.cs_atan2_host(float(p), float(q)) for p, q in zip(y[::7], x[::7])])
    assert out[::7].tobytes() == host.tobytes()
    orc = np.array([O.orc_atan2_portable(float(p), float(q)) for p, q in zip(y[::7], x[::7])])
    assert out[::7].tobytes() == orc.tobytes()
    finite = np.isfinite(y) & np.isfinite(x)
    assert _ulps(out[finite], np.arctan2(y[finite], x[finite])).max() <= 1
    ctx.close()
