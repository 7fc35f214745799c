"""The arithmetic of the descriptor / matcher kernels without a GPU.

cube_slam_b200/csrc/cs_lbd_core.h holds every formula of k_lbd_describe and k_lbd_match as functions that compile for the device and for
the host.  tests/host_core/lbd_core_host.cpp builds them with g++ (-ffp-contract=off) and drives them exactly as the kernels do -- thread
index by thread index, phase after phase -- and the product library's own host side (cs_lbd_debug_prepare, cs_keylines_from_lines: host-only
entry points of libcubeslam_b200.so) prepares the inputs.  The results must equal the oracle (which tests/test_oracle_ref_lbd.py pins to
the compiled reference) bit for bit.  A second harness (tests/host_core/lbd_kernels_emu.cpp) compiles the kernels' own source,
cs_lbd_kernels.cuh, against an emulation of the CUDA execution model -- a std::thread per CUDA thread, a std::barrier for __syncthreads,
function-local statics for __shared__, shuffles through a block-wide array -- and runs whole launches: the kernels' index arithmetic,
phase split and barrier placement give the oracle's bytes too.  What is left for the GPU box (tests/test_z_gpu_lbd_parity.py) is the
host-side launch code: allocations, copies, grid sizes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def core():
    src = os.path.join(HERE, "host_core", "lbd_core_host.cpp")
    hdr = os.path.join(HERE, "..", "cube_slam_b200", "csrc", "cs_lbd_core.h")
    out = os.path.join(HERE, "host_core", "_build", "liblbdcore.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", out, src])
    L = C.CDLL(out)
    L.host_lbd_round.argtypes = [C.c_float]
    L.host_lbd_key_dist.argtypes = [C.c_uint64]
    L.host_lbd_key_train.argtypes = [C.c_uint64]
    L.host_lbd_key_train.restype = C.c_uint32
    return L


@pytest.fixture(scope="module")
def emu():
    """cs_lbd_kernels.cuh itself, compiled against an emulation of the CUDA execution model (tests/host_core/lbd_kernels_emu.cpp)."""
    src = os.path.join(HERE, "host_core", "lbd_kernels_emu.cpp")
    deps = [src] + [os.path.join(HERE, "..", "cube_slam_b200", "csrc", f) for f in ("cs_lbd_core.h", "cs_lbd_kernels.cuh")]
    out = os.path.join(HERE, "host_core", "_build", "liblbdemu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def product():
    from cube_slam_b200 import _lib
    return _lib, _lib.load()


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _describe(core, product, oracle, img, kl, entry="host_lbd_describe"):
    """prepare on the product's host side, run the kernel arithmetic on the host, return (desc32, desc72)."""
    _lib, L = product
    n = len(kl)
    klp = np.ascontiguousarray(kl).view(_lib.KEYLINE_DTYPE)     # same 40-byte layout as the oracle's record
    lines = np.zeros((n, 6), np.float32)
    g, l = np.zeros(63, np.float32), np.zeros(21, np.float32)
    assert L.cs_lbd_debug_prepare(klp.ctypes.data, n, lines.ctypes.data, _p(g, C.c_float), _p(l, C.c_float)) == 0
    st = oracle.edl_detect(img, 15.0, want_stages=True)["stages"]      # the Sobel maps (the EDLines front end == computeSobel)
    dx, dy = np.ascontiguousarray(st["dx"]), np.ascontiguousarray(st["dy"])
    h, w = dx.shape
    desc, fdesc = np.zeros((n, 32), np.uint8), np.zeros((n, 72), np.float32)
    coef = np.concatenate([g, l]).astype(np.float32)
    getattr(core, entry)(lines.ctypes.data, n, _p(dx, C.c_int16), _p(dy, C.c_int16), w, h, _p(coef, C.c_float), _p(desc, C.c_uint8), _p(fdesc, C.c_float))
    return desc, fdesc


def test_weights_equal_the_oracles(product, oracle):
    _lib, L = product
    g, l = np.zeros(63, np.float32), np.zeros(21, np.float32)
    assert L.cs_lbd_debug_prepare(None, 0, None, _p(g, C.c_float), _p(l, C.c_float)) == 0
    G, Lw, _ = oracle.lbd_tables()
    np.testing.assert_array_equal(g, G.astype(np.float32))
    np.testing.assert_array_equal(l, Lw.astype(np.float32))


@pytest.mark.parametrize("use_lsd", [True, False])
def test_descriptors_of_the_demo_frame(core, product, oracle, fixture_a, use_lsd):
    img = fixture_a["img"]
    kl = oracle.lbd_detect_keylines(img, use_lsd, 15.0)
    want, fwant = oracle.lbd_compute(img, kl, want_float=True)
    got, fgot = _describe(core, product, oracle, img, kl)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(fgot, fwant)


def test_descriptors_of_sequence_and_synthetic_frames(core, product, oracle, fixture_b):
    from cube_slam_b200 import synthetic as S
    imgs = [fixture_b["frames"][i][0] for i in (0, 17, 40)] + list(S.make_batch(21, 2, 640, 480, 3)[0]) + list(S.make_batch(22, 1, 1242, 375, 3, kind="kitti")[0])
    for img in imgs:
        for use_lsd in (True, False):
            kl = oracle.lbd_detect_keylines(img, use_lsd, 15.0)
            want, fwant = oracle.lbd_compute(img, kl, want_float=True)
            got, fgot = _describe(core, product, oracle, img, kl)
            np.testing.assert_array_equal(got, want)
            np.testing.assert_array_equal(fgot, fwant)


def test_border_lines_short_lines_and_gray_input(core, product, oracle):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    img[30:90, 40:120] = 220
    rows = np.array([[2, 2, 150, 3], [5, 110, 5, 4], [158, 1, 158, 118], [0, 0, 159, 119], [80, 60, 80.4, 60.2], [10, 117, 150, 118.5],
                     [40, 30, 120, 30], [40.5, 90.2, 120.3, 89.7]], np.float32)
    rows = np.concatenate([rows, rng.uniform(0, 1, (200, 4)).astype(np.float32) * np.array([159, 119, 159, 119], np.float32)])
    _lib, L = product
    kl = oracle.lbd_keylines_from_lsd(rows, 160, 120)
    klp = np.zeros(len(rows), _lib.KEYLINE_DTYPE)
    assert L.cs_keylines_from_lines(_p(rows, C.c_float), len(rows), 160, 120, klp.ctypes.data) == 0
    for a, b in zip(("sx", "sy", "ex", "ey", "angle", "line_length", "response", "size", "num_pixels", "class_id"), _lib.KEYLINE_DTYPE.names):
        np.testing.assert_array_equal(kl[a], klp[b], err_msg=b)
    want, fwant = oracle.lbd_compute(img, kl, want_float=True)
    got, fgot = _describe(core, product, oracle, img, kl)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(fgot, fwant)        # NaNs included (assert_array_equal pairs them): a one-pixel line's variance can round below
    assert np.isnan(fwant).any()                       # zero, sqrt gives NaN and the normalisation spreads it -- in the reference as well


def test_edlines_key_lines_assembled_on_the_host(product, oracle, fixture_a, fixture_b):
    """cs_detect_descrip_lines' host step for the EDLines flavour: end points + {direction, numOfPixels} from the kernels -> KeyLine fields."""
    _lib, L = product
    for img in (fixture_a["img"], fixture_b["frames"][12][0]):
        h, w = img.shape[:2]
        want = oracle.lbd_detect_keylines(img, False, 15.0)
        rows = np.ascontiguousarray(np.stack([want["sx"], want["sy"], want["ex"], want["ey"]], 1), np.float32)
        extra = np.zeros((len(want), 2), np.float32)
        extra[:, 0] = want["angle"]
        extra[:, 1] = want["num_pixels"].astype(np.int32).view(np.float32)
        got = np.zeros(len(want), _lib.KEYLINE_DTYPE)
        assert L.cs_lbd_debug_keylines_edl(_p(rows, C.c_float), _p(extra, C.c_float), len(want), w, h, got.ctypes.data) == 0
        for a, b in zip(("sx", "sy", "ex", "ey", "angle", "line_length", "response", "size", "num_pixels", "class_id"), _lib.KEYLINE_DTYPE.names):
            np.testing.assert_array_equal(got[b], want[a], err_msg=b)


def test_round_is_half_away_from_zero(core):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-2000, 2000, 200000), np.arange(-50, 50) + 0.5, np.arange(-50, 50) - 0.5,
                         np.nextafter(np.float32(0.5), np.float32(0)) * np.ones(1), [0.49999997, -0.49999997, 1e-30, -1e-30, 0.0, 8388607.5]]).astype(np.float32)
    want = np.where(xs >= 0, np.floor(xs.astype(np.float64) + 0.5), -np.floor(-xs.astype(np.float64) + 0.5)).astype(np.int64)
    got = np.array([core.host_lbd_round(float(x)) for x in xs], np.int64)
    np.testing.assert_array_equal(got, want)


def _match(core, qs, ts, thres):
    """k_lbd_match's arithmetic on the host + the filter cs_match_line_descrip_batch applies to the keys."""
    qo = np.concatenate([[0], np.cumsum([len(q) for q in qs])]).astype(np.int32)
    to = np.concatenate([[0], np.cumsum([len(t) for t in ts])]).astype(np.int32)
    q = np.ascontiguousarray(np.concatenate(qs))
    t = np.ascontiguousarray(np.concatenate(ts))
    pq = np.concatenate([np.full(len(x), p, np.int32) for p, x in enumerate(qs)])
    keys = np.zeros(len(q), np.uint64)
    core.host_lbd_match(_p(q, C.c_uint8), _p(t, C.c_uint8), _p(pq, C.c_int32), _p(to, C.c_int32), len(q), _p(keys, C.c_uint64))
    res = []
    for p in range(len(qs)):
        qi, ti, di = [], [], []
        for i in range(qo[p], qo[p + 1]):
            key = int(keys[i])
            if key == 0xFFFFFFFFFFFFFFFF or to[p + 1] == to[p]:
                continue
            d = core.host_lbd_key_dist(C.c_uint64(key))
            if not np.float32(d) < np.float32(thres):
                continue
            qi.append(i - qo[p])
            ti.append(core.host_lbd_key_train(C.c_uint64(key)) if d <= 128 else -1)
            di.append(float(d))
        res.append((np.array(qi, np.int32), np.array(ti, np.int32), np.array(di, np.float32)))
    return res


def test_matcher_keys_reproduce_the_hash_order(core, oracle):
    rng = np.random.default_rng(11)

    def flip(c, bits):
        c = c.copy()
        for b in bits:
            c[b // 8] ^= np.uint8(1 << (b % 8))
        return c

    qs, ts = [], []
    for trial in range(25):
        nq, nt = int(rng.integers(1, 40)), int(rng.integers(6, 300))
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        q = np.stack([flip(t[int(rng.integers(0, nt))], rng.integers(0, 256, int(rng.integers(0, 40)))) for _ in range(nq)])
        k = int(rng.integers(1, 7))
        for j in range(min(6, nt)):                       # ties at the same distance from query 0, in different bytes / patterns
            t[(j * 7) % nt] = flip(q[0], [int(x) for x in rng.choice(256, k, replace=False)])
        t[nt - 1] = t[0]
        qs.append(q)
        ts.append(t)
    qs.append(rng.integers(0, 256, (100, 32), dtype=np.uint8))       # unrelated codes: far matches, some beyond 128
    ts.append(rng.integers(0, 256, (5, 32), dtype=np.uint8))
    for thres in (25.0, 300.0):
        got = _match(core, qs, ts, thres)                 # all pairs in one "launch", as the batch entry point does
        for (a, b, c), q, t in zip(got, qs, ts):
            wa, wb, wc = oracle.lbd_match(q, t, thres)
            np.testing.assert_array_equal(a, wa)
            np.testing.assert_array_equal(b, wb)
            np.testing.assert_array_equal(c, wc)


@pytest.mark.parametrize("use_lsd", [True, False])
def test_kernel_source_under_the_cuda_model_emulation_describe(emu, product, oracle, fixture_a, use_lsd):
    """k_lbd_describe as written (blockIdx / threadIdx indexing, shared arrays, four barriers), one std::thread per CUDA thread."""
    img = fixture_a["img"]
    kl = oracle.lbd_detect_keylines(img, use_lsd, 30.0 if use_lsd else 15.0)
    want, fwant = oracle.lbd_compute(img, kl, want_float=True)
    got, fgot = _describe(emu, product, oracle, img, kl, entry="emu_lbd_describe")
    assert len(kl) > 60
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(fgot, fwant)


def test_kernel_source_under_the_cuda_model_emulation_match(emu, core, oracle):
    """k_lbd_match as written (strided train loop, shuffle reduction, per-warp shared slots), several pairs in one launch."""
    rng = np.random.default_rng(4)
    qs, ts = [], []
    for nq, nt in ((5, 300), (3, 7), (4, 129), (2, 128)):
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        q = t[rng.integers(0, nt, nq)].copy()
        for i in range(nq):
            for b in rng.integers(0, 256, int(rng.integers(0, 30))):
                q[i, b // 8] ^= np.uint8(1 << (b % 8))
        t[nt - 1] = t[0]
        qs.append(q)
        ts.append(t)
    qo = np.concatenate([[0], np.cumsum([len(q) for q in qs])]).astype(np.int32)
    to = np.concatenate([[0], np.cumsum([len(t) for t in ts])]).astype(np.int32)
    q, t = np.ascontiguousarray(np.concatenate(qs)), np.ascontiguousarray(np.concatenate(ts))
    assert q.ctypes.data % 16 == 0 and t.ctypes.data % 16 == 0
    pq = np.concatenate([np.full(len(x), p, np.int32) for p, x in enumerate(qs)])
    keys_emu, keys_core = np.zeros(len(q), np.uint64), np.zeros(len(q), np.uint64)
    emu.emu_lbd_match(q.ctypes.data, t.ctypes.data, _p(pq, C.c_int32), _p(to, C.c_int32), len(q), _p(keys_emu, C.c_uint64))
    core.host_lbd_match(_p(q, C.c_uint8), _p(t, C.c_uint8), _p(pq, C.c_int32), _p(to, C.c_int32), len(q), _p(keys_core, C.c_uint64))
    np.testing.assert_array_equal(keys_emu, keys_core)
    for p_, (qq, tt) in enumerate(zip(qs, ts)):              # and the keys mean what the oracle says
        wq, wt, wd = oracle.lbd_match(qq, tt, 300.0)
        k = keys_emu[qo[p_]:qo[p_ + 1]]
        np.testing.assert_array_equal((k >> np.uint64(48)).astype(np.int64)[wq], wd.astype(np.int64))
        np.testing.assert_array_equal((k & np.uint64(0xFFFFFFFF)).astype(np.int64)[wq][wd <= 128], wt[wd <= 128])


def test_racecheck_of_the_kernels_under_threadsanitizer():
    """compute-sanitizer's racecheck, without a GPU: the kernels' source under the thread-per-CUDA-thread emulation, built with
    -fsanitize=thread.  No report as the kernels are; and -- the control that shows the check can see anything -- a report whenever any one of
    k_lbd_describe's four barriers (k_lbd_match's one) is skipped."""
    src = os.path.join(HERE, "host_core", "lbd_racecheck_main.cpp")
    bdir = os.path.join(HERE, "host_core", "_build")
    os.makedirs(bdir, exist_ok=True)

    def build_and_run(extra, name):
        exe = os.path.join(bdir, name)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-ffp-contract=off", "-o", exe, src] + extra)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        return r.stdout + r.stderr

    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", os.path.join(bdir, "tsan_probe")], input="int main(){return 0;}", text=True, capture_output=True)
    if probe.returncode != 0:
        pytest.skip("this g++ cannot link ThreadSanitizer")
    out = build_and_run([], "lbd_racecheck")
    assert "racecheck done" in out and "ThreadSanitizer" not in out, out[-2000:]
    for n in (1, 2, 3, 4):
        out = build_and_run(["-DCS_EMU_DROP_BARRIER=%d" % n], "lbd_racecheck_neg")
        assert "WARNING: ThreadSanitizer: data race" in out, "dropping barrier %d went unnoticed" % n


def test_pattern_order_is_numeric_order(oracle):
    """cs_lbd_match_key orders the s-bit xor patterns of a byte by their value; Mihasher::query's enumeration (restated in the oracle, which
    is pinned to the reference's matches) visits them in exactly that order."""
    _, _, rank = oracle.lbd_tables()
    for s in range(5):
        pats = [i for i in range(256) if bin(i).count("1") == s]
        assert sorted(pats, key=lambda i: rank[i]) == sorted(pats)
