"""The Python mirror of class line_lbd_detect (cube_slam_b200/line_lbd.py), descriptor / matcher methods, WITHOUT a GPU: the ctypes
marshalling -- record layouts, per-frame slots, CSR offsets, the threshold juggling of the Mat overload, the octaves variant's swap -- is
exercised against a stand-in for the five device entry points that answers from the CPU oracle through the very same C signatures
(pointers and sizes in, buffers filled).  Everything else (cs_default_line_params, cs_keylines_from_lines) is the real library.  What the
kernels compute is not tested here (tests/test_lbd_host_core.py, tests/test_z_gpu_lbd_parity.py); what the GPU test would trip over in the
Python layer is."""
import ctypes as C

import numpy as np
import pytest


def _view(addr, dtype, count):
    if hasattr(addr, "contents") or isinstance(addr, C._Pointer):
        addr = C.cast(addr, C.c_void_p).value
    if isinstance(addr, C.c_void_p):
        addr = addr.value
    dtype = np.dtype(dtype)
    buf = (C.c_char * (dtype.itemsize * count)).from_address(int(addr))
    return np.frombuffer(buf, dtype=dtype, count=count)


class FakeLib(object):
    """Delegates to the real libcubeslam_b200.so except for the calls that need a device."""

    def __init__(self, real, oracle, _lib):
        self._real, self._o, self._lib = real, oracle, _lib
        self.calls = []

    def __getattr__(self, name):
        return getattr(self._real, name)

    def _frames(self, imgs, F, W, H, stride, ch):
        a = _view(imgs, np.uint8, F * H * stride).reshape(F, H, stride)[:, :, :W * ch]
        return a.reshape(F, H, W, ch) if ch == 3 else a.reshape(F, H, W)

    def cs_detect_lines_batch(self, h, imgs, F, W, H, stride, ch, params, out, cap, n):
        p = params._obj
        frames = self._frames(imgs, F, W, H, stride, ch)
        o, nn = _view(out, np.float32, F * cap * 4).reshape(F, cap, 4), _view(n, np.int32, F)
        for f in range(F):
            lines = (self._o.lsd_detect if p.use_LSD else self._o.edl_detect)(frames[f], float(p.line_length_thres))["lines"]
            o[f, :len(lines)] = lines
            nn[f] = len(lines)
        self.calls.append(("detect_lines", float(p.line_length_thres)))
        return 0

    def cs_detect_descrip_lines_batch(self, h, imgs, F, W, H, stride, ch, params, kl, desc, cap, n):
        p = params._obj
        frames = self._frames(imgs, F, W, H, stride, ch)
        k = _view(kl, self._lib.KEYLINE_DTYPE, F * cap).reshape(F, cap)
        d, nn = _view(desc, np.uint8, F * cap * 32).reshape(F, cap, 32), _view(n, np.int32, F)
        for f in range(F):
            want = self._o.lbd_detect_keylines(frames[f], bool(p.use_LSD), float(p.line_length_thres))
            k[f, :len(want)] = want.view(self._lib.KEYLINE_DTYPE)
            d[f, :len(want)] = self._o.lbd_compute(frames[f], want)
            nn[f] = len(want)
        self.calls.append(("detect_descrip", float(p.line_length_thres)))
        return 0

    def cs_lbd_compute(self, h, img, W, H, stride, ch, kl, n, desc, fdesc):
        frame = self._frames(img, 1, W, H, stride, ch)[0]
        k = _view(kl, self._lib.KEYLINE_DTYPE, n).view(self._o.KEYLINE_DTYPE) if n else np.zeros(0, self._o.KEYLINE_DTYPE)
        if n:
            dd, ff = self._o.lbd_compute(frame, k, want_float=True)
            _view(desc, np.uint8, n * 32).reshape(n, 32)[:] = dd
            if fdesc:
                _view(fdesc, np.float32, n * 72).reshape(n, 72)[:] = ff
        return 0

    def cs_match_line_descrip_batch(self, h, q, qo, t, to, n_pairs, thres, out, n):
        qo, to = _view(qo, np.int32, n_pairs + 1), _view(to, np.int32, n_pairs + 1)
        qq, tt = _view(q, np.uint8, max(int(qo[-1]), 1) * 32).reshape(-1, 32), _view(t, np.uint8, max(int(to[-1]), 1) * 32).reshape(-1, 32)
        o, nn = _view(out, self._lib.DMATCH_DTYPE, max(int(qo[-1]), 1)), _view(n, np.int32, n_pairs)
        for p in range(n_pairs):
            a, b, c = self._o.lbd_match(qq[qo[p]:qo[p + 1]], tt[to[p]:to[p + 1]], thres.value)
            m = o[qo[p]:qo[p] + len(a)]
            m["query_idx"], m["train_idx"], m["img_idx"], m["distance"] = a, b, 0, c
            nn[p] = len(a)
        return 0

    def cs_match_line_descrip(self, h, q, nq, t, nt, thres, out, n):
        qo, to, nn = np.array([0, nq], np.int32), np.array([0, nt], np.int32), np.zeros(1, np.int32)
        self.cs_match_line_descrip_batch(h, q, qo.ctypes.data, t, to.ctypes.data, 1, thres, out, nn.ctypes.data)
        n._obj.value = int(nn[0])
        return 0


class FakeContext(object):
    def __init__(self, L):
        self.L, self.h = L, None

    def check(self, rc):
        assert rc == 0


@pytest.fixture()
def det(oracle):
    import cube_slam_b200 as cs
    from cube_slam_b200 import _lib
    d = cs.line_lbd_detect(context=FakeContext(FakeLib(_lib.load(), oracle, _lib)))
    d.line_length_thres = 15
    return d


def test_detect_descrip_lines_batch_slots_and_layout(det, oracle, fixture_b):
    imgs = np.stack([fixture_b["frames"][i][0] for i in (0, 9, 33)])
    for use_lsd in (True, False):
        det.use_LSD = use_lsd
        out = det.detect_descrip_lines_batch(imgs, cap=512)
        for f, (kl, desc) in enumerate(out):
            want = oracle.lbd_detect_keylines(imgs[f], use_lsd, 15.0)
            assert len(kl) == len(want) and kl.dtype.itemsize == 40
            for a, b in zip(kl.dtype.names, want.dtype.names):
                np.testing.assert_array_equal(kl[a], want[b])
            np.testing.assert_array_equal(desc, oracle.lbd_compute(imgs[f], want))


def test_mat_overload_octaves_variant_raw_lines_and_given_lines(det, oracle, fixture_a):
    img = fixture_a["img"]
    det.use_LSD = True
    lines, desc = det.detect_descrip_lines(img, as_mat=True)
    want = oracle.lbd_detect_keylines(img, True, -1.0)
    assert det.line_length_thres == 15 and det._ctx.L.calls[-1] == ("detect_descrip", -1.0) and lines.shape == (445, 4) and lines.dtype == np.float32
    np.testing.assert_array_equal(lines, np.stack([want["sx"], want["sy"], want["ex"], want["ey"]], 1))
    np.testing.assert_array_equal(desc, oracle.lbd_compute(img, want))
    kls, descs = det.detect_descrip_lines_octaves(img)
    w15 = oracle.lbd_detect_keylines(img, True, 15.0)
    wz = oracle.lbd_order_keylines(w15)
    assert len(kls) == len(descs) == 1 and (wz["sx"] != w15["sx"]).any()
    for a, b in zip(kls[0].dtype.names, wz.dtype.names):
        if a != "class_id":
            np.testing.assert_array_equal(kls[0][a], wz[b], err_msg=a)
    np.testing.assert_array_equal(kls[0]["class_id"], np.arange(len(wz)))
    np.testing.assert_array_equal(descs[0], oracle.lbd_compute(img, w15))
    np.testing.assert_array_equal(det.detect_raw_lines(img), oracle.lsd_detect(img, -1.0)["lines"])
    assert det.line_length_thres == 15
    # get_line_descriptors: rows -> key lines (real host code) -> descriptors
    rows = np.stack([w15["sx"], w15["sy"], w15["ex"], w15["ey"]], 1)
    np.testing.assert_array_equal(det.get_line_descriptors(img, rows), oracle.lbd_compute(img, oracle.lbd_keylines_from_lsd(rows, img.shape[1], img.shape[0])))
    d, f = det.compute_descriptors(img, det.keylines_from_lines(rows, img.shape[1], img.shape[0]), want_float=True)
    assert d.shape == (len(rows), 32) and f.shape == (len(rows), 72) and f.dtype == np.float32
    assert det.compute_descriptors(img, np.zeros(0, kls[0].dtype)).shape == (0, 32)
    det.numoctaves_ = 2
    import cube_slam_b200 as cs
    with pytest.raises(cs.CubeSlamError):
        det.detect_descrip_lines_octaves(img)


def test_match_marshalling(det, oracle):
    rng = np.random.default_rng(5)
    qs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (7, 1, 30)]
    ts = [np.concatenate([q[::2], rng.integers(0, 256, (5, 32), dtype=np.uint8)]) for q in qs]
    batch = det.match_line_descrip_batch(qs, ts, 300.0)
    for m, q, t in zip(batch, qs, ts):
        a, b, c = oracle.lbd_match(q, t, 300.0)
        assert m.dtype.itemsize == 16
        np.testing.assert_array_equal(m["query_idx"], a)
        np.testing.assert_array_equal(m["train_idx"], b)
        np.testing.assert_array_equal(m["distance"], c)
    one = det.match_line_descrip(qs[2], ts[2], 300.0)
    np.testing.assert_array_equal(one, batch[2])
    assert len(det.match_line_descrip(qs[0][:0], ts[0])) == 0
