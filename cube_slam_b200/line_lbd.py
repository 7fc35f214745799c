"""Host-side mirror of `class line_lbd_detect` (line_lbd/include/line_lbd/line_lbd_allclass.h:22-70), detection half.

detect_filter_lines(img) -> n x 4 float32 [x1 y1 x2 y2]: what the reference writes into its `cv::Mat& linesmat_out`
(line_lbd/class/line_lbd_allclass.cpp:216-221).  Descriptor / matcher methods are out of scope."""
import ctypes as C

import numpy as np

from . import _lib
from .detect_3d_cuboid import Context, CubeSlamError


class line_lbd_detect(object):
    def __init__(self, numoctaves=1, octaveratio=1.0, device=0, max_width=2048, max_height=2048, context=None):
        self.numoctaves_ = int(numoctaves)
        self.octaveratio_ = float(octaveratio)
        self.use_LSD = False            # line_lbd_allclass.cpp:121
        self.line_length_thres = 50.0   # :122
        self._ctx = context if context is not None else Context(device, max_width, max_height, 1, 1, 1)

    def params(self):
        p = _lib.LineParams()
        self._ctx.L.cs_default_line_params(C.byref(p))
        p.use_LSD = int(bool(self.use_LSD))
        p.numoctaves = self.numoctaves_
        p.octaveratio = self.octaveratio_
        p.line_length_thres = float(self.line_length_thres)
        return p

    def detect_filter_lines(self, gray_img, cap=8192):
        """One frame (H x W or H x W x 3 uint8) -> n x 4 float32."""
        return self.detect_filter_lines_batch(np.asarray(gray_img)[None], cap)[0]

    def detect_filter_lines_batch(self, imgs, cap=4096):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        if imgs.ndim == 3:
            F, H, W = imgs.shape
            ch = 1
        else:
            F, H, W, ch = imgs.shape
        out = np.zeros((F, cap, 4), np.float32)
        n = np.zeros(F, np.int32)
        p = self.params()
        rc = self._ctx.L.cs_detect_lines_batch(self._ctx.h, imgs.ctypes.data, F, W, H, W * ch, ch, C.byref(p), _lib.ptr(out, C.c_float), cap,
                                               _lib.ptr(n, C.c_int32))
        if rc != 0:
            raise CubeSlamError("%s: %s" % (_lib.STATUS_NAMES.get(rc, rc), self._ctx.L.cs_last_error(self._ctx.h).decode()))
        return [out[f, :n[f]].copy() for f in range(F)]

    def debug_frame(self, frame=0, cap=8192):
        L = self._ctx.L
        wh = np.zeros(2, np.int32)
        self._ctx.check(L.cs_debug_lsd(self._ctx.h, frame, _lib.ptr(wh, C.c_int32), None, None, None, None, None, None, None, 0))
        W, H = int(wh[0]), int(wh[1])
        sc, mg, an = np.zeros((H, W)), np.zeros((H, W)), np.zeros((H, W))
        lst = np.zeros(W * H, np.int32)
        ll, nr = C.c_int32(), C.c_int32()
        raw = np.zeros((cap, 4), np.float32)
        self._ctx.check(L.cs_debug_lsd(self._ctx.h, frame, _lib.ptr(wh, C.c_int32), _lib.ptr(sc, C.c_double), _lib.ptr(mg, C.c_double),
                                       _lib.ptr(an, C.c_double), _lib.ptr(lst, C.c_int32), C.byref(ll), _lib.ptr(raw, C.c_float), C.byref(nr), cap))
        return dict(scaled=sc, modgrad=mg, angles=an, list=lst[:ll.value].copy(), raw_lines=raw[:nr.value].copy())

    def seed_loop_stats(self, n_frames):
        """Diagnostics of the last LSD run: (n_frames x 4 int32 {rounds, processed, refused, re-grown}, redo flags)."""
        st, redo = np.zeros((n_frames, 4), np.int32), np.zeros(n_frames, np.int32)
        self._ctx.check(self._ctx.L.cs_debug_lsd_stats(self._ctx.h, _lib.ptr(st, C.c_int32), _lib.ptr(redo, C.c_int32), n_frames))
        return st, redo

    def debug_frame_edlines(self, width, height, frame=0, cap=8192):
        """EDLineDetector's intermediate maps of one frame of the last use_LSD = False run (tests)."""
        L = self._ctx.L
        H, W = int(height), int(width)
        blur, dirm, edge = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)
        dx, dy, g = np.zeros((H, W), np.int16), np.zeros((H, W), np.int16), np.zeros((H, W), np.int16)
        anchors = np.zeros(W * H // 5 + 1, np.int32)
        na, nr = C.c_int32(), C.c_int32()
        raw = np.zeros((cap, 4), np.float32)
        self._ctx.check(L.cs_debug_edlines(self._ctx.h, frame, _lib.ptr(blur, C.c_uint8), _lib.ptr(dx, C.c_int16), _lib.ptr(dy, C.c_int16),
                                           _lib.ptr(g, C.c_int16), _lib.ptr(dirm, C.c_uint8), _lib.ptr(anchors, C.c_int32), C.byref(na),
                                           _lib.ptr(edge, C.c_uint8), _lib.ptr(raw, C.c_float), C.byref(nr), cap))
        return dict(blur=blur, dx=dx, dy=dy, g=g, dir=dirm, anchors=anchors[:na.value].copy(), edge=edge, raw_lines=raw[:nr.value].copy())
