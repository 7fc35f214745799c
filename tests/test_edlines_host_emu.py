"""The EDLines translation unit (cube_slam_b200/csrc/cs_edlines.cu: kernels and host code) on the CPU, under an emulation of the CUDA
execution model -- a rehearsal of the one change this session made to kernels that had already been verified on the GPU: the two extra
stores per kept segment (direction, numOfPixels) that the descriptor needs, and of cs_edl_sobel_maps.

The source is compiled as it is, after three textual substitutions made here: `kernel<<<grid, block, smem, stream>>>(args)` becomes
`EMU_LAUNCH(kernel, grid, block)(args)`, `extern __shared__ T name[]` becomes a pointer to a block-wide byte array, and cs_tma.cuh gives way
to the inert interface in tests/host_core/cuda_emu_full.h (which also supplies threads for CUDA threads, per-warp barriers for the shuffle /
ballot primitives, atomics and the runtime calls).  Kernel transcendentals come from the host's libm here, so this is a check of logic and
indexing; on the frames used the segments nevertheless equal the oracle's exactly."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "cube_slam_b200", "csrc")


def _split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _preprocess(text):
    text = text.replace('#include "cs_tma.cuh"', "/* cs_tma.cuh: the inert interface of cuda_emu_full.h */")
    text, n_dyn = re.subn(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1 *\2 = (\1 *)g_blk.dyn;", text)
    assert n_dyn == 1
    out, pos, n = "", 0, 0
    for m in re.finditer(r"(\bk_\w+(?:<\w+>)?)<<<", text):
        end = text.index(">>>", m.end())
        cfg = _split_top_level(text[m.end():end])
        assert 2 <= len(cfg) <= 4, cfg
        out += text[pos:m.start()] + "EMU_LAUNCH(%s, %s, %s)" % (m.group(1), cfg[0], cfg[1])
        pos = end + 3
        n += 1
    assert n == 19, n
    return out + text[pos:]


PROLOGUE = r'''
#include "cuda_emu_full.h"
#include <cstdarg>
#include <cstdio>
#include <float.h>
#include <math.h>
#define __CUDACC__ 1 /* after the standard headers: the library's own headers show their device-side declarations */
struct cs_ctx {
    void *edl = nullptr;
    char err[512] = {0};
    int seq = 0;
};
cudaStream_t cs_ctx_stream(cs_ctx *) { return nullptr; }
'''
CTX_FUNCS = r'''
int cs_ctx_device(cs_ctx *) { return 0; }
int cs_ctx_fail(cs_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof c->err, fmt, ap);
    va_end(ap);
    return code;
}
void **cs_ctx_edl_slot(cs_ctx *c) { return &c->edl; }
void cs_ctx_count_launches(cs_ctx *, int64_t) {}
int cs_ctx_seq_lines(cs_ctx *c) { return c->seq; }
int cs_ctx_use_tma(cs_ctx *) { return 0; }
'''


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(CSRC, "cs_edlines.cu")
    glue = os.path.join(HERE, "host_core", "edlines_emu_glue.inc")
    hdr = os.path.join(HERE, "host_core", "cuda_emu_full.h")
    bdir = os.path.join(HERE, "host_core", "_build")
    os.makedirs(bdir, exist_ok=True)
    gen = os.path.join(bdir, "cs_edlines_emu.cpp")
    out = os.path.join(bdir, "libedlinesemu.so")
    deps = [src, glue, hdr, os.path.abspath(__file__)] + [os.path.join(CSRC, f) for f in ("cs_internal.h", "cs_nfa.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        body = _preprocess(open(src).read())
        i = body.index('#include "cs_internal.h"') + len('#include "cs_internal.h"')       # the context's narrow view is declared there
        open(gen, "w").write(PROLOGUE + body[:i] + CTX_FUNCS + body[i:] + open(glue).read())
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-w", "-ffp-contract=off", "-fno-fast-math", "-I", os.path.join(HERE, "host_core"),
                               "-I", os.path.join(HERE, "host_core", "fake_cuda_full"), "-I", CSRC, "-o", out, gen])
    L = C.CDLL(out)
    L.emu_ctx_new.restype = C.c_void_p
    L.emu_last_error.restype = C.c_char_p
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _run(emu, imgs, thres, cap, want, seq=0):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    F, H, W = imgs.shape[:3]
    ch = 1 if imgs.ndim == 3 else imgs.shape[3]
    ctx = C.c_void_p(emu.emu_ctx_new())
    lines, counts = np.zeros((F, cap, 4), np.float32), np.zeros(F, np.int32)
    extra, dx, dy = np.zeros((F, cap, 2), np.float32), np.zeros((F, H, W), np.int16), np.zeros((F, H, W), np.int16)
    rc = emu.emu_edl_run_keylines(ctx, _p(imgs, C.c_uint8), F, W, H, W * ch, ch, C.c_float(thres), cap, int(want), int(seq), _p(lines, C.c_float), _p(counts, C.c_int32),
                                  _p(extra, C.c_float), _p(dx, C.c_int16), _p(dy, C.c_int16))
    assert rc == 0, emu.emu_last_error(ctx)
    return lines, counts, extra, dx, dy


@pytest.mark.parametrize("seq", [0, 1])
def test_key_line_extras_of_the_edlines_kernels(emu, oracle, seq):
    """seq = 0: the walk-graph / warp-per-chain kernels (k_ed_fit -> k_ed_emit); seq = 1: the pixel-map kernel k_ed_route_fit."""
    rng = np.random.default_rng(2)
    img = np.full((96, 128, 3), 40, np.uint8)
    img[20:70, 25:100] = 190
    img[35:55, 50:80] = 90
    for y in range(96):                                       # a slanted edge as well
        img[y, max(0, 110 - y // 2):, :] = 230
    img = (img.astype(np.int16) + rng.integers(0, 5, img.shape)).clip(0, 255).astype(np.uint8)
    imgs = np.stack([img, img[::-1].copy()])
    cap = 256
    lines, counts, extra, dx, dy = _run(emu, imgs, 15.0, cap, True, seq)
    plain = _run(emu, imgs, 15.0, cap, False, seq)
    for f in range(2):
        want = oracle.lbd_detect_keylines(imgs[f], False, 15.0)
        st = oracle.edl_detect(imgs[f], 15.0, want_stages=True)["stages"]
        assert counts[f] == len(want) >= 3
        n = counts[f]
        np.testing.assert_array_equal(lines[f, :n], np.stack([want["sx"], want["sy"], want["ex"], want["ey"]], 1))
        np.testing.assert_array_equal(extra[f, :n, 0], want["angle"])                               # KeyLine::angle = lineDirection_
        np.testing.assert_array_equal(extra[f, :n, 1].view(np.int32), want["num_pixels"])            # KeyLine::numOfPixels, an integer's bits
        np.testing.assert_array_equal(dx[f], st["dx"])
        np.testing.assert_array_equal(dy[f], st["dy"])
        # without the request the detection is what it was: same segments, same counts
        assert plain[1][f] == n
        np.testing.assert_array_equal(plain[0][f, :n], lines[f, :n])


def test_key_line_extras_on_a_real_image_crop(emu, oracle, fixture_a):
    """A 320 x 240 crop of the reference's demo frame: a dozen and more segments per frame, chains that split into several lines."""
    img = np.ascontiguousarray(fixture_a["img"][150:390, 200:520])
    lines, counts, extra, dx, dy = _run(emu, img[None], 15.0, 512, True, 0)
    want = oracle.lbd_detect_keylines(img, False, 15.0)
    n = int(counts[0])
    assert n == len(want) >= 10
    np.testing.assert_array_equal(lines[0, :n], np.stack([want["sx"], want["sy"], want["ex"], want["ey"]], 1))
    np.testing.assert_array_equal(extra[0, :n, 0], want["angle"])
    np.testing.assert_array_equal(extra[0, :n, 1].view(np.int32), want["num_pixels"])


def test_sobel_maps_entry_point(emu, oracle, fixture_b):
    imgs = np.stack([fixture_b["frames"][0][0][:160, :200], fixture_b["frames"][9][0][:160, :200]])
    for data in (imgs, np.ascontiguousarray(imgs[..., 1])):                                          # BGR and one channel
        F, H, W = data.shape[:3]
        ch = 1 if data.ndim == 3 else 3
        data = np.ascontiguousarray(data)
        ctx = C.c_void_p(emu.emu_ctx_new())
        dx, dy = np.zeros((F, H, W), np.int16), np.zeros((F, H, W), np.int16)
        assert emu.emu_edl_sobel_maps(ctx, _p(data, C.c_uint8), F, W, H, W * ch, ch, _p(dx, C.c_int16), _p(dy, C.c_int16)) == 0
        for f in range(F):
            st = oracle.edl_detect(data[f], 15.0, want_stages=True)["stages"]
            np.testing.assert_array_equal(dx[f], st["dx"])
            np.testing.assert_array_equal(dy[f], st["dy"])
