"""cs_lbd.cu's own host code AND kernels on the CPU: the GPU parity test's assertions, run against an emulated device.

tests/host_core/lbd_host_emu.cpp compiles cube_slam_b200/csrc/cs_lbd.cu itself with g++: the CUDA execution model and the few runtime calls
are emulated (tests/host_core/cuda_emu.h), the detector entry points it calls in other translation units answer from the oracle in the
layout the real ones leave in HBM.  The library's real entry points -- cs_detect_descrip_lines_batch, cs_lbd_compute[_batch],
cs_match_line_descrip[_batch] -- are then driven through the unchanged Python mirror, and the test functions of
tests/test_z_gpu_lbd_parity.py are called on it as they are.  So the first GPU run of that file has already been rehearsed down to the
host-side slotting, offsets, copies and launches; what only the GPU can show is the real runtime and the detector kernels' two new stores."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DEVICE_ENTRY_POINTS = ("cs_detect_descrip_lines", "cs_detect_descrip_lines_batch", "cs_lbd_compute", "cs_lbd_compute_batch", "cs_match_line_descrip",
                       "cs_match_line_descrip_batch")


class EmuLib(object):
    """The real libcubeslam_b200.so for everything host-only, the emulated build of cs_lbd.cu for its device entry points."""

    def __init__(self, real, emu):
        self._real, self._emu = real, emu
        for name in DEVICE_ENTRY_POINTS:
            fn = getattr(emu, name)
            fn.argtypes = getattr(real, name).argtypes
            fn.restype = C.c_int

    def __getattr__(self, name):
        return getattr(self._emu if name in DEVICE_ENTRY_POINTS else self._real, name)


class EmuContext(object):
    def __init__(self, L, emu):
        self.L, self._emu = L, emu
        emu.emu_ctx_new.restype = C.c_void_p
        emu.emu_last_error.restype = C.c_char_p
        emu.emu_last_error.argtypes = [C.c_void_p]
        self.h = emu.emu_ctx_new()

    def check(self, rc):
        if rc != 0:
            import cube_slam_b200 as cs
            raise cs.CubeSlamError("%d: %s" % (rc, self._emu.emu_last_error(self.h).decode()))


@pytest.fixture(scope="module")
def det(oracle):
    import cube_slam_b200 as cs
    from cube_slam_b200 import _lib
    src = os.path.join(HERE, "host_core", "lbd_host_emu.cpp")
    csrc = os.path.join(HERE, "..", "cube_slam_b200", "csrc")
    deps = [src, os.path.join(HERE, "host_core", "cuda_emu.h")] + [os.path.join(csrc, f) for f in ("cs_lbd.cu", "cs_lbd_core.h", "cs_lbd_kernels.cuh", "cs_internal.h")]
    out = os.path.join(HERE, "host_core", "_build", "liblbdhostemu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    orc = os.path.join(HERE, "..", "oracle", "_build")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", "-o", out, src,
                               "-L", orc, "-loracle", "-Wl,-rpath," + os.path.abspath(orc)])
    emu = C.CDLL(out)
    d = cs.line_lbd_detect(context=EmuContext(EmuLib(_lib.load(), emu), emu))
    d.line_length_thres = 15
    return d


import test_z_gpu_lbd_parity as G  # noqa: E402  (the GPU test file: its functions take the detector as an argument)


@pytest.mark.parametrize("use_lsd,n", [(True, 271), (False, 88)])
def test_detect_descrip_lines_demo_frame(det, oracle, fixture_a, use_lsd, n):
    G.test_detect_descrip_lines_demo_frame(det, oracle, fixture_a, use_lsd, n)


@pytest.mark.parametrize("use_lsd", [True, False])
def test_batches(det, oracle, fixture_b, use_lsd):
    G.test_batches_of_sequence_and_synthetic_frames(det, oracle, fixture_b, use_lsd)


@pytest.mark.parametrize("use_lsd", [True, False])
def test_mat_overload_octaves_variant_and_gray_input(det, oracle, fixture_a, use_lsd):
    G.test_mat_overload_octaves_variant_and_gray_input(det, oracle, fixture_a, use_lsd)


def test_given_keylines_border_and_degenerate_lines(det, oracle):
    G.test_given_keylines_border_and_degenerate_lines(det, oracle)


def test_match_line_descrip(det, oracle, fixture_b):
    G.test_match_line_descrip(det, oracle, fixture_b)


@pytest.mark.parametrize("flavour", ["lsd", "edlines"])
def test_against_the_committed_goldens(det, flavour):
    G.test_against_the_committed_goldens(det, flavour)


def test_cpp_shim_members_linked_against_the_emulated_library(oracle, fixture_a, fixture_b, monkeypatch):
    """shim/line_lbd_b200.cpp + shim/test/line_shim_driver.cpp -- the C++ a maintainer compiles into the reference's line_lbd package -- built
    against the reference's own class header as always, but linked against the emulated build of the library instead of libcubeslam_b200.so:
    the GPU test of the shim's descriptor / matcher members, run on the CPU.  Needs the reference checkout (its headers), like oracle/_ref."""
    ref_inc = "/root/reference/line_lbd/include"
    if not os.path.isdir(ref_inc):
        pytest.skip("the reference's headers are not on this machine")
    root = os.path.join(HERE, "..")
    out = os.path.join(HERE, "host_core", "_build", "libshim_line_emu.so")
    orc = os.path.abspath(os.path.join(root, "oracle", "_build"))
    srcs = [os.path.join(root, "shim", "line_lbd_b200.cpp"), os.path.join(root, "shim", "test", "line_shim_driver.cpp"), os.path.join(HERE, "host_core", "lbd_host_emu.cpp")]
    deps = srcs + [os.path.join(root, "cube_slam_b200", "csrc", f) for f in ("cs_lbd.cu", "cs_lbd_core.h", "cs_lbd_kernels.cuh")] + [os.path.join(root, "include", "cube_slam_b200.h")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-w", "-ffp-contract=off", "-fno-fast-math", "-DCS_EMU_WITH_SHIM_GLUE",
                               "-I", os.path.join(root, "include"), "-I", os.path.join(root, "oracle", "ref", "fakecv"), "-I", ref_inc, "-o", out] + srcs +
                              ["-L", orc, "-loracle", "-Wl,-rpath," + orc])
    monkeypatch.setattr(G, "SHIM", out)
    for use_lsd in (1, 0):
        G.test_cpp_shim_descriptor_and_matcher_members(oracle, fixture_a, fixture_b, use_lsd)


def test_detect_raw_lines_is_not_part_of_this_rehearsal(det):
    """detect_raw_lines goes through cs_detect_lines_batch (cs_lsd.cu / cs_edlines.cu, verified on the GPU since round 1), not cs_lbd.cu."""
    assert "cs_detect_lines_batch" not in DEVICE_ENTRY_POINTS


def test_errors_and_capacity(det, fixture_a):
    import cube_slam_b200 as cs
    det.use_LSD = True
    with pytest.raises(cs.CubeSlamError, match="exceed max_lines_per_frame"):
        det.detect_descrip_lines_batch(fixture_a["img"][None], cap=3)
    before = det._ctx._emu.emu_launches(C.c_void_p(det._ctx.h))
    det.use_LSD = False
    det.detect_descrip_lines(fixture_a["img"])
    assert det._ctx._emu.emu_launches(C.c_void_p(det._ctx.h)) == before + 1      # one descriptor launch; the detector is a stand-in here
