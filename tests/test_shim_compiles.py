"""The C++ shims a maintainer adds to the reference packages (shim/*.cpp) are real translation units, not prose: they must compile.
Where Eigen / OpenCV C++ headers and the reference's own headers are on the include path the whole body is compiled; on this image
(none of them installed) the guard leaves an empty translation unit, which still proves the guard and the file are well-formed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["detect_3d_cuboid_b200.cpp", "line_lbd_b200.cpp"])
def test_shim_translation_unit_compiles(name, tmp_path):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    src = os.path.join(ROOT, "shim", name)
    inc = ["-I", os.path.join(ROOT, "include")]
    for d in ("/usr/include/eigen3", "/usr/include/opencv4"):
        if os.path.isdir(d):
            inc += ["-I", d]
    for env in ("CUBE_SLAM_REFERENCE_DETECT_INCLUDE", "CUBE_SLAM_REFERENCE_LINE_LBD_INCLUDE"):
        if os.environ.get(env):
            inc += ["-I", os.environ[env]]
    out = str(tmp_path / (name + ".o"))
    r = subprocess.run([cxx, "-std=c++14", "-Wall", "-c", src, "-o", out] + inc, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(out) > 0
