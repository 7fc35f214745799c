import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """tests/test_z_gpu_lbd_parity.py holds the GPU tests of code that has not run on a GPU yet.  Within it, the cases that depend on the one
    change made to kernels that HAVE been verified -- the EDLines kernels' two extra stores per kept segment (direction, numOfPixels; parameter
    values False / 0 / "edlines") -- go last, so that with -x a problem there cannot hide the results of everything else."""
    z = [i for i, it in enumerate(items) if it.fspath.basename == "test_z_gpu_lbd_parity.py"]
    if not z:
        return

    def late(it):
        cs = getattr(it, "callspec", None)
        return cs is not None and any(v is False or v == "edlines" or (v == 0 and not isinstance(v, bool)) for v in cs.params.values())
    block = [items[i] for i in z]
    block = [it for it in block if not late(it)] + [it for it in block if late(it)]
    for i, it in zip(z, block):
        items[i] = it


def _imread(path):
    import cv2
    img = cv2.imread(path, 1)
    assert img is not None, path
    return img


@pytest.fixture(scope="session")
def fixture_a():
    """Inputs of the reference's single-frame demo (detect_3d_cuboid/src/main.cpp:35-48)."""
    d = os.path.join(GOLD, "fixture_a")
    meta = json.load(open(os.path.join(d, "meta.json")))
    return dict(img=_imread(os.path.join(d, "0000_rgb_raw.jpg")), K=np.array(meta["K"], float), T=np.array(meta["T"], float),
                boxes=np.array(meta["boxes"], float), lines=np.loadtxt(os.path.join(d, "0000_edge.txt")),
                expected=json.load(open(os.path.join(GOLD, "expected_fixture_a.json"))))


@pytest.fixture(scope="session")
def fixture_b():
    """The 58-frame object_slam/data sequence (inputs only; object_slam/src/main_obj.cpp:392-450)."""
    d = os.path.join(GOLD, "fixture_b")
    meta = json.load(open(os.path.join(d, "meta.json")))
    frames = []
    for i in range(meta["n_frames"]):
        img = _imread(os.path.join(d, "raw_imgs", "%04d_rgb_raw.jpg" % i))
        txt = os.path.join(d, "filter_2d_obj_txts", "%04d_yolo2_0.15.txt" % i)
        boxes = np.loadtxt(txt, ndmin=2) if os.path.getsize(txt) > 0 else np.zeros((0, 5))
        boxes = boxes.reshape(-1, 5).copy()
        boxes[:, :2] -= 1  # matlab -> c++ coordinates (main_obj.cpp:439)
        frames.append((img, boxes))
    return dict(K=np.array(meta["K"], float), T=np.array(meta["T"], float), frames=frames)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle
