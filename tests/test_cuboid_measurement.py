"""cs_cuboid_measurement (the step after the path in object_slam/src/main_obj.cpp:455-473,505): the best cuboid as
a camera-frame measurement.  g2o / Eigen are not in this image, so the pin is scipy's independent rotation algebra (1e-12), plus
the algebraic identities of g2o::cuboid::transform_to / transform_from.  Host-only function: runs without a GPU."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial.transform import Rotation


def _call(rec, cam_t, cam_q, euler_raw=None):
    from cube_slam_b200 import _lib
    L = _lib.load()
    t, q, s = np.zeros(3), np.zeros(4), np.zeros(3)
    qual = C.c_double()
    e = None if euler_raw is None else _lib.ptr(np.ascontiguousarray(euler_raw, float), C.c_double)
    rc = L.cs_cuboid_measurement(rec.ctypes.data, _lib.ptr(np.ascontiguousarray(cam_t, float), C.c_double),
                                 _lib.ptr(np.ascontiguousarray(cam_q, float), C.c_double), e, _lib.ptr(t, C.c_double), _lib.ptr(q, C.c_double),
                                 _lib.ptr(s, C.c_double), C.byref(qual))
    assert rc == 0
    return t, q, s, qual.value


def _record(rng):
    import cube_slam_b200 as cs
    rec = np.zeros(1, cs.CUBOID_DTYPE)
    rec["pos"] = rng.normal(0, 3, 3)
    rec["scale"] = rng.uniform(0.1, 2, 3)
    rec["rotY"] = rng.uniform(-np.pi, np.pi)
    rec["normalized_error"] = rng.uniform(0, 1)
    rec["camera_roll_delta"] = rng.choice([-6, -3, 0, 3, 6]) / 180 * np.pi
    rec["camera_pitch_delta"] = rng.choice([-6, -3, 0, 3, 6]) / 180 * np.pi
    return rec


def test_measurement_against_scipy():
    rng = np.random.default_rng(5)
    for _ in range(200):
        rec = _record(rng)
        cam_q = rng.normal(0, 1, 4) * rng.choice([1.0, 3.0])       # SE3Quat(Vector7d) normalises; sign arbitrary
        cam_t = rng.normal(0, 2, 3)
        t, q, s, qual = _call(rec, cam_t, cam_q)
        Rc = Rotation.from_quat(cam_q / np.linalg.norm(cam_q))
        Ro = Rotation.from_euler("z", float(rec["rotY"][0]))
        t_ref = Rc.inv().apply(rec["pos"][0] - cam_t)
        R_ref = (Rc.inv() * Ro).as_matrix()
        np.testing.assert_allclose(t, t_ref, rtol=0, atol=1e-12)
        np.testing.assert_allclose(Rotation.from_quat(q).as_matrix(), R_ref, rtol=0, atol=1e-12)
        assert q[3] >= 0 and abs(np.linalg.norm(q) - 1) < 1e-15      # SE3Quat::normalizeRotation
        np.testing.assert_array_equal(s, rec["scale"][0])
        assert qual == (1 - float(rec["normalized_error"][0]) + 0.5) / 2
        # transform_from(Twc) of the measurement gives the ground-frame cuboid back (g2o_Object.h:118-133)
        back = Rc.apply(t) + cam_t
        np.testing.assert_allclose(back, rec["pos"][0], rtol=0, atol=1e-11)


def test_measurement_with_sampled_roll_pitch():
    """main_obj.cpp:463-471: the camera frame of the measurement is the raw pose with the winning roll / pitch deltas."""
    rng = np.random.default_rng(6)
    for _ in range(100):
        rec = _record(rng)
        euler = np.array([rng.uniform(-2.2, -1.6), rng.normal(0, 0.05), rng.uniform(-3, 3)])
        cam_t = rng.normal(0, 2, 3)
        t, q, s, qual = _call(rec, cam_t, [0, 0, 0, 1], euler)      # the quaternion argument is ignored in this mode
        e = euler + [float(rec["camera_roll_delta"][0]), float(rec["camera_pitch_delta"][0]), 0]
        Rc = Rotation.from_euler("ZYX", [e[2], e[1], e[0]])          # euler_zyx_to_rot: Rz(yaw) Ry(pitch) Rx(roll)
        t_ref = Rc.inv().apply(rec["pos"][0] - cam_t)
        R_ref = (Rc.inv() * Rotation.from_euler("z", float(rec["rotY"][0]))).as_matrix()
        np.testing.assert_allclose(t, t_ref, rtol=0, atol=1e-11)
        np.testing.assert_allclose(Rotation.from_quat(q).as_matrix(), R_ref, rtol=0, atol=1e-11)


def test_orb_variant_against_scipy():
    """orb_object_slam/src/Tracking.cc:1636-1647,1680-1687."""
    from cube_slam_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(9)
    for _ in range(100):
        rec = _record(rng)
        rec["pos"] = [rng.normal(0, 3), rng.normal(0, 3), rng.uniform(0, 40)]
        Rc = Rotation.random(random_state=int(rng.integers(1 << 30)))
        T = np.eye(4)
        T[:3, :3] = Rc.as_matrix()
        T[:3, 3] = rng.normal(0, 2, 3)
        conf = float(rng.choice([0.0, 0.4, 0.9]))
        t, q, s = np.zeros(3), np.zeros(4), np.zeros(3)
        qual = C.c_double()
        assert L.cs_cuboid_measurement_orb(rec.ctypes.data, _lib.ptr(np.ascontiguousarray(T), C.c_double), conf, _lib.ptr(t, C.c_double),
                                           _lib.ptr(q, C.c_double), _lib.ptr(s, C.c_double), C.byref(qual)) == 0
        t_ref = Rc.inv().apply(rec["pos"][0] - T[:3, 3])
        R_ref = (Rc.inv() * Rotation.from_euler("z", float(rec["rotY"][0]))).as_matrix()
        np.testing.assert_allclose(t, t_ref, rtol=0, atol=1e-11)
        np.testing.assert_allclose(Rotation.from_quat(q).as_matrix(), R_ref, rtol=0, atol=1e-11)
        ref_q = (60.0 - min(max(t[2], 10.0), 30.0)) / 40.0 * (conf if conf > 0 else 1.0)
        assert abs(qual.value - ref_q) < 1e-15


def test_python_wrapper_matches():
    import cube_slam_b200 as cs
    rng = np.random.default_rng(8)
    rec = _record(rng)
    cam_q, cam_t = rng.normal(0, 1, 4), rng.normal(0, 1, 3)
    a = _call(rec, cam_t, cam_q)
    b = cs.cuboid_measurement(rec, cam_t, cam_q)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_bad_arguments():
    from cube_slam_b200 import _lib
    L = _lib.load()
    assert L.cs_cuboid_measurement(None, None, None, None, None, None, None, None) != 0
