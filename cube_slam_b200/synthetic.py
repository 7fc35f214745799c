"""Deterministic synthetic frames for tests and bench.py.

Each frame: mid-grey noisy background, the tile joints of a ground-plane grid, 25-50 clutter segments and one rendered shaded cuboid per 2D box
(standing on the ground plane, seen by a camera shaped like the reference's demo camera,
detect_3d_cuboid/src/main.cpp:35-44, or the KITTI camera of orb_object_slam's launch files).
Returned per frame: BGR image, camera-to-world T, N x 5 boxes [x y w h prob] (0-based, integer),
M x 4 line segments (cuboid edges + clutter with sub-pixel jitter -- the stand-in for a line detector's
output when lines are an input, as in orb_object_slam's online mode, Tracking.cc:1583-1590).
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

FIXTURE_K = np.array([[529.5, 0, 365.0], [0, 529.5, 265.0], [0, 0, 1.0]])
FIXTURE_T = np.array([[1, 0.0011, 0.0004, 0], [0, -0.3376, 0.9413, 0], [0.0011, -0.9413, -0.3376, 1.35], [0, 0, 0, 1.0]])
KITTI_K = np.array([[721.5377, 0, 609.5593], [0, 721.5377, 172.854], [0, 0, 1.0]])


def euler_zyx_to_rot(roll, pitch, yaw):
    cp, sp, sr, cr, sy, cy = np.cos(pitch), np.sin(pitch), np.sin(roll), np.cos(roll), np.sin(yaw), np.cos(yaw)
    return np.array([[cp * cy, sr * sp * cy - cr * sy, cr * sp * cy + sr * sy],
                     [cp * sy, sr * sp * sy + cr * cy, cr * sp * sy - sr * cy],
                     [-sp, sr * cp, cr * cp]])


def camera_for(width, height, kind="indoor"):
    if kind == "kitti":
        K = KITTI_K.copy()
        K[0, 2] = KITTI_K[0, 2] * width / 1242.0
        K[1, 2] = KITTI_K[1, 2] * height / 375.0
        return K
    f = 529.5 * width / 730.0
    return np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]])


def _pose(rng, kind):
    if kind == "kitti":
        # quaternion (-0.7071, 0, 0, 0.7071) (x y z w), height 1.7: camera looks along +y, level
        roll, pitch, yaw, hgt = -np.pi / 2, 0.0, 0.0, 1.7
    else:
        roll, pitch, yaw, hgt = -1.9152, -0.0011, -5e-5, 1.35
    roll += rng.uniform(-0.05, 0.05)
    pitch += rng.normal(0, 0.01)
    yaw += rng.normal(0, 0.01)
    hgt += rng.uniform(-0.2, 0.2)
    T = np.eye(4)
    T[:3, :3] = euler_zyx_to_rot(roll, pitch, yaw)
    T[2, 3] = hgt
    return T


def _project(K, T, pts_w):
    Tcw = np.linalg.inv(T)
    pc = (Tcw[:3, :3] @ pts_w.T + Tcw[:3, 3:4])
    uv = K @ pc
    return (uv[:2] / uv[2]).T, pc[2]


_BODY = np.array([[1, 1, -1, -1, 1, 1, -1, -1], [1, -1, -1, 1, 1, -1, -1, 1], [-1, -1, -1, -1, 1, 1, 1, 1]], float).T
_FACES = [((0, 1, 2, 3), 90), ((4, 5, 6, 7), 200), ((0, 1, 5, 4), 140), ((2, 3, 7, 6), 140), ((1, 2, 6, 5), 60), ((3, 0, 4, 7), 60)]
_EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def _floor_grid(rng, img, K, T, kind, lines):
    """Tile joints of the ground plane (z = 0): two families of parallel world lines, drawn thin and low-contrast before the objects.
    They give the line detectors the long converging segments of a real indoor / road scene (the fixture frame has 271 segments)."""
    height, width = img.shape
    yaw = rng.uniform(-np.pi / 4, np.pi / 4)
    step = rng.uniform(0.55, 0.95) if kind != "kitti" else rng.uniform(2.0, 3.5)
    far = 6.5 if kind != "kitti" else 35.0
    near = 0.9 if kind != "kitti" else 3.0
    c, s = np.cos(yaw), np.sin(yaw)
    val = int(np.clip(110 + rng.choice([-1, 1]) * rng.uniform(28, 60), 0, 255))
    for fam in range(2):
        d = np.array([c, s, 0.0]) if fam == 0 else np.array([-s, c, 0.0])
        n = np.array([-d[1], d[0], 0.0])
        for k in np.arange(-far, far, step):
            p0, p1 = n * (k + rng.uniform(-0.02, 0.02)) - d * far, n * k + d * far
            pts = np.stack([p0 + (p1 - p0) * t for t in np.linspace(0, 1, 65)])
            uv, depth = _project(K, T, pts)
            ok = (depth > near) & (depth < far)
            if ok.sum() < 2:
                continue
            idx = np.nonzero(ok)[0]
            a, b = uv[idx[0]], uv[idx[-1]]
            hit, q0, q1 = cv2.clipLine((0, 0, width, height), (int(round(a[0])), int(round(a[1]))), (int(round(b[0])), int(round(b[1]))))
            if not hit:
                continue
            cv2.line(img, q0, q1, val, 1, cv2.LINE_AA)
            lines.append([q0[0], q0[1], q1[0], q1[1]])


def make_frame(rng, width=640, height=480, n_boxes=3, kind="indoor", min_box=110, clutter=(25, 51), floor=True):
    """One synthetic frame -> (bgr uint8 HxWx3, T 4x4, boxes Nx5, lines Mx4)."""
    K = camera_for(width, height, kind)
    T = _pose(rng, kind)
    img = np.clip(110 + rng.normal(0, 4, (height, width)), 0, 255).astype(np.uint8)
    lines = []
    if floor:
        _floor_grid(rng, img, K, T, kind, lines)
    for _ in range(int(rng.integers(clutter[0], clutter[1]))):
        x0, y0 = rng.uniform(0, width - 1), rng.uniform(0, height - 1)
        ang, ln = rng.uniform(0, np.pi), rng.uniform(20, 200)
        x1, y1 = np.clip(x0 + ln * np.cos(ang), 0, width - 1), np.clip(y0 + ln * np.sin(ang), 0, height - 1)
        val = int(np.clip(110 + rng.choice([-1, 1]) * rng.uniform(30, 120), 0, 255))
        cv2.line(img, (int(round(x0)), int(round(y0))), (int(round(x1)), int(round(y1))), val, 1, cv2.LINE_AA)
        lines.append([x0, y0, x1, y1])
    objs = []
    tries = 0
    while len(objs) < n_boxes and tries < 400:
        tries += 1
        if kind == "kitti":
            dims = np.array([rng.uniform(3.5, 4.5), rng.uniform(1.5, 1.9), rng.uniform(1.2, 1.5)]) / 2  # lower than the 1.7 m camera so a top face is visible
            centre = np.array([rng.uniform(-6, 6), rng.uniform(5, 16), 0.0])
        else:
            dims = np.array([rng.uniform(0.45, 1.2), rng.uniform(0.45, 1.2), rng.uniform(0.7, 1.5)]) / 2
            centre = np.array([rng.uniform(-1.3, 1.3), rng.uniform(1.7, 3.6), 0.0])
        yaw = rng.uniform(-np.pi, np.pi)
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        pts = (Rz @ (_BODY * dims).T).T + centre + np.array([0, 0, dims[2]])
        uv, depth = _project(K, T, pts)
        if depth.min() < 0.3:
            continue
        l, t = np.floor(uv.min(0))
        r, b = np.ceil(uv.max(0))
        if l < 21 or r > width - 22 or t < 1 or b > height - 2 or (r - l) < min_box or (b - t) < min_box:
            continue
        objs.append((depth.mean(), pts, uv, (l, t, r, b)))
    objs.sort(key=lambda o: -o[0])  # far first
    boxes = []
    for _, pts, uv, (l, t, r, b) in objs:
        _, depth = _project(K, T, pts)
        order = sorted(_FACES, key=lambda fc: -np.mean([depth[i] for i in fc[0]]))
        for idx, shade in order:
            poly = np.round(uv[list(idx)] * 16).astype(np.int32)
            cv2.fillConvexPoly(img, poly, int(shade), cv2.LINE_AA, 4)
        for a, c in _EDGES:
            lines.append([uv[a, 0], uv[a, 1], uv[c, 0], uv[c, 1]])
        boxes.append([l, t, r - l, b - t, rng.uniform(0.3, 0.99)])
    lines = np.asarray(lines, np.float64).reshape(-1, 4)
    lines += rng.normal(0, 0.4, lines.shape)
    lines[:, [0, 2]] = np.clip(lines[:, [0, 2]], 0, width - 1)
    lines[:, [1, 3]] = np.clip(lines[:, [1, 3]], 0, height - 1)
    keep = np.hypot(lines[:, 2] - lines[:, 0], lines[:, 3] - lines[:, 1]) > 15
    lines = lines[keep]
    bgr = np.repeat(img[:, :, None], 3, axis=2)
    bgr[:, :, 0] = np.clip(bgr[:, :, 0].astype(int) - 6, 0, 255)
    bgr[:, :, 2] = np.clip(bgr[:, :, 2].astype(int) + 5, 0, 255)
    return np.ascontiguousarray(bgr), T, np.asarray(boxes, np.float64).reshape(-1, 5), lines, K


def make_batch(seed, n_frames, width=640, height=480, boxes_per_frame=3, kind="indoor", poisson=False, distinct=None):
    """A batch of frames sharing one K.  `distinct` limits how many different frames are rendered (the rest repeat
    them cyclically) so large benchmark batches are cheap to generate; content still exceeds L2."""
    rng = np.random.default_rng(seed)
    distinct = n_frames if distinct is None else min(distinct, n_frames)
    base = []
    for _ in range(distinct):
        nb = boxes_per_frame
        if poisson:
            nb = int(np.clip(rng.poisson(boxes_per_frame), 1, 6))
        for _attempt in range(20):
            fr = make_frame(rng, width, height, nb, kind)
            if len(fr[2]) >= 1:
                break
        base.append(fr)
    imgs = np.stack([base[i % distinct][0] for i in range(n_frames)])
    Ts = np.stack([base[i % distinct][1] for i in range(n_frames)])
    boxes = [base[i % distinct][2] for i in range(n_frames)]
    lines = [base[i % distinct][3] for i in range(n_frames)]
    return imgs, Ts, boxes, lines, base[0][4]
