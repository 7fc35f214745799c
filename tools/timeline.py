"""Timeline of K batches in flight (bench.py's --inflight loop): per-stage start / end of the last round of steps, from CUDA events.
python tools/timeline.py [K] [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cube_slam_b200 as cs
from cube_slam_b200 import _lib
import bench

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = bench.make_workload("c3", 0)
params = cs.default_params(**wl["over"])
F, h, w = wl["F"], wl["h"], wl["w"]
ctxs = []
for _ in range(K):
    c = cs.Context(0, w, h, F, 16, 8192)
    c.set_calibration(wl["K"])
    c.upload(wl["imgs"], wl["Ts"], wl["boxes"], wl["lines"], params)
    c.set_profiling(1)
    ctxs.append(c)
for i in range(steps * K):
    ctxs[i % K].run_async()
torch.cuda.synchronize()
names = ["lsd", "gray", "canny", "hyst", "dt", "lines", "sweep", "fuse", "end"]
rows = []
for k, c in enumerate(ctxs):
    off = np.zeros(9, np.float32)
    c.check(c.L.cs_debug_stage_offsets(c.h, ctxs[0].h, _lib.ptr(off, C.c_float)))
    rows.append(off)
    print("ctx %d: " % k + "  ".join("%s@%.3f" % (n, o) for n, o in zip(names, off)))
# merged event list
ev = []
for k, off in enumerate(rows):
    for s in range(8):
        if off[s + 1] - off[s] > 0.004:
            ev.append((off[s], off[s + 1], k, names[s]))
for a, b, k, n in sorted(ev):
    print("%8.3f -> %8.3f  (%.3f ms)  ctx%d %s" % (a, b, b - a, k, n))
