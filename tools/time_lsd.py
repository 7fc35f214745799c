"""Times cs_detect_lines_batch (host buffers in/out) and reports frames/s; python tools/time_lsd.py [frames] [lsd|edlines]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cube_slam_b200 as cs
from cube_slam_b200 import synthetic as S
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
imgs, Ts, boxes, lines, K = S.make_batch(5, F, 640, 480, 3, poisson=True, distinct=32)
ctx = cs.Context(0, 640, 480, F, 16, 8192)
det = cs.line_lbd_detect(context=ctx); det.use_LSD = (len(sys.argv) <= 2 or sys.argv[2] == 'lsd'); det.line_length_thres = 15
for rep in range(3):
    t = time.perf_counter(); out = det.detect_filter_lines_batch(imgs, cap=1024); dt = time.perf_counter() - t
    print("frames %d  %.1f ms  %.0f frames/s  lines/frame %.1f" % (F, dt * 1e3, F / dt, np.mean([len(o) for o in out])))
