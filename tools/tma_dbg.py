import sys; sys.path.insert(0,".")
import numpy as np, cube_slam_b200 as cs
from cube_slam_b200 import synthetic as S
imgs, Ts, boxes, lines, K = S.make_batch(61, 1, 640, 480, 2, poisson=False)
ctx=cs.Context(0,640,480,1,16,4096); ctx.set_calibration(K)
try:
    out,counts=ctx.detect_batch_host(imgs,Ts,boxes,lines,cs.default_params())
    print("ok",counts)
except Exception as e:
    print("ERR",e)
