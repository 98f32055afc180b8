"""Per-call (unbatched) use of the operators, as a MATLAB loop calling the MEX gateways would: one
detectSIFTFeatures+extractFeatures per image, five matchFeatures, triangulate, estworldpose per frame."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import vo_b200
from vo_b200 import synth, vo
ctx = vo_b200.Context(0)
left, right = synth.shift_stream(24, seed=20260)
g = vo.VisualOdometry(synth.KITTI_P0, synth.KITTI_P1, vo.CudaOps(ctx=ctx, seed=1))
for i in range(4):
    g.step(left[i], right[i])
t0 = time.time()
for i in range(4, 24):
    g.step(left[i], right[i])
dt = (time.time() - t0) / 20
t1 = time.time()
for i in range(20):
    vo_b200.detectSIFTFeatures(left[i], ctx=ctx)
ds = (time.time() - t1) / 20
f = [vo_b200.detectSIFTFeatures(left[i], ctx=ctx)._features for i in range(2)]
t2 = time.time()
for i in range(20):
    vo_b200.matchFeatures(f[0], f[1], ctx=ctx)
dm = (time.time() - t2) / 20
print(f"per-call loop: {1e3 * dt:.2f} ms per stereo frame ({1 / dt:.0f} frames/s); vo_sift {1e3 * ds:.2f} ms per 1241x376 image "
      f"({len(f[0])} keypoints); vo_match {1e3 * dm:.2f} ms for {len(f[0])} x {len(f[1])}")
