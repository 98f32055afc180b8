"""Device-decode pipeline: frames/s over (frame-loop slots, decode slots) on a 1153-frame file list."""
import os, sys, tempfile, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cv2, torch, numpy as np
from vo_b200 import io, synth
import bench
l, r, _ = bench.street_frames(385)
d = tempfile.mkdtemp()
lf, rf = [], []
for i in range(385):
    for k, (a, lst) in enumerate(((l, lf), (r, rf))):
        p = os.path.join(d, f"{k}_{i:04d}.png"); cv2.imwrite(p, a[i]); lst.append(p)
lf3, rf3 = lf + lf[::-1] + lf, rf + rf[::-1] + rf          # 1155 frames (content does not matter for the rate)
for depth, dec in ((3, 6), (3, 10), (3, 16), (2, 10), (4, 12), (3, 24)):
    pipe = io.DevicePngPipeline(376, 1241, batch=32, depth=depth, decoders=dec)
    pipe.run(lf, rf, synth.KITTI_P0, synth.KITTI_P1, seed=1)
    t0 = time.perf_counter()
    pipe.run(lf3, rf3, synth.KITTI_P0, synth.KITTI_P1, seed=1)
    dt = time.perf_counter() - t0
    pipe.close()
    print(f"frame slots {depth} decode slots {dec}: {(len(lf3) - 1) / dt:.0f} frames/s", flush=True)
