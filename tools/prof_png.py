"""Profiling target: device-side PNG decode of one batch (66 images of 1241x376 written by OpenCV)."""
import os, sys, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cv2, torch, numpy as np
import vo_b200
from vo_b200 import io
import bench
l, r, _ = bench.street_frames(33)
d = tempfile.mkdtemp()
paths = []
for i in range(33):
    for k, a in enumerate((l, r)):
        p = os.path.join(d, f"{k}_{i:04d}.png"); cv2.imwrite(p, a[i]); paths.append(p)
ctx = vo_b200.Context(0)
buf = torch.empty((66, 376, 1241), dtype=torch.uint8, device="cuda")
for _ in range(2):
    io.read_batch_dev(paths, 376, 1241, buf, ctx)
ctx.profile_enable(True)
io.read_batch_dev(paths, 376, 1241, buf, ctx)
print(ctx.profile())
