"""Per-gateway latency through the stand-in MATLAB host (what a MATLAB loop calling the MEX files one by one would see)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from vo_b200 import mexhost, synth
import bench
l, r, _ = bench.street_frames(8)
h = mexhost.Host()
def t(f, n=20):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n): out = f()
    return 1e3 * (time.perf_counter() - t0) / n, out
ms, o = t(lambda: h.call("vo_sift_mex", 2, l[0])); d0, p0 = o
print(f"vo_sift_mex(I)                 {ms:.2f} ms  ({len(d0)} keypoints)")
ms, o8 = t(lambda: h.call("vo_sift_mex", 8, np.stack([l[0], r[0]], axis=2)))
print(f"vo_sift_mex(cat(3, lf, rf))    {ms:.2f} ms")
ms, _ = t(lambda: h.call("vo_sift_mex", 8, np.stack([l[0], r[0], l[1], r[1], l[2], r[2], l[3], r[3]], axis=2)), 10)
print(f"vo_sift_mex(8 images)          {ms:.2f} ms")
d1, p1 = h.call("vo_sift_mex", 2, r[0])
ms, pr = t(lambda: h.call("vo_match_mex", 1, d0, d1))
print(f"vo_match_mex                   {ms:.2f} ms  ({len(pr[0])} pairs)")
pairs = pr[0].astype(np.int64) - 1
a, b = p0[pairs[:, 0]].astype(np.float64), p1[pairs[:, 1]].astype(np.float64)
ms, xyz = t(lambda: h.call("vo_triangulate_mex", 1, a, b, synth.KITTI_P0, synth.KITTI_P1))
print(f"vo_triangulate_mex             {ms:.2f} ms  ({len(a)} points)")
ms, _ = t(lambda: h.call("vo_p3p_mex", 3, a, xyz[0], synth.KITTI_K4, "Seed", np.uint64(3)))
print(f"vo_p3p_mex                     {ms:.2f} ms")
# the host's own share: building the mxArray inputs
t0 = time.perf_counter()
for _ in range(20): x = h.mx(d0); h.free(x)
print(f"mx(desc) copy                  {1e3 * (time.perf_counter() - t0) / 20:.2f} ms")
